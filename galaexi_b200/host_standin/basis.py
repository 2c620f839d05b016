"""1-D polynomial basis: nodes, weights, differentiation and DG operator building blocks.

Host-side (init-time) mirror of the reference's interpolation layer. Everything here runs once
per run on the host and produces the small operator tables the device kernels consume
(`D_T, D_Hat_T, DVolSurf, L_Minus/Plus, L_HatMinus/Plus`).

Follows (reference paths relative to /root/reference):
  * src/interpolation/basis.f90:195-235   LegendrePolynomialAndDerivative
  * src/interpolation/basis.f90:417-484   LegendreGaussNodesAndWeights
  * src/interpolation/basis.f90:534-629   qAndLEvaluation / LegGaussLobNodesAndWeights
  * src/interpolation/basis.f90:269-290   ChebyGaussLobNodesAndWeights
  * src/interpolation/basis.f90:638-655   BarycentricWeights
  * src/interpolation/basis.f90:664-686   PolynomialDerivativeMatrix
  * src/interpolation/basis.f90:720-752   LagrangeInterpolationPolys
  * src/interpolation/basis.f90:100-160   buildLegendreVdm
  * src/interpolation/interpolation.f90:199-316 GetNodesAndWeights, :324-395 GetVandermonde
  * src/dg/dg.f90:181-242                 InitDGbasis (D_Hat, DVolSurf, L_Hat)

The loops are kept scalar and in the reference's operation order so the tables agree with the
reference's unit-test goldens (unitTests/NodesAndWeights.bin, DerivativeMatrix.bin) to a few ulp.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # PP_RealTolerance (src/globals/preprocessing.f90:25)
PI = math.acos(-1.0)  # PP_Pi (src/globals/preprocessing.f90:26)

NODETYPE_G = "GAUSS"
NODETYPE_GL = "GAUSS-LOBATTO"
NODETYPE_CL = "CHEBYSHEV-GAUSS-LOBATTO"
NODETYPE_VISU = "VISU"


def legendre_poly_and_deriv(n: int, x: float) -> tuple[float, float]:
    """Normalised Legendre polynomial L_n(x) and derivative (basis.f90:195-235)."""
    if n == 0:
        L, Ld = 1.0, 0.0
    elif n == 1:
        L, Ld = x, 1.0
    else:
        L_nm2, L_nm1 = 1.0, x
        Ld_nm1 = 1.0
        L = Ld = 0.0
        for i in range(2, n + 1):
            L = (float(2 * i - 1) * x * L_nm1 - float(i - 1) * L_nm2) / float(i)
            Ld = i * L_nm1 + x * Ld_nm1
            L_nm2 = L_nm1
            L_nm1 = L
            Ld_nm1 = Ld
    s = math.sqrt(float(n) + 0.5)
    return L * s, Ld * s


def legendre_gauss_nodes_and_weights(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Legendre-Gauss nodes and weights by Newton iteration (basis.f90:417-484)."""
    x = np.zeros(n + 1)
    w = np.zeros(n + 1)
    if n == 0:
        x[0] = 0.0
        w[0] = 2.0
        return x, w
    if n == 1:
        x[0] = -math.sqrt(1.0 / 3.0)
        x[1] = -x[0]
        w[:] = 1.0
        return x, w
    cheb = 2.0 * math.atan(1.0) / float(n + 1)
    tol = 1.0e-15
    for i in range((n + 1) // 2):
        xi = -math.cos(cheb * float(2 * i + 1))
        for _ in range(11):
            L, Ld = legendre_poly_and_deriv(n + 1, xi)
            dx = -L / Ld
            xi = xi + dx
            if abs(dx) < tol * abs(xi):
                break
        else:
            raise RuntimeError("Legendre Gauss nodes could not be computed up to desired precision")
        L, Ld = legendre_poly_and_deriv(n + 1, xi)
        x[i] = xi
        x[n - i] = -xi
        w[i] = (2.0 * n + 3) / ((1.0 - xi * xi) * Ld * Ld)
        w[n - i] = w[i]
    if n % 2 == 0:
        x[n // 2] = 0.0
        L, Ld = legendre_poly_and_deriv(n + 1, 0.0)
        w[n // 2] = (2.0 * n + 3) / (Ld * Ld)
    return x, w


def _q_and_l_evaluation(n: int, x: float) -> tuple[float, float, float]:
    """q = L_{n+1} - L_{n-1}, q', L_n (basis.f90:534-559)."""
    L_nm2, L_nm1 = 1.0, x
    L = x
    for i in range(2, n + 1):
        L = (float(2 * i - 1) * x * L_nm1 - float(i - 1) * L_nm2) / float(i)
        L_nm2 = L_nm1
        L_nm1 = L
    q = float(2 * n + 1) / float(n + 1) * (x * L - L_nm2)
    qder = float(2 * n + 1) * L
    return q, qder, L


def legendre_gauss_lobatto_nodes_and_weights(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Legendre-Gauss-Lobatto nodes and weights (basis.f90:567-629)."""
    if n < 1:
        raise ValueError("Gauss-Lobatto needs N >= 1")
    x = np.zeros(n + 1)
    w = np.zeros(n + 1)
    x[0], x[n] = -1.0, 1.0
    w[0] = 2.0 / float(n * (n + 1))
    w[n] = w[0]
    tol = 1.0e-15
    if n > 1:
        cont1 = PI / float(n)
        cont2 = 3.0 / (float(8 * n) * PI)
        for i in range(1, (n + 1) // 2):
            xi = -math.cos(cont1 * (float(i) + 0.25) - cont2 / (float(i) + 0.25))
            for _ in range(11):
                q, qder, L = _q_and_l_evaluation(n, xi)
                dx = -q / qder
                xi = xi + dx
                if abs(dx) < tol * abs(xi):
                    break
            else:
                raise RuntimeError("Legendre Gauss Lobatto nodes could not be computed")
            q, qder, L = _q_and_l_evaluation(n, xi)
            x[i] = xi
            x[n - i] = -xi
            w[i] = w[0] / (L * L)
            w[n - i] = w[i]
    if n % 2 == 0:
        x[n // 2] = 0.0
        q, qder, L = _q_and_l_evaluation(n, 0.0)
        w[n // 2] = w[0] / (L * L)
    return x, w


def chebyshev_gauss_lobatto_nodes_and_weights(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Chebyshev-Gauss-Lobatto nodes and weights (basis.f90:269-290)."""
    x = np.array([-math.cos(i / float(n) * PI) for i in range(n + 1)])
    w = np.full(n + 1, PI / float(n))
    w[0] *= 0.5
    w[n] *= 0.5
    return x, w


def barycentric_weights(x: np.ndarray) -> np.ndarray:
    """Barycentric interpolation weights (basis.f90:638-655)."""
    n = len(x) - 1
    wb = np.ones(n + 1)
    for i in range(1, n + 1):
        for j in range(i):
            wb[j] = wb[j] * (x[j] - x[i])
            wb[i] = wb[i] * (x[i] - x[j])
    return 1.0 / wb


def get_nodes_and_weights(n: int, node_type: str) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Nodes, weights and barycentric weights for a node type (interpolation.f90:199-316)."""
    t = node_type.strip().lower()
    if t == "gauss":
        x, w = legendre_gauss_nodes_and_weights(n)
    elif t == "gauss-lobatto":
        x, w = legendre_gauss_lobatto_nodes_and_weights(n)
    elif t == "chebyshev-gauss-lobatto":
        x, w = chebyshev_gauss_lobatto_nodes_and_weights(n)
    elif t == "visu":
        x = np.array([2.0 * float(i) / float(n) - 1.0 for i in range(n + 1)])
        w = np.full(n + 1, 2.0 / float(n))
        w[0] *= 0.5
        w[n] *= 0.5
    elif t == "visu_inner":
        x = np.array([1.0 / float(n + 1) + 2.0 * float(i) / float(n + 1) - 1.0 for i in range(n + 1)])
        w = np.full(n + 1, 2.0 / float(n + 1))
    else:
        raise ValueError(f'NodeType "{node_type}" in get_nodes_and_weights not found!')
    return x, w, barycentric_weights(x)


def polynomial_derivative_matrix(x: np.ndarray) -> np.ndarray:
    """Differentiation matrix D(iGP,iLagrange) (basis.f90:664-686)."""
    n = len(x) - 1
    wb = barycentric_weights(x)
    D = np.zeros((n + 1, n + 1))
    for il in range(n + 1):
        for ig in range(n + 1):
            if il != ig:
                D[ig, il] = wb[il] / (wb[ig] * (x[ig] - x[il]))
                D[ig, ig] = D[ig, ig] - D[ig, il]
    return D


def almost_equal(x: float, y: float) -> bool:
    """ALMOSTEQUAL (basis.f90:694-710)."""
    if x == 0.0 or y == 0.0:
        return abs(x - y) <= 2.0 * EPS
    return abs(x - y) <= EPS * abs(x) and abs(x - y) <= EPS * abs(y)


def lagrange_interpolation_polys(x: float, xgp: np.ndarray, wbary: np.ndarray) -> np.ndarray:
    """All Lagrange polynomials at x (basis.f90:720-752)."""
    n = len(xgp) - 1
    L = np.zeros(n + 1)
    hit = False
    for i in range(n + 1):
        if almost_equal(x, float(xgp[i])):
            L[i] = 1.0
            hit = True
    if hit:
        return L
    s = 0.0
    for i in range(n + 1):
        L[i] = wbary[i] / (x - xgp[i])
        s = s + L[i]
    for i in range(n + 1):
        L[i] = L[i] / s
    return L


def initialize_vandermonde(x_in: np.ndarray, wbary_in: np.ndarray, x_out: np.ndarray) -> np.ndarray:
    """Nodal Vandermonde Vdm(0:N_out,0:N_in) (basis.f90:168-186)."""
    V = np.zeros((len(x_out), len(x_in)))
    for i, xo in enumerate(x_out):
        V[i, :] = lagrange_interpolation_polys(float(xo), x_in, wbary_in)
    return V


def build_legendre_vdm(x_in: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Modal<->nodal Vandermonde (basis.f90:100-160).

    The reference inverts with LAPACK by default; no LAPACK binding is needed here because
    numpy.linalg.inv is LAPACK's dgetrf/dgetri as well.
    """
    n = len(x_in) - 1
    V = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(n + 1):
            V[i, j], _ = legendre_poly_and_deriv(j, float(x_in[i]))
    sV = np.linalg.inv(V)
    chk = abs(np.sum(np.abs(sV @ V)) / (n + 1.0) - 1.0)
    if chk > 10.0 * EPS * 100:
        raise RuntimeError(f"problems in MODAL<->NODAL Vandermonde {chk}")
    return V, sV


def get_vandermonde(n_in: int, type_in: str, n_out: int, type_out: str, modal: bool = False) -> np.ndarray:
    """Vdm_In_Out(0:N_out,0:N_in) (interpolation.f90:324-395)."""
    if type_in.strip().lower() == type_out.strip().lower() and n_in == n_out:
        return np.eye(n_in + 1)
    x_in, _, wb_in = get_nodes_and_weights(n_in, type_in)
    x_out, _, _ = get_nodes_and_weights(n_out, type_out)
    if modal and n_out < n_in:
        _, sV_in = build_legendre_vdm(x_in)
        V_out, _ = build_legendre_vdm(x_out)
        return V_out[: n_out + 1, : n_out + 1] @ sV_in[: n_out + 1, : n_in + 1]
    return initialize_vandermonde(x_in, wb_in, x_out)


@dataclass
class DGBasis:
    """Operator tables of InitInterpolationBasis + InitDGbasis (dg.f90:181-242).

    Matrices are stored with the reference's index meaning, numpy index = Fortran index:
    ``D[i, l]`` = Fortran ``D(i,l)``. ``D_T = D^T``, ``D_Hat = -Minv D^T M``, ``D_Hat_T = D_Hat^T``,
    ``DVolSurf = D_T`` with the two corner corrections of the strong split form.
    """

    N: int
    node_type: str
    xGP: np.ndarray
    wGP: np.ndarray
    wBary: np.ndarray
    L_Minus: np.ndarray
    L_Plus: np.ndarray
    D: np.ndarray
    D_T: np.ndarray
    D_Hat: np.ndarray
    D_Hat_T: np.ndarray
    DVolSurf: np.ndarray
    L_HatMinus: np.ndarray
    L_HatPlus: np.ndarray


def polynomial_mass_matrix(N: int, x: np.ndarray, w: np.ndarray, exact_mm: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """PolynomialMassMatrix (basis.f90:762-807): diag(wGP) and its inverse; with EXACT_MM on Gauss-Lobatto nodes the exact mass
    matrix and its inverse as rank-one updates with the Legendre polynomial of degree N (Teukolsky, JCP 2015)."""
    M = np.diag(np.asarray(w, dtype=np.float64))
    Minv = np.diag(1.0 / np.asarray(w, dtype=np.float64))
    if exact_mm:
        hN = 2.0 / (2.0 * N + 1.0)
        gammaN = 2.0 / N
        norm = legendre_poly_and_deriv(N, 1.0)[0]
        alpha = (hN - gammaN) / (gammaN ** 2 * norm ** 2)
        beta = -(hN - gammaN) / (gammaN * hN * norm ** 2)
        pN = np.array([legendre_poly_and_deriv(N, float(xi))[0] for xi in x])
        for i in range(N + 1):
            for j in range(N + 1):
                M[i, j] = M[i, j] + alpha * w[i] * w[j] * pN[i] * pN[j]
                Minv[i, j] = Minv[i, j] + beta * pN[i] * pN[j]
    return M, Minv


def init_dg_basis(N: int, node_type: str, exact_mm: bool = False) -> DGBasis:
    """exact_mm: the build option FLEXI_EXACT_MASSMATRIX (-DEXACT_MM, src/CMakeLists.txt:150-156), Gauss-Lobatto nodes only."""
    if exact_mm and node_type.strip().upper() != NODETYPE_GL:
        raise ValueError("FLEXI_EXACT_MASSMATRIX only works on FLEXI_NODETYPE==GAUSS-LOBATTO points.")
    x, w, wb = get_nodes_and_weights(N, node_type)
    L_plus = lagrange_interpolation_polys(1.0, x, wb)
    L_minus = lagrange_interpolation_polys(-1.0, x, wb)
    D = polynomial_derivative_matrix(x)
    D_T = D.T.copy()
    M, Minv = polynomial_mass_matrix(N, x, w, exact_mm)
    # D_Hat = -MATMUL(Minv, MATMUL(TRANSPOSE(D), M))  (dg.f90:222)
    D_Hat = -(Minv @ (D.T @ M))
    D_Hat_T = D_Hat.T.copy()
    DVolSurf = D_T.copy()
    DVolSurf[0, 0] = DVolSurf[0, 0] + 1.0 / (2.0 * w[0])
    DVolSurf[N, N] = DVolSurf[N, N] - 1.0 / (2.0 * w[N])
    L_HatPlus = Minv @ L_plus
    L_HatMinus = Minv @ L_minus
    return DGBasis(N, node_type, x, w, wb, L_minus, L_plus, D, D_T, D_Hat, D_Hat_T, DVolSurf,
                   L_HatMinus, L_HatPlus)
