"""1-D operators of non-conforming (mortar) interfaces and the mortar surface metrics (host, init time).

Follows the host FLEXI code in the reference tree (GALAEXI has no GPU mortar path), paths relative
to /root/reference/src:
  * mortar/mortar.f90:111-187   MortarBasis_BigToSmall  M_0_1, M_0_2 (interpolation [-1,1] -> [-1,0], [0,1])
  * mortar/mortar.f90:195-256   MortarBasis_SmallToBig  M_1_0, M_2_0 (projection, without the interval Jacobian 1/2:
                                the small-side surface element already carries it, mortar_metrics.f90)
  * mortar/mortar.f90:88-101    mean-value self check
  * mortar/mortar_metrics.f90   Mortar_CalcSurfMetrics: big-side Ja / xGP interpolated to the small sides

Storage: like the reference, the four matrices are kept TRANSPOSED ("ATTENTION" note mortar.f90:178-184), numpy index
== Fortran index: ``M_0_1[l, p]`` multiplies big-side node l for small-side node p; ``M_1_0[l, p]`` multiplies
small-side node l for big-side node p.
"""
from __future__ import annotations

import numpy as np

from . import basis as bs


def mortar_basis_big_to_small(N: int, node_type: str):
    xi, _, wb = bs.get_nodes_and_weights(N, node_type)
    M_0_1 = bs.initialize_vandermonde(xi, wb, 0.5 * (xi - 1.0))
    M_0_2 = bs.initialize_vandermonde(xi, wb, 0.5 * (xi + 1.0))
    return np.ascontiguousarray(M_0_1.T), np.ascontiguousarray(M_0_2.T)


def mortar_basis_small_to_big(N: int, node_type: str):
    xi, _, _ = bs.get_nodes_and_weights(N, node_type)
    xg, wg, _ = bs.get_nodes_and_weights(N, bs.NODETYPE_G)
    VGP = bs.get_vandermonde(N, node_type, N, bs.NODETYPE_G)
    VGP = np.diag(wg) @ VGP
    n = N + 1
    Vleg, Vphi1, Vphi2 = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            Vleg[i, j] = bs.legendre_poly_and_deriv(j, float(xi[i]))[0]
            Vphi1[i, j] = bs.legendre_poly_and_deriv(j, 0.5 * (float(xg[i]) - 1.0))[0]
            Vphi2[i, j] = bs.legendre_poly_and_deriv(j, 0.5 * (float(xg[i]) + 1.0))[0]
    M_1_0 = Vleg @ (Vphi1.T @ VGP)
    M_2_0 = Vleg @ (Vphi2.T @ VGP)
    return np.ascontiguousarray(M_1_0.T), np.ascontiguousarray(M_2_0.T)


def init_mortar(N: int, node_type: str) -> dict:
    """InitMortar (mortar.f90:49-105) incl. the mean-value check 0.5*(0.5+1.5)=1 (the reference runs it for Gauss
    nodes only; it holds for Gauss-Lobatto as well because the check integrates with Gauss weights after an exact
    change of basis -- evaluated here for both)."""
    M_0_1, M_0_2 = mortar_basis_big_to_small(N, node_type)
    M_1_0, M_2_0 = mortar_basis_small_to_big(N, node_type)
    xi, w, wb = bs.get_nodes_and_weights(N, node_type)
    t1, t2 = np.full(N + 1, 0.5), np.full(N + 1, 1.5)
    err = abs(0.25 * np.sum((M_1_0.T @ t1 + M_2_0.T @ t2) * w) - 1.0)
    if err > 100.0 * bs.EPS * 100:
        raise RuntimeError(f"problems in building Mortar {err}")
    return dict(M_0_1=M_0_1, M_0_2=M_0_2, M_1_0=M_1_0, M_2_0=M_2_0)


def mortar_surf_metrics(mtype: int, N: int, node_type: str, Ja_face: np.ndarray, xGP_face: np.ndarray):
    """Mortar_CalcSurfMetrics: Ja_face[q,p,d,c], xGP_face[q,p,c] of one big side -> lists (4 or 2 entries, index
    iMortar-1) of small-side Mortar_Ja[q,p,d,c] and Mortar_xGP[q,p,c]."""
    A1, A2 = mortar_basis_big_to_small(N, node_type)
    V = [A1.T, A2.T]                       # un-transposed again: V[out, in]
    Vh = [0.5 * V[0], 0.5 * V[1]]
    ja, xg = {}, {}
    if mtype == 1:
        for iNb in range(2):
            ja2 = np.einsum("Pp,qpdc->qPdc", Vh[iNb], Ja_face)
            xg2 = np.einsum("Pp,qpc->qPc", V[iNb], xGP_face)
            for jNb in range(2):
                ind = iNb + 2 * jNb
                ja[ind] = np.einsum("Qq,qpdc->Qpdc", Vh[jNb], ja2)
                xg[ind] = np.einsum("Qq,qpc->Qpc", V[jNb], xg2)
        cnt = 4
    elif mtype == 2:
        for jNb in range(2):
            ja[jNb] = np.einsum("Qq,qpdc->Qpdc", Vh[jNb], Ja_face)
            xg[jNb] = np.einsum("Qq,qpc->Qpc", V[jNb], xGP_face)
        cnt = 2
    elif mtype == 3:
        for iNb in range(2):
            ja[iNb] = np.einsum("Pp,qpdc->qPdc", Vh[iNb], Ja_face)
            xg[iNb] = np.einsum("Pp,qpc->qPc", V[iNb], xGP_face)
        cnt = 2
    else:
        raise ValueError(mtype)
    return [ja[i] for i in range(cnt)], [xg[i] for i in range(cnt)]
