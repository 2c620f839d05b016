"""Analysis basis of the reference (host, init time): the integration rule AnalyzeTestcase and CalcErrorNorms use.

Follows /root/reference/src/analyze/analyze.f90:
  * InitAnalyze        :120-197  NAnalyze = 2 (N+1) by default, Vol = sum wGPVol / sJ over the solution nodes
  * InitAnalyzeBasis   :236-272  Gauss-Lobatto nodes of degree NAnalyze, Vandermonde solution nodes -> analysis nodes,
                                 tensor-product weights wGPVolAnalyze
"""
from __future__ import annotations

import numpy as np

from . import basis as bs


def init_analyze_basis(N: int, node_type: str, NAnalyze: int | None = None):
    """Returns (NAnalyze, Vdm_GaussN_NAnalyze [NAnalyze+1, N+1], wAnalyze [NAnalyze+1])."""
    NA = 2 * (N + 1) if NAnalyze is None else int(NAnalyze)
    xGP, _, wBary = bs.get_nodes_and_weights(N, node_type)
    xiA, wA, _ = bs.get_nodes_and_weights(NA, bs.NODETYPE_GL)
    return NA, bs.initialize_vandermonde(xGP, wBary, xiA), wA


def volume(case) -> float:
    """Vol of this rank's elements (analyze.f90:160-168); the caller sums over the ranks."""
    w = case.basis.wGP
    W = w[:, None, None] * w[None, :, None] * w[None, None, :]
    return float(np.sum(W[None] / case.geo["sJ"]))


def bc_surfaces(case) -> np.ndarray:
    """Surf(nBCs) of this rank's sides (analyze.f90:168-194): sum wGPSurf SurfElem over the sides with AnalyzeSide = iBC;
    the caller sums over the ranks. Boundary conditions without sides get HUGE (the reference's guard against 0-division)."""
    m = case.mesh
    w = case.basis.wGP
    wS = w[:, None] * w[None, :]
    nBCs = int(m.BoundaryType.shape[0])
    S = np.zeros(nBCs)
    has = np.zeros(nBCs, dtype=bool)
    az = m.AnalyzeSide if m.AnalyzeSide is not None else np.concatenate([m.BC[:m.nBCSides], np.zeros(m.nSides - m.nBCSides, dtype=np.int64)])
    for s in range(m.nSides):
        b = int(az[s])
        if b == 0:
            continue
        has[b - 1] = True
        S[b - 1] += float(np.sum(wS * case.geo["SurfElem"][s]))
    S[~has] = np.finfo(np.float64).max
    return S


def calc_error_norms(case, U: np.ndarray, t: float, exact, NAnalyze: int | None = None, Vol: float | None = None,
                     reduce=None) -> tuple[np.ndarray, np.ndarray]:
    """CalcErrorNorms(Time,L_2_Error,L_Inf_Error) (analyze.f90:383-470): the solution U [e,k,j,i,var], the node coordinates and
    the Jacobian 1/sJ are interpolated to the NAnalyze Gauss-Lobatto points, ``exact(x, t)`` (ExactFunc with AnalyzeExactFunc) is
    evaluated there; L2 = sqrt(sum (U - U_exact)^2 wGPVolAnalyze J / Vol), Linf = max |U - U_exact|, per variable.
    ``reduce(sum_array, max_array)`` combines the partial results of several ranks (MPI_REDUCE SUM / MAX, :456-464)."""
    NA, V, wA = init_analyze_basis(case.N, case.node_type, NAnalyze)

    def up(X):
        Y = np.einsum("Ii,ekjic->ekjIc", V, X)
        Y = np.einsum("Jj,ekjIc->ekJIc", V, Y)
        return np.einsum("Kk,ekJIc->eKJIc", V, Y)
    Ua, xa, Ja = up(U), up(case.geo["Elem_xGP"]), up((1.0 / case.geo["sJ"])[..., None])[..., 0]
    w3 = wA[:, None, None] * wA[None, :, None] * wA[None, None, :]
    d = Ua - exact(xa, t)
    l2 = np.sum((w3[None] * Ja)[..., None] * d * d, axis=(0, 1, 2, 3))
    linf = np.abs(d).max(axis=(0, 1, 2, 3)) if d.size else np.full(U.shape[-1], -1.0e10)
    vol = volume(case) if Vol is None else Vol
    if reduce is not None:
        l2, linf = reduce(l2, linf)
    return np.sqrt(l2 / vol), linf


# nTGVvars = 15 columns of the *_TGVAnalysis file in the order of testcase/taylorgreenvortex/testcase.f90:497-498 (what
# dgx_analyze_tgv / DGSolver.AnalyzeTestcase return)
TGV_COLUMNS = ("Dissipation Rate Incompressible", "Dissipation Rate Compressible", "Ekin incomp", "Ekin comp", "Enstrophy comp",
               "DR_u", "DR_S", "DR_Sd", "DR_p", "Maximum Vorticity", "Mean Temperature", "uprime", "Mean Entropy", "ED_S", "ED_D")
