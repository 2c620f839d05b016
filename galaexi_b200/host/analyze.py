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
