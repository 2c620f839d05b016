"""Modal filter matrix (host, init time), mirror of /root/reference/src/filter/filter.f90:
  * InitFilter :95-212  FilterType 1 (cut-off at NFilter) and 2 (Hesthaven exponential filter); FilterMat =
                        Vdm_Leg * diag * sVdm_Leg (:203)
  * HestFilter :219-248
FilterType 3 (LAF, an adaptive per-element blend with host-side state, :140-178) is not part of the device path here.
The matrix is applied to U at the start of every DGTimeDerivative_weakForm (dg/dg.f90:331, Filter :272-306)."""
from __future__ import annotations

import math

import numpy as np

from . import basis as bs

FILTERTYPE_NONE, FILTERTYPE_CUTOFF, FILTERTYPE_MODAL, FILTERTYPE_LAF = 0, 1, 2, 3


def filter_matrix(N: int, node_type: str, filter_type: int | str, NFilter: int | None = None,
                  HestFilterParam=(36.0, 12.0, 1.0)) -> np.ndarray:
    """FilterMat(0:N,0:N) with numpy index == Fortran index (``F[i, l]`` multiplies nodal value l for node i)."""
    ft = {"none": 0, "cutoff": 1, "modal": 2, "laf": 3}.get(str(filter_type).lower(), filter_type)
    ft = int(ft)
    diag = np.zeros(N + 1)
    if ft == FILTERTYPE_CUTOFF:
        if NFilter is None:
            raise ValueError("NFilter needed for the cut-off filter")
        diag[: int(NFilter) + 1] = 1.0
    elif ft == FILTERTYPE_MODAL:
        alpha, s, etac = float(HestFilterParam[0]), float(HestFilterParam[1]), float(HestFilterParam[2]) / float(N + 1)
        for iDeg in range(0, min(int(HestFilterParam[2]) - 1, N) + 1):
            diag[iDeg] = 1.0
        if alpha >= 0.0:
            for iDeg in range(int(HestFilterParam[2]), N + 1):
                eta = float(iDeg + 1) / float(N + 1)
                diag[iDeg] = math.exp(-alpha * ((eta - etac) / (1.0 - etac)) ** s)
    elif ft == FILTERTYPE_LAF:
        raise NotImplementedError("FilterType LAF keeps adaptive per-element state on the host (filter.f90:140-178); not supported")
    else:
        raise ValueError("FilterType unknown!")
    x, _, _ = bs.get_nodes_and_weights(N, node_type)
    Vdm_Leg, sVdm_Leg = bs.build_legendre_vdm(x)
    return (Vdm_Leg @ np.diag(diag)) @ sVdm_Leg
