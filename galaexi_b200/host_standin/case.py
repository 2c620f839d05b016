"""Assemble everything the Fortran host hands to the hot path at init time.

In the reference this is the sequence InitInterpolation -> InitMesh -> InitEquation -> InitDG ->
InitLifting -> InitTimeDisc (src/flexilib.f90:188-214); the result is the set of module-global arrays
that `DGTimeDerivative_weakForm` reads. Here the same arrays are collected in a
`Case` whose buffers are in the reference's memory layout, ready to be passed through the C ABI
(include/dgx.h) by the Python driver that stands in for the Fortran host.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import basis as bs
from . import equation as eq
from . import filter as fl
from . import mappings as mp
from . import mesh as ms
from . import metrics as mt
from . import mortar as mo
from . import timedisc as td

SPLIT_IDS = {None: -1, "NONE": -1, "SD": 0, "MO": 1, "DU": 2, "KG": 3, "PI": 4}      # SPLIT_DG
RIEMANN_IDS = {"LF": 0, "ROE": 1, "ROEL2": 2, "ROEENTROPYFIX": 3, "HLL": 4, "HLLC": 5, "HLLE": 6, "HLLEM": 7,
               "FLUXAVERAGE": 9, "AVG": 9, "CENTRAL": 9}   # RIEMANN (src/CMakeLists.txt:97-130); 9: Riemann_FluxAverage riemann.f90:1239


@dataclass
class Case:
    N: int
    node_type: str
    basis: bs.DGBasis
    mesh: ms.Mesh
    geo: dict
    maps: dict
    eos: eq.Eos
    RefStatePrim: np.ndarray
    BCSides: np.ndarray
    timedisc: td.TimeDisc
    split: int
    riemann: int
    parabolic: bool
    hopr: dict = field(repr=False, default=None)
    lifting: int = 1            # 1: BR1 (GALAEXI default, src/CMakeLists.txt:161-162), 2: BR2 (host FLEXI, lifting_br2.t90)
    etaBR2: float = 2.0         # lifting.f90:86-91
    etaBR2_wall: float = -1.0   # -1: use etaBR2 (lifting.f90:89-91)
    mortar: dict = field(repr=False, default=None)   # M_0_1, M_0_2, M_1_0, M_2_0 (mortar/mortar.f90)
    FilterMat: np.ndarray = None                     # (N+1,N+1) or None: FilterType 0 (filter/filter.f90)
    SpongeMat: np.ndarray = None                     # [e,k,j,i] damping*sigma/sJ (sponge.f90:259-457) or None
    SpBaseFlow: np.ndarray = None                    # [e,k,j,i,5] initial sponge base flow
    doWeakLifting: bool = False                      # lifting.f90:81-85, 139-141 (BR2 is always strong)
    doConservativeLifting: bool = False
    IniExactFunc: int = 0                            # selects the source term of CalcSource (exactfunc.f90:665-926): 4 or 0
    AdvVel: tuple = (0.0, 0.0, 0.0)
    exact_mm: bool = False                           # FLEXI_EXACT_MASSMATRIX (GL nodes, exact mass matrix): see op_node_type
    # overintegration of JU_t (dg/overintegration.f90:92-165): 0 none, 1 cut-off filter, 2 conservative cut-off
    OverintegrationType: int = 0
    NUnder: int = -1
    OverintegrationMat: np.ndarray = None            # (N+1,N+1), type 1
    Vdm_N_NUnder: np.ndarray = None                  # (NUnder+1,N+1), type 2
    Vdm_NUnder_N: np.ndarray = None                  # (N+1,NUnder+1), type 2
    sJNUnder: np.ndarray = None                      # [e,kU,jU,iU], type 2

    @property
    def op_node_type(self) -> int:
        """PP_NodeType as the operator applications see it (dgx_config.nodeType): 1 = interpolating prolongation and full
        L_Hat surface integral, 2 = Gauss-Lobatto collocation short cuts. The reference compiles the first form for
        `PP_NodeType==1 || (PP_NodeType==2 && defined(EXACT_MM))` (surfint.t90:74-104, lifting_br1.t90:238-270); on GL nodes
        L_Minus / L_Plus are unit vectors, so the interpolating prolongation returns the boundary node values exactly."""
        return 2 if (self.node_type == bs.NODETYPE_GL and not self.exact_mm) else 1

    @property
    def n(self):
        return self.N + 1

    @property
    def nDOF(self):
        return self.mesh.nElems * self.n ** 3


def build_case(hopr: dict, N: int, node_type: str = bs.NODETYPE_GL, split: str | None = "PI",
               riemann: str = "RoeEntropyFix", parabolic: bool = True, eos: eq.Eos | None = None,
               refstates=((1.0, 1.0, 0.0, 0.0, 17194.8345650329),), user_bcs: dict | None = None,
               nProcs: int = 1, myRank: int = 0, timedisc: str = "carpenterrk4-5", CFLScale: float = 0.9,
               DFLScale: float = 0.9, useCurveds: bool = True, crossProductMetrics: bool = False,
               lifting: str = "br1", etaBR2: float = 2.0, etaBR2_wall: float = -1.0,
               FilterType: int | str = 0, NFilter: int | None = None, HestFilterParam=(36.0, 12.0, 1.0),
               IniExactFunc: int = 0, AdvVel=(0.0, 0.0, 0.0), doWeakLifting: bool = False,
               doConservativeLifting: bool = False, exact_mm: bool = False,
               OverintegrationType: int | str = 0, NUnder: int | None = None) -> Case:
    eos = eos or eq.Eos()
    node_type = node_type.upper()
    split_id = SPLIT_IDS[split.upper() if isinstance(split, str) else split]
    riem_id = RIEMANN_IDS[riemann.upper()]
    if split_id >= 0 and node_type != bs.NODETYPE_GL:
        # splitflux.f90:116-119
        raise ValueError("Wrong Pointset: Gauss-Lobatto-Points are mandatory for using SplitDG !")
    if split_id >= 0 and riem_id in (4, 5, 6, 7):
        # src/CMakeLists.txt:108-127
        raise ValueError("HLL-type Riemann solvers are not available with SplitDG")
    if split_id < 0 and riem_id == 9:
        # riemann.f90:1236-1252: Riemann_FluxAverage only exists inside #ifdef SPLIT_DG
        raise ValueError("The flux-average Riemann solver is only available with SplitDG")
    if exact_mm and split_id >= 0:
        raise ValueError("EXACT_MM with SplitDG is not built (the split-form kernels use the collocation surface integral)")
    basis = bs.init_dg_basis(N, node_type, exact_mm)
    mesh = ms.prepare_mesh(hopr, nProcs=nProcs, myRank=myRank, useCurveds=useCurveds, user_bcs=user_bcs)
    geo = mt.calc_metrics(mesh, N, node_type, crossProductMetrics=crossProductMetrics, hopr=hopr)
    maps = mp.build_mappings(N)
    refprim = eq.init_bc_refstates(eq.refstate_prim(refstates, eos), mesh.BoundaryType)
    bcs = eq.bc_sides(mesh)
    # overintegration.f90:92-165 InitOverintegration
    ot = {"none": 0, "cutoff": 1, "conscutoff": 2}.get(str(OverintegrationType).lower(), OverintegrationType)
    ot = int(ot)
    if ot not in (0, 1, 2):
        raise ValueError("Unknown OverintegrationType!")
    nunder, omat, vdn, vup, sjn = N, None, None, None, None
    if ot > 0:
        if NUnder is None:
            raise ValueError("NUnder needed for OverintegrationType cutoff / conscutoff")
        nunder = int(NUnder)
    if ot == 1:
        omat = fl.filter_matrix(N, node_type, "cutoff", nunder)     # Vdm_Leg * diag(1 up to NUnder) * sVdm_Leg (:120-131)
    if ot == 2:
        if nunder < N:
            vdn = bs.get_vandermonde(N, node_type, nunder, node_type, modal=True)
            vup = bs.get_vandermonde(nunder, node_type, N, node_type, modal=True)
            djr = mt.det_jac_ref(mesh.NodeCoords, mesh.NGeo, node_type)          # on NGeoRef = 3 NGeo
            v3 = bs.get_vandermonde(3 * mesh.NGeo, node_type, nunder, node_type, modal=True)
            sjn = 1.0 / mt.change_basis_volume(v3, djr[..., None])[..., 0]
        else:
            ot, nunder = 0, N      # "Overintegration is disabled for NUnder >= N" (:158-160)
    # timedisc_func.f90:171-173: NEff = MIN(PP_N,NFilter,NUnder), FilterType > 2 has no time step effect
    ftype = {"none": 0, "cutoff": 1, "modal": 2, "laf": 3}.get(str(FilterType).lower(), FilterType)
    neff = min(N, nunder, int(NFilter) if (NFilter is not None and int(ftype) in (1, 2)) else N)
    tdisc = td.set_timedisc(timedisc, N, node_type, CFLScale, DFLScale, overintegration=ot > 0, exact_mm=exact_mm, NEff=neff)
    lift_id = {"br1": 1, "br2": 2}[lifting.lower()]
    if etaBR2_wall == -1.0:
        etaBR2_wall = etaBR2   # lifting.f90:158-159
    return Case(N, node_type, basis, mesh, geo, maps, eos, refprim, bcs, tdisc, split_id, riem_id, bool(parabolic), hopr,
                lift_id, float(etaBR2), float(etaBR2_wall), mo.init_mortar(N, node_type),
                None if str(FilterType).lower() in ("0", "none") else fl.filter_matrix(N, node_type, FilterType, NFilter, HestFilterParam),
                None, None, bool(doWeakLifting), bool(doConservativeLifting) and not bool(doWeakLifting),
                int(IniExactFunc), tuple(float(v) for v in AdvVel), bool(exact_mm), ot, nunder, omat, vdn, vup, sjn)
