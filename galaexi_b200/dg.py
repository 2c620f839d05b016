"""Host-side mirror of the reference's DG module interface over the C ABI (include/dgx.h).

The reference toolchain (nvfortran) is absent here, so this Python driver plays the role of the
Fortran host: it feeds the arrays built by ``galaexi_b200.host_standin`` (the stand-in for InitMesh /
InitInterpolation / InitEquation) through ``dgx_create`` and calls the same procedures, with the same
names and argument meaning, that ``src/timedisc`` calls in the reference:

    InitDG()                         dg/dg.f90:67          -> DGSolver(case)
    DGTimeDerivative_weakForm(t)     dg/dg.f90:255         -> DGSolver.DGTimeDerivative_weakForm(t)
    TimeStepByLSERKW2(t)             timedisc/timestep.f90:49 -> DGSolver.TimeStepByLSERKW2(t, dt)
    CalcTimeStep(errType)            calctimestep.f90:98   -> DGSolver.CalcTimeStep()
    FinalizeDG()                     dg/dg.f90:464         -> DGSolver.FinalizeDG()

There is no CPU fallback: if ``libdgx.so`` (hand-written sm_100a kernels) is missing or no B200 is
visible, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdgx.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

EXPORTS = ("dgx_create", "dgx_destroy", "dgx_last_error", "dgx_set_state", "dgx_get_state", "dgx_get_ut",
           "dgx_get_gradients", "dgx_get_face_array", "dgx_set_keep_gradients", "dgx_time_derivative", "dgx_rk_stage", "dgx_rk_step", "dgx_calc_timestep",
           "dgx_analyze_tgv", "dgx_calc_bulk_velocity", "dgx_calc_body_forces", "dgx_calc_wall_velocity", "dgx_set_channel_forcing", "dgx_temp_filter_time_deriv", "dgx_get_baseflow", "dgx_sync", "dgx_run_steps", "dgx_get_dt_history", "dgx_step_graph_active", "dgx_profile_stage", "dgx_nccl_unique_id", "dgx_launch_count", "dgx_sizeof_config", "dgx_halo_plan")


class DgxConfig(C.Structure):
    """Mirror of ``struct dgx_config`` (include/dgx.h)."""
    _fields_ = (
        [(k, C.c_int) for k in ("N", "nodeType", "splitDG", "riemann", "parabolic", "viscLaw", "nElems", "nSides",
                                "nBCSides", "firstInnerSide", "lastInnerSide", "firstMPISide_MINE", "lastMPISide_MINE",
                                "firstMPISide_YOUR", "lastMPISide_YOUR")]
        + [("EOS_Vars", C.c_double * 8), ("nRefState", C.c_int), ("RefStatePrim", _dp), ("BCSides", _ip)]
        + [(k, _dp) for k in ("D_T", "D_Hat_T", "DVolSurf", "L_Minus", "L_Plus", "L_HatMinus", "L_HatPlus")]
        + [(k, _ip) for k in ("ElemToSide", "S2V2", "S2V2_inv")]
        + [(k, _dp) for k in ("Metrics_fTilde", "Metrics_gTilde", "Metrics_hTilde", "sJ", "NormVec", "TangVec1",
                              "TangVec2", "SurfElem")]
        + [("nRKStages", C.c_int), ("RKA", _dp), ("RKb", _dp), ("RKc", _dp), ("CFLScale", C.c_double),
           ("DFLScale", C.c_double), ("myRank", C.c_int), ("nRanks", C.c_int), ("nNbProcs", C.c_int)]
        + [(k, _ip) for k in ("NbProc", "nMPISides_MINE_Proc", "nMPISides_YOUR_Proc", "offsetMPISides_MINE",
                              "offsetMPISides_YOUR")]
        + [("ncclUniqueId", C.c_char_p), ("device", C.c_int)]
        + [("lifting", C.c_int), ("etaBR2", C.c_double), ("etaBR2_wall", C.c_double)]
        + [(k, C.c_int) for k in ("nMortarSides", "firstMortarInnerSide", "lastMortarInnerSide", "firstMortarMPISide",
                                  "lastMortarMPISide")]
        + [("MortarType", _ip), ("MortarInfo", _ip)] + [(k, _dp) for k in ("M_0_1", "M_0_2", "M_1_0", "M_2_0", "FilterMat")]
        + [("IniExactFunc", C.c_int), ("AdvVel", C.c_double * 3), ("Elem_xGP", _dp)]
        + [("doWeakLifting", C.c_int), ("doConservativeLifting", C.c_int)]
        + [("SpongeMat", _dp), ("SpBaseFlow", _dp)]
        + [(k, _dp) for k in ("RKdelta", "RKg1", "RKg2", "RKg3")]
        + [("OverintegrationType", C.c_int), ("NUnder", C.c_int)]
        + [(k, _dp) for k in ("OverintegrationMat", "Vdm_N_NUnder", "Vdm_NUnder_N", "sJNUnder")]
    )


_lib = None


def load_library():
    """Load libdgx.so; raises (no fallback) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DGX_LIB", LIB_PATH)    # A/B runs of tuning builds; the default is the in-tree library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA extension first "
                           f"(python -c 'import __graft_entry__ as g; g.build()' or make -C galaexi_b200/csrc)")
    lib = C.CDLL(path)
    h = C.c_void_p
    lib.dgx_create.argtypes = [C.POINTER(h), C.POINTER(DgxConfig)]
    lib.dgx_destroy.argtypes = [h]
    lib.dgx_destroy.restype = None
    lib.dgx_last_error.argtypes = [h]
    lib.dgx_last_error.restype = C.c_char_p
    for nm in ("dgx_set_state", "dgx_get_state", "dgx_get_ut"):
        getattr(lib, nm).argtypes = [h, _dp]
    lib.dgx_get_gradients.argtypes = [h, _dp, _dp, _dp]
    lib.dgx_get_face_array.argtypes = [h, C.c_int, _dp]
    lib.dgx_set_keep_gradients.argtypes = [h, C.c_int]
    lib.dgx_time_derivative.argtypes = [h, C.c_double]
    lib.dgx_rk_stage.argtypes = [h, C.c_int, C.c_double, C.c_double]
    lib.dgx_rk_step.argtypes = [h, C.c_double, C.c_double]
    lib.dgx_calc_timestep.argtypes = [h, _dp, _ip]
    lib.dgx_analyze_tgv.argtypes = [h, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.dgx_temp_filter_time_deriv.argtypes = [h, C.c_double, C.c_double]
    lib.dgx_get_baseflow.argtypes = [h, _dp]
    lib.dgx_calc_bulk_velocity.argtypes = [h, _dp, C.c_double, _dp]
    lib.dgx_calc_body_forces.argtypes = [h, _dp, C.POINTER(C.c_int), C.c_int, _dp, _dp]
    lib.dgx_calc_wall_velocity.argtypes = [h, _dp, C.POINTER(C.c_int), C.c_int, _dp, _dp, _dp, _dp]
    lib.dgx_set_channel_forcing.argtypes = [h, C.c_int, C.c_double, C.c_double]
    lib.dgx_sync.argtypes = [h]
    lib.dgx_run_steps.argtypes = [h, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    lib.dgx_get_dt_history.argtypes = [h, C.c_int, _dp, _ip]
    lib.dgx_step_graph_active.argtypes = [h]
    lib.dgx_profile_stage.argtypes = [h, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), _ip]
    lib.dgx_nccl_unique_id.argtypes = [C.c_char_p]
    lib.dgx_launch_count.argtypes = [h]
    lib.dgx_launch_count.restype = C.c_longlong
    lib.dgx_sizeof_config.restype = C.c_ulong
    if lib.dgx_sizeof_config() != C.sizeof(DgxConfig):
        raise RuntimeError(f"struct dgx_config mismatch: library {lib.dgx_sizeof_config()} B, binding {C.sizeof(DgxConfig)} B")
    _lib = lib
    return lib


def halo_plan(mesh):
    """The message plan one exchange phase executes (list of (peer, isSend, slaveArray, firstSide0, nSides))."""
    lib = load_library()
    if not mesh.nNbProcs:
        return []
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    tabs = [i32(mesh.NbProc), i32(mesh.nMPISides_MINE_Proc), i32(mesh.nMPISides_YOUR_Proc), i32(mesh.offsetMPISides_MINE),
            i32(mesh.offsetMPISides_YOUR)]
    cap = 4 * mesh.nNbProcs
    out = np.zeros((cap, 5), dtype=np.int32)
    lib.dgx_halo_plan.argtypes = [C.c_int] + [_ip] * 5 + [C.c_int, _ip]
    cnt = lib.dgx_halo_plan(mesh.nNbProcs, *[t.ctypes.data_as(_ip) for t in tabs], cap, out.ctypes.data_as(_ip))
    return [tuple(int(x) for x in r) for r in out[:cnt]]


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if load_library().dgx_nccl_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed (libnccl not loadable?)")
    return buf.raw


class DGError(RuntimeError):
    """Raised where the reference would CALL Abort(__STAMP__, ...)."""


class DGSolver:
    """One rank's DG operator + LSERK time integrator on one B200 (see module docstring)."""

    def __init__(self, case, device: int = 0, nccl_id: bytes | None = None):
        self.lib = load_library()
        self.case = case
        m, b, g = case.mesh, case.basis, case.geo
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        td = case.timedisc
        k = dict(
            D_T=f64(b.D_T.T), D_Hat_T=f64(b.D_Hat_T.T), DVolSurf=f64(b.DVolSurf.T), L_Minus=f64(b.L_Minus),
            L_Plus=f64(b.L_Plus), L_HatMinus=f64(b.L_HatMinus), L_HatPlus=f64(b.L_HatPlus),
            ElemToSide=i32(m.ElemToSide), S2V2=i32(case.maps["S2V2"]), S2V2_inv=i32(case.maps["S2V2_inv"]),
            BCSides=i32(case.BCSides if case.BCSides.size else np.zeros((1, 2))),
            Metrics_fTilde=f64(g["Metrics_fTilde"]), Metrics_gTilde=f64(g["Metrics_gTilde"]),
            Metrics_hTilde=f64(g["Metrics_hTilde"]), sJ=f64(g["sJ"]), NormVec=f64(g["NormVec"]),
            TangVec1=f64(g["TangVec1"]), TangVec2=f64(g["TangVec2"]), SurfElem=f64(g["SurfElem"]),
            RefStatePrim=f64(case.RefStatePrim), RKA=f64(td.RKA), RKb=f64(td.RKb), RKc=f64(td.RKc),
            NbProc=i32(m.NbProc if m.nNbProcs else np.zeros(1)),
            nMPISides_MINE_Proc=i32(m.nMPISides_MINE_Proc if m.nNbProcs else np.zeros(1)),
            nMPISides_YOUR_Proc=i32(m.nMPISides_YOUR_Proc if m.nNbProcs else np.zeros(1)),
            offsetMPISides_MINE=i32(m.offsetMPISides_MINE if m.nNbProcs else np.zeros(2)),
            offsetMPISides_YOUR=i32(m.offsetMPISides_YOUR if m.nNbProcs else np.zeros(2)),
            MortarType=i32(m.MortarType), MortarInfo=i32(m.MortarInfo),
            # mortar operators: Fortran M(l,p) at [l + n*p] == C array M.T
            M_0_1=f64(case.mortar["M_0_1"].T), M_0_2=f64(case.mortar["M_0_2"].T),
            M_1_0=f64(case.mortar["M_1_0"].T), M_2_0=f64(case.mortar["M_2_0"].T),
        )
        self._keep = k
        c = DgxConfig()
        c.N = case.N
        c.nodeType = getattr(case, "op_node_type", 2 if case.node_type == "GAUSS-LOBATTO" else 1)
        c.splitDG, c.riemann, c.parabolic, c.viscLaw = case.split, case.riemann, int(case.parabolic), case.eos.visc_law
        c.nElems, c.nSides, c.nBCSides = m.nElems, m.nSides, m.nBCSides
        c.firstInnerSide, c.lastInnerSide = m.firstInnerSide, m.lastInnerSide
        c.firstMPISide_MINE, c.lastMPISide_MINE = m.firstMPISide_MINE, m.lastMPISide_MINE
        c.firstMPISide_YOUR, c.lastMPISide_YOUR = m.firstMPISide_YOUR, m.lastMPISide_YOUR
        for i, v in enumerate(case.eos.eos_vars()):
            c.EOS_Vars[i] = v
        c.nRefState = case.RefStatePrim.shape[0]
        for nm in ("RefStatePrim", "D_T", "D_Hat_T", "DVolSurf", "L_Minus", "L_Plus", "L_HatMinus", "L_HatPlus",
                   "Metrics_fTilde", "Metrics_gTilde", "Metrics_hTilde", "sJ", "NormVec", "TangVec1", "TangVec2",
                   "SurfElem", "RKA", "RKb", "RKc"):
            setattr(c, nm, k[nm].ctypes.data_as(_dp))
        for nm in ("BCSides", "ElemToSide", "S2V2", "S2V2_inv", "NbProc", "nMPISides_MINE_Proc",
                   "nMPISides_YOUR_Proc", "offsetMPISides_MINE", "offsetMPISides_YOUR", "MortarType", "MortarInfo"):
            setattr(c, nm, k[nm].ctypes.data_as(_ip))
        for nm in ("M_0_1", "M_0_2", "M_1_0", "M_2_0"):
            setattr(c, nm, k[nm].ctypes.data_as(_dp))
        c.doWeakLifting, c.doConservativeLifting = int(case.doWeakLifting), int(case.doConservativeLifting)
        if getattr(td, "RKg1", None) is not None:   # three-register scheme (TimeStepByLSERKK3)
            for nm in ("RKdelta", "RKg1", "RKg2", "RKg3"):
                if getattr(td, nm) is not None:
                    k[nm] = f64(getattr(td, nm))
                    setattr(c, nm, k[nm].ctypes.data_as(_dp))
        if case.SpongeMat is not None:
            k["SpongeMat"], k["SpBaseFlow"] = f64(case.SpongeMat), f64(case.SpBaseFlow)
            c.SpongeMat, c.SpBaseFlow = k["SpongeMat"].ctypes.data_as(_dp), k["SpBaseFlow"].ctypes.data_as(_dp)
        if case.IniExactFunc:
            k["Elem_xGP"] = f64(g["Elem_xGP"])
            c.Elem_xGP = k["Elem_xGP"].ctypes.data_as(_dp)
            c.IniExactFunc = int(case.IniExactFunc)
            for i_, v_ in enumerate(case.AdvVel):
                c.AdvVel[i_] = v_
        if getattr(case, "OverintegrationType", 0):   # Fortran M(a,b) at [a + na*b] == C array M.T
            c.OverintegrationType, c.NUnder = int(case.OverintegrationType), int(case.NUnder)
            if case.OverintegrationType == 1:
                k["OverintegrationMat"] = f64(np.asarray(case.OverintegrationMat).T)
                c.OverintegrationMat = k["OverintegrationMat"].ctypes.data_as(_dp)
            else:
                k["Vdm_N_NUnder"], k["Vdm_NUnder_N"] = f64(np.asarray(case.Vdm_N_NUnder).T), f64(np.asarray(case.Vdm_NUnder_N).T)
                k["sJNUnder"] = f64(case.sJNUnder)
                for nm in ("Vdm_N_NUnder", "Vdm_NUnder_N", "sJNUnder"):
                    setattr(c, nm, k[nm].ctypes.data_as(_dp))
        if case.FilterMat is not None:
            k["FilterMat"] = f64(np.asarray(case.FilterMat).T)      # Fortran FilterMat(i,l) at [i + n*l]
            c.FilterMat = k["FilterMat"].ctypes.data_as(_dp)
        c.lifting, c.etaBR2, c.etaBR2_wall = case.lifting, case.etaBR2, case.etaBR2_wall
        c.nMortarSides = m.nMortarSides
        c.firstMortarInnerSide, c.lastMortarInnerSide = m.firstMortarInnerSide, m.lastMortarInnerSide
        c.firstMortarMPISide, c.lastMortarMPISide = m.firstMortarMPISide, m.lastMortarMPISide
        c.nRKStages, c.CFLScale, c.DFLScale = td.nRKStages, td.CFLScale, td.DFLScale
        c.myRank, c.nRanks, c.nNbProcs = m.myRank, m.nProcs, m.nNbProcs
        self._id = nccl_id
        c.ncclUniqueId = nccl_id if (m.nProcs > 1 and nccl_id is not None) else None
        c.device = device
        self._cfg = c
        self.h = C.c_void_p()
        rc = self.lib.dgx_create(C.byref(self.h), C.byref(c))
        if rc:
            msg = self.lib.dgx_last_error(self.h).decode() if self.h else "dgx_create failed"
            if self.h:
                self.lib.dgx_destroy(self.h)
                self.h = None
            raise DGError(msg)
        n = case.N + 1
        self.shape_U = (m.nElems, n, n, n, 5)
        self.shape_grad = (m.nElems, n, n, n, 4)

    def _need_global(self, who: str, what: str, value):
        """The device sums are reduced over all ranks: the normalising volume / surface has to be the global one as well."""
        if value is None and self.case.mesh.nProcs > 1:
            raise DGError(f"{who}: with {self.case.mesh.nProcs} ranks the global {what} has to be given ({what}=...); "
                          f"this rank's local {what} would scale the rank-reduced sums wrongly")

    # ---- error mapping: non-zero return -> Abort ---------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise DGError(self.lib.dgx_last_error(self.h).decode())

    # ---- state transfer (reference layout U(nVar,i,j,k,iElem) == numpy [e,k,j,i,v]) -----------------------
    def set_state(self, U: np.ndarray):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.shape == self.shape_U, (U.shape, self.shape_U)
        self._ck(self.lib.dgx_set_state(self.h, U.ctypes.data_as(_dp)))

    def get_state(self, out: np.ndarray | None = None) -> np.ndarray:
        U = out if out is not None else np.empty(self.shape_U)
        self._ck(self.lib.dgx_get_state(self.h, U.ctypes.data_as(_dp)))
        return U

    def get_ut(self) -> np.ndarray:
        Ut = np.empty(self.shape_U)
        self._ck(self.lib.dgx_get_ut(self.h, Ut.ctypes.data_as(_dp)))
        return Ut

    def get_gradients(self):
        g = [np.empty(self.shape_grad) for _ in range(3)]
        self._ck(self.lib.dgx_get_gradients(self.h, *[x.ctypes.data_as(_dp) for x in g]))
        return g

    FACE_ARRAYS = ("U_master", "U_slave", "Flux_master", "gradUx_master", "gradUy_master", "gradUz_master", "gradUx_slave",
                   "gradUy_slave", "gradUz_slave")

    def get_face_array(self, name: str) -> np.ndarray:
        """A face array of the last RHS evaluation in the reference layout: numpy [side, q, p, var] == Fortran (var,p,q,side)."""
        which = self.FACE_ARRAYS.index(name)
        n = self.case.N + 1
        out = np.empty((self.case.mesh.nSides, n, n, 5 if which < 3 else 4))
        self._ck(self.lib.dgx_get_face_array(self.h, which, out.ctypes.data_as(_dp)))
        return out

    def set_keep_gradients(self, on: bool):
        """on (default): the last stage of every RK step stores the volume gradients, as the reference's d_gradUx/y/z hold them
        at analyze steps (testcase.f90:361-364); off: RK stages never write them (steps no analysis follows)."""
        self._ck(self.lib.dgx_set_keep_gradients(self.h, int(bool(on))))

    # ---- the reference's procedures ------------------------------------------------------------------------
    def DGTimeDerivative_weakForm(self, t: float = 0.0):
        self._ck(self.lib.dgx_time_derivative(self.h, float(t)))

    def TimeStepByLSERKW2(self, t: float, dt: float):
        self._ck(self.lib.dgx_rk_step(self.h, float(t), float(dt)))

    # the reference binds one of the two through the procedure pointer TimeStep (timedisc_func.f90:143-151); the library picks the
    # update from the tables it was created with, so both names run dgx_rk_step
    TimeStepByLSERKK3 = TimeStepByLSERKW2
    TimeStep = TimeStepByLSERKW2

    def rk_stage(self, iStage: int, t: float, dt: float):
        self._ck(self.lib.dgx_rk_stage(self.h, int(iStage), float(t), float(dt)))

    def CalcTimeStep(self):
        dt, err = C.c_double(), C.c_int()
        self._ck(self.lib.dgx_calc_timestep(self.h, C.byref(dt), C.byref(err)))
        return dt.value, err.value

    def AnalyzeTestcase(self, NAnalyze: int | None = None, Vol: float | None = None, rho0: float = 1.0) -> np.ndarray:
        """The 15 TGVAnalysis columns (testcase/taylorgreenvortex/testcase.f90:283-515), integrated on the device.
        Vol: global volume (all ranks); defaults to this rank's volume (single-rank runs)."""
        from .host_standin import analyze as an
        self._need_global("AnalyzeTestcase", "Vol", Vol)
        key = (NAnalyze,)
        if getattr(self, "_an_key", None) != key:
            NA, V, wA = an.init_analyze_basis(self.case.N, self.case.node_type, NAnalyze)
            # Fortran Vdm(0:NA,0:N) at [I + (NA+1) i] == C array V.T
            self._an = (NA, np.ascontiguousarray(V.T, dtype=np.float64), np.ascontiguousarray(wA, dtype=np.float64))
            self._an_vol = an.volume(self.case)
            self._an_key = key
        NA, Vt, wA = self._an
        out = np.zeros(15)
        self._ck(self.lib.dgx_analyze_tgv(self.h, NA, Vt.ctypes.data_as(_dp), wA.ctypes.data_as(_dp),
                                          float(self._an_vol if Vol is None else Vol), float(rho0), out.ctypes.data_as(_dp)))
        return out

    def TempFilterTimeDeriv(self, dt: float, tempFilterWidth: float):
        """Pruett temporal filter of the sponge base flow (sponge/pruettdamping.f90:69-92), once per time step."""
        self._ck(self.lib.dgx_temp_filter_time_deriv(self.h, float(dt), float(tempFilterWidth)))

    def get_baseflow(self) -> np.ndarray:
        B = np.empty(self.shape_U)
        self._ck(self.lib.dgx_get_baseflow(self.h, B.ctypes.data_as(_dp)))
        return B

    def CalcForcing(self, Vol: float | None = None) -> float:
        """CalcForcing of the channel testcase (testcase/channel/testcase.f90:241-271): the bulk velocity."""
        from .host_standin import analyze as an
        self._need_global("CalcForcing", "Vol", Vol)
        w = np.ascontiguousarray(self.case.basis.wGP, dtype=np.float64)
        b = C.c_double()
        self._ck(self.lib.dgx_calc_bulk_velocity(self.h, w.ctypes.data_as(_dp), float(an.volume(self.case) if Vol is None else Vol),
                                                 C.byref(b)))
        return b.value

    def CalcErrorNorms(self, Time: float, exact, NAnalyze: int | None = None, Vol: float | None = None, reduce=None):
        """CalcErrorNorms(Time,L_2_Error,L_Inf_Error) (analyze.f90:383-470) on the downloaded state; ``exact(x, t)`` plays
        ExactFunc(AnalyzeExactFunc, ...). Host arithmetic like the reference (it analyses the host copy U = d_U)."""
        from .host_standin import analyze as an
        return an.calc_error_norms(self.case, self.get_state(), Time, exact, NAnalyze, Vol, reduce)

    def CalcBodyForces(self):
        """CalcBodyForces(BodyForce,Fp,Fv) (equations/navierstokes/calcbodyforces.f90:41-110) from the face data of the last
        DGTimeDerivative_weakForm: arrays (nBCs,3) == Fortran (3,nBCs), zero for boundary conditions that are not walls."""
        m = self.case.mesh
        nBCs = int(m.BoundaryType.shape[0])
        w = np.ascontiguousarray(self.case.basis.wGP, dtype=np.float64)
        bc = np.ascontiguousarray(m.BC[:m.nBCSides] if m.nBCSides else np.zeros(1), dtype=np.int32)
        Fp, Fv = np.zeros((nBCs, 3)), np.zeros((nBCs, 3))
        self._ck(self.lib.dgx_calc_body_forces(self.h, w.ctypes.data_as(_dp), bc.ctypes.data_as(C.POINTER(C.c_int)), nBCs,
                                               Fp.ctypes.data_as(_dp), Fv.ctypes.data_as(_dp)))
        return Fp + Fv, Fp, Fv

    def CalcWallVelocity(self, Surf: np.ndarray | None = None):
        """CalcWallVelocity(maxV,minV,meanV) (analyze_equation.f90:435-499); Surf(nBCs): global surface per boundary condition
        (default: integrated from this rank's SurfElem, i.e. single-rank)."""
        from .host_standin import analyze as an
        self._need_global("CalcWallVelocity", "Surf", Surf)
        m = self.case.mesh
        nBCs = int(m.BoundaryType.shape[0])
        w = np.ascontiguousarray(self.case.basis.wGP, dtype=np.float64)
        bc = np.ascontiguousarray(m.BC[:m.nBCSides] if m.nBCSides else np.zeros(1), dtype=np.int32)
        S = np.ascontiguousarray(an.bc_surfaces(self.case) if Surf is None else Surf, dtype=np.float64)
        out = [np.zeros(nBCs) for _ in range(3)]
        self._ck(self.lib.dgx_calc_wall_velocity(self.h, w.ctypes.data_as(_dp), bc.ctypes.data_as(C.POINTER(C.c_int)), nBCs,
                                                 S.ctypes.data_as(_dp), *[o.ctypes.data_as(_dp) for o in out]))
        return tuple(out)

    def set_channel_forcing(self, dpdx: float, BulkVel: float, on: bool = True):
        """Parameters of TestcaseSource (testcase/channel/testcase.f90:277-296)."""
        self._ck(self.lib.dgx_set_channel_forcing(self.h, int(on), float(dpdx), float(BulkVel)))

    def WriteState(self, MeshFileName: str, OutputTime: float, FutureTime: float, ProjectName: str, out_dir: str = ".",
                   NOut: int | None = None, dt: float | None = None, ini_text: str = "", isErrorFile: bool = False, barrier=None) -> str:
        """WriteState (io_hdf5/hdf5_output.f90:84-211): D2H of U through dgx_get_state and a state file in the reference
        layout (readable by posti and by the reference's Restart). All ranks call it; each writes its own element range."""
        from .host_standin import state_io
        m = self.case.mesh
        ed = {"myRank": float(m.myRank)}
        if dt is not None:
            ed["dt"] = float(dt)
        return state_io.write_state(self.get_state(), self.case.N, self.case.node_type, ProjectName, MeshFileName, OutputTime,
                                    FutureTime, out_dir=out_dir, sJ=self.case.geo["sJ"], NOut=NOut, elem_data=ed, ini_text=ini_text,
                                    is_error_file=isErrorFile, offsetElem=m.offsetElem, nGlobalElems=m.nGlobalElems,
                                    rank=m.myRank, barrier=barrier)

    def WriteBaseFlow(self, MeshFileName: str, OutputTime: float, FutureTime: float, ProjectName: str, out_dir: str = ".", barrier=None) -> str:
        """WriteBaseflow (io_hdf5/hdf5_output.f90:527-603) of the sponge base flow held on the device (dgx_get_baseflow)."""
        from .host_standin import state_io
        m = self.case.mesh
        return state_io.write_baseflow(self.get_baseflow(), self.case.N, self.case.node_type, ProjectName, MeshFileName, OutputTime,
                                       FutureTime, out_dir=out_dir, offsetElem=m.offsetElem, nGlobalElems=m.nGlobalElems, rank=m.myRank,
                                       barrier=barrier)

    def Restart(self, RestartFile: str, ResetTime: bool = False) -> float:
        """InitRestart + Restart (restart/restart.f90:60-135, 304-560): this rank's element range of DG_Solution, interpolated
        when the file's degree / node type differ, H2D through dgx_set_state; returns RestartTime."""
        from .host_standin import metrics, state_io
        m = self.case.mesh
        info = state_io.read_state_attrs(RestartFile)
        dj = metrics.det_jac_ref(m.NodeCoords, m.NGeo, self.case.node_type) if info["N"] > self.case.N else None
        U, t = state_io.restart(RestartFile, self.case.N, self.case.node_type, sJ=self.case.geo["sJ"], detJac_Ref=dj, NGeo=m.NGeo,
                                offsetElem=m.offsetElem, nElems=m.nElems, nGlobalElems=m.nGlobalElems, ResetTime=ResetTime)
        self.set_state(U)
        return t

    def FinalizeDG(self):
        if getattr(self, "h", None):
            self.lib.dgx_destroy(self.h)
            self.h = None

    # ---- adapters used by host.timeloop.advance ---------------------------------------------------------------
    def calc_timestep(self):
        dt, err = self.CalcTimeStep()
        if err:
            raise DGError("Error: (1) density, (2) convective / (3) viscous timestep is NaN.")
        return dt, None, None

    def rk_step(self, t, dt):
        self.TimeStepByLSERKW2(t, dt)

    # ---- measurement ---------------------------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.dgx_sync(self.h))

    def run_steps(self, nSteps: int, t: float, dt: float, adaptive: bool = False, forcing: bool = False, device_paced: bool = False,
                  graph: bool = False):
        """The time loop between two analyze points on the device (dgx_run_steps): returns (device ms, kernels launched).
        device_paced: dt never leaves the device (no host synchronisation inside the call); graph: replay step pairs as a
        CUDA graph. The step sizes of a device-paced call are read with dt_history()."""
        ms, ln = C.c_float(), C.c_longlong()
        flags = int(bool(adaptive)) | 2 * int(bool(forcing)) | 4 * int(bool(device_paced)) | 8 * int(bool(graph))
        self._ck(self.lib.dgx_run_steps(self.h, nSteps, float(t), float(dt), flags, C.byref(ms), C.byref(ln)))
        return ms.value, ln.value

    def dt_history(self) -> np.ndarray:
        cnt = C.c_int()
        self._ck(self.lib.dgx_get_dt_history(self.h, 0, None, C.byref(cnt)))
        out = np.zeros(max(cnt.value, 1))
        self._ck(self.lib.dgx_get_dt_history(self.h, cnt.value, out.ctypes.data_as(_dp), C.byref(cnt)))
        return out[:cnt.value]

    def step_graph_active(self) -> bool:
        return bool(self.lib.dgx_step_graph_active(self.h))

    def profile_stage(self, t: float, dt: float):
        names = (C.c_char_p * 8)()
        ms = (C.c_float * 8)()
        cnt = C.c_int()
        self._ck(self.lib.dgx_profile_stage(self.h, float(t), float(dt), 8, names, ms, C.byref(cnt)))
        return {names[i].decode(): ms[i] for i in range(cnt.value)}

    def launch_count(self) -> int:
        return int(self.lib.dgx_launch_count(self.h))

    def __del__(self):
        try:
            self.FinalizeDG()
        except Exception:
            pass
