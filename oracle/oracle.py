"""ctypes loader for the CPU parity oracle (oracle/dg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package (galaexi_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ip = C.POINTER(C.c_int)


class _Prec:
    """One build of dg_oracle.c: FP64 ("double") or 80-bit extended ("extended", -DDGO_EXTENDED)."""

    def __init__(self, name):
        self.name = name
        ext = name == "extended"
        self.real = C.c_longdouble if ext else C.c_double
        self.np = np.longdouble if ext else np.float64
        self.rp = C.POINTER(self.real)
        self.libname = "libdgoracle_ld.so" if ext else "libdgoracle.so"
        rp = self.rp

        class Config(C.Structure):
            _fields_ = [(k, C.c_int) for k in ("N", "nElems", "nSides", "nBCSides", "firstInnerSide", "lastInnerSide",
                                               "firstMPISide_MINE", "lastMPISide_MINE", "firstMPISide_YOUR",
                                               "lastMPISide_YOUR", "nodeType", "splitDG", "riemann", "parabolic",
                                               "viscLaw", "nRefState")] + \
                       [("EOS", self.real * 8)] + \
                       [(k, rp) for k in ("D_T", "D_Hat_T", "DVolSurf", "L_Minus", "L_Plus", "L_HatMinus", "L_HatPlus")] + \
                       [(k, _ip) for k in ("ElemToSide", "S2V2", "S2V2_inv", "BCSides")] + \
                       [(k, rp) for k in ("Metrics_fTilde", "Metrics_gTilde", "Metrics_hTilde", "sJ", "NormVec",
                                          "TangVec1", "TangVec2", "SurfElem", "RefStatePrim")] + \
                       [("lifting", C.c_int), ("etaBR2", self.real), ("etaBR2_wall", self.real)] + \
                       [(k, C.c_int) for k in ("firstMortarInnerSide", "lastMortarInnerSide", "firstMortarMPISide",
                                               "lastMortarMPISide")] + \
                       [(k, _ip) for k in ("MortarType", "MortarInfo", "FS2M", "SideToElem")] + \
                       [(k, rp) for k in ("M_0_1", "M_0_2", "M_1_0", "M_2_0", "FilterMat")] + \
                       [("iniExactFunc", C.c_int), ("AdvVel", self.real * 3), ("Elem_xGP", rp)] + \
                       [("tcSource", C.c_int), ("dpdx", self.real), ("BulkVel", self.real)] + \
                       [("doWeakLifting", C.c_int), ("doConservativeLifting", C.c_int)] + \
                       [("SpongeMat", rp)] + \
                       [("OverintegrationType", C.c_int), ("NUnder", C.c_int)] + \
                       [(k, rp) for k in ("OverintegrationMat", "Vdm_N_NUnder", "Vdm_NUnder_N", "sJNUnder")]
        self.Config = Config
        self._lib = None

    def lib(self):
        if self._lib is None:
            path = os.path.join(_HERE, self.libname)
            src = os.path.join(_HERE, "dg_oracle.c")
            if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
                subprocess.check_call(["make", "-s", "-C", _HERE, self.libname])
            L = C.CDLL(path)
            r, rp = self.real, self.rp
            L.dgo_create.restype = C.c_void_p
            L.dgo_create.argtypes = [C.POINTER(self.Config)]
            L.dgo_destroy.argtypes = [C.c_void_p]
            L.dgo_array.restype = rp
            L.dgo_array.argtypes = [C.c_void_p, C.c_char_p]
            L.dgo_time_derivative.argtypes = [C.c_void_p, r]
            L.dgo_rk_stage.argtypes = [C.c_void_p, r, r, r]
            L.dgo_rk_step.argtypes = [C.c_void_p, r, r, C.c_int, rp, rp, rp]
            L.dgo_calc_timestep.restype = r
            L.dgo_calc_timestep.argtypes = [C.c_void_p, r, r, rp, rp]
            L.dgo_prolong_to_face.argtypes = [C.c_void_p, C.c_int, rp, rp, rp]
            L.dgo_surf_int.argtypes = [C.c_void_p, C.c_int, rp, rp, rp]
            L.dgo_lifting.argtypes = [C.c_void_p]
            L.dgo_filter.argtypes = [C.c_void_p]
            L.dgo_temp_filter_time_deriv.argtypes = [C.c_void_p, r, r]
            L.dgo_set_forcing.argtypes = [C.c_void_p, C.c_int, r, r]
            L.dgo_bulk_velocity.restype = r
            L.dgo_bulk_velocity.argtypes = [C.c_void_p, rp, r]
            L.dgo_rhs_phase.argtypes = [C.c_void_p, C.c_int]
            L.dgo_rk_update.argtypes = [C.c_void_p, r, r]
            L.dgo_sizeof_config.restype = C.c_size_t
            assert L.dgo_sizeof_config() == C.sizeof(self.Config)
            self._lib = L
        return self._lib

    def d(self, a):
        return a.ctypes.data_as(self.rp)


_PREC = {"double": _Prec("double"), "extended": _Prec("extended")}


def build(force: bool = False) -> str:
    if force:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE])
    _PREC["double"].lib()
    return os.path.join(_HERE, "libdgoracle.so")


def lib():
    return _PREC["double"].lib()


def _d(a):
    return _PREC["double"].d(a)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """One single-rank DG operator instance built from a galaexi_b200.host_standin.case.Case."""

    def __init__(self, case, precision: str = "double"):
        self.prec = _PREC[precision]
        L = self.prec.lib()
        _d = self.prec.d
        m, b, g = case.mesh, case.basis, case.geo
        self.case = case
        self.n = case.N + 1
        # optional pieces (hand-built single-element cases of the unit tests do not carry them)
        nS_ = m.nSides
        MortarType = getattr(m, "MortarType", None)
        MortarType = np.zeros((nS_, 2)) if MortarType is None else MortarType
        MortarInfo = getattr(m, "MortarInfo", None)
        MortarInfo = -np.ones((1, 4, 2)) if MortarInfo is None else MortarInfo
        SideToElem = getattr(m, "SideToElem", None)
        SideToElem = -np.ones((nS_, 5)) if SideToElem is None else SideToElem
        mortar = getattr(case, "mortar", None)
        if mortar is None:
            from galaexi_b200.host_standin import mortar as _mo
            mortar = _mo.init_mortar(case.N, case.node_type)
        # keep references: the C side stores raw pointers
        f64 = lambda a: np.ascontiguousarray(a, dtype=self.prec.np)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        # operator matrices: Fortran M(a,b) at [a + n*b] == C array M.T
        self._keep = dict(
            D_T=f64(b.D_T.T), D_Hat_T=f64(b.D_Hat_T.T), DVolSurf=f64(b.DVolSurf.T), L_Minus=f64(b.L_Minus), L_Plus=f64(b.L_Plus),
            L_HatMinus=f64(b.L_HatMinus), L_HatPlus=f64(b.L_HatPlus),
            ElemToSide=i32(m.ElemToSide), S2V2=i32(case.maps["S2V2"]), S2V2_inv=i32(case.maps["S2V2_inv"]),
            BCSides=i32(case.BCSides if case.BCSides.size else np.zeros((1, 2))),
            Metrics_fTilde=f64(g["Metrics_fTilde"]), Metrics_gTilde=f64(g["Metrics_gTilde"]), Metrics_hTilde=f64(g["Metrics_hTilde"]),
            sJ=f64(g["sJ"]), NormVec=f64(g["NormVec"]), TangVec1=f64(g["TangVec1"]), TangVec2=f64(g["TangVec2"]),
            SurfElem=f64(g["SurfElem"]), RefStatePrim=f64(case.RefStatePrim),
            MortarType=i32(MortarType), MortarInfo=i32(MortarInfo), FS2M=i32(case.maps["FS2M"]), SideToElem=i32(SideToElem),
            # mortar operators: Fortran M(l,p) at [l + n*p] == C array M.T
            M_0_1=f64(mortar["M_0_1"].T), M_0_2=f64(mortar["M_0_2"].T), M_1_0=f64(mortar["M_1_0"].T),
            M_2_0=f64(mortar["M_2_0"].T))
        c = self.prec.Config()
        c.N, c.nElems, c.nSides = case.N, m.nElems, m.nSides
        c.nBCSides, c.firstInnerSide, c.lastInnerSide = m.nBCSides, m.firstInnerSide, m.lastInnerSide
        c.firstMPISide_MINE, c.lastMPISide_MINE = m.firstMPISide_MINE, m.lastMPISide_MINE
        c.firstMPISide_YOUR, c.lastMPISide_YOUR = m.firstMPISide_YOUR, m.lastMPISide_YOUR
        c.nodeType = getattr(case, "op_node_type", 2 if case.node_type == "GAUSS-LOBATTO" else 1)
        c.splitDG, c.riemann, c.parabolic = case.split, case.riemann, int(case.parabolic)
        c.viscLaw, c.nRefState = case.eos.visc_law, case.RefStatePrim.shape[0]
        for k, v in enumerate(case.eos.eos_vars()):
            c.EOS[k] = v
        for k in ("D_T", "D_Hat_T", "DVolSurf", "L_Minus", "L_Plus", "L_HatMinus", "L_HatPlus", "Metrics_fTilde",
                  "Metrics_gTilde", "Metrics_hTilde", "sJ", "NormVec", "TangVec1", "TangVec2", "SurfElem", "RefStatePrim",
                  "M_0_1", "M_0_2", "M_1_0", "M_2_0"):
            setattr(c, k, _d(self._keep[k]))
        for k in ("ElemToSide", "S2V2", "S2V2_inv", "BCSides", "MortarType", "MortarInfo", "FS2M", "SideToElem"):
            setattr(c, k, _i(self._keep[k]))
        if getattr(case, "IniExactFunc", 0):
            self._keep["Elem_xGP"] = f64(g["Elem_xGP"])
            c.Elem_xGP = _d(self._keep["Elem_xGP"])
            c.iniExactFunc = int(case.IniExactFunc)
            for k_, v_ in enumerate(case.AdvVel):
                c.AdvVel[k_] = v_
        if getattr(case, "SpongeMat", None) is not None:
            self._keep["SpongeMat"] = f64(case.SpongeMat)
            c.SpongeMat = _d(self._keep["SpongeMat"])
        ot = int(getattr(case, "OverintegrationType", 0))
        if ot:   # dg/overintegration.f90: Fortran M(a,b) at [a + na*b] == C array M.T
            c.OverintegrationType, c.NUnder = ot, int(case.NUnder)
            if ot == 1:
                self._keep["OverintegrationMat"] = f64(np.asarray(case.OverintegrationMat).T)
                c.OverintegrationMat = _d(self._keep["OverintegrationMat"])
            else:
                self._keep["Vdm_N_NUnder"] = f64(np.asarray(case.Vdm_N_NUnder).T)
                self._keep["Vdm_NUnder_N"] = f64(np.asarray(case.Vdm_NUnder_N).T)
                self._keep["sJNUnder"] = f64(case.sJNUnder)
                for k in ("Vdm_N_NUnder", "Vdm_NUnder_N", "sJNUnder"):
                    setattr(c, k, _d(self._keep[k]))
        c.doWeakLifting = int(getattr(case, "doWeakLifting", False))
        c.doConservativeLifting = int(getattr(case, "doConservativeLifting", False))
        FilterMat = getattr(case, "FilterMat", None)
        if FilterMat is not None:
            self._keep["FilterMat"] = f64(np.asarray(FilterMat).T)   # Fortran FilterMat(i,l) at [i + n*l]
            c.FilterMat = _d(self._keep["FilterMat"])
        c.lifting, c.etaBR2 = getattr(case, "lifting", 1), getattr(case, "etaBR2", 2.0)
        c.etaBR2_wall = getattr(case, "etaBR2_wall", c.etaBR2)
        c.firstMortarInnerSide = getattr(m, "firstMortarInnerSide", m.nBCSides + 1)
        c.lastMortarInnerSide = getattr(m, "lastMortarInnerSide", m.nBCSides)
        c.firstMortarMPISide = getattr(m, "firstMortarMPISide", m.nSides + 1)
        c.lastMortarMPISide = getattr(m, "lastMortarMPISide", m.nSides)
        self._cfg = c
        self.h = L.dgo_create(C.byref(c))
        n, nE, nS = self.n, m.nElems, m.nSides
        self._shapes = {}
        for nm in ("U", "Ut", "Ut_tmp", "SpBaseFlow"):
            self._shapes[nm] = (nE, n, n, n, 5)
        if getattr(case, "SpBaseFlow", None) is not None:
            self.array("SpBaseFlow")[...] = case.SpBaseFlow
        self._shapes["UPrim"] = (nE, n, n, n, 6)
        for nm in ("gradUx", "gradUy", "gradUz"):
            self._shapes[nm] = (nE, n, n, n, 5)
        for nm in ("U_master", "U_slave", "Flux_master", "Flux_slave", "gradUx_master", "gradUy_master", "gradUz_master",
                   "gradUx_slave", "gradUy_slave", "gradUz_slave"):
            self._shapes[nm] = (nS, n, n, 5)
        for nm in ("UPrim_master", "UPrim_slave"):
            self._shapes[nm] = (nS, n, n, 6)

    def array(self, name: str) -> np.ndarray:
        """Writable numpy view (reference memory layout, C-order reversed index list) of an oracle array."""
        p = self.prec.lib().dgo_array(self.h, name.encode())
        shp = self._shapes[name]
        cnt = int(np.prod(shp))
        buf = (self.prec.real * cnt).from_address(C.addressof(p.contents))
        return np.frombuffer(buf, dtype=self.prec.np, count=cnt).reshape(shp)

    def set_state(self, U: np.ndarray):
        self.array("U")[...] = U
        self.array("Ut_tmp")[...] = 0.0

    def time_derivative(self, t: float = 0.0) -> np.ndarray:
        err = self.prec.lib().dgo_time_derivative(self.h, float(t))
        if err:
            raise RuntimeError("oracle: unsupported boundary condition type")
        return self.array("Ut")

    def rhs_phase(self, phase: int):
        """One of the five pieces of DGTimeDerivative_weakForm between the reference's halo exchanges."""
        if self.prec.lib().dgo_rhs_phase(self.h, int(phase)):
            raise RuntimeError("oracle: unsupported boundary condition type")

    def rk_update(self, mRKA: float, b_dt: float):
        self.prec.lib().dgo_rk_update(self.h, float(mRKA), float(b_dt))

    def rk_step(self, t: float, dt: float):
        td = self.case.timedisc
        if getattr(td, "kind", "LSERKW2") == "LSERKK3":
            return self._rk_step_k3(t, dt)
        _d = self.prec.d
        A, b, c = (np.ascontiguousarray(x, dtype=self.prec.np) for x in (td.RKA, td.RKb, td.RKc))
        err = self.prec.lib().dgo_rk_step(self.h, float(t), float(dt), td.nRKStages, _d(A), _d(b), _d(c))
        if err:
            raise RuntimeError("oracle: unsupported boundary condition type")

    def _rk_step_k3(self, t: float, dt: float):
        """TimeStepByLSERKK3 (timedisc/timestep.f90:129-200): S1 == U, S2, S3 == UPrev; one operation per VAXPBY call."""
        td = self.case.timedisc
        U = self.array("U")
        b_dt = td.RKb * dt
        S2 = UPrev = None
        for i in range(td.nRKStages):
            tStage = t if i == 0 else t + td.RKc[i] * dt
            Ut = self.time_derivative(tStage)
            if i == 0:
                UPrev = U.copy()
                S2 = U.copy()
            else:
                S2 = S2 + U * td.RKdelta[i]
                U[...] = U * td.RKg1[i] + S2 * td.RKg2[i]
                U[...] = U + UPrev * td.RKg3[i]
            U[...] = U + Ut * b_dt[i]

    def temp_filter_time_deriv(self, dt: float, tempFilterWidth: float):
        self.prec.lib().dgo_temp_filter_time_deriv(self.h, float(dt), float(tempFilterWidth))

    def set_forcing(self, dpdx: float, BulkVel: float, on: bool = True):
        self.prec.lib().dgo_set_forcing(self.h, int(on), float(dpdx), float(BulkVel))

    def bulk_velocity(self, Vol: float) -> float:
        w = np.ascontiguousarray(self.case.basis.wGP, dtype=self.prec.np)
        return float(self.prec.lib().dgo_bulk_velocity(self.h, self.prec.d(w), float(Vol)))

    def calc_timestep(self):
        tc, tv = self.prec.real(), self.prec.real()
        td = self.case.timedisc
        dt = self.prec.lib().dgo_calc_timestep(self.h, td.CFLScale, td.DFLScale, C.byref(tc), C.byref(tv))
        return float(dt), float(tc.value), float(tv.value)

    def close(self):
        if self.h:
            self.prec.lib().dgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
