"""Equation-system setup on the host: EOS constants, reference states, BC table, initial conditions.

Mirrors (relative to /root/reference/src):
  * equations/navierstokes/idealgas/eos.f90:99-207   InitEos  (EOS_Vars(1:8), eos.h:44-62)
  * equations/navierstokes/equation.f90:82-180        InitEquation (RefState -> RefStatePrim with T)
  * equations/navierstokes/idealgas/getboundaryflux.f90:244-252  BCSides(2,nBCSides) = [type, state]
  * equations/navierstokes/idealgas/exactfunc.f90:250-502        ExactFunc cases 1 (refstate), 7 (Shu vortex)
  * testcase/taylorgreenvortex/testcase.f90:107-160, 220-262     TGV constants and initial condition
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .basis import PI


@dataclass
class Eos:
    kappa: float = 1.4
    R: float = 287.058
    Pr: float = 0.72
    mu0: float = 0.0
    visc_law: int = 0          # PP_VISC: 0 constant, 1 Sutherland
    Ts: float = 110.4          # Sutherland (ini values, dimensional)
    Tref: float = 280.0
    ExpoSuth: float = 1.5

    def eos_vars(self) -> np.ndarray:
        """EOS_Vars(1:8): kappa, R, Pr, mu0, Ts, Tref, ExpoSuth, cSuth (eos.f90:150-203)."""
        v = np.zeros(8)
        v[0], v[1], v[2], v[3] = self.kappa, self.R, self.Pr, self.mu0
        if self.visc_law == 1:
            Tref = 1.0 / self.Tref
            Ts = self.Ts * Tref
            v[4] = Ts
            v[5] = Tref
            v[6] = self.ExpoSuth
            v[7] = Ts ** self.ExpoSuth * (1 + Ts) / (2 * Ts * Ts)
        return v


def prim_to_cons(prim: np.ndarray, kappa: float) -> np.ndarray:
    """eos.f90:467-489 on arrays [..., 6] -> [..., 5]."""
    cons = np.empty(prim.shape[:-1] + (5,))
    cons[..., 0] = prim[..., 0]
    cons[..., 1] = prim[..., 1] * prim[..., 0]
    cons[..., 2] = prim[..., 2] * prim[..., 0]
    cons[..., 3] = prim[..., 3] * prim[..., 0]
    cons[..., 4] = prim[..., 4] / (kappa - 1.0) + 0.5 * (cons[..., 1] * prim[..., 1] + cons[..., 2] * prim[..., 2]
                                                        + cons[..., 3] * prim[..., 3])
    return cons


def cons_to_prim(cons: np.ndarray, kappa: float, R: float) -> np.ndarray:
    """eos.f90:212-239 on arrays [..., 5] -> [..., 6]."""
    prim = np.empty(cons.shape[:-1] + (6,))
    srho = 1.0 / cons[..., 0]
    prim[..., 0] = cons[..., 0]
    prim[..., 1] = cons[..., 1] * srho
    prim[..., 2] = cons[..., 2] * srho
    prim[..., 3] = cons[..., 3] * srho
    prim[..., 4] = (kappa - 1.0) * (cons[..., 4] - 0.5 * (cons[..., 1] * prim[..., 1] + cons[..., 2] * prim[..., 2]
                                                         + cons[..., 3] * prim[..., 3]))
    prim[..., 5] = prim[..., 4] * srho / R
    return prim


def refstate_prim(refstates, eos: Eos) -> np.ndarray:
    """RefStatePrim(6,nRefState) as C array [nRefState,6] (equation.f90:127-143)."""
    rs = np.atleast_2d(np.asarray(refstates, dtype=np.float64))
    out = np.zeros((rs.shape[0], 6))
    out[:, :5] = rs[:, :5]
    out[:, 5] = rs[:, 4] * (1.0 / rs[:, 0]) / eos.R
    return out


def init_bc_refstates(refprim: np.ndarray, BoundaryType: np.ndarray) -> np.ndarray:
    """InitBC (getboundaryflux.f90:194-212): for every boundary of type 27 (subsonic inflow) the refstate
    (Tt, alpha, beta, <empty>, pt) is rewritten in place to (Tt, a1, a2, a3, pt) with the unit direction vector a.
    Also the refstate sanity checks of :121-165 (Abort -> ValueError)."""
    ref = np.array(refprim, dtype=np.float64, copy=True)
    need_state = {2: "No refstate (rho,velx,vely,velz,p) defined for BC_TYPE",
                  4: "No refstate (rho,x,x,x,p) defined to compute temperature from density and pressure for BC_TYPE",
                  23: "No outflow Mach number in refstate (x,Ma,x,x,x) defined for BC_TYPE",
                  24: "No outflow pressure in refstate (x,x,x,x,p) defined for BC_TYPE",
                  25: "No outflow pressure in refstate (x,x,x,x,p) defined for BC_TYPE",
                  27: "No inflow refstate (Tt,alpha,beta,<empty>,pT) in refstate defined for BC_TYPE"}
    for i in range(BoundaryType.shape[0]):
        t, st = int(BoundaryType[i, 0]), int(BoundaryType[i, 1])
        if t in need_state and st < 1:
            raise ValueError(f"{need_state[t]} {t}")
        if t in need_state and st > ref.shape[0]:
            raise ValueError(f"ERROR: Boundary RefState not defined! (MaxBCState,nRefState): {st} {ref.shape[0]}")
        if t == 27:
            talpha = math.tan(PI / 180.0 * ref[st - 1, 1])
            tbeta = math.tan(PI / 180.0 * ref[st - 1, 2])
            ref[st - 1, 1] = 1.0 / math.sqrt((1.0 + talpha ** 2 + tbeta ** 2))
            ref[st - 1, 2] = talpha / math.sqrt((1.0 + talpha ** 2 + tbeta ** 2))
            ref[st - 1, 3] = tbeta / math.sqrt((1.0 + talpha ** 2 + tbeta ** 2))
    return ref


def bc_sides(mesh) -> np.ndarray:
    """BCSides(2,nBCSides) as C array [nBCSides,2] (getboundaryflux.f90:244-252)."""
    out = np.zeros((mesh.nBCSides, 2), dtype=np.int32)
    if mesh.nBCSides:
        out[:, 0] = mesh.BoundaryType[mesh.BC - 1, 0]
        out[:, 1] = mesh.BoundaryType[mesh.BC - 1, 1]
    return out


# --------------------------------------------------------------------------------------------------
# initial conditions (FillIni, dg.f90:434-458): U[e,k,j,i,5] from Elem_xGP[e,k,j,i,3]
# --------------------------------------------------------------------------------------------------
def ini_refstate(x: np.ndarray, refprim: np.ndarray, eos: Eos) -> np.ndarray:
    cons = prim_to_cons(refprim[None, :], eos.kappa)[0]
    return np.broadcast_to(cons, x.shape[:-1] + (5,)).copy()


def ini_tgv(x: np.ndarray, eos: Eos, mach: float = 0.1, ini_const_dens: bool = True) -> np.ndarray:
    """Taylor-Green vortex (testcase.f90:241-258); rho0 = U0 = 1."""
    rho0, U0 = 1.0, 1.0
    p0 = (U0 / mach) ** 2 / eos.kappa * rho0
    T0 = p0 / (rho0 * eos.R)
    prim = np.zeros(x.shape[:-1] + (6,))
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    prim[..., 1] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
    prim[..., 2] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
    prim[..., 3] = 0.0
    prim[..., 4] = p0 + (rho0 * U0 ** 2) / 16.0 * (np.cos(2 * X) * np.cos(2.0 * Z) + 2.0 * np.cos(2.0 * Y)
                                                  + 2.0 * np.cos(2.0 * X) + np.cos(2 * Y) * np.cos(2.0 * Z))
    if ini_const_dens:
        prim[..., 0] = rho0
        prim[..., 5] = prim[..., 4] / (prim[..., 0] * eos.R)
    else:
        prim[..., 5] = T0
        prim[..., 0] = prim[..., 4] / (prim[..., 5] * eos.R)
    return prim_to_cons(prim, eos.kappa)


def ini_shu_vortex(x: np.ndarray, refprim: np.ndarray, eos: Eos, t: float = 0.0, center=(0.0, 0.0, 0.0),
                   axis=(0.0, 0.0, 1.0), amplitude: float = 0.2, halfwidth: float = 0.2) -> np.ndarray:
    """Isentropic vortex, ExactFunc case 7 (exactfunc.f90:482-502)."""
    kappa = eos.kappa
    prim = np.broadcast_to(refprim, x.shape[:-1] + (6,)).copy()
    vel = refprim[1:4]
    RT = refprim[4] / refprim[0]
    cent = x - (np.asarray(center) + vel * t)
    cent = np.cross(np.asarray(axis, dtype=np.float64), cent)
    cent = cent / halfwidth
    r2 = np.sum(cent * cent, axis=-1)
    du = amplitude / (2.0 * PI) * np.exp(0.5 * (1.0 - r2))
    dTemp = -(kappa - 1.0) / (2.0 * kappa * RT) * du ** 2
    prim[..., 0] = prim[..., 0] * (1.0 + dTemp) ** (1.0 / (kappa - 1.0))
    prim[..., 1:4] = prim[..., 1:4] + du[..., None] * cent
    prim[..., 4] = prim[..., 4] * (1.0 + dTemp) ** (kappa / (kappa - 1.0))
    prim[..., 5] = prim[..., 4] / (prim[..., 0] * eos.R)
    return prim_to_cons(prim, kappa)


def perturb(U: np.ndarray, amp: float = 1e-3, seed: int = 12345) -> np.ndarray:
    """Optional deterministic perturbation to defeat value-dependent shortcuts; keeps rho, p > 0."""
    rng = np.random.default_rng(seed)
    return U * (1.0 + amp * rng.standard_normal(U.shape))


def exact_func_4(x: np.ndarray, t: float, adv_vel=(0.3, 0.3, 0.3)) -> np.ndarray:
    """IniExactFunc = 4 (idealgas/exactfunc.f90:326-333): oblique sine wave, the manufactured solution of the convergence
    tests; needs the source term of CalcSource (exactfunc.f90:991-1023)."""
    omega = PI * 1.0
    a = adv_vel[0] * 2.0 * PI
    r = 2.0 + 0.1 * np.sin(omega * np.sum(x, axis=-1) - a * t)
    U = np.empty(x.shape[:-1] + (5,))
    U[..., 0:4] = r[..., None]
    U[..., 4] = r * r
    return U
