"""Shared test cases: the reference's regression configurations + synthetic meshes (seeded, smooth fields)."""
import os

import numpy as np

from galaexi_b200.host_standin import basis as bs
from galaexi_b200.host_standin import case as cs
from galaexi_b200.host_standin import equation as eq
from galaexi_b200.host_standin import mesh as ms

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_mesh(name):
    m = np.load(os.path.join(GOLD, name))
    return dict(NGeo=int(m["NGeo"]), ElemInfo=m["ElemInfo"], SideInfo=m["SideInfo"], NodeCoords=m["NodeCoords"],
                BCNames=[str(s) for s in m["BCNames"]], BCType=m["BCType"])


TGV_EOS = dict(kappa=1.4, R=71.42857, Pr=0.72, mu0=6.25e-4)
TGV_REF = ((1.0, 1.0, 0.0, 0.0, 17194.8345650329),)


def tgv_split_case(nProcs=1, myRank=0, hopr=None, N=7, **kw):
    """regressioncheck/checks/tgv/split: N=7 GL, PI split form, RoeEntropyFix, BR1, 8^3 periodic."""
    h = hopr or load_mesh("tgv_split_mesh.npz")
    eos = eq.Eos(**TGV_EOS)
    args = dict(split="PI", riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=TGV_REF, CFLScale=0.8,
                DFLScale=0.8, useCurveds=False, nProcs=nProcs, myRank=myRank)
    args.update(kw)
    c = cs.build_case(h, N, bs.NODETYPE_GL, **args)
    U0 = eq.ini_tgv(c.geo["Elem_xGP"], eos, mach=0.1, ini_const_dens=True)
    return c, U0


def tgv_box_case(E=4, N=5, NGeo=1, deform=0.0, nProcs=1, myRank=0, perturb=0.0, **kw):
    """Synthetic TGV box [0,2pi]^3, optionally curved."""
    h = ms.make_box_mesh((E, E, E), x0=(0.0, 0.0, 0.0), x1=(2 * np.pi,) * 3, NGeo=NGeo, deform=deform)
    eos = eq.Eos(**TGV_EOS)
    args = dict(split="PI", riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=TGV_REF, nProcs=nProcs,
                myRank=myRank)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_GL)
    c = cs.build_case(h, N, nt, **args)
    U0 = eq.ini_tgv(c.geo["Elem_xGP"], eos)
    if perturb:
        U0 = eq.perturb(U0, perturb, seed=12345 + myRank)
    return c, U0


def cavity_case(nProcs=1, myRank=0, riemann="RoeEntropyFix", N=2, node_type=bs.NODETYPE_G, **kw):
    """regressioncheck/checks/parabolic/cavity_3D: N=2 Gauss, weak form, BR1, walls (4) + Dirichlet lid (2)."""
    h = load_mesh("cavity3d_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=1.0, Pr=0.72, mu0=0.01)
    c = cs.build_case(h, N, node_type, split=None, riemann=riemann, parabolic=True, eos=eos,
                      refstates=((1.0, 1.0, 0.0, 0.0, 71.4285714286), (1.0, 0.0, 0.0, 0.0, 71.4285714286)),
                      user_bcs={"BC_wall_left": (4, 1), "BC_wall_right": (4, 1), "BC_free": (2, 1)}, CFLScale=0.99,
                      DFLScale=0.4, useCurveds=False, nProcs=nProcs, myRank=myRank, **kw)
    U0 = eq.ini_refstate(c.geo["Elem_xGP"], c.RefStatePrim[1], eos)
    return c, U0


def shu_vortex_case(E=8, N=3, nProcs=1, myRank=0, **kw):
    """BASELINE config #1: Shu vortex, Euler, N=3, 8^3 periodic Cartesian box (ini/shuVortex)."""
    h = ms.make_box_mesh((E, E, E), x0=(-1.0, -1.0, -1.0), x1=(1.0, 1.0, 1.0))
    eos = eq.Eos(kappa=1.4, R=287.058)
    args = dict(split=None, riemann="LF", parabolic=False, eos=eos, refstates=((1.0, 1.0, 0.0, 0.0, 2.85714),),
                nProcs=nProcs, myRank=myRank)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    U0 = eq.ini_shu_vortex(c.geo["Elem_xGP"], c.RefStatePrim[0], eos, amplitude=0.2, halfwidth=0.5)
    return c, U0


def naca_case(N=3, nProcs=1, myRank=0, **kw):
    """regressioncheck/checks/naca/3D mesh: curved NGeo=2, BC 2 (refstate) + 3 (adiabatic wall) + periodic z."""
    h = load_mesh("naca_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=287.058, Pr=0.72, mu0=0.0002)
    args = dict(split=None, riemann="RoeEntropyFix", parabolic=True, eos=eos,
                refstates=((1.0, 0.990268069, 0.139173101, 0.0, 4.4642857),), nProcs=nProcs, myRank=myRank,
                # regressioncheck/checks/naca/3D/parameter.ini:50-53
                user_bcs={"BC_inflow": (2, 1), "BC_outflow": (2, 1)})
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    x = c.geo["Elem_xGP"]
    U0 = eq.ini_refstate(x, c.RefStatePrim[0], eos)
    # smooth deterministic perturbation so that every term of the operator is exercised
    s = 1.0 + 0.02 * np.sin(3.0 * x[..., 0]) * np.cos(2.0 * x[..., 1]) * np.cos(5.0 * x[..., 2] + 0.3)
    U0 = U0 * s[..., None]
    return c, U0


def mortar_case(mesh="002", N=3, nProcs=1, myRank=0, bc=None, **kw):
    """tutorials/convtest/CART_HEX_PERIODIC_MORTAR_*: Cartesian box with non-conforming interfaces of all three mortar
    types (1->4, 1->2 in eta, 1->2 in xi). bc: optional (BCType, BCState) put on all six (otherwise periodic) boundaries."""
    h = load_mesh(f"cart_mortar_{mesh}_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=287.058, Pr=0.72, mu0=0.01)
    args = dict(split=None, riemann="Roe", parabolic=True, eos=eos, refstates=((1.0, 0.3, 0.2, -0.1, 2.0),),
                nProcs=nProcs, myRank=myRank)
    if bc is not None:
        args["user_bcs"] = {nm: bc for nm in ("BC_z-", "BC_y-", "BC_x+", "BC_y+", "BC_x-", "BC_z+")}
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    x = c.geo["Elem_xGP"]
    L = h["NodeCoords"].max(axis=0) - h["NodeCoords"].min(axis=0)   # global extents: the same field on every rank
    kx = 2.0 * np.pi / np.where(L > 0, L, 1.0)
    prim = np.broadcast_to(c.RefStatePrim[0], x.shape[:-1] + (6,)).copy()
    ph = kx[0] * x[..., 0] + 0.3, kx[1] * x[..., 1] - 0.2, kx[2] * x[..., 2] + 0.1
    prim[..., 0] *= 1.0 + 0.1 * np.sin(ph[0]) * np.cos(ph[1]) * np.cos(ph[2])
    prim[..., 1] += 0.1 * np.cos(ph[0]) * np.sin(ph[1])
    prim[..., 2] += 0.1 * np.sin(ph[1] + ph[2])
    prim[..., 3] += 0.1 * np.cos(ph[2]) * np.sin(ph[0])
    prim[..., 4] *= 1.0 + 0.1 * np.cos(ph[0] + ph[1]) * np.sin(ph[2])
    U0 = eq.prim_to_cons(prim, eos.kappa)
    return c, U0


CONV_ADV = (0.3, 0.3, 0.3)     # AdvVel of regressioncheck/checks/convtest/h_3D/parameter.ini


def exact_sine(x, t, kappa=1.4, adv=CONV_ADV):
    """IniExactFunc = 2 (idealgas/exactfunc.f90:254-268): density sine wave advected with AdvVel -- an exact solution of
    the Euler equations on the periodic [-1,1]^3 box."""
    adv = np.asarray(adv)
    cent = x - adv * t
    rho = 1.0 * (1.0 + 0.3 * np.sin(2.0 * np.pi * 0.5 * np.sum(cent, axis=-1)))
    U = np.empty(x.shape[:-1] + (5,))
    U[..., 0] = rho
    U[..., 1:4] = rho[..., None] * adv
    U[..., 4] = 1.0 / (kappa - 1.0) + 0.5 * rho * np.sum(adv * adv)
    return U


def convtest_case(mesh="cart_periodic_004", N=3, nProcs=1, myRank=0, **kw):
    """tutorials/convtest + regressioncheck/checks/convtest: periodic [-1,1]^3, conforming (cart_periodic_*) or
    non-conforming (cart_mortar_*) meshes, Euler, exact function 2."""
    h = load_mesh(f"{mesh}_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=287.058)
    args = dict(split=None, riemann="RoeEntropyFix", parabolic=False, eos=eos, refstates=((1.0, 0.3, 0.0, 0.0, 0.71428571),),
                nProcs=nProcs, myRank=myRank, CFLScale=0.7, useCurveds=False)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    return c, exact_sine(c.geo["Elem_xGP"], 0.0)


def manufactured_case(mesh="cart_periodic_004", N=3, nProcs=1, myRank=0, **kw):
    """regressioncheck/checks/convtest/h_3D: Navier-Stokes, IniExactFunc=4 (oblique sine wave) with the source term of
    CalcSource, AdvVel=(0.3,0.3,0.3), mu0=1e-3, CFLScale=DFLScale=0.7 (parameter.ini)."""
    h = kw.pop("hopr", None) or load_mesh(f"{mesh}_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=287.058, Pr=0.72, mu0=1.0e-3)
    args = dict(split=None, riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=((1.0, 0.3, 0.0, 0.0, 0.71428571),),
                nProcs=nProcs, myRank=myRank, CFLScale=0.7, DFLScale=0.7, useCurveds=False, IniExactFunc=4, AdvVel=CONV_ADV)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    return c, eq.exact_func_4(c.geo["Elem_xGP"], 0.0, CONV_ADV)


def l2_error(c, U, t, NAnalyze=None, exact=None):
    """CalcErrorNorms (analyze/analyze.f90:383-470): L2 error against the exact function on the analysis nodes."""
    from galaexi_b200.host_standin import analyze as an
    return an.calc_error_norms(c, U, t, exact or exact_sine, NAnalyze)[0]


def naca_regression_case(nProcs=1, myRank=0):
    """regressioncheck/checks/naca/3D/parameter.ini: NACA0012, Re=5000, AoA 8 deg; N=3 Gauss, weak form, BR1, build-default
    RoeEntropyFix, curved NGeo=2 mesh, BCs 2 (refstate) / 3 (adiabatic wall) / periodic z, sponge ramp from x=2 over a distance
    of 3 with the Pruett base flow (tempFilterWidth 2). Returns case, initial state, tempFilterWidth."""
    from galaexi_b200.host_standin import sponge as sp
    h = load_mesh("naca_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=2.857142857, Pr=0.72, mu0=0.0002)
    c = cs.build_case(h, 3, bs.NODETYPE_G, split=None, riemann="RoeEntropyFix", parabolic=True, eos=eos,
                      refstates=((1.0, 0.990268069, 0.139173101, 0.0, 4.4642857),), nProcs=nProcs, myRank=myRank,
                      user_bcs={"BC_inflow": (2, 1), "BC_outflow": (2, 1)}, CFLScale=0.9, DFLScale=0.9)
    U0 = eq.ini_refstate(c.geo["Elem_xGP"], c.RefStatePrim[0], eos)          # IniExactFunc = 1
    c.SpongeMat = sp.sponge_mat(c, [dict(shape=1, xStart=(2.0, 0.0, 0.0), dir=(1.0, 0.0, 0.0), distance=3.0)], damping=1.0)
    c.SpBaseFlow = U0.copy()                                                  # Pruett from scratch: the exact function
    return c, U0, 2.0


def channel_case(E=4, N=5, nProcs=1, myRank=0, **kw):
    """BASELINE config #4-like: plane channel, isothermal walls (4) at y+-, periodic x,z, Roe flux, y-stretched."""
    def stretch(d, s):
        if d != 1:
            return s
        return 0.5 * (1.0 + np.tanh(1.5 * (2.0 * s - 1.0)) / np.tanh(1.5))
    h = ms.make_box_mesh((E, E, E), x0=(0.0, -1.0, -np.pi / 2), x1=(2 * np.pi, 1.0, np.pi / 2),
                         bctype=["periodic", (4, 1), "periodic", (4, 1), "periodic", "periodic"], stretch=stretch)
    eos = eq.Eos(kappa=1.4, R=287.058, Pr=0.71, mu0=5.0e-4)
    args = dict(split="PI", riemann="Roe", parabolic=True, eos=eos, refstates=((1.0, 1.0, 0.0, 0.0, 71.4285714),),
                nProcs=nProcs, myRank=myRank)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_GL)
    c = cs.build_case(h, N, nt, **args)
    x = c.geo["Elem_xGP"]
    prim = np.broadcast_to(c.RefStatePrim[0], x.shape[:-1] + (6,)).copy()
    prim[..., 1] = 1.5 * (1.0 - x[..., 1] ** 2) * (1.0 + 0.1 * np.sin(2 * x[..., 0]) * np.cos(2 * x[..., 2]))
    prim[..., 2] = 0.05 * np.sin(x[..., 0]) * (1.0 - x[..., 1] ** 2)
    prim[..., 3] = 0.05 * np.cos(2 * x[..., 2]) * (1.0 - x[..., 1] ** 2)
    U0 = eq.prim_to_cons(prim, eos.kappa)
    return c, U0


# uniform subsonic duct flow and the matching total conditions (BC types 23/24/25/27)
DUCT_KAPPA, DUCT_R = 1.4, 287.058
DUCT_RHO, DUCT_U, DUCT_P = 1.2, 60.0, 1.0e5
DUCT_MACH = DUCT_U / np.sqrt(DUCT_KAPPA * DUCT_P / DUCT_RHO)
DUCT_T = DUCT_P / (DUCT_RHO * DUCT_R)
DUCT_TT = DUCT_T * (1.0 + 0.5 * (DUCT_KAPPA - 1.0) * DUCT_MACH ** 2)
DUCT_PT = DUCT_P * (1.0 + 0.5 * (DUCT_KAPPA - 1.0) * DUCT_MACH ** 2) ** (DUCT_KAPPA / (DUCT_KAPPA - 1.0))
DUCT_REFS = ((DUCT_RHO, DUCT_U, 0.0, 0.0, DUCT_P),   # 1: the uniform flow itself (Dirichlet state / outflow pressure)
             (1.0, DUCT_MACH, 0.0, 0.0, 1.0),        # 2: type 23, outflow Mach number in the second slot
             (DUCT_TT, 0.0, 0.0, 0.0, DUCT_PT))      # 3: type 27 (Tt, alpha, beta, -, pt)


def duct_case(inflow, outflow, wall, N=3, parabolic=True, node_type=bs.NODETYPE_G, split=None, riemann="Roe",
              nelems=(3, 2, 2), deform=0.0, nProcs=1, myRank=0):
    """Duct along x: x- inflow, x+ outflow, y+- walls, z periodic (local-side order z-, y-, x+, y+, x-, z+).
    Returns the case, the uniform state and a sheared, perturbed state."""
    h = ms.make_box_mesh(nelems, bctype=["periodic", wall, outflow, wall, inflow, "periodic"], NGeo=2, deform=deform)
    eos = eq.Eos(kappa=DUCT_KAPPA, R=DUCT_R, Pr=0.72, mu0=1e-3 if parabolic else 0.0)
    c = cs.build_case(h, N, node_type, split=split, riemann=riemann, parabolic=parabolic, eos=eos, refstates=DUCT_REFS,
                      nProcs=nProcs, myRank=myRank)
    x = c.geo["Elem_xGP"]
    base = np.array([DUCT_RHO, DUCT_U, 0.0, 0.0, DUCT_P, DUCT_T])
    U_uniform = eq.ini_refstate(x, base, eos)
    prim = np.broadcast_to(base, x.shape[:-1] + (6,)).copy()
    prim[..., 0] *= 1.0 + 0.03 * np.sin(2.0 * x[..., 0]) * np.cos(x[..., 2] * np.pi)
    prim[..., 1] *= 1.0 + 0.2 * np.sin(2.0 * x[..., 1]) * np.cos(x[..., 0])
    prim[..., 2] = 5.0 * np.sin(x[..., 0] + x[..., 1])
    prim[..., 3] = 3.0 * np.cos(np.pi * x[..., 2]) * np.sin(x[..., 0])
    prim[..., 4] *= 1.0 + 0.02 * np.cos(x[..., 0]) * np.cos(x[..., 1])
    return c, U_uniform, eq.prim_to_cons(prim, eos.kappa)


def rel_l2(a, b):
    return float(np.sqrt(np.sum((a - b) ** 2)) / max(np.sqrt(np.sum(b ** 2)), 1e-300))


def tgv_oint_case(nProcs=1, myRank=0, N=11, NUnder=7, OverintegrationType="cutoff", node_type=bs.NODETYPE_G, **kw):
    """regressioncheck/checks/tgv/oInt: N=11, weak form (FLEXI_SPLIT_DG=ON is excluded), RoeEntropyFix, BR1, 4^3 periodic elements,
    OverintegrationType=1 (cut-off filter on JU_t) with NUnder=7, CFLscale = DFLscale = 0.9."""
    h = load_mesh("tgv_oint_mesh.npz")
    eos = eq.Eos(**TGV_EOS)
    args = dict(split=None, riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=TGV_REF, CFLScale=0.9, DFLScale=0.9,
                useCurveds=False, nProcs=nProcs, myRank=myRank, OverintegrationType=OverintegrationType, NUnder=NUnder)
    args.update(kw)
    c = cs.build_case(h, N, node_type, **args)
    U0 = eq.ini_tgv(c.geo["Elem_xGP"], eos, mach=0.1, ini_const_dens=True)
    return c, U0
