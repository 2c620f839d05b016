"""The BASELINE.json configurations as the reference's parameter files describe them (stand-in for reading the .ini / mesh
files; used by bench.py, tools/ and tests/cases.py so that the measured and the tested workloads are the same objects).

    #2  tutorials/taylorgreenvortex + regressioncheck/checks/tgv/split/parameter.ini   TGV, N=7 GL, SplitDG=PI, RoeEntropyFix, BR1
    #3  the same at N=5 on the 64^3 box (weak: 32^3 per GPU; strong: 64^3 in total)
    #4  tutorials/plane_turbulent_channel_flow/parameter_flexi.ini + parameter_hopr.ini  channel Re_tau=180, N=5, walls (4)
    #5  tutorials/naca0012/parameter_flexi_navierstokes.ini                              NACA0012_652_Ng2 mesh, N=4 here
"""
from __future__ import annotations

import os

import numpy as np

from . import basis as bs
from . import case as cs
from . import equation as eq
from . import mesh as ms

TGV_EOS = dict(kappa=1.4, R=71.42857, Pr=0.72, mu0=6.25e-4)
TGV_REF = ((1.0, 1.0, 0.0, 0.0, 17194.8345650329),)
# CFLscale / DFLscale of the split-form TGV (regressioncheck/checks/tgv/split/parameter.ini:62-63). The tutorial's 0.9 belongs
# to its Gauss + overintegration set-up: at N=7 on Gauss-Lobatto nodes the explicit scheme is unstable with 0.9 (the CPU
# restatement of the reference blows up after ~30 steps on the 8^3 .. 32^3 boxes; 0.8 is stable).
TGV_CFL = 0.8

FIXTURES = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")


def load_mesh(name: str) -> dict:
    """A HOPR mesh converted to .npz by tools/make_golden.py (ElemInfo / SideInfo / NodeCoords / BC tables)."""
    m = np.load(os.path.join(FIXTURES, name))
    return dict(NGeo=int(m["NGeo"]), ElemInfo=m["ElemInfo"], SideInfo=m["SideInfo"], NodeCoords=m["NodeCoords"],
                BCNames=[str(s) for s in m["BCNames"]], BCType=m["BCType"])


def box_dims(ngpus: int, per_gpu: int):
    """Global element counts of the weak-scaling boxes: per_gpu^3 elements per GPU."""
    e = per_gpu
    return {1: (e, e, e), 2: (2 * e, e, e), 4: (2 * e, 2 * e, e), 8: (2 * e, 2 * e, 2 * e)}[ngpus]


def tgv(dims, N: int, nProcs: int = 1, myRank: int = 0, curved: bool = False, cfl: float = TGV_CFL):
    """Taylor-Green vortex, Navier-Stokes Re 1600 Ma 0.1 on the periodic box of dims elements, edge 2 pi per min(dims)."""
    L = tuple(2 * np.pi * d / min(dims) for d in dims)
    h = (ms.make_box_mesh(dims, x0=(0.0, 0.0, 0.0), x1=L, NGeo=2, deform=0.1) if curved
         else ms.make_box_mesh(dims, x0=(0.0, 0.0, 0.0), x1=L, NGeo=1))
    eos = eq.Eos(**TGV_EOS)
    c = cs.build_case(h, N, bs.NODETYPE_GL, split="PI", riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=TGV_REF,
                      nProcs=nProcs, myRank=myRank, CFLScale=cfl, DFLScale=cfl)
    return c, eq.ini_tgv(c.geo["Elem_xGP"], eos)


def channel_stretch(d, s):
    """Wall-normal grading (parameter_hopr.ini: StretchType 3 in y, DxMaxToDxMin 8): a tanh bell with the same max/min ratio."""
    if d != 1:
        return s
    return 0.5 * (1.0 + np.tanh(1.5 * (2.0 * s - 1.0)) / np.tanh(1.5))


CHANNEL_DPDX = -1.0   # testcase/channel/testcase.f90:139 (ChannelFlow: dpdx = -1)


def channel(dims, N: int = 5, nProcs: int = 1, myRank: int = 0, cfl: float = 0.5, **kw):
    """Plane turbulent channel, Re_tau = 180: 2 pi x 2 x pi box, isothermal walls (BC type 4) at y = +-1, periodic x and z,
    SplitDG PI, RoeEntropyFix, mu0 = 1/180, CFLscale = DFLscale = 0.5 (parameter_flexi.ini). The initial state is a laminar
    profile with a smooth 3-D disturbance (the reference's Reichardt profile + random-phase modes need its testcase init)."""
    h = ms.make_box_mesh(dims, x0=(0.0, -1.0, -np.pi / 2), x1=(2 * np.pi, 1.0, np.pi / 2),
                         bctype=["periodic", (4, 1), "periodic", (4, 1), "periodic", "periodic"], stretch=channel_stretch)
    eos = eq.Eos(kappa=1.4, R=71.42857, Pr=0.72, mu0=5.555555556e-3)
    args = dict(split="PI", riemann="RoeEntropyFix", parabolic=True, eos=eos, refstates=((1.0, 1.0, 0.0, 0.0, 17194.8345650329),),
                nProcs=nProcs, myRank=myRank, CFLScale=cfl, DFLScale=cfl)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_GL)
    c = cs.build_case(h, N, nt, **args)
    x = c.geo["Elem_xGP"]
    prim = np.broadcast_to(c.RefStatePrim[0], x.shape[:-1] + (6,)).copy()
    prim[..., 1] = 1.5 * (1.0 - x[..., 1] ** 2) * (1.0 + 0.1 * np.sin(2 * x[..., 0]) * np.cos(2 * x[..., 2]))
    prim[..., 2] = 0.05 * np.sin(x[..., 0]) * (1.0 - x[..., 1] ** 2)
    prim[..., 3] = 0.05 * np.cos(2 * x[..., 2]) * (1.0 - x[..., 1] ** 2)
    return c, eq.prim_to_cons(prim, args["eos"].kappa)


def naca(N: int = 4, nProcs: int = 1, myRank: int = 0, **kw):
    """NACA0012, Re = 5000, AoA 8 deg on the tutorial's curved NGeo=2 mesh of 652 elements (weak form on Gauss nodes, BR1,
    RoeEntropyFix; BC 2 with the reference state, adiabatic wall 3, periodic z), from the free stream with a smooth disturbance."""
    h = load_mesh("naca_mesh.npz")
    eos = eq.Eos(kappa=1.4, R=2.857142857, Pr=0.72, mu0=0.0002)
    args = dict(split=None, riemann="RoeEntropyFix", parabolic=True, eos=eos,
                refstates=((1.0, 0.990268069, 0.139173101, 0.0, 4.4642857),), nProcs=nProcs, myRank=myRank,
                user_bcs={"BC_inflow": (2, 1), "BC_outflow": (2, 1)}, CFLScale=0.9, DFLScale=0.9)
    args.update(kw)
    nt = args.pop("node_type", bs.NODETYPE_G)
    c = cs.build_case(h, N, nt, **args)
    x = c.geo["Elem_xGP"]
    U0 = eq.ini_refstate(x, c.RefStatePrim[0], args["eos"])
    s = 1.0 + 0.02 * np.sin(3.0 * x[..., 0]) * np.cos(2.0 * x[..., 1]) * np.cos(5.0 * x[..., 2] + 0.3)
    return c, U0 * s[..., None]
