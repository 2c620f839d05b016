"""CalcBodyForces (equations/navierstokes/calcbodyforces.f90): numpy oracle properties + the reference's tutorial CSV, and
GPU parity of dgx_calc_body_forces against the oracle."""
import os

import numpy as np
import pytest

import cases
from galaexi_b200.host_standin import equation as eq
from oracle.analyze_body_forces import calc_body_forces, calc_wall_velocity


def _oracle_forces(c, U, t=0.0):
    from oracle.oracle import Oracle
    o = Oracle(c, "double")
    o.set_state(U)
    o.time_derivative(t)
    g = [o.array(nm).copy() for nm in ("gradUx_master", "gradUy_master", "gradUz_master")]
    out = calc_body_forces(c, o.array("UPrim_master").copy(), *g)
    o.close()
    return out


def test_uniform_state_pressure_force_is_p_times_area_vector():
    """Uniform state at rest in the lid-driven cavity box [0,1]^3 (walls: BC type 4): Fp of a wall BC = p * sum n dA over its
    sides; non-wall BCs (the type-2 lid) stay zero."""
    c, _ = cases.cavity_case()
    prim = np.array([1.0, 0.0, 0.0, 0.0, 71.4285714286, 0.0])
    U = eq.ini_refstate(c.geo["Elem_xGP"], prim, c.eos)
    Fp, Fv = _oracle_forces(c, U)
    m = c.mesh
    w = c.basis.wGP
    wS = w[:, None] * w[None, :]
    area_vec = np.zeros_like(Fp)
    for s in range(m.nBCSides):
        area_vec[int(m.BC[s]) - 1] += np.einsum("qp,qpd->d", wS * c.geo["SurfElem"][s], c.geo["NormVec"][s])
    for b in range(m.BoundaryType.shape[0]):
        if int(m.BoundaryType[b, 0]) in (3, 4, 9):
            assert np.allclose(Fp[b], prim[4] * area_vec[b], rtol=0, atol=1e-12)
        else:
            assert not Fp[b].any() and not Fv[b].any()
    # (Fv is not zero here: the moving lid's Dirichlet state enters the lifted gradients of the corner elements)
    walls = [b for b in range(m.BoundaryType.shape[0]) if int(m.BoundaryType[b, 0]) == 4]
    assert walls and np.abs(Fp[walls]).max() > 1.0


def test_naca_forces_against_reference_tutorial_csv():
    """tutorials/naca0012/NACA0012_Re5000_AoA8_BodyForces_BC_wall_Reference.csv (t = 10): Fp = (1.3836969744113e-2,
    1.7266366997869e-1), Fv = (1.9135540981598e-2, 3.0381410536723e-3). That run is the tutorial setup (no sponge, 2-D build);
    here the forces are evaluated on the reference's regression state of the same case (naca/3D, with sponge, 3-D, spanwise
    extent 1), so the comparison is a physical pin (sign conventions, magnitudes: within 6 %), not a bit-level one."""
    c = cases.naca_regression_case()
    c = c[0] if isinstance(c, tuple) else c
    U = np.load(os.path.join(cases.GOLD, "naca3d_state.npz"))["DG_Solution"]
    Fp, Fv = _oracle_forces(c, U, 10.0)
    b = c.mesh.BoundaryName.index("BC_wall")
    ref_p = np.array([0.13836969744113e-1, 0.17266366997869e0])
    ref_v = np.array([0.19135540981598e-1, 0.30381410536723e-2])
    assert np.all(np.abs(Fp[b, :2] / ref_p - 1.0) < 0.06), Fp[b]
    assert np.all(np.abs(Fv[b, :2] / ref_v - 1.0) < 0.06), Fv[b]
    assert abs(Fp[b, 2]) < 1e-12 and abs(Fv[b, 2]) < 1e-12          # spanwise symmetric
    others = [i for i in range(len(c.mesh.BoundaryName)) if i != b]
    assert not Fp[others].any() and not Fv[others].any()


def _gpu_forces(c, U, t=0.0):
    from galaexi_b200.dg import DGSolver
    s = DGSolver(c)
    s.set_state(U)
    s.DGTimeDerivative_weakForm(t)
    out = s.CalcBodyForces() + (s.CalcWallVelocity(),)
    s.FinalizeDG()
    return out


def _oracle_wall_velocity(c, U):
    from galaexi_b200.host_standin import analyze as an
    from oracle.oracle import Oracle
    o = Oracle(c, "double")
    o.set_state(U)
    o.time_derivative(0.0)
    out = calc_wall_velocity(c, o.array("UPrim_master").copy(), an.bc_surfaces(c))
    o.close()
    return out


def test_wall_velocity_oracle_uniform_flow():
    """Uniform |v| = 3 along slip walls of the duct: max = min = mean = 3 on the wall BCs, the reference's initial values
    (-1e14, 1e14, 0) elsewhere; Surf of a BC without sides is HUGE."""
    from galaexi_b200.host_standin import analyze as an
    c, U, _ = cases.duct_case((2, 1), (24, 1), (9, 0), parabolic=False)
    v0 = np.sqrt(np.sum((U[..., 1:4] / U[..., :1]) ** 2, axis=-1)).max()
    mx, mn, me = _oracle_wall_velocity(c, U)
    S = an.bc_surfaces(c)
    for b in range(c.mesh.BoundaryType.shape[0]):
        if int(c.mesh.BoundaryType[b, 0]) == 9:
            assert abs(mx[b] - v0) < 1e-12 * v0 and abs(mn[b] - v0) < 1e-12 * v0 and abs(me[b] - v0) < 1e-12 * v0 and S[b] < 1e3
        else:
            assert mx[b] == -1.e14 and mn[b] == 1.e14 and me[b] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["naca", "cavity", "channel", "naca_br2", "channel_sutherland", "euler_slip"])
def test_gpu_body_forces_match_oracle(name):
    if name == "naca":
        c = cases.naca_regression_case()
        c = c[0] if isinstance(c, tuple) else c
        U = np.load(os.path.join(cases.GOLD, "naca3d_state.npz"))["DG_Solution"]
    elif name == "naca_br2":
        c, U = cases.naca_case(lifting="br2")
    elif name == "cavity":
        c, U = cases.cavity_case()
    elif name == "channel_sutherland":
        eos = eq.Eos(kappa=1.4, R=287.058, Pr=0.71, mu0=5.0e-4, visc_law=1, Ts=0.4, Tref=4.0, ExpoSuth=1.5)
        c, U = cases.channel_case(eos=eos)
    elif name == "channel":
        c, U = cases.channel_case()
    else:
        c, _, U = cases.duct_case((2, 1), (24, 1), (9, 0), parabolic=False)
    rng = np.random.default_rng(7)
    U = U * (1.0 + 1e-2 * rng.standard_normal(U.shape))         # non-trivial velocity gradients at the walls
    Fp_o, Fv_o = _oracle_forces(c, U)
    Ft, Fp, Fv, (mx, mn, me) = _gpu_forces(c, U)
    mx_o, mn_o, me_o = _oracle_wall_velocity(c, U)
    assert np.allclose(mx, mx_o, rtol=1e-13, atol=0) and np.allclose(mn, mn_o, rtol=1e-13, atol=1e-300)
    assert np.allclose(me, me_o, rtol=1e-12, atol=1e-14 * max(np.abs(me_o).max(), 1e-300))
    scale = max(np.abs(Fp_o).max(), 1e-300)
    assert np.abs(Fp - Fp_o).max() <= 1e-12 * scale
    if c.parabolic:
        assert np.abs(Fv_o).max() > 0
        assert np.abs(Fv - Fv_o).max() <= 1e-11 * np.abs(Fv_o).max()
    else:
        assert not Fv.any()
    assert np.array_equal(Ft, Fp + Fv)
    nonwall = [b for b in range(c.mesh.BoundaryType.shape[0]) if int(c.mesh.BoundaryType[b, 0]) not in (3, 4, 9)]
    assert not Fp[nonwall].any() and not Fv[nonwall].any()


@pytest.mark.gpu
def test_gpu_body_forces_bad_arguments():
    import ctypes as C
    from galaexi_b200.dg import DGSolver, _dp
    c, U = cases.cavity_case()
    s = DGSolver(c)
    s.set_state(U)
    s.DGTimeDerivative_weakForm(0.0)
    w = np.ascontiguousarray(c.basis.wGP)
    bc = np.full(c.mesh.nBCSides, 99, dtype=np.int32)
    F = np.zeros((3, 3))
    rc = s.lib.dgx_calc_body_forces(s.h, w.ctypes.data_as(_dp), bc.ctypes.data_as(C.POINTER(C.c_int)), 3, F.ctypes.data_as(_dp),
                                    F.ctypes.data_as(_dp))
    assert rc != 0 and b"outside 1..nBCs" in s.lib.dgx_last_error(s.h)
    s.FinalizeDG()
