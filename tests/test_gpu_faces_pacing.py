"""GPU parity beyond Ut / U: (a) the face arrays the kernels produce (face states, lifted gradient traces, numerical flux)
against the oracle's arrays on curved, unstructured and non-conforming meshes -- pins orientation (S2V2 / flip), master/slave
assignment and sign conventions on the KERNELS, not only on the oracle (unitTests/ProlongToFace.f90:70-114,
unitTests/SurfInt.f90:79-132); (b) device-paced / CUDA-graph time stepping (dgx_run_steps flags 4, 8) is bit-identical to the
host-paced call sequence CalcTimeStep + TimeStepByLSERKW2; (c) the oracle comparison of Ut at BASELINE config #2's full size."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _solver(c):
    from galaexi_b200.dg import DGSolver
    return DGSolver(c)


def _oracle(c):
    from oracle.oracle import Oracle
    return Oracle(c)


def _face_cases():
    return {
        "tgv_curved_gl": lambda: cases.tgv_box_case(E=3, N=4, NGeo=2, deform=0.05, perturb=1e-3),
        "tgv_curved_gauss_weak": lambda: cases.tgv_box_case(E=3, N=3, NGeo=2, deform=0.05, perturb=1e-3, split=None, riemann="Roe",
                                                            node_type="GAUSS"),
        "naca": lambda: cases.naca_case(N=3),
        "mortar004": lambda: cases.mortar_case("004", N=3),
        "mortar002_br2": lambda: cases.mortar_case("002", N=3, lifting="br2"),
        "channel_walls": lambda: cases.channel_case(E=3, N=4),
    }


@pytest.mark.parametrize("name", sorted(_face_cases()))
def test_face_arrays_match_the_oracle(name):
    c, U0 = _face_cases()[name]()
    # a different smooth factor per conserved variable: velocity and temperature vary in all three directions on every mesh
    # (the NACA start field scales all variables alike, which leaves the lifted variables constant)
    x = c.geo["Elem_xGP"]
    for v, (a, b, cc) in enumerate(((1.3, 0.7, 2.1), (0.9, 1.7, 1.1), (2.3, 0.5, 1.9), (1.1, 1.3, 0.6), (0.4, 2.2, 1.5))):
        U0[..., v] *= 1.0 + 0.01 * np.sin(a * x[..., 0] + 0.1 * v) * np.cos(b * x[..., 1] - 0.2) * np.sin(cc * x[..., 2] + 0.3 + v)
    m = c.mesh
    o, s = _oracle(c), _solver(c)
    o.set_state(U0)
    s.set_state(U0)
    o.time_derivative(0.0)
    s.DGTimeDerivative_weakForm(0.0)
    sides = np.arange(m.nSides)
    inner = (sides >= m.firstInnerSide - 1) & (sides <= m.lastInnerSide - 1)          # both sides hold element data
    notbig = (sides < m.nBCSides) | (sides >= m.firstInnerSide - 1)                   # everything but big mortar sides
    for nm in s.FACE_ARRAYS:
        ref = o.array(nm)
        if nm.startswith("grad"):
            ref = ref[..., 1:]          # the oracle lifts (rho,u,v,w,T), the library the four variables the viscous flux needs
        got = s.get_face_array(nm)
        if nm.endswith("_slave"):
            sel = inner                 # BC sides and big mortar sides have no slave data (dg.f90:118-121)
        elif nm == "Flux_master":
            sel = np.ones(m.nSides, bool)
        elif nm.startswith("grad"):
            sel = notbig                # gradient traces of big sides live on their small sides (lifting_br2.t90:126-137)
        else:
            sel = np.ones(m.nSides, bool)
        scale = max(float(np.abs(ref[sel]).max()), 1e-300)
        assert scale > 1e-6, (name, nm, scale)      # the comparison must not be one of round-off against round-off
        err = float(np.abs(got[sel] - ref[sel]).max()) / scale
        assert err <= 1e-11, (name, nm, err)
    s.FinalizeDG()
    o.close()


def _pacing_cases():
    def channel():
        c, U0 = cases.channel_case(E=4, N=5)
        return c, U0, True
    return {
        "tgv_split_N7": lambda: cases.tgv_split_case() + (False,),
        "tgv_curved_N5": lambda: cases.tgv_box_case(E=4, N=5, NGeo=2, deform=0.05, perturb=1e-3) + (False,),
        "shu_euler_gauss": lambda: cases.shu_vortex_case(E=4, N=3) + (False,),
        "naca_N4": lambda: cases.naca_case(N=4) + (False,),
        "mortar004": lambda: cases.mortar_case("004", N=3) + (False,),
        "channel_forcing": channel,
    }


@pytest.mark.parametrize("name", sorted(_pacing_cases()))
def test_device_paced_and_graph_steps_are_bit_identical(name):
    """timedisc.f90:176-200 executed three ways: host-paced (dt through the host every step), device-paced, device-paced with
    CUDA-graph replay. Same dt sequence, same state, same stored gradients, bit for bit."""
    from galaexi_b200.host_standin import analyze as an
    c, U0, forcing = _pacing_cases()[name]()
    nsteps = 7
    results = []
    for mode in ("host", "device", "graph"):
        s = _solver(c)
        s.set_state(U0)
        if forcing:
            bv = s.CalcForcing(Vol=an.volume(c))
            s.set_channel_forcing(-1.0, bv)
        dt0, err = s.CalcTimeStep()
        assert err == 0
        if mode == "host":
            dts, t = [], 0.0
            for _ in range(nsteps):
                if forcing:
                    s.set_channel_forcing(-1.0, s.CalcForcing(Vol=an.volume(c)))
                dt, err = s.CalcTimeStep()
                assert err == 0
                s.TimeStepByLSERKW2(t, dt)
                t += dt
                dts.append(dt)
            dts = np.array(dts)
        else:
            # two calls: the second one replays the graph captured at the end of the first
            s.run_steps(3, 0.0, dt0, adaptive=True, forcing=forcing, device_paced=True, graph=mode == "graph")
            d1 = s.dt_history().copy()
            s.run_steps(nsteps - 3, 0.0, dt0, adaptive=True, forcing=forcing, device_paced=True, graph=mode == "graph")
            dts = np.concatenate([d1, s.dt_history()])
            if mode == "graph":
                assert s.step_graph_active()
        g = s.get_gradients() if c.parabolic else None
        results.append((dts, s.get_state(), g))
        s.FinalizeDG()
    ref = results[0]
    for mode, r in zip(("device", "graph"), results[1:]):
        assert np.array_equal(r[0], ref[0]), (name, mode, r[0], ref[0])
        assert np.array_equal(r[1], ref[1]), (name, mode, float(np.abs(r[1] - ref[1]).max()))
        if ref[2] is not None:
            for a, b in zip(r[2], ref[2]):
                assert np.array_equal(a, b), (name, mode)


def test_host_paced_run_steps_with_forcing_matches_the_call_sequence():
    from galaexi_b200.host_standin import analyze as an
    c, U0 = cases.channel_case(E=4, N=5)
    vol = an.volume(c)
    s1, s2 = _solver(c), _solver(c)
    for s in (s1, s2):
        s.set_state(U0)
        s.set_channel_forcing(-1.0, s.CalcForcing(Vol=vol))
    t = 0.0
    for _ in range(3):
        s1.set_channel_forcing(-1.0, s1.CalcForcing(Vol=vol))
        dt, _ = s1.CalcTimeStep()
        s1.TimeStepByLSERKW2(t, dt)
        t += dt
    s2.run_steps(3, 0.0, 0.0, adaptive=True, forcing=True)
    assert np.array_equal(s1.get_state(), s2.get_state())
    s1.FinalizeDG()
    s2.FinalizeDG()


def test_device_paced_reports_an_inadmissible_state():
    """calctimestep.f90:134-146: the device-paced loop cannot abort mid-call; the flag is raised when the call returns."""
    from galaexi_b200.dg import DGError
    c, U0 = cases.tgv_box_case(E=2, N=3)
    s = _solver(c)
    U = U0.copy()
    U[1, 2, 1, 0, 0] = -1.0
    s.set_state(U)
    with pytest.raises(DGError, match="timestep is NaN"):
        s.run_steps(2, 0.0, 1e-4, adaptive=True, device_paced=True)
    s.FinalizeDG()


def test_full_size_ut_vs_oracle():
    """BASELINE config #2 at its full size (32^3 elements, N=7: 16.8 M DOF, 32 768 elements, 98 304 sides): Ut of the CUDA path
    against the OpenMP oracle on the same state -- index arithmetic at size, not only conservation properties."""
    import ctypes
    import os
    from galaexi_b200.host_standin import workloads as wl
    from oracle import parity
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(len(os.sched_getaffinity(0)))
    except OSError:
        pass
    c, U0 = wl.tgv((32, 32, 32), 7)
    rng = np.random.default_rng(7)
    U0 = U0 * (1.0 + 1e-3 * rng.standard_normal(U0.shape[:-1]))[..., None]     # every node different: no symmetry can hide a mix-up
    s = _solver(c)
    s.set_state(U0)
    s.DGTimeDerivative_weakForm(0.0)
    Ut = s.get_ut()
    s.FinalizeDG()
    o = _oracle(c)
    o.set_state(U0)
    Ut_ref = o.time_derivative(0.0).copy()
    o.close()
    err = parity.rel_l2(Ut, Ut_ref)
    worst = float(np.abs(Ut - Ut_ref).max() / np.abs(Ut_ref).max())
    assert err <= 1e-12 and worst <= 1e-11, (err, worst)


# ---- overintegration of JU_t (dg/overintegration.f90; the oracle's restatement is pinned by the reference's tgv/oInt CSV at N=11,
# tests/test_oracle_goldens.py; the kernels exist for N <= 9) -------------------------------------------------------------------------
def _oint_cases():
    return {
        "cutoff_gauss_curved": lambda: cases.tgv_box_case(E=3, N=6, NGeo=2, deform=0.05, perturb=1e-3, split=None, riemann="Roe",
                                                          node_type="GAUSS", OverintegrationType="cutoff", NUnder=3),
        "conscutoff_gauss_curved": lambda: cases.tgv_box_case(E=3, N=6, NGeo=2, deform=0.05, perturb=1e-3, split=None, riemann="Roe",
                                                              node_type="GAUSS", OverintegrationType="conscutoff", NUnder=4),
        "conscutoff_gl_curved": lambda: cases.tgv_box_case(E=3, N=5, NGeo=2, deform=0.05, perturb=1e-3, split=None,
                                                           riemann="RoeEntropyFix", OverintegrationType="conscutoff", NUnder=3),
        "cutoff_split_n7": lambda: cases.tgv_box_case(E=3, N=7, NGeo=2, deform=0.05, perturb=1e-3, OverintegrationType="cutoff", NUnder=5),
        "cutoff_source_manufactured": lambda: cases.manufactured_case("cart_periodic_002", N=4, OverintegrationType="cutoff", NUnder=3),
        "conscutoff_walls_channel": lambda: cases.channel_case(E=3, N=4, OverintegrationType="conscutoff", NUnder=2),
        "oint_reference_setup_n9": lambda: cases.tgv_oint_case(N=9, NUnder=6),
    }


@pytest.mark.parametrize("name", sorted(_oint_cases()))
def test_overintegration_parity(name):
    from oracle import parity
    c, U0 = _oint_cases()[name]()
    assert c.OverintegrationType in (1, 2)
    o, s = _oracle(c), _solver(c)
    o.set_state(U0)
    s.set_state(U0)
    Ut_ref = o.time_derivative(0.0).copy()
    s.DGTimeDerivative_weakForm(0.0)
    r = parity.ut_error(c, U0, s.get_ut(), Ut_ref, label="oint_" + name)
    # only the unperturbed low-Mach TGV start field of the reference's own set-up is cancellation-dominated (as in tgv/split)
    assert r["ok"] and (not r["used_extended"] or name == "oint_reference_setup_n9"), r
    dt_ref = o.calc_timestep()[0]
    dt, err = s.CalcTimeStep()
    assert err == 0 and abs(dt - dt_ref) <= 1e-13 * dt_ref
    t = 0.0
    for _ in range(2):
        o.rk_step(t, dt_ref)
        s.TimeStepByLSERKW2(t, dt_ref)
        t += dt_ref
    assert parity.rel_l2(s.get_state(), o.array("U")) <= 1e-10
    # the device-paced / graph form of the same steps
    s2 = _solver(c)
    s2.set_state(U0)
    if not c.IniExactFunc:
        s2.run_steps(2, 0.0, dt_ref, adaptive=True, device_paced=True, graph=True)
        s3 = _solver(c)
        s3.set_state(U0)
        tt = 0.0
        for _ in range(2):
            d_, _e = s3.CalcTimeStep()
            s3.TimeStepByLSERKW2(tt, d_)
            tt += d_
        assert np.array_equal(s2.get_state(), s3.get_state())
        s3.FinalizeDG()
    s2.FinalizeDG()
    s.FinalizeDG()
    o.close()
