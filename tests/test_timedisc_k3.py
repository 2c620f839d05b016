"""Three-register low-storage Runge-Kutta (TimeDiscType LSERKK3: timedisc_vars.f90:464-773, timestep.f90:129-200).

No reference golden exists for these schemes beyond the free-stream criterion of regressioncheck/checks/timedisc/freestream_3D
(analyze_L2 = 1e-1), so the transcribed tables are pinned by their design properties: temporal order of accuracy on a
non-autonomous ODE system (a single wrong digit in any of the six coefficient arrays destroys it) and agreement of the
oracle's DG solution with the Carpenter RK4-5 solution."""
import numpy as np
import pytest

import cases
from galaexi_b200.host_standin import timedisc as td


def _step_k3(f, u, t, dt, T):
    b_dt = T.RKb * dt
    for i in range(T.nRKStages):
        ts = t if i == 0 else t + T.RKc[i] * dt
        ut = f(ts, u)
        if i == 0:
            up, s2 = u.copy(), u.copy()
        else:
            s2 = s2 + u * T.RKdelta[i]
            u = u * T.RKg1[i] + s2 * T.RKg2[i]
            u = u + up * T.RKg3[i]
        u = u + ut * b_dt[i]
    return u


@pytest.mark.parametrize("name,stages,order", [("ketchesonrk4-20", 20, 5), ("ketchesonrk4-18", 18, 4)])
def test_tables_have_their_design_order(name, stages, order):
    T = td.set_timedisc(name, 3, "GAUSS", 0.9, 0.9)
    assert T.kind == "LSERKK3" and T.nRKStages == stages and all(len(getattr(T, k)) == stages for k in ("RKb", "RKc", "RKdelta", "RKg1", "RKg2", "RKg3"))
    assert T.RKc[0] == 0.0 and T.RKg1[0] == 0.0 and T.RKg2[0] == 1.0 and T.RKdelta[0] == 1.0 and T.RKdelta[-1] == 0.0
    # CFL / DFL scaling: the reference uses the Niegemann RK4-14 calibration (timedisc_vars.f90:475-497)
    Tn = td.set_timedisc("niegemannrk4-14", 3, "GAUSS", 0.9, 0.9)
    assert T.CFLScale == Tn.CFLScale and T.DFLScale == Tn.DFLScale
    f = lambda t, u: np.array([u[1] * np.cos(t), -u[0] * (1 + 0.5 * np.sin(2 * t)) + 0.1 * u[1] ** 2])

    def run(n):
        u, t, dt = np.array([1.0, 0.3]), 0.0, 2.0 / n
        for _ in range(n):
            u = _step_k3(f, u, t, dt, T)
            t += dt
        return u
    ref = run(4096)
    e = [np.abs(run(n) - ref).max() for n in (8, 16, 32)]
    rates = [np.log2(e[i] / e[i + 1]) for i in range(2)]
    assert min(rates) > order - 0.25, (e, rates)


def test_unknown_scheme_is_rejected():
    with pytest.raises(ValueError, match="Unknown method of time discretization"):
        td.set_timedisc("ketchesonrk4-19", 3, "GAUSS", 0.9, 0.9)


@pytest.mark.parametrize("name", ["ketchesonrk4-20", "ketchesonrk4-18"])
def test_oracle_k3_free_stream_and_agreement_with_carpenter(name):
    """Oracle DG run: the free stream is preserved to round-off (timedisc/freestream_3D), and at a small fixed dt the solution
    equals the Carpenter RK4-5 one up to the (tiny) time-integration errors of both."""
    from galaexi_b200.host_standin import equation as eq
    from oracle.oracle import Oracle
    c, U0 = cases.tgv_box_case(E=2, N=3, NGeo=2, deform=0.05, perturb=1e-3, timedisc=name)
    o = Oracle(c)
    Ufs = eq.ini_refstate(c.geo["Elem_xGP"], c.RefStatePrim[0], c.eos)
    o.set_state(Ufs)
    dt = o.calc_timestep()[0]
    o.rk_step(0.0, dt)
    assert np.abs(o.array("U") - Ufs).max() <= 1e-10 * np.abs(Ufs).max()
    o.set_state(U0)
    dts = 0.05 * dt
    for k in range(2):
        o.rk_step(k * dts, dts)
    Uk = o.array("U").copy()
    o.close()
    c2, _ = cases.tgv_box_case(E=2, N=3, NGeo=2, deform=0.05, perturb=1e-3, timedisc="carpenterrk4-5")
    o2 = Oracle(c2)
    o2.set_state(U0)
    for k in range(2):
        o2.rk_step(k * dts, dts)
    Uc = o2.array("U").copy()
    o2.close()
    assert np.abs(Uc - U0).max() > 1e-9 * np.abs(U0).max()          # the steps did something
    assert np.abs(Uk - Uc).max() <= 1e-6 * np.abs(Uc - U0).max()     # ... and both schemes agree on what
