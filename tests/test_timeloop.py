"""The TimeDisc control flow (timedisc.f90:36-203, timedisc_func.f90:246-400) around a stub operator and around the oracle."""
import numpy as np

import cases
from galaexi_b200.host_standin import timeloop


class _Stub:
    """dt_CFL = 0.3 always; counts the calls."""

    def __init__(self, dt=0.3):
        self.dt, self.steps, self.rhs = dt, [], []

    def calc_timestep(self):
        return (self.dt, None, None)

    def rk_step(self, t, dt):
        self.steps.append((t, dt))

    def DGTimeDerivative_weakForm(self, t):
        self.rhs.append(t)


def test_analyze_and_write_cadence():
    op = _Stub(0.3)
    ev = []
    t, it = timeloop.time_disc(op, 0.0, 2.0, 0.5, nWriteData=2, nAnalyzeTestCase=3,
                               on_testcase=lambda t, fin: ev.append(("tc", round(t, 12), fin)),
                               on_analyze=lambda t, i: ev.append(("an", round(t, 12), i)),
                               on_write=lambda t, tw: ev.append(("wr", round(t, 12), round(tw, 12))),
                               on_step=lambda t, dt: ev.append(("st", round(t, 12))))
    assert t == 2.0 and it == 8
    # steps: 0.3, 0.2 (clipped to tAnalyze = 0.5), 0.3, 0.2, ...
    assert np.allclose([d for _, d in op.steps], [0.3, 0.2] * 4)
    an = [e for e in ev if e[0] == "an"]
    assert [a[1] for a in an] == [0.0, 0.5, 1.0, 1.5, 2.0]
    wr = [e for e in ev if e[0] == "wr"]
    assert wr == [("wr", 0.0, 1.0), ("wr", 1.0, 2.0), ("wr", 2.0, 2.0)]           # initial state, every 2nd analyze time, the end
    tc = [e for e in ev if e[0] == "tc"]
    assert [x[1] for x in tc] == [0.0, 0.5, 0.8, 1.0, 1.5, 2.0] and tc[-1][2] is True and not any(x[2] for x in tc[:-1])
    assert len([e for e in ev if e[0] == "st"]) == 8
    assert op.rhs == [0.0, 0.5, 1.0, 1.5, 2.0]                                   # refreshed faces / gradients before every analysis


def test_one_percent_rule_and_max_iter():
    op = _Stub(0.3)
    t, it = timeloop.time_disc(op, 0.0, 0.601, 10.0)          # 0.3 + 0.301: the rest within 1 % of dt is taken in one step
    assert it == 2 and t == 0.601 and abs(op.steps[1][1] - 0.301) < 1e-15
    op = _Stub(0.3)
    ev = []
    t, it = timeloop.time_disc(op, 0.0, 100.0, 50.0, maxIter=3, on_write=lambda t, tw: ev.append((t, tw)))
    assert it == 3 and abs(t - 0.9) < 1e-15 and ev[-1] == (t, t)
    assert timeloop.time_disc(_Stub(), 1.0, 1.0, 0.5) == (1.0, 0)


def test_time_disc_equals_advance_on_the_oracle():
    """Without analyze times inside (Analyze_dt > tEnd) the loop takes exactly the steps of timeloop.advance."""
    from oracle.oracle import Oracle
    c, U0 = cases.tgv_box_case(E=2, N=3, NGeo=2, deform=0.05, perturb=1e-3)

    class _Op:
        def __init__(self):
            self.o = Oracle(c)
            self.o.set_state(U0)

        def calc_timestep(self):
            return self.o.calc_timestep()

        def rk_step(self, t, dt):
            self.o.rk_step(t, dt)
    a, b = _Op(), _Op()
    dt0 = a.o.calc_timestep()[0]
    tEnd = 3.4 * dt0
    ta, ia = timeloop.advance(a, 0.0, tEnd)
    tb, ib = timeloop.time_disc(b, 0.0, tEnd, 10.0 * tEnd, initial_output=False)
    assert (ta, ia) == (tb, ib) and np.array_equal(a.o.array("U"), b.o.array("U"))
    a.o.close(); b.o.close()
