"""Time loop control: dt selection and end-time clipping around the RK step.

Mirrors the control flow of /root/reference/src/timedisc/timedisc.f90:36-203 (TimeDisc) and
src/timedisc/timedisc_func.f90:246-300 (UpdateTimeStep) for the part that touches the hot path:
dt = min(CalcTimeStep, tAnalyze-t, tEnd-t) with the "within 1% of the end -> take the rest" rule.
Works with any operator exposing ``calc_timestep() -> (dt, dt_conv, dt_visc)`` and ``rk_step(t, dt)``
(the CUDA solver galaexi_b200.dg.DGSolver or the CPU oracle used by the tests).
"""
from __future__ import annotations


def advance(op, t0: float, tEnd: float, maxIter: int | None = None, fixed_dt: float | None = None):
    """Advance ``op`` from t0 to tEnd; returns (t, nTimeSteps)."""
    t = t0
    it = 0
    while True:
        if maxIter is not None and it >= maxIter:
            break
        dt_min = fixed_dt if fixed_dt is not None else op.calc_timestep()[0]
        dt_end = tEnd - t
        dt = min(dt_min, dt_end)
        finalize = dt == dt_end
        if dt_end - dt < dt / 100.0 and dt_end > 0:
            dt = dt_end
            finalize = True
        op.rk_step(t, dt)
        t += dt
        it += 1
        if finalize:
            t = tEnd
            break
    return t, it
