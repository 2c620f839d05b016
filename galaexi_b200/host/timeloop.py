"""Time loop control: dt selection and end-time clipping around the RK step.

Mirrors the control flow of /root/reference/src/timedisc/timedisc.f90:36-203 (TimeDisc) and
src/timedisc/timedisc_func.f90:246-300 (UpdateTimeStep) for the part that touches the hot path:
dt = min(CalcTimeStep, tAnalyze-t, tEnd-t) with the "within 1% of the end -> take the rest" rule.
Works with any operator exposing ``calc_timestep() -> (dt, dt_conv, dt_visc)`` and ``rk_step(t, dt)``
(the CUDA solver galaexi_b200.dg.DGSolver or the CPU oracle used by the tests).
"""
from __future__ import annotations


def advance(op, t0: float, tEnd: float, maxIter: int | None = None, fixed_dt: float | None = None,
            nCalcTimeStepMax: int = 1, after_step=None):
    """Advance ``op`` from t0 to tEnd; returns (t, nTimeSteps).

    nCalcTimeStepMax (ini key NCalcTimeStepMax, default 1 = every step): the reference re-evaluates dt only every
    n-th step, n = min(floor(|log10((dt_old/dt - 1)^2 * 100 + eps)|), nCalcTimeStepMax), i.e. less often the slower dt
    changes (timedisc_func.f90:265-280); the end-time clipping is applied when dt is evaluated, as in the reference.
    after_step(t_new, dt): the per-step part of AnalyzeTimeStep (timedisc_func.f90:351-357), e.g. the Pruett filter."""
    import math
    import sys
    t = t0
    it = 0
    nCalc = 0
    dt_old = -999.0
    dt_keep = None
    while True:
        if maxIter is not None and it >= maxIter:
            break
        if nCalc >= 1 and dt_keep is not None and tEnd - t > dt_keep * 1.01:
            nCalc -= 1
            op.rk_step(t, dt_keep)
            t += dt_keep
            it += 1
            if after_step is not None:
                after_step(t, dt_keep)
            continue
        dt_min = fixed_dt if fixed_dt is not None else op.calc_timestep()[0]
        if nCalcTimeStepMax > 1:
            arg = abs(dt_old / dt_min - 1.0) ** 2 * 100.0 + sys.float_info.epsilon
            nCalc = min(int(math.floor(abs(math.log10(arg)))), nCalcTimeStepMax) - 1
            dt_old = dt_min
            dt_keep = dt_min
        dt_end = tEnd - t
        dt = min(dt_min, dt_end)
        finalize = dt == dt_end
        if dt_end - dt < dt / 100.0 and dt_end > 0:
            dt = dt_end
            finalize = True
        op.rk_step(t, dt)
        t += dt
        it += 1
        if after_step is not None:
            after_step(t, dt)
        if finalize:
            t = tEnd
            break
    return t, it
