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


def time_disc(op, t0: float, tEnd: float, Analyze_dt: float, *, nWriteData: int = 1, maxIter: int = -1, nCalcTimeStepMax: int = 1,
              nAnalyzeTestCase: int = 10, initial_output: bool = True, on_testcase=None, on_step=None, on_analyze=None,
              on_write=None, on_error=None):
    """The reference's TimeDisc loop (timedisc.f90:36-203) with UpdateTimeStep (timedisc_func.f90:246-300) and AnalyzeTimeStep
    (:305-400) around an operator with ``calc_timestep() -> (dt, ...)``, ``rk_step(t, dt)`` and (optional)
    ``DGTimeDerivative_weakForm(t)``.

    Hooks (all optional), in the reference's order:
      on_testcase(t, doFinalize)  AnalyzeTestCase: every nAnalyzeTestCase-th step and at every analyze time (:351)
      on_step(t, dt)              per step after TimeStep: RecordPoints, TempFilterTimeDeriv (:352-357)
      on_write(t, tWriteData)     WriteBaseFlow / WriteState / Visualize every nWriteData-th analyze time and at the end (:370-388)
      on_analyze(t, iter)         Analyze: error norms, body forces ... at every analyze time (:390)
      on_error(t, tWriteData)     WriteState(isErrorFile=.TRUE.) before the NaN abort (:282)
    dt = min(dt_CFL, tAnalyze - t, tEnd - t), a remaining interval within 1 % of dt is taken in one step; dt is re-evaluated only
    every n-th step for nCalcTimeStepMax > 1. Returns (t, iter)."""
    import math
    import sys
    t = float(t0)
    tAnalyze = min(t + Analyze_dt, tEnd)
    WriteData_dt = Analyze_dt * nWriteData
    tWriteData = min(t + WriteData_dt, tEnd)
    it = iter_analyze = writeCounter = nCalc = 0
    doAnalyze = doFinalize = False
    dt_minOld = -999.0
    dt = None
    refresh = getattr(op, "DGTimeDerivative_weakForm", None)
    if initial_output:
        if refresh is not None:
            refresh(t)
        if on_testcase is not None:
            on_testcase(t, False)
        if on_write is not None:
            on_write(t, tWriteData)
    if t >= tEnd or maxIter == 0:
        return t, it
    if initial_output and on_analyze is not None:
        on_analyze(t, it)
    while True:
        # ---- UpdateTimeStep
        if nCalc >= 1:
            nCalc -= 1
        else:
            try:
                dt_cfl = op.calc_timestep()[0]
            except Exception:
                if on_error is not None:
                    on_error(t, tWriteData)
                raise
            dt_an, dt_end = tAnalyze - t, tEnd - t
            dt = min(dt_cfl, dt_an, dt_end)
            if dt == dt_an:
                doAnalyze = True
            if dt == dt_end:
                doAnalyze = doFinalize = True
            dt = min(x for x in (dt_cfl, dt_an, dt_end) if x > 0)
            arg = abs(dt_minOld / dt - 1.0) ** 2 * 100.0 + sys.float_info.epsilon
            nCalc = min(int(math.floor(abs(math.log10(arg)))), nCalcTimeStepMax) - 1
            dt_minOld = dt
            if dt_an - dt < dt / 100.0 and dt_an > 0:
                dt, doAnalyze = dt_an, True
            if dt_end - dt < dt / 100.0 and dt_end > 0:
                dt, doAnalyze, doFinalize = dt_end, True, True
        # ---- TimeStep
        op.rk_step(t, dt)
        it += 1
        iter_analyze += 1
        t += dt
        # ---- AnalyzeTimeStep
        if it == maxIter:
            tEnd = tAnalyze = tWriteData = t
            doAnalyze = doFinalize = True
        if doAnalyze and refresh is not None:
            refresh(t)
        if on_testcase is not None and (it % nAnalyzeTestCase == 0 or doAnalyze):
            on_testcase(t, doFinalize)
        if on_step is not None:
            on_step(t, dt)
        if doAnalyze:
            writeCounter += 1
            if writeCounter == nWriteData or doFinalize:
                tWriteData = min(tAnalyze + WriteData_dt, tEnd)
                if on_write is not None:
                    on_write(t, tWriteData)
            if on_analyze is not None:
                on_analyze(t, it)
            if writeCounter == nWriteData or doFinalize:
                writeCounter = 0
            iter_analyze = 0
            tAnalyze = min(tAnalyze + Analyze_dt, tEnd)
            doAnalyze = False
        if doFinalize:
            break
    return t, it
