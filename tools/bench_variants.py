#!/usr/bin/env python
"""Per-kernel stage timings of other configurations than the headline one (weak form on Gauss nodes, Euler, other degrees).
usage: python tools/bench_variants.py  (one B200)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cases  # noqa: E402
from galaexi_b200.dg import DGSolver  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0
out = []
for name, kw in (("N3 Gauss weak Euler (config #1 physics)", dict(N=3, node_type="GAUSS", split=None, riemann="LF", parabolic=False)),
                 ("N4 Gauss weak NS (config #5 physics)", dict(N=4, node_type="GAUSS", split=None, riemann="RoeEntropyFix")),
                 ("N5 GL split NS (config #3/#4 physics)", dict(N=5)),
                 ("N7 GL weak NS", dict(N=7, split=None, riemann="Roe"))):
    E = 32
    c, U0 = cases.tgv_box_case(E=E, **kw)
    s = DGSolver(c)
    s.set_state(U0)
    dt, _ = s.CalcTimeStep()
    s.run_steps(3, 0.0, dt, adaptive=True)
    ms, _ = s.run_steps(10, 0.0, dt, adaptive=True)
    prof = {}
    for _ in range(5):
        for k, v in s.profile_stage(0.0, dt).items():
            prof.setdefault(k, []).append(v)
    prof = {k: round(float(np.mean(v)), 4) for k, v in prof.items()}
    n = c.N + 1
    par = c.parabolic
    b_stage = 8.0 * (75.0 + 360.0 / n) if par else 8.0 * (30.0 + 135.0 / n)
    pid = ms * 1e-3 / (c.nDOF * 10 * 5)
    rec = dict(case=name, dof=c.nDOF, ms_per_step=ms / 10, gdof_s=c.nDOF * 50 / (ms * 1e-3) / 1e9, pid_ns=pid * 1e9,
               stage_frac_of_hbm_roof=b_stage / pid / 1e9 / HBM, kernel_ms=prof)
    print(json.dumps(rec), flush=True)
    out.append(rec)
    s.FinalizeDG()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_variants.json"), "w"), indent=1)
