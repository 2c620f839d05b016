#!/usr/bin/env python
"""Launch-bound configurations: time per RK step with host-paced, device-paced and CUDA-graph stepping (dgx_run_steps) on one GPU.
BASELINE config #1 (Shu vortex, Euler, N=3, 8^3 elements) and config #5 (NACA0012, N=4, 652 elements)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from galaexi_b200 import dg
from galaexi_b200.host_standin import workloads as wl

out = {}
for name, build in (("config1_shu_vortex_N3_8x8x8", lambda: cases.shu_vortex_case(E=8, N=3)), ("config5_naca_N4_652", lambda: wl.naca(4))):
    c, U0 = build()
    res = {}
    for mode in ("host", "device", "graph"):
        s = dg.DGSolver(c)
        s.set_state(U0)
        dt0, _ = s.CalcTimeStep()
        kw = dict(adaptive=True, device_paced=mode != "host", graph=mode == "graph")
        s.run_steps(20, 0.0, dt0, **kw)
        best = min(s.run_steps(200, 0.0, dt0, **kw)[0] for _ in range(3))
        res[mode] = dict(ms_per_step=best / 200, pid_s=best * 1e-3 / 200 / (c.nDOF * c.timedisc.nRKStages), dof_updates_per_s=c.nDOF * c.timedisc.nRKStages * 200 / (best * 1e-3))
        s.FinalizeDG()
    out[name] = dict(nDOF=c.nDOF, **res)
    print(name, {k: round(v["ms_per_step"], 4) for k, v in res.items()}, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "paced_small.json"), "w"), indent=1)
