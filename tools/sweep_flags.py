#!/usr/bin/env python
"""A/B sweep of the kernel tuning switches (KParams::flags via DGX_FLAGS, and any other DGX_* env switch) on the bench
workload, in ONE process: per setting, per-kernel CUDA-event times of an RK stage and the device time of K full steps.
usage: python tools/sweep_flags.py [--elems 32] [--N 7] [--steps 5] ENV1=a,b,c [ENV2=x,y] ...   (cartesian product)"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elems", type=int, default=32)
    ap.add_argument("--N", type=int, default=7)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("sweep", nargs="*")
    a = ap.parse_args()
    from galaexi_b200 import dg
    wl = bench.make_workload(2, 'weak', 1, 0, degree=a.N, elems=a.elems)
    c, U0 = wl['c'], wl['U0']
    keys = [s.split("=")[0] for s in a.sweep]
    vals = [s.split("=")[1].split(",") for s in a.sweep]
    rows = []
    for combo in itertools.product(*vals) if keys else [()]:
        for k, v in zip(keys, combo):
            os.environ[k] = v
        s = dg.DGSolver(c)
        s.set_state(U0)
        dt0, _ = s.CalcTimeStep()
        s.run_steps(2, 0.0, dt0, adaptive=True)
        prof = {}
        for _ in range(5):
            for k, v in s.profile_stage(0.0, dt0).items():
                prof.setdefault(k, []).append(v)
        prof = {k: round(float(np.mean(v)), 4) for k, v in prof.items()}
        ms, _ = s.run_steps(a.steps, 0.0, dt0, adaptive=True)
        s.FinalizeDG()
        row = dict(zip(keys, combo), ms_per_step=round(ms / a.steps, 3), gdof_s=round(c.nDOF * 5 * a.steps / ms / 1e6, 3), **prof)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
