#!/usr/bin/env python
"""A/B timing of tuning builds of libdgx.so in ONE process on one B200: the case tables are built once, every library
(galaexi_b200/csrc/libdgx.so and the variants given) runs the same warm-up + K time steps + per-kernel stage profile, and the
state after the steps is compared with the first library's (a variant that changes results is flagged).

    python tools/ab_bench.py [--degree 7] [--elems 32] [--steps 10] [--config tgv|channel] lib1.so lib2.so ...
Prints one JSON line per library (appended to gpurun_out/ab_<tag>.jsonl when --tag is given)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--degree", type=int, default=7)
    ap.add_argument("--elems", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--curved", action="store_true")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--mode", default="host", choices=("host", "device", "graph"), help="time-step pacing of dgx_run_steps")
    args = ap.parse_args()
    import bench
    from galaexi_b200 import dg
    wl = bench.make_workload(args.config, "weak", 1, 0, degree=args.degree, elems=args.elems, curved=args.curved)
    c, U0 = wl["c"], wl["U0"]
    kw = dict(adaptive=True, device_paced=args.mode != "host", graph=args.mode == "graph")
    libs = args.libs or [dg.LIB_PATH]
    ref = None
    n = args.degree + 1
    for path in libs:
        dg._lib = None
        os.environ["DGX_LIB"] = os.path.abspath(path)
        s = dg.DGSolver(c, device=0)
        s.set_state(U0)
        dt0, _ = s.CalcTimeStep()
        s.run_steps(args.warmup, 0.0, dt0, **kw)
        s.sync()
        best = None
        for _ in range(3):
            ms, launches = s.run_steps(args.steps, 0.0, dt0, **kw)
            best = ms if best is None else min(best, ms)
        prof = {}
        for _ in range(5):
            for k, v in s.profile_stage(0.0, dt0).items():
                prof.setdefault(k, []).append(v)
        prof = {k: round(float(np.mean(v)), 4) for k, v in prof.items()}
        U = s.get_state()
        if ref is None:
            ref = U
            dev = 0.0
        else:
            dev = float(np.max(np.abs(U - ref)) / np.max(np.abs(ref)))
        s.FinalizeDG()
        line = dict(lib=os.path.basename(path), mode=args.mode, degree=args.degree, elems=args.elems, ms_per_step=round(best / args.steps, 4),
                    gdof_per_s=round(c.nDOF * 5 * args.steps / (best * 1e-3) / 1e9, 4), kernels_ms=prof, launches=int(launches),
                    finite=bool(np.isfinite(U).all()), rel_dev_vs_first=dev)
        print(json.dumps(line), flush=True)
        if args.tag:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"ab_{args.tag}.jsonl"), "a") as f:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
