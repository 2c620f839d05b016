#!/usr/bin/env python
"""Executed FP64 flops and DRAM bytes per DOF of the three stage kernels for the BASELINE configurations, measured with ncu on
one RK step (bench.py --ncu-child under `ncu --metrics ...`), merged into profiles/kernel_counts.json (the fallback bench.py
uses when its own live ncu pass is not possible, e.g. N>1 or counters not permitted).

    python tools/ncu_counts.py 2 3 4 5        # configurations to capture (one GPU)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    scaling, degree, elems, curved, ncu_timeout = "weak", None, None, False, 600


def main():
    path = os.path.join(ROOT, "profiles", "kernel_counts.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for cfg in [int(x) for x in sys.argv[1:]] or [2]:
        a = A()
        a.config = cfg
        a.scaling = "strong" if cfg == 5 else "weak"
        d, _ = bench.workload_dims(cfg, a.scaling, 1, None)
        N = {2: 7, 3: 5, 4: 5, 5: 4}[cfg]
        ndof = (652 if cfg == 5 else d[0] * d[1] * d[2]) * (N + 1) ** 3
        c, note = bench.live_ncu_counts(a, ndof)
        if not c:
            print(cfg, "FAILED", note)
            continue
        c["source"] = f"tools/ncu_counts.py on one B200: ncu --metrics of one RK step of {bench.workload_desc(cfg, a.scaling, 1)}"
        out[f"cfg{cfg}_N{N}"] = c
        print(cfg, json.dumps(c))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
