#!/bin/bash
# GPU session: all parity tests (no -x), tuning sweep, bench. usage: bash tools/gpu_r2.sh tag "SWEEP ARGS"
TAG=${1:-r2}; SWEEP=${2:-DGX_FLAGS=0,1,2,4,8,15}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -25 $OUT/pytest_$TAG.log
timeout 600 python tools/sweep_flags.py --out $OUT/sweep_$TAG.json $SWEEP > $OUT/sweep_$TAG.log 2>&1; echo "sweep exit $?"; cat $OUT/sweep_$TAG.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["roofline"]["kernel_ms_per_stage"], "frac", d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
