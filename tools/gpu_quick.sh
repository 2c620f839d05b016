#!/bin/bash
# quick GPU iteration: parity tests + bench (no ncu). usage: bash tools/gpu_quick.sh tag [pytest-args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["roofline"]["kernel_ms_per_stage"], "frac", d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
