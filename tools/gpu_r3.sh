#!/bin/bash
# usage: bash tools/gpu_r3.sh tag [pytest-expr]
TAG=${1:-r3}; EXPR=${2:-}
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$EXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -k "$EXPR" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
else
  timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
fi
grep -v "^$" $OUT/pytest_$TAG.log | tail -40
for mode in host device graph; do
  timeout 300 python tools/ab_bench.py --degree 7 --elems 32 --mode $mode --tag ${TAG}_N7 2>> $OUT/ab_$TAG.err | cut -c1-330
done
for mode in host graph; do
  timeout 300 python tools/ab_bench.py --config 5 --degree 4 --mode $mode --steps 50 --tag ${TAG}_naca 2>> $OUT/ab_$TAG.err | cut -c1-330
done
tail -5 $OUT/ab_$TAG.err
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["config"]["step_pacing"], "e2e %.3e" % d["e2e"]["value"], d["clocks"])
    for k,v in d["roofline"]["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
    print("roof", d["roofline"]["kernel"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"].get("ncu_note"), d["roofline"]["counts_source"][:60] if d["roofline"].get("counts_source") else None)
    print("extras", json.dumps(d.get("extras"))[:600])
    print("cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
except Exception as ex:
    print("bench parse failed", ex)
PY
