#!/bin/bash
TAG=${1:-r3g}; NG=${2:-2}; CASES=${3:-"tgv mortar001 channel"}
OUT=gpurun_out; mkdir -p $OUT
export MR_CHECK_WATCHDOG=110
PORT=29810
for c in $CASES; do
  PORT=$((PORT+1))
  timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT tools/mr_check.py $c > $OUT/mr${NG}_${c}_$TAG.log 2>&1
  echo "mr_check $c exit $?"; grep MRCHECK $OUT/mr${NG}_${c}_$TAG.log | cut -c1-700
  grep -q MRCHECK $OUT/mr${NG}_${c}_$TAG.log || (grep -v "^\s*$" $OUT/mr${NG}_${c}_$TAG.log | grep "Error\|error\|failed\|watchdog" | head -5)
done
PORT=$((PORT+1))
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 20 --warmup 3 --watchdog 300 > $OUT/bench_n${NG}_$TAG.json 2> $OUT/bench_n${NG}_$TAG.err
echo "bench exit $?"; tail -3 $OUT/bench_n${NG}_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["config"]["step_pacing"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_stage"].items()}, "e2e %.3e" % d["e2e"]["value"])
    print("parity", json.dumps(d.get("parity"))[:1200])
    print("x3", json.dumps((d.get("extras") or {}).get("config3_weak"))[:600])
except Exception as ex:
    print("bench parse failed", ex)
PY
