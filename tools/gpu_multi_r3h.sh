#!/bin/bash
TAG=${1:-r3h}; NG=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
export MR_CHECK_WATCHDOG=50 MR_TRACE=1
PORT=29910
for c in mortar001 tgv; do
  PORT=$((PORT+1))
  timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT tools/mr_check.py $c > $OUT/mr${NG}_${c}_$TAG.log 2>&1
  echo "mr_check $c exit $?"; grep MRCHECK $OUT/mr${NG}_${c}_$TAG.log | cut -c1-700
  grep "^\[rank" $OUT/mr${NG}_${c}_$TAG.log | tail -12
done
PORT=$((PORT+1))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 5 --warmup 3 --no-extras --e2e-steps 1 --watchdog 180 > $OUT/bench_n${NG}_$TAG.json 2> $OUT/bench_n${NG}_$TAG.err
echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("value %.4e ms/step %.3f" % (d["value"], d["ms_per_step"]), d["config"]["step_pacing"])
    print("parity", json.dumps(d.get("parity"))[:1000])
except Exception as ex:
    print("bench parse failed", ex)
PY
