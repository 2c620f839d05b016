#!/bin/bash
TAG=${1:-r3t}; NG=${2:-4}
OUT=gpurun_out; mkdir -p $OUT
PORT=30010
run() { # name, args
  PORT=$((PORT+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 20 --warmup 4 --no-extras --no-parity --e2e-steps 1 --watchdog 130 $2 > $OUT/bench_n${NG}_$1_$TAG.json 2> $OUT/bench_n${NG}_$1_$TAG.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$1_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$1: value %.4e ms/step %.4f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["measurement"]["step_pacing"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_stage"].items()}, "launches", d["gpu_launches"])
except Exception as ex:
    print("$1 parse failed", ex); print(open("$OUT/bench_n${NG}_$1_$TAG.err").read()[-500:])
PY
}
run c2_device "--pacing device"
run c3_device "--config 3 --pacing device"
run c3_graph "--config 3 --pacing graph"
