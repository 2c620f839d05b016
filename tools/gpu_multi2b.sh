#!/bin/bash
# 2-GPU session: parity checks + weak-scaling bench with and without the second compute stream
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
for c in tgv cavity shu channel naca mortar002; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mr_check.py $c > $OUT/mr2_${c}_$TAG.log 2>&1
  echo "$c exit $?"; grep MRCHECK $OUT/mr2_${c}_$TAG.log | cut -c1-200 || tail -15 $OUT/mr2_${c}_$TAG.log
done
for v in split nosplit; do
  if [ $v = nosplit ]; then export DGX_NO_SPLIT_STREAM=1; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 1 > $OUT/bench_n2_${v}_$TAG.json 2> $OUT/bench_n2_${v}_$TAG.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_n2_${v}_$TAG.json").read().strip().splitlines()[-1])
print("$v", d["n_gpus"], d["value"], d["ms_per_step"], d["pid_s"])
PY
done
