#!/bin/bash
# 4-GPU session: 4-rank parity checks (several neighbours per rank), then the weak-scaling bench at 4 ranks.
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
for c in tgv naca mortar004; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/mr_check.py $c > $OUT/mr4_${c}_$TAG.log 2>&1
  echo "$c exit $?"; grep MRCHECK $OUT/mr4_${c}_$TAG.log || tail -15 $OUT/mr4_${c}_$TAG.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > $OUT/bench_n4_$TAG.json 2> $OUT/bench_n4_$TAG.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_n4_$TAG.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["pid_s"], d["setup_s"], d["roofline"]["kernel_ms_per_stage"])
PY
tail -3 $OUT/bench_n4_$TAG.err
