#!/bin/bash
# 8-GPU session: one 8-rank parity check and the weak-scaling bench at 8 ranks.
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/mr_check.py tgv > $OUT/mr8_tgv_$TAG.log 2>&1
echo "tgv exit $?"; grep MRCHECK $OUT/mr8_tgv_$TAG.log || tail -15 $OUT/mr8_tgv_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 1 > $OUT/bench_n8_$TAG.json 2> $OUT/bench_n8_$TAG.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_n8_$TAG.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["pid_s"], d["setup_s"], d["roofline"]["kernel_ms_per_stage"])
PY
tail -3 $OUT/bench_n8_$TAG.err
