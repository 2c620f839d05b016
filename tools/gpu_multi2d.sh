#!/bin/bash
# 2-rank parity incl. wall diagnostics and the collectively written state file. usage: bash tools/gpu_multi2d.sh [tag]
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
for c in cavity naca channel tgv mortar002; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mr_check.py $c > $OUT/mr_${c}_$TAG.log 2>&1
  echo "$c exit $?"; grep MRCHECK $OUT/mr_${c}_$TAG.log || tail -15 $OUT/mr_${c}_$TAG.log
done
