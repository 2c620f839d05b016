#!/bin/bash
# 2-GPU session: rank-reduced diagnostics, filter, source term and mortar cases through NCCL
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
for c in tgv channel tgv_filter manufactured mortar004_br2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/mr_check.py $c > $OUT/mr2c_${c}_$TAG.log 2>&1
  echo "$c exit $?"; grep MRCHECK $OUT/mr2c_${c}_$TAG.log | cut -c1-330 || tail -15 $OUT/mr2c_${c}_$TAG.log
done
