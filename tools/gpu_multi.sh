#!/bin/bash
# Multi-GPU session: 2-rank parity checks, then the weak-scaling bench at N ranks. usage: bash tools/gpu_multi.sh N [tag]
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
export NCCL_DEBUG=WARN
for c in tgv cavity channel shu naca tgv_br2 mortar001 mortar002 mortar004_br2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mr_check.py $c > $OUT/mr_${c}_$TAG.log 2>&1
  echo "$c exit $?"; grep MRCHECK $OUT/mr_${c}_$TAG.log || tail -15 $OUT/mr_${c}_$TAG.log
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_$TAG.json 2> $OUT/bench_n1_$TAG.err; cat $OUT/bench_n1_$TAG.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
echo "bench exit $?"; cat $OUT/bench_n${N}_$TAG.json; tail -5 $OUT/bench_n${N}_$TAG.err
