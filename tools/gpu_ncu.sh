#!/bin/bash
# ncu --set full of one kernel (regex) inside a short bench run. usage: bash tools/gpu_ncu.sh <kernel-regex> <tag> [skip] [count]
K=$1; TAG=$2; SKIP=${3:-6}; CNT=${4:-1}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o $OUT/prof_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_$TAG.log | cut -c1-300
