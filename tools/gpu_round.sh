#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list and a full capture of the dominant kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" >> $OUT/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_volsurf -s 6 -c 2 -f -o $OUT/prof_volsurf_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full_$TAG.log 2>&1
tail -5 $OUT/pytest_$TAG.log; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
