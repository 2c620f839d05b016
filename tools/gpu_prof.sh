#!/bin/bash
# Profiling session for profiles/: ncu launch list of a short bench run + one --set full capture of each stage kernel.
# usage (under gpurun): bash tools/gpu_prof.sh tag
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_lifting|k_sideflux|k_volsurf' -s 9 -c 3 -f -o $OUT/prof_stage_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full_$TAG.log 2>&1
echo "full exit $?"; tail -3 $OUT/ncu_full_$TAG.log | cut -c1-300
