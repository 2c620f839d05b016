#!/bin/bash
# GPU session: A/B of tuning builds + parity tests. usage: bash tools/gpu_ab.sh tag "lib1 lib2 ..." [pytest: 1/0]
TAG=${1:-ab}; LIBS=${2:-galaexi_b200/csrc/libdgx.so}; PYT=${3:-1}
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/ab_${TAG}_*.jsonl
timeout 600 python tools/ab_bench.py --degree 7 --elems 32 --tag ${TAG}_N7 $LIBS 2> $OUT/ab_${TAG}_N7.err | cut -c1-400
timeout 600 python tools/ab_bench.py --degree 5 --elems 32 --tag ${TAG}_N5 $LIBS 2> $OUT/ab_${TAG}_N5.err | cut -c1-400
if [ "$PYT" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
fi
