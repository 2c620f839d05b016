#!/bin/bash
TAG=${1:-r3e}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "overintegration or face_arrays or restart_from_reference or h_convergence_manufactured or paced" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
grep -v "^$" $OUT/pytest_$TAG.log | tail -25
L=galaexi_b200/csrc
timeout 600 python tools/ab_bench.py --degree 7 --elems 32 --mode graph --tag ${TAG}_N7 $L/libdgx.so $L/libdgx_skipdiag.so 2>> $OUT/ab_$TAG.err | cut -c1-330
timeout 600 python tools/ab_bench.py --degree 5 --elems 32 --mode graph --tag ${TAG}_N5 $L/libdgx.so $L/libdgx_skipdiag.so $L/libdgx_epb1m5.so $L/libdgx_epb1m6.so 2>> $OUT/ab_$TAG.err | cut -c1-330
timeout 300 python tools/ab_bench.py --degree 7 --elems 32 --mode host --tag ${TAG}_N7 2>> $OUT/ab_$TAG.err | cut -c1-330
tail -5 $OUT/ab_$TAG.err
# ncu: N=5 kernels, full set with source
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_volsurf2|k_lifting" -s 12 -c 2 -f -o $OUT/prof_n5_$TAG python tools/ab_bench.py --degree 5 --elems 32 --mode host --steps 2 --warmup 1 > $OUT/ncu_n5_$TAG.log 2>&1; echo "ncu exit $?"
for cfg in 4 5; do
  timeout 400 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline --ncu-timeout 150 > $OUT/bench_c${cfg}_$TAG.json 2> $OUT/bench_c${cfg}_$TAG.err; echo "bench config $cfg exit $?"; tail -2 $OUT/bench_c${cfg}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_c${cfg}_$TAG.json").read().strip().splitlines()[-1])
    print("cfg$cfg value %.4e ms/step %.4f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["config"]["step_pacing"], "launches", d["gpu_launches"])
    for k,v in d["roofline"]["kernels"].items(): print("  ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as ex:
    print("parse failed", ex)
PY
done
