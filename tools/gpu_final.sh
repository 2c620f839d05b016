#!/bin/bash
# final 1-GPU session of the round: full parity suite, diagnostics, bench lines of all configurations, ncu evidence
TAG=${1:-r3z}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
grep -v "^$" $OUT/pytest_$TAG.log | tail -12
timeout 200 python tools/diag_n7.py 2>&1 | tee $OUT/diag_n7_$TAG.log | tail -30
L=galaexi_b200/csrc
timeout 200 python tools/ab_bench.py --degree 7 --elems 32 --mode graph --tag ${TAG}_N7 $L/libdgx.so $L/libdgx_d8.so 2>> $OUT/ab_$TAG.err | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; tail -2 $OUT/bench_$TAG.err
for cfg in 3 4 5; do
  timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline --ncu-timeout 120 > $OUT/bench_c${cfg}_$TAG.json 2> $OUT/bench_c${cfg}_$TAG.err; echo "bench config $cfg exit $?"
done
timeout 200 python bench.py --curved --steps 10 --warmup 3 --no-cpu-baseline --no-ncu > $OUT/bench_curved_$TAG.json 2> $OUT/bench_curved_$TAG.err; echo "bench curved exit $?"
python - <<PY
import json
for nm in ("bench_$TAG", "bench_c3_$TAG", "bench_c4_$TAG", "bench_c5_$TAG", "bench_curved_$TAG"):
    try:
        d=json.loads(open("$OUT/%s.json" % nm).read().strip().splitlines()[-1])
        print(nm, "value %.4e ms/step %.4f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["measurement"]["step_pacing"], "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
        for k,v in d["roofline"]["kernels"].items(): print("    ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms","hbm_frac","fp64_frac","bound","dram_bytes_per_dof","algorithmic_bytes_per_dof")})
        print("    roof", d["roofline"]["kernel"], d["roofline"]["frac"], "stage", round(d["roofline"]["stage"]["frac"],4), d["roofline"].get("ncu_note"), "cpu", (d["cpu_baseline"] or {}).get("value"))
        if d.get("extras"): print("    extras", json.dumps(d["extras"])[:400])
    except Exception as ex:
        print(nm, "parse failed", ex)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras --no-ncu --pacing device > $OUT/ncu_launches_$TAG.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_volsurf2|k_lifting|k_sideflux" -s 9 -c 3 -f -o $OUT/prof_stage_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras --no-ncu --pacing device > $OUT/ncu_full_$TAG.log 2>&1; echo "full exit $?"
