#!/bin/bash
# usage: bash tools/gpu_multi_r3.sh tag NGPUS "case1 case2 ..." [bench-extra-args]
TAG=${1:-mr3}; NG=${2:-2}; CASES=${3:-"tgv channel"}; BARGS=${4:-}
OUT=gpurun_out; mkdir -p $OUT
PORT=29610
for c in $CASES; do
  PORT=$((PORT+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT tools/mr_check.py $c > $OUT/mr${NG}_${c}_$TAG.log 2>&1
  echo "mr_check $c exit $?"; grep MRCHECK $OUT/mr${NG}_${c}_$TAG.log | cut -c1-700 || tail -5 $OUT/mr${NG}_${c}_$TAG.log
  grep -q MRCHECK $OUT/mr${NG}_${c}_$TAG.log || tail -15 $OUT/mr${NG}_${c}_$TAG.log
done
run_bench() { # name, env, args
  PORT=$((PORT+1))
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 20 --warmup 3 $3 > $OUT/bench_n${NG}_$1_$TAG.json 2> $OUT/bench_n${NG}_$1_$TAG.err
  echo "bench $1 exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$1_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$1: value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["config"]["step_pacing"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_stage"].items()}, "e2e %.3e" % d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("ut_rel_l2"), "x3", (d.get("extras") or {}).get("config3_weak", {}).get("value"))
except Exception as ex:
    print("bench parse failed", ex); print(open("$OUT/bench_n${NG}_$1_$TAG.err").read()[-1500:])
PY
}
run_bench default "X=1" "$BARGS"
run_bench noearly "DGX_NO_EARLY_HALO=1" "--no-extras --no-parity $BARGS"
run_bench host "X=1" "--no-extras --no-parity --pacing host $BARGS"
run_bench c3 "X=1" "--config 3 --no-parity $BARGS"
