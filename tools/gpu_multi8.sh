#!/bin/bash
# 8-GPU session: default bench line (config #2 weak + extras.config3_weak + parity), NACA (config #5), a 2-rank mortar check
TAG=${1:-r3s}; NG=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
PORT=29950
run() { # name, args, timeout
  PORT=$((PORT+1))
  timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG $2 > $OUT/bench_n${NG}_$1_$TAG.json 2> $OUT/bench_n${NG}_$1_$TAG.err
  echo "bench $1 exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$1_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$1: value %.4e ms/step %.4f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["measurement"]["step_pacing"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_stage"].items()}, "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"])
    p=d.get("parity")
    if p: print("   parity ok", p["ok"], "ut", p["ut_rel_l2"], "u", p["u_rel_l2"], "dt", p["dt_rel"], {k:(v["ut_rel_l2"], v["ok"]) for k,v in p["cases"].items()})
    x=(d.get("extras") or {}).get("config3_weak")
    if x: print("   config3_weak value %.4e pid %.4e ms/step %.3f" % (x["value"], x["pid_s"], x["ms_per_step"]), {k: round(v["ms"],4) for k,v in x["kernels"].items()})
except Exception as ex:
    print("$1 parse failed", ex); print(open("$OUT/bench_n${NG}_$1_$TAG.err").read()[-800:])
PY
}
run default "--steps 20 --warmup 3 --watchdog 200" 220
run naca "--config 5 --steps 50 --warmup 5 --no-parity --e2e-steps 2 --watchdog 100" 120
export MR_CHECK_WATCHDOG=50
PORT=$((PORT+1))
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tools/mr_check.py mortar001 > $OUT/mr2_mortar001_$TAG.log 2>&1
echo "mr_check mortar001 exit $?"; grep MRCHECK $OUT/mr2_mortar001_$TAG.log | cut -c1-500
