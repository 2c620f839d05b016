#!/bin/bash
# 2-GPU verification with tight timeouts. usage: bash tools/gpu_multi_r3f.sh tag NG "cases"
TAG=${1:-r3f}; NG=${2:-2}; CASES=${3:-"tgv channel"}
OUT=gpurun_out; mkdir -p $OUT
L=galaexi_b200/csrc
timeout 300 python tools/ab_bench.py --degree 5 --elems 32 --mode graph --tag ${TAG}_N5 $L/libdgx.so $L/libdgx_va.so $L/libdgx_vb.so $L/libdgx_vc.so $L/libdgx_vd.so $L/libdgx_ve.so 2>> $OUT/ab_$TAG.err | cut -c1-300
timeout 200 python -m pytest tests -m gpu -q -k "overintegration or channel or paced" 2>&1 | tail -4
PORT=29710
for c in $CASES; do
  PORT=$((PORT+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT tools/mr_check.py $c > $OUT/mr${NG}_${c}_$TAG.log 2>&1
  echo "mr_check $c exit $?"; grep MRCHECK $OUT/mr${NG}_${c}_$TAG.log | cut -c1-600
  grep -q MRCHECK $OUT/mr${NG}_${c}_$TAG.log || (grep -v "^\s*$" $OUT/mr${NG}_${c}_$TAG.log | grep "Error\|error\|failed" | head -5)
done
PORT=$((PORT+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/bench_n${NG}_$TAG.json 2> $OUT/bench_n${NG}_$TAG.err
echo "bench exit $?"; tail -3 $OUT/bench_n${NG}_$TAG.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${NG}_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("value %.4e ms/step %.3f pid %.4e" % (d["value"], d["ms_per_step"], d["pid_s"]), d["config"]["step_pacing"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_stage"].items()}, "e2e %.3e" % d["e2e"]["value"])
    print("parity", json.dumps(d.get("parity"))[:900])
    print("x3", json.dumps((d.get("extras") or {}).get("config3_weak"))[:500])
except Exception as ex:
    print("bench parse failed", ex)
PY
PORT=$((PORT+1))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NG --steps 20 --warmup 3 --pacing host --no-extras --no-parity > $OUT/bench_n${NG}_host_$TAG.json 2> $OUT/bench_n${NG}_host_$TAG.err
echo "bench host exit $?"; python -c "
import json
d=json.loads([l for l in open('$OUT/bench_n${NG}_host_$TAG.json').read().strip().splitlines() if l.startswith('{')][-1]); print('host pacing: value %.4e ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v,4) for k,v in d['roofline']['kernel_ms_per_stage'].items()})"
