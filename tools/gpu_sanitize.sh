#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over small cases of every kernel family.
# usage (under gpurun): bash tools/gpu_sanitize.sh tag
TAG=${1:-san}; OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cases
from galaexi_b200.dg import DGSolver
def run(c, U0, analyze=False):
    s = DGSolver(c); s.set_state(U0); s.DGTimeDerivative_weakForm(0.0); dt, _ = s.CalcTimeStep(); s.TimeStepByLSERKW2(0.0, dt)
    if analyze: s.AnalyzeTestcase()
    u = s.get_state(); assert np.isfinite(u).all(); s.FinalizeDG()
run(*cases.tgv_box_case(E=2, N=7, NGeo=2, deform=0.05, perturb=1e-3), analyze=True)      # DMMA lifting + volsurf2 + analysis
run(*cases.tgv_box_case(E=2, N=5, NGeo=2, deform=0.05, perturb=1e-3))                     # TMA lifting + volsurf2 (2 elements per CTA)
run(*cases.tgv_box_case(E=2, N=3, node_type="GAUSS", split=None, riemann="Roe"))          # weak form, odd n
run(*cases.mortar_case("002", N=3, lifting="br2"))                                        # mortars + BR2
run(*cases.mortar_case("001", N=7, node_type="GAUSS-LOBATTO", split="PI", riemann="RoeEntropyFix"))
c, U0 = cases.cavity_case(); run(c, U0)                                                   # wall BCs
print("sanitize cases ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $OUT/sanitize_${tool}_$TAG.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize cases ok|Error|hazard" $OUT/sanitize_${tool}_$TAG.log | sort | uniq -c | head -8
done
