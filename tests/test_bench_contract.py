"""bench.py contract on a machine without a GPU: the reference arm (the CPU restatement of the reference algorithm on the host
cores) prints one JSON line with the contract's keys, and the product arm refuses to run without a B200 (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "DOF-updates/s" and line["unit"] == "DOF*stage/s" and line["dtype"] == "f64"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["config"]["dof_global"] == 32 ** 3 * 512 and line["config"]["rk_stages"] == 5     # the same keys and values the product arm prints


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_product_arm_has_no_cpu_fallback():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CPU fallback" in out.stderr
