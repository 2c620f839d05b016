"""N>1 on real GPUs: SFC partition + NCCL face halos reproduce the single-rank result (skipped on 1-GPU boxes)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case", ["tgv", "cavity", "channel", "shu", "tgv_br2", "mortar001", "mortar004_br2", "tgv_filter", "manufactured", "tgv_oint"])
def test_two_ranks_match_single_rank(case):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mr_check.py"), case]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MRCHECK ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    r = json.loads(lines[-1][8:])
    assert r["ut_ok"] and r["paced_bitwise"], r
    assert r["u_rel_l2"] <= 1e-10 and r["dt_rel"] <= 1e-13 and r["diag_rel"] <= 1e-9 and r["bulk_rel"] <= 1e-10


@pytest.mark.parametrize("case,key,tol", [("cavity_regression", "max_abs_vs_reference_state", 1e-12), ("tgv_csv", "worst_column_deviation", 1e-8)])
def test_reference_regressions_on_two_ranks(case, key, tol):
    """parabolic/cavity_3D (MPI=1,2 in the reference) and the first 41 rows of tgv/split (MPI=6) on two ranks."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29800 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mr_check.py"), case]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MRCHECK ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    assert json.loads(lines[-1][8:])[key] <= tol


def test_naca_regression_on_two_ranks():
    """regressioncheck/checks/naca/3D is run with MPI=6 by the reference: the whole run to t=10 on two ranks (NCCL halos, sponge,
    Pruett base flow) against the reference's state file, abs 5e-11."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mr_check.py"), "naca_regression"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MRCHECK ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    r = json.loads(lines[-1][8:])
    assert r["steps"] > 20000 and r["max_abs_vs_reference_state"] <= 5e-11
