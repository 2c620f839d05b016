import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "unit_goldens.npz"))


def pytest_terminal_summary(terminalreporter):
    """Which residual comparisons met 1e-12 against the FP64 oracle directly and which needed the extended-precision criterion."""
    try:
        from oracle import parity
    except Exception:
        return
    if not parity.UT_LOG:
        return
    import json
    ext = [r for r in parity.UT_LOG if r["used_extended"]]
    terminalreporter.write_line(f"Ut parity: {len(parity.UT_LOG)} comparisons, {len(ext)} used the extended-precision criterion, "
                                f"worst rel-L2 among those that met 1e-12 directly {max([r['err_fp64'] for r in parity.UT_LOG if not r['used_extended']] or [0.0]):.3e}")
    for r in ext:
        terminalreporter.write_line(f"  extended: {r['case']}: vs FP64 oracle {r['err_fp64']:.3e}, FP64 round-off floor {r['floor']:.3e}, vs exact {r['err_exact']:.3e}")
    try:
        import torch
        if not torch.cuda.is_available():      # the log is evidence of the GPU suite; CPU runs only exercise the checker itself
            return
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ut_parity_log.json"), "w") as f:
            json.dump(parity.UT_LOG, f, indent=1)
    except OSError:
        pass
