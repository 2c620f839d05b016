#!/usr/bin/env python
"""Where does the Ut deviation of the N=7 split-form path sit? Single rank, the disturbed TGV of bench.py's parity leg."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from galaexi_b200 import dg
from galaexi_b200.host_standin import workloads as wl
from oracle.oracle import Oracle

bench.host_threads()
for N, E in ((7, 8), (5, 8), (7, 4)):
    c, U = wl.tgv((E, E, E), N)
    x = c.geo["Elem_xGP"]
    for v in range(5):
        U[..., v] *= 1.0 + 0.05 * np.sin((2.0 + v) * x[..., 0] + 0.3 * v) * np.cos(3.0 * x[..., 1] - 0.1 * v) * np.sin(2.0 * x[..., 2] + 0.5)
    s = dg.DGSolver(c)
    s.set_state(U)
    s.DGTimeDerivative_weakForm(0.0)
    Ut = s.get_ut()
    s.FinalizeDG()
    o = Oracle(c); o.set_state(U); ref = o.time_derivative(0.0).copy(); o.close()
    xo = Oracle(c, "extended"); xo.set_state(U); ex = np.asarray(xo.time_derivative(0.0), dtype=np.float64); xo.close()
    rl2 = lambda a, b: float(np.sqrt(np.sum((a - b) ** 2) / np.sum(b ** 2)))
    print(f"N={N} E={E}: GPU vs FP64 oracle {rl2(Ut, ref):.3e}, GPU vs exact {rl2(Ut, ex):.3e}, FP64 oracle vs exact {rl2(ref, ex):.3e}")
    for v in range(5):
        print(f"   var {v}: GPU-exact {rl2(Ut[..., v], ex[..., v]):.3e}  oracle-exact {rl2(ref[..., v], ex[..., v]):.3e}  scale {np.abs(ex[..., v]).max():.3e}")
    err = np.abs(Ut - ex)[..., 4]
    n = N + 1
    idx = np.indices(err.shape)
    onface = np.zeros(err.shape, bool)
    for ax in (1, 2, 3):
        onface |= (idx[ax] == 0) | (idx[ax] == n - 1)
    print(f"   energy: rms err interior nodes {np.sqrt(np.mean(err[~onface] ** 2)):.3e}, face nodes {np.sqrt(np.mean(err[onface] ** 2)):.3e}, max {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}")
