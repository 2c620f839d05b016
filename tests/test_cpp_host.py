"""The compiled-language host above the C ABI (include/dgx.hpp + tests/cpp/host_driver.cpp): the reference's call sequence
(SetState, DGTimeDerivative_weakForm, CalcTimeStep, TimeDisc) driven from C++ on a case dumped to a flat file, compared with
the CPU oracle. CPU part: the driver compiles against include/ and links against libdgx.so (no compute call)."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "galaexi_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "host_driver")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_driver.cpp")
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "dgx.hpp")),
                                                                os.path.getmtime(os.path.join(ROOT, "include", "dgx.h"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE, "-L", CSRC, "-ldgx",
                               f"-Wl,-rpath,{CSRC}"])
    return EXE


def _dump(path, c, U0, Ut_ref, U_ref, tEnd, dt0, nsteps):
    m, b, g, td = c.mesh, c.basis, c.geo, c.timedisc
    rec = []

    def add(name, arr, kind):
        a = np.ascontiguousarray(arr, dtype=np.float64 if kind else np.int32)
        rec.append(struct.pack("<i", len(name)) + name.encode() + struct.pack("<iq", kind, a.size) + a.tobytes())
    si = [c.N, 2 if c.node_type == "GAUSS-LOBATTO" else 1, c.split, c.riemann, int(c.parabolic), c.eos.visc_law, m.nElems, m.nSides, m.nBCSides,
          m.firstInnerSide, m.lastInnerSide, m.firstMPISide_MINE, m.lastMPISide_MINE, m.firstMPISide_YOUR, m.lastMPISide_YOUR,
          c.RefStatePrim.shape[0], td.nRKStages, c.lifting, m.nMortarSides, m.firstMortarInnerSide, m.lastMortarInnerSide,
          m.firstMortarMPISide, m.lastMortarMPISide, nsteps]
    sd = list(c.eos.eos_vars()) + [td.CFLScale, td.DFLScale, c.etaBR2, c.etaBR2_wall, tEnd, dt0]
    add("scalars_int", si, 0)
    add("scalars_real", sd, 1)
    add("RefStatePrim", c.RefStatePrim, 1)
    add("BCSides", c.BCSides if c.BCSides.size else np.zeros((1, 2)), 0)
    for nm, a in (("D_T", b.D_T.T), ("D_Hat_T", b.D_Hat_T.T), ("DVolSurf", b.DVolSurf.T), ("L_Minus", b.L_Minus), ("L_Plus", b.L_Plus),
                  ("L_HatMinus", b.L_HatMinus), ("L_HatPlus", b.L_HatPlus), ("RKA", td.RKA), ("RKb", td.RKb), ("RKc", td.RKc)):
        add(nm, a, 1)
    add("ElemToSide", m.ElemToSide, 0)
    add("S2V2", c.maps["S2V2"], 0)
    add("S2V2_inv", c.maps["S2V2_inv"], 0)
    for nm in ("Metrics_fTilde", "Metrics_gTilde", "Metrics_hTilde", "sJ", "NormVec", "TangVec1", "TangVec2", "SurfElem"):
        add(nm, g[nm], 1)
    add("MortarType", m.MortarType, 0)
    add("MortarInfo", m.MortarInfo, 0)
    for nm in ("M_0_1", "M_0_2", "M_1_0", "M_2_0"):
        add(nm, c.mortar[nm].T, 1)
    add("U0", U0, 1)
    add("Ut_ref", Ut_ref, 1)
    add("U_ref", U_ref, 1)
    with open(path, "wb") as f:
        for r in rec:
            f.write(r)


def test_cpp_host_compiles_and_links():
    exe = _build()
    out = subprocess.run([exe, "/dev/null", "--link-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    from galaexi_b200 import dg
    import ctypes as C
    assert r["linked"] and r["sizeof_config"] == C.sizeof(dg.DgxConfig)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tgv_curved", "mortar_br2"])
def test_cpp_host_runs_the_reference_call_sequence(name, tmp_path):
    from galaexi_b200.host_standin import timeloop
    from oracle.oracle import Oracle
    if name == "tgv_curved":
        c, U0 = cases.tgv_box_case(E=3, N=4, NGeo=2, deform=0.05, perturb=1e-3)
    else:
        c, U0 = cases.mortar_case("002", N=3, lifting="br2")
    o = Oracle(c)
    o.set_state(U0)
    Ut_ref = o.time_derivative(0.0).copy()
    dt0 = o.calc_timestep()[0]
    tEnd = 3.4 * dt0

    class _Op:
        def calc_timestep(self):
            return o.calc_timestep()

        def rk_step(self, t, dt):
            o.rk_step(t, dt)
    _, nsteps = timeloop.advance(_Op(), 0.0, tEnd)
    path = str(tmp_path / "case.bin")
    _dump(path, c, U0, Ut_ref, o.array("U"), tEnd, dt0, nsteps)
    o.close()
    out = subprocess.run([_build(), path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["ut_rel_l2"] <= 1e-12 and r["u_rel_l2"] <= 1e-10, r
    assert abs(r["dt0"] - r["dt0_ref"]) <= 1e-13 * r["dt0_ref"] and r["errType"] == 0
    assert r["nsteps"] == r["nsteps_ref"] == 4 and r["abort_ok"] and r["launches"] > 0
