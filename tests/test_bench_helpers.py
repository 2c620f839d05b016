"""bench.py / oracle.parity helpers that run without a GPU: both arms describe the same workload, the ncu counter parser, the
N-rank parity verdict and the measured metric round-off sensitivity behind it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_both_arms_print_the_same_workload_string():
    for cfg, scaling in ((2, "weak"), (3, "weak"), (3, "strong"), (4, "weak"), (5, "strong")):
        for world in (1, 2, 4, 8):
            a = bench.workload_desc(cfg, scaling, world)
            assert a == bench.workload_desc(cfg, scaling, world) and ("elements" in a)
    # weak: the box grows with the GPU count; strong: it does not
    assert "64x64x64" in bench.workload_desc(3, "weak", 8) and "64x64x64" in bench.workload_desc(3, "strong", 1)
    assert "32x32x32" in bench.workload_desc(2, "weak", 1) and "64x32x32" in bench.workload_desc(2, "weak", 2)
    assert bench.workload_dofs(2, "weak", 1) == 32 ** 3 * 512 and bench.workload_dofs(3, "weak", 8) == 64 ** 3 * 216 == bench.workload_dofs(3, "strong", 2)
    assert bench.workload_dofs(5, "strong", 4) == 652 * 125
    assert "CFLscale=DFLscale=0.8" in bench.workload_desc(2, "weak", 1)      # tgv/split/parameter.ini:62-63 (0.9 is unstable at N=7 GL)


def test_algorithmic_bytes_add_up():
    for n in (4, 6, 8):
        assert abs(bench.b_alg_stage_fused(n) - (bench.b_alg_lifting(n) + bench.b_alg_sideflux(n) + bench.b_alg_volsurf(n))) < 1e-12
        assert bench.b_alg_stage_fused(n) < bench.b_alg_stage(n)             # the volume gradients do not round-trip HBM
    assert bench.b_alg_stage_fused(8) == 787.0 and bench.b_alg_stage(8) == 960.0


def test_ncu_counter_parser_takes_the_last_launch_of_each_kernel():
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    def row(i, k, m, u, v):
        return f'"{i}","1","python","h","{k}","1","7","(128, 1, 1)","(10, 1, 1)","0","10.0","Command line profiler metrics","{m}","{u}","{v}"'
    k1, k2 = "void dgx::k_lifting<8, 2, 0>(dgx::KParams, int)", "void dgx::k_volsurf2<8, 1, 4>(dgx::KParams, int, double, double, int)"
    lines = [hdr]
    for i, k, scale in ((0, k1, 1.0), (1, k2, 1.0), (2, k1, 2.0)):
        lines += [row(i, k, "dram__bytes_read.sum", "Gbyte", f"{1.0 * scale}"), row(i, k, "dram__bytes_write.sum", "Mbyte", "500"),
                  row(i, k, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "inst", "1,000"),
                  row(i, k, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "inst", "10"),
                  row(i, k, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "inst", "20"),
                  row(i, k, "sm__ops_path_tensor_src_fp64.sum", "", "512")]
    c = bench.parse_ncu_counts("\n".join(["==PROF== noise"] + lines), ndof=1000)
    assert set(c) == {"k_lifting", "k_volsurf"}
    assert abs(c["k_lifting"]["dram_bytes_per_dof"] - (2.0e9 + 5.0e8) / 1000) < 1e-6        # launch ID 2, not 0
    assert abs(c["k_volsurf"]["flops_per_dof"] - (2 * 1000 + 10 + 20 + 512) / 1000) < 1e-12


def test_n_rank_parity_verdict_and_metric_roundoff_sensitivity():
    """oracle/parity.compare on the oracle's own results: exact agreement passes; a deviation above 1e-12 passes on N ranks only
    if it stays inside what 1e-14 of metric round-off does to the oracle's own Ut, and never on one rank."""
    from galaexi_b200.host_standin import workloads as wl
    from oracle import parity
    c, U0 = wl.channel((4, 4, 4), 5)
    Ut, dt, U = parity.oracle_rhs_and_steps(c, U0, 1)
    sens = parity.geometry_roundoff_sensitivity(c, U0, Ut)
    assert 2e-12 < sens < 1e-10                          # the surface operator amplifies 1e-14 on the normals (N=5: L_hat = 15)
    assert np.array_equal(c.geo["NormVec"], wl.channel((4, 4, 4), 5)[0].geo["NormVec"])     # the case is left untouched
    ok = parity.compare(c, U0, Ut, dt, U, nsteps=1, label="self", nranks=2)
    assert ok["ok"] and ok["ut_rel_l2"] == 0.0 and ok["ut_geometry_roundoff_sensitivity"] == sens
    rng = np.random.default_rng(0)
    noise = rng.standard_normal(Ut.shape) * np.sqrt(np.mean(Ut ** 2))
    small = Ut + 0.6 * sens * noise                      # above 1e-12, inside the sensitivity
    big = Ut + 30.0 * sens * noise
    assert parity.rel_l2(small, Ut) > parity.TOL_UT
    assert parity.compare(c, U0, small, dt, U, nsteps=1, label="small@2", nranks=2)["ok"]
    assert not parity.compare(c, U0, small, dt, U, nsteps=1, label="small@1", nranks=1)["ok"]
    assert not parity.compare(c, U0, big, dt, U, nsteps=1, label="big@2", nranks=2)["ok"]
    assert not parity.compare(c, U0, Ut, dt * (1 + 1e-10), U, nsteps=1, label="dt", nranks=2)["ok"]
