#!/usr/bin/env python
"""Headline benchmark: PID / DOF-updates per second of the DG RHS + LSERK stage (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       (one rank per GPU, N>1)

Workload (config.workload): BASELINE config #2 -- Taylor-Green vortex, Navier-Stokes, N=7 Gauss-Lobatto,
split form (Pirozzoli) + BR1 lifting, RoeEntropyFix, CarpenterRK4-5, 32^3 elements PER GPU (weak scaling:
the global box grows with the GPU count and is cut along the Hilbert curve like the reference does).
A "step" is one full RK time step (5 stages) including CalcTimeStep and its min-reduction, exactly the
window the reference's PID covers (src/output/output.f90:430).

One JSON line on rank 0; see the task contract for the keys. `value` = DOF-updates/s (DOF x RK stages / s,
whole job) with the state resident in HBM; `e2e` = the same metric through the C ABI with HOST buffers
(dgx_set_state + dgx_rk_step + dgx_get_state every step, H2D/D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_POLY = 7
CURVED = False
ELEMS_PER_GPU = 32          # 32^3 elements per GPU
NSTAGES = 5


def box_dims(ngpus: int):
    """Global element counts for weak scaling: 32^3 per GPU."""
    e = ELEMS_PER_GPU
    return {1: (e, e, e), 2: (2 * e, e, e), 4: (2 * e, 2 * e, e), 8: (2 * e, 2 * e, 2 * e)}[ngpus]


def b_alg_stage(n: int) -> float:
    """Algorithmic bytes per DOF per NS RK stage (SURVEY.md 8d / BASELINE.md 3)."""
    return 8.0 * (75.0 + 360.0 / n)


def b_alg_volsurf(n: int) -> float:
    """Algorithmic bytes per DOF of the dominant kernel k_volsurf2: reads U 5, the viscous volume integral
    k_lifting left in Ut 4, metrics 9, sJ 1, Ut_tmp 5, face fluxes 30/n; writes Ut_tmp 5, U 5, next-stage face states 30/n."""
    return 8.0 * (34.0 + 60.0 / n)


def b_alg_lifting(n: int) -> float:
    """k_lifting: reads U 5, metrics 9, sJ 1, face states 10 + geometry 4 per element face node (6/n per DOF); writes the viscous
    volume integral 4 and the gradient traces 12 per element face node."""
    return 8.0 * (19.0 + 156.0 / n)


def build_case(ngpus: int, rank: int, elems=None, N=N_POLY):
    from galaexi_b200.host_standin import basis as bs, case as cs, equation as eq, mesh as ms
    import cases
    dims = elems or box_dims(ngpus)
    L = tuple(2 * np.pi * d / min(dims) for d in dims)
    h = ms.make_box_mesh(dims, x0=(0.0, 0.0, 0.0), x1=L, NGeo=2, deform=0.1) if CURVED else ms.make_box_mesh(dims, x0=(0.0, 0.0, 0.0), x1=L, NGeo=1)
    eos = eq.Eos(**cases.TGV_EOS)
    c = cs.build_case(h, N, bs.NODETYPE_GL, split="PI", riemann="RoeEntropyFix", parabolic=True, eos=eos,
                      refstates=cases.TGV_REF, nProcs=ngpus, myRank=rank, CFLScale=0.9, DFLScale=0.9)
    U0 = eq.ini_tgv(c.geo["Elem_xGP"], eos)
    return c, U0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Use every host core this process may run on, also under torchrun (which exports OMP_NUM_THREADS=1): the oracle's
    OpenMP runtime is told directly."""
    import ctypes
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        n = int(os.environ.get("OMP_NUM_THREADS", n))
    return n


def cpu_baseline(sample_elems=12, steps=1, N=N_POLY):
    """The CPU oracle (a port of the reference algorithm, NOT the reference binary) on the host cores, on a bounded
    sample of the same workload: TGV N=7 split-form NS on sample_elems^3 elements, `steps` RK steps."""
    from oracle.oracle import Oracle
    cores = host_threads()
    c, U0 = build_case(1, 0, elems=(sample_elems,) * 3, N=N)
    o = Oracle(c)
    o.set_state(U0)
    dt = o.calc_timestep()[0]
    o.rk_step(0.0, dt)  # warm-up (page faults, OpenMP team)
    t0 = time.perf_counter()
    t = dt
    for _ in range(steps):
        dt = o.calc_timestep()[0]
        o.rk_step(t, dt)
        t += dt
    wall = time.perf_counter() - t0
    ndof = c.nDOF
    o.close()
    return dict(value=ndof * NSTAGES * steps / wall, unit="DOF*stage/s", cores=cores, kind="port",
                sample=f"TGV N={N} GL split-PI NS+BR1, {sample_elems}^3 elements ({ndof} DOF), {steps} RK step(s) of 5 stages, "
                       f"OpenMP over elements; restatement of the reference algorithm (oracle/dg_oracle.c), not the reference binary",
                pid_s=wall * cores / (ndof * NSTAGES * steps), wall_s=wall), c


def run_reference(args):
    """--impl reference: the reference's CPU path cannot be built here (CUDA Fortran + HDF5 + MPI), so the arm times
    the oracle port with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    sample = 12
    vals = []
    from oracle.oracle import Oracle
    cores = host_threads()
    c, U0 = build_case(1, 0, elems=(sample,) * 3)
    o = Oracle(c)
    o.set_state(U0)
    t = 0.0
    for it in range(args.warmup + args.steps):
        s0 = time.perf_counter()
        dt = o.calc_timestep()[0]
        o.rk_step(t, dt)
        t += dt
        if it >= args.warmup:
            vals.append(time.perf_counter() - s0)
    wall = float(np.sum(vals))
    ndof = c.nDOF
    value = ndof * NSTAGES * args.steps / wall
    line = dict(metric="DOF-updates/s", value=value, unit="DOF*stage/s", impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=workload_name(args.gpus), sample=f"{sample}^3 elements per step"),
                pid_s=wall * cores / (ndof * NSTAGES * args.steps),
                cpu_baseline=dict(value=value, unit="DOF*stage/s", cores=cores, kind="port",
                                  sample=f"TGV N=7 GL split-PI NS+BR1 on {sample}^3 elements ({ndof} DOF), one RK step (5 stages) per bench step"),
                e2e=dict(value=value, unit="DOF*stage/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="reference binary not buildable offline (nvfortran/HDF5/MPI absent); this arm is the C restatement "
                     "(oracle/dg_oracle.c) on all host threads", wall_total_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


def workload_name(ngpus):
    d = box_dims(ngpus)
    return (f"TGV Navier-Stokes Re1600 Ma0.1, N={N_POLY} Gauss-Lobatto, split-form PI + BR1, RoeEntropyFix, CarpenterRK4-5, "
            f"{d[0]}x{d[1]}x{d[2]} elements ({ELEMS_PER_GPU}^3 per GPU){', curved NGeo=2 mesh (sine deformation 0.1)' if CURVED else ''}, adaptive dt")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--degree", type=int, default=N_POLY,
                    help="polynomial degree N (default 7 = BASELINE config #2, the headline; 5 = per-GPU load of config #3)")
    ap.add_argument("--curved", action="store_true", help="NGeo=2 mesh deformed by the reference's meshdeform sine (mesh.f90:224-235) instead of the Cartesian one")
    ap.add_argument("--elems", type=int, default=ELEMS_PER_GPU, help="elements per direction and GPU (default 32; 64 = config #3 on one GPU)")
    args = ap.parse_args()
    globals()["N_POLY"] = args.degree
    globals()["ELEMS_PER_GPU"] = args.elems
    globals()["CURVED"] = bool(args.curved)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    from galaexi_b200 import dg
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [dg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]

    t_build = time.perf_counter()
    c, U0 = build_case(world, rank, N=N_POLY)
    s = dg.DGSolver(c, device=local, nccl_id=nccl_id)
    t_build = time.perf_counter() - t_build
    s.set_state(U0)
    ndof_local = c.nDOF
    ndof_global = ndof_local
    if world > 1:
        tt = torch.tensor([ndof_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt)
        ndof_global = int(tt.item())

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident run: W warm-up + K timed steps (adaptive dt every step, as in the reference's PID window)
    dt0, err = s.CalcTimeStep()
    ms_w, _ = s.run_steps(args.warmup, 0.0, dt0, adaptive=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = s.launch_count()
    ms, launches = s.run_steps(args.steps, 0.0, dt0, adaptive=True)
    barrier()
    if rank == 0:
        sampler.stop()
    ms = max_over_ranks(ms)
    launches = s.launch_count() - l0
    sec = ms * 1e-3
    value = ndof_global * NSTAGES * args.steps / sec
    pid = sec * world / (ndof_global * args.steps * NSTAGES)

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline of the dominant kernel
    prof = {}
    for _ in range(5):
        for k, v in s.profile_stage(0.0, dt0).items():
            prof.setdefault(k, []).append(v)
    prof = {k: float(np.mean(v)) for k, v in prof.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    n = N_POLY + 1
    vs_ms = prof.get("volsurf_rk", float("nan"))
    ach = b_alg_volsurf(n) * ndof_local / (vs_ms * 1e-3) / 1e9
    # every stage kernel against the HBM roof (algorithmic bytes per DOF from DESIGN.md 4)
    b_k = {"halo+lifting": b_alg_lifting(n), "sideflux": 8.0 * 147.0 / n, "volsurf_rk": b_alg_volsurf(n)}
    per_kernel = {k: dict(ms=prof[k], algorithmic_bytes_per_dof=b_k[k], achieved_gbs=b_k[k] * ndof_local / (prof[k] * 1e-3) / 1e9,
                          frac=b_k[k] * ndof_local / (prof[k] * 1e-3) / 1e9 / hbm_peak) for k in b_k if k in prof}
    roofline = dict(bound="hbm", kernel=f"k_volsurf2<{n},RK>", achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak,
                    traffic=None, peak_source=peak_src, algorithmic_bytes_per_dof=b_alg_volsurf(n), ms_per_launch=vs_ms,
                    kernel_ms_per_stage=prof, kernels=per_kernel,
                    fp64_peak_tflops=dict(dfma=36.0, dmma_m8n8k4=37.1, source="profiles/r02_fp64_rates_b200.txt (tools/microbench/fp64_rates.cu on this pool's B200)"),
                    stage=dict(algorithmic_bytes_per_dof=b_alg_stage(n), achieved=b_alg_stage(n) / pid / 1e9,
                               frac=b_alg_stage(n) / pid / 1e9 / hbm_peak, note="whole RK stage incl. CalcTimeStep: B_alg,NS(N)/PID vs HBM peak"))
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if N_POLY == 7 and world == 1:   # the ncu capture is of the headline workload on one GPU
            roofline["traffic"] = tr.get("k_volsurf_bytes_per_launch")
            roofline["traffic_source"] = tr.get("source")
    except Exception:
        pass

    # ---- end to end through the C ABI with host buffers: H2D state + RK step + D2H state, every step
    pinned_in = torch.from_numpy(U0).pin_memory()
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    uin = pinned_in.numpy()
    uout = pinned_out.numpy()
    e2e_steps = max(1, args.e2e_steps)
    s.set_state(uin)
    s.TimeStepByLSERKW2(0.0, dt0)
    s.get_state(uout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.set_state(uin)
        dt_e, _ = s.CalcTimeStep()
        s.TimeStepByLSERKW2(0.0, dt_e)
        s.get_state(uout)
    barrier()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_val = ndof_global * NSTAGES * e2e_steps / e2e_sec
    nbytes = int(U0.nbytes)

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            cpu, _ = cpu_baseline()
        line = dict(metric="DOF-updates/s", value=value, unit="DOF*stage/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", pid_s=pid, pid_floor_s=b_alg_stage(n) / (hbm_peak * 1e9),
                    config=dict(workload=workload_name(world), dof_global=ndof_global, dof_per_gpu=ndof_local, rk_stages=NSTAGES,
                                l2_policy=f"inputs larger than L2 (state {U0.nbytes / 1e6:.0f} MB per GPU >> 126 MB L2), no explicit flush",
                                timing="CUDA events on the launching stream around K steps, max over ranks",
                                parallelism=f"dd{world} (SFC element decomposition, NCCL face halos)"),
                    roofline=roofline, cpu_baseline=cpu, clocks=sampler.summary(),
                    e2e=dict(value=e2e_val, unit="DOF*stage/s", h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes + 8,
                             steps=e2e_steps, ms_per_step=1e3 * e2e_sec / e2e_steps,
                             call="dgx_set_state(U_host) + dgx_calc_timestep + dgx_rk_step + dgx_get_state(U_host) per step, pinned host buffers"),
                    gpu_launches=int(launches), setup_s=t_build)
        print(json.dumps(line), flush=True)
    s.FinalizeDG()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
