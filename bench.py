#!/usr/bin/env python
"""Headline benchmark: PID / DOF-updates per second of the DG RHS + LSERK stage (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 2|3|4|5] [--scaling weak|strong] [--degree N] [--elems E] [--curved]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       (one rank per GPU, N>1)

Default line (config.workload): BASELINE config #2 -- Taylor-Green vortex, Navier-Stokes, N=7 Gauss-Lobatto, split form
(Pirozzoli) + BR1 lifting, RoeEntropyFix, CarpenterRK4-5, 32^3 elements PER GPU (weak scaling: the global box grows with the
GPU count and is cut along the Hilbert curve like the reference does). A "step" is one full RK time step (5 stages)
including CalcTimeStep and its min-reduction, exactly the window the reference's PID covers (src/output/output.f90:430).

--config selects the other BASELINE configurations (3: TGV N=5 on the 64^3 box, weak = 32^3 per GPU / strong = 64^3 in total;
4: turbulent channel N=5 with isothermal walls and the pressure-gradient forcing; 5: NACA0012 N=4 on the tutorial's curved
mesh). The default invocation also measures config #3 (weak) with a few steps and attaches it as `extras.config3_weak`.

One JSON line on rank 0; see the task contract for the keys. `value` = DOF-updates/s (DOF x RK stages / s, whole job) with the
state resident in HBM; `e2e` = the same metric through the C ABI with HOST buffers (dgx_set_state + dgx_calc_timestep +
dgx_rk_step + dgx_get_state every step, H2D/D2H inside the timed region). With N>1 the line carries `parity`: a small N-rank
run (TGV 8^3 N=7 and a channel with walls) checked against the single-rank CPU oracle outside the timed region.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NSTAGES = 5
FP64_PEAK_TFLOPS = 36.0   # measured DFMA rate of this pool's B200s (profiles/r02_fp64_rates_b200.txt, tools/microbench/fp64_rates.cu)


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def workload_dims(cfg: int, scaling: str, world: int, elems):
    from galaexi_b200.host_standin import workloads as wl
    e = elems if elems is not None else 32
    if cfg == 5:
        return None, e
    if scaling == "strong" and cfg in (3, 4):
        return (2 * e, 2 * e, 2 * e), e
    return wl.box_dims(world, e), e


def workload_desc(cfg: int, scaling: str, world: int, degree=None, elems=None, curved=False, dims=None) -> str:
    """config.workload (both arms print the same string; no mesh is built here)."""
    from galaexi_b200.host_standin import workloads as wl
    d, e = workload_dims(cfg, scaling, world, elems)
    d = dims or d
    if cfg != 5:
        size = f"{d[0]}x{d[1]}x{d[2]} elements" + (" (fixed total)" if (scaling == "strong" and cfg in (3, 4)) else f" ({e}^3 per GPU)")
        if dims is not None:
            size = f"{d[0]}x{d[1]}x{d[2]} elements"
    if cfg in (2, 3):
        N = degree if degree is not None else (7 if cfg == 2 else 5)
        return (f"TGV Navier-Stokes Re1600 Ma0.1, N={N} Gauss-Lobatto, split-form PI + BR1, RoeEntropyFix, CarpenterRK4-5, {size}"
                + (", curved NGeo=2 mesh (sine deformation 0.1)" if curved else "") + f", adaptive dt, CFLscale=DFLscale={wl.TGV_CFL}")
    if cfg == 4:
        N = degree if degree is not None else 5
        return (f"plane turbulent channel Re_tau=180, N={N} Gauss-Lobatto, split-form PI + BR1, RoeEntropyFix, isothermal walls (BC 4) at "
                f"y=+-1, dp/dx forcing with CalcForcing every step, {size}, y-stretched, adaptive dt, CFLscale=DFLscale=0.5")
    N = degree if degree is not None else 4
    return (f"NACA0012 Re5000 AoA8, N={N} Gauss, weak form + BR1, RoeEntropyFix, tutorial mesh NACA0012_652_Ng2 (652 curved elements, "
            f"BC 2 / adiabatic wall 3 / periodic z), fixed total, adaptive dt, CFLscale=DFLscale=0.9")


def workload_dofs(cfg: int, scaling: str, world: int, degree=None, elems=None) -> int:
    """Global number of DOF of the configuration on `world` GPUs (both arms print it in `config`)."""
    d, _ = workload_dims(cfg, scaling, world, elems)
    N = degree if degree is not None else {2: 7, 3: 5, 4: 5, 5: 4}[cfg]
    return (652 if cfg == 5 else d[0] * d[1] * d[2]) * (N + 1) ** 3


def l2_policy(state_bytes_per_gpu: float) -> str:
    """Timing rule: inputs larger than L2, or say that they are not (a property of the workload; both arms print it in `config`)."""
    mb = state_bytes_per_gpu / 1e6
    if state_bytes_per_gpu > 2 * 126e6:
        return f"inputs larger than L2 (state {mb:.0f} MB per GPU vs 126 MB L2), no explicit flush"
    return f"state {mb:.1f} MB per GPU fits L2: launch-bound configuration, no explicit flush"


def make_workload(cfg: int, scaling: str, world: int, rank: int, degree=None, elems=None, curved=False, dims=None):
    """This rank's slice of BASELINE config `cfg` on `world` ranks."""
    from galaexi_b200.host_standin import workloads as wl
    d, _ = workload_dims(cfg, scaling, world, elems)
    d = dims or d
    desc = workload_desc(cfg, scaling, world, degree, elems, curved, dims)
    if cfg in (2, 3):
        N = degree if degree is not None else (7 if cfg == 2 else 5)
        c, U0 = wl.tgv(d, N, nProcs=world, myRank=rank, curved=curved)
    elif cfg == 4:
        N = degree if degree is not None else 5
        c, U0 = wl.channel(d, N, nProcs=world, myRank=rank)
    elif cfg == 5:
        N = degree if degree is not None else 4
        c, U0 = wl.naca(N, nProcs=world, myRank=rank)
    else:
        raise SystemExit(f"unknown --config {cfg}")
    return dict(c=c, U0=U0, N=N, forcing=cfg == 4, desc=desc, dims=d, cfg=cfg)


# algorithmic bytes per DOF, n = N+1, Navier-Stokes
def b_alg_stage(n: int) -> float:
    """SURVEY.md 8d / BASELINE.md 3: whole NS RK stage as the reference's data flow needs it (volume gradients materialised)."""
    return 8.0 * (75.0 + 360.0 / n)


def b_alg_volsurf(n: int) -> float:
    """k_volsurf2: reads U 5, the viscous volume integral k_lifting left in Ut 4, metrics 9, sJ 1, Ut_tmp 5, face fluxes 30/n;
    writes Ut_tmp 5, U 5, next-stage face states 30/n."""
    return 8.0 * (34.0 + 60.0 / n)


def b_alg_lifting(n: int) -> float:
    """k_lifting: reads U 5, metrics 9, sJ 1, face states 10 + geometry 4 per element face node (6/n per DOF); writes the viscous
    volume integral 4 and the gradient traces 12 per element face node."""
    return 8.0 * (19.0 + 156.0 / n)


def b_alg_sideflux(n: int) -> float:
    """k_sideflux per volume DOF (3/n sides per DOF): face states 10, gradient traces 24, geometry 10 in, flux 5 out."""
    return 8.0 * 147.0 / n


def b_alg_stage_fused(n: int) -> float:
    """What the three fused kernels of this library have to move per DOF and stage (sum of the per-kernel figures)."""
    return b_alg_lifting(n) + b_alg_sideflux(n) + b_alg_volsurf(n)


KERNEL_BYTES = {"halo+lifting": b_alg_lifting, "sideflux": b_alg_sideflux, "volsurf_rk": b_alg_volsurf}
KERNEL_OF = {"halo+lifting": "k_lifting", "sideflux": "k_sideflux", "volsurf_rk": "k_volsurf"}


def load_counts():
    """Executed FP64 flops and DRAM bytes per DOF of each kernel from committed ncu captures (profiles/kernel_counts.json, written
    by tools/ncu_counts.py); the live ncu pass of this run (below) replaces them when it succeeds."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_counts.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------------
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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the oracle is the checker / the reported baseline, never the product path)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg=2, degree=None, sample_elems=16, steps=1):
    """The CPU oracle (a port of the reference algorithm, NOT the reference binary) on the host cores, on a bounded sample of
    the same workload: sample_elems^3 elements, `steps` RK steps (the per-DOF cost does not depend on the mesh size)."""
    from oracle.oracle import Oracle
    cores = host_threads()
    wl = make_workload(cfg, "weak", 1, 0, degree=degree, dims=(sample_elems,) * 3 if cfg != 5 else None)
    c, U0 = wl["c"], wl["U0"]
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
                sample=f"{wl['desc']} -- {ndof} DOF, {steps} RK step(s) of 5 stages, OpenMP over elements; restatement of the reference "
                       f"algorithm (oracle/dg_oracle.c), not the reference binary",
                pid_s=wall * cores / (ndof * NSTAGES * steps), wall_s=wall)


def run_reference(args):
    """--impl reference: the reference's CPU path cannot be built here (CUDA Fortran + HDF5 + MPI), so the arm times the oracle
    port with all host threads. Each bench step is one RK step (CalcTimeStep + 5 stages) of ONE GPU's share of the workload --
    at the default that is the whole 32^3-element box of the N=1 product arm (same mesh, same state). The per-DOF cost of the
    CPU path does not depend on how many such boxes the job has, so the value is the CPU's DOF-updates/s for the workload at
    every N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    from oracle.oracle import Oracle
    cores = host_threads()
    wl = make_workload(args.config, "weak" if args.config != 5 else "strong", 1, 0, degree=args.degree, elems=args.elems, curved=args.curved)
    c, U0 = wl["c"], wl["U0"]
    o = Oracle(c)
    o.set_state(U0)
    t = 0.0
    vals = []
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
    sample = (f"one GPU's share of the workload ({wl['desc']}): {ndof} DOF, one RK step (CalcTimeStep + 5 stages) per bench step, "
              f"OpenMP over elements on {cores} host threads")
    line = dict(metric="DOF-updates/s", value=value, unit="DOF*stage/s", impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=workload_desc(args.config, args.scaling, args.gpus, args.degree, args.elems, args.curved), rk_stages=NSTAGES,
                            dof_global=workload_dofs(args.config, args.scaling, args.gpus, args.degree, args.elems),
                            l2_policy=l2_policy(40.0 * workload_dofs(args.config, args.scaling, args.gpus, args.degree, args.elems) / args.gpus)),
                pid_s=wall * cores / (ndof * NSTAGES * args.steps),
                cpu_baseline=dict(value=value, unit="DOF*stage/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="DOF*stage/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="reference binary not buildable offline (nvfortran/HDF5/MPI absent); this arm is the C restatement "
                     "(oracle/dg_oracle.c) on all host threads", wall_total_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------------------------
class Comm:
    """torch.distributed plumbing of one bench process."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def init(self):
        torch = self.torch
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def new_nccl_id(self):
        """A fresh NCCL id for one DGSolver (the library owns its communicator)."""
        from galaexi_b200 import dg
        if self.world == 1:
            return None
        ids = [dg.nccl_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier(self, s=None):
        if s is not None:
            s.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max(self, x):
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def sum(self, x):
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tt)
        return float(tt.item())


def measure(comm: Comm, wl: dict, steps: int, warmup: int, sample_clocks: bool = True, pacing: str = "graph"):
    """W warm-up + K timed RK steps with the state resident in HBM (adaptive dt every step; CalcForcing every step for the
    channel), then the per-kernel stage profile. Returns the solver (kept for the e2e leg) and the numbers."""
    from galaexi_b200 import dg
    from galaexi_b200.host_standin import analyze as an
    c, U0 = wl["c"], wl["U0"]
    t_build = time.perf_counter()
    s = dg.DGSolver(c, device=comm.local, nccl_id=comm.new_nccl_id())
    t_build = time.perf_counter() - t_build
    s.set_state(U0)
    ndof_local = c.nDOF
    ndof_global = int(round(comm.sum(float(ndof_local))))
    forcing = bool(wl["forcing"])
    if forcing:
        from galaexi_b200.host_standin import workloads as wls
        vol = comm.sum(an.volume(c))
        bv = s.CalcForcing(Vol=vol)                       # also leaves the quadrature of the per-step CalcForcing in the library
        s.set_channel_forcing(wls.CHANNEL_DPDX, bv)
    dt0, err = s.CalcTimeStep()
    if err:
        raise SystemExit("initial state not admissible")
    kw = dict(adaptive=True, forcing=forcing, device_paced=pacing != "host", graph=pacing == "graph")
    s.run_steps(warmup, 0.0, dt0, **kw)
    comm.barrier(s)
    sampler = ClockSampler(comm.local)
    if comm.rank == 0 and sample_clocks:
        sampler.start()
        time.sleep(0.3)
    comm.barrier(s)
    l0 = s.launch_count()
    ms, _ = s.run_steps(steps, 0.0, dt0, **kw)   # fails when the state leaves the admissible set
    comm.barrier(s)
    if comm.rank == 0 and sample_clocks:
        time.sleep(0.15)
        sampler.stop()
    ms = comm.max(ms)
    launches = s.launch_count() - l0
    sec = ms * 1e-3
    value = ndof_global * NSTAGES * steps / sec
    pid = sec * comm.world / (ndof_global * steps * NSTAGES)
    prof = {}
    for _ in range(5):
        for k, v in s.profile_stage(0.0, dt0).items():
            prof.setdefault(k, []).append(v)
    prof = {k: comm.max(float(np.mean(v))) for k, v in prof.items()}
    return s, dict(value=value, ms_per_step=ms / steps, pid_s=pid, launches=int(launches), kernel_ms_per_stage=prof, dt0=dt0,
                   ndof_local=ndof_local, ndof_global=ndof_global, setup_s=t_build,
                   pacing=pacing + (" (graph active)" if pacing == "graph" and s.step_graph_active() else (" (graph capture failed: stream launches)" if pacing == "graph" else "")), clocks=sampler.summary() if sample_clocks else None)


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(p.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in p else "fallback 6.65 TB/s (B200_PROFILING.md)"
    return hbm, src


def roofline_of(wl, m, counts_live=None):
    """Per kernel: achieved algorithmic GB/s vs the measured HBM peak, executed FP64 TFLOP/s vs the measured DFMA peak, the
    governing (slower) roof and the fraction of it. Top level = the kernel with the longest launch."""
    n = wl["N"] + 1
    hbm_peak, peak_src = peaks()
    prof, ndof = m["kernel_ms_per_stage"], m["ndof_local"]
    counts = load_counts().get(f"cfg{wl['cfg']}_N{wl['N']}", {})
    if counts_live:
        counts = counts_live
    kernels = {}
    for k, ms in prof.items():
        if k not in KERNEL_BYTES or not ms > 0.0:
            continue
        b = KERNEL_BYTES[k](n)
        gbs = b * ndof / (ms * 1e-3) / 1e9
        ent = dict(ms=ms, algorithmic_bytes_per_dof=b, achieved_gbs=gbs, hbm_frac=gbs / hbm_peak)
        kc = counts.get(KERNEL_OF[k])
        if kc:
            tf = kc["flops_per_dof"] * ndof / (ms * 1e-3) / 1e12
            t_hbm, t_f64 = b / (hbm_peak * 1e9), kc["flops_per_dof"] / (FP64_PEAK_TFLOPS * 1e12)
            ent.update(flops_per_dof=kc["flops_per_dof"], achieved_tflops=tf, fp64_frac=tf / FP64_PEAK_TFLOPS,
                       bound="hbm" if t_hbm >= t_f64 else "fp64", dram_bytes_per_dof=kc.get("dram_bytes_per_dof"))
            ent["frac"] = ent["hbm_frac"] if ent["bound"] == "hbm" else ent["fp64_frac"]
        else:
            ent.update(fp64_frac=None, bound="hbm", frac=ent["hbm_frac"])
        kernels[k] = ent
    top = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    r = dict(bound="hbm", kernel=None, achieved=None, peak=hbm_peak, unit="GB/s", frac=None, traffic=None, peak_source=peak_src)
    if top:
        e = kernels[top]
        traffic = e.get("dram_bytes_per_dof")
        hb = e["bound"] == "hbm"
        r.update(bound=e["bound"], kernel=f"{KERNEL_OF[top]}<{n}> ({top})", achieved=e["achieved_gbs"] if hb else e["achieved_tflops"],
                 peak=hbm_peak if hb else FP64_PEAK_TFLOPS, unit="GB/s" if hb else "TFLOP/s", frac=e["frac"],
                 traffic=None if traffic is None else traffic * ndof, algorithmic_bytes_per_dof=e["algorithmic_bytes_per_dof"],
                 ms_per_launch=e["ms"])
    r.update(kernel_ms_per_stage=prof, kernels=kernels, counts_source=counts.get("source"),
             fp64_peak_tflops=dict(dfma=FP64_PEAK_TFLOPS, dmma_m8n8k4=37.1,
                                   source="profiles/r02_fp64_rates_b200.txt (tools/microbench/fp64_rates.cu on this pool's B200)"),
             stage=dict(algorithmic_bytes_per_dof=b_alg_stage_fused(n), achieved=b_alg_stage_fused(n) / m["pid_s"] / 1e9,
                        frac=b_alg_stage_fused(n) / m["pid_s"] / 1e9 / hbm_peak, reference_dataflow_bytes_per_dof=b_alg_stage(n),
                        note="whole RK stage incl. CalcTimeStep: bytes the three fused kernels must move / PID vs HBM peak"))
    return r


NCU_METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,"
               "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
               "sm__ops_path_tensor_src_fp64.sum")


def parse_ncu_counts(text: str, ndof: int):
    """ncu --csv (one row per launch and metric) -> {kernel: flops_per_dof, dram_bytes_per_dof}, from the LAST launch of each
    kernel family (a mid-step RK stage, not the first-touch one)."""
    rows = list(csv.reader(io.StringIO(text)))
    hdr = None
    per = {}
    for r in rows:
        if "Kernel Name" in r and "Metric Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        nm = d["Kernel Name"]
        key = "k_lifting" if "k_lifting" in nm else ("k_sideflux" if "k_sideflux" in nm else ("k_volsurf" if "k_volsurf" in nm else None))
        if key is None:
            continue
        try:
            val = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(d.get("Metric Unit", ""), 1.0)
        per.setdefault(key, {}).setdefault(int(d["ID"]), {})[d["Metric Name"]] = val * scale
    out = {}
    for key, launches in per.items():
        last = launches[max(launches)]
        flops = (2.0 * last.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
                 + last.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
                 + last.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0) + last.get("sm__ops_path_tensor_src_fp64.sum", 0.0))
        out[key] = dict(flops_per_dof=flops / ndof,
                        dram_bytes_per_dof=(last.get("dram__bytes_read.sum", 0.0) + last.get("dram__bytes_write.sum", 0.0)) / ndof)
    return out


def live_ncu_counts(args, ndof):
    """DRAM bytes and executed FP64 flops of this run's own kernels: one RK step of the same workload under `ncu --metrics ...`
    in a child process (counters only -- no time is taken from it)."""
    cmd = ["ncu", "--metrics", NCU_METRICS, "--clock-control", "none", "--csv", "-k", "regex:k_lifting|k_sideflux|k_volsurf",
           sys.executable, os.path.abspath(__file__), "--ncu-child", "--config", str(args.config), "--scaling", args.scaling]
    for k in ("degree", "elems"):
        if getattr(args, k) is not None:
            cmd += [f"--{k}", str(getattr(args, k))]
    if args.curved:
        cmd.append("--curved")
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=args.ncu_timeout)
        c = parse_ncu_counts(p.stdout, ndof)
        if not c:
            return None, f"ncu pass gave no counters (rc {p.returncode}): {(p.stderr or p.stdout)[-300:]}"
        c["source"] = ("live: ncu --metrics dram__bytes_{read,write}.sum, smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on.sum, "
                       "sm__ops_path_tensor_src_fp64.sum on one RK step of this workload in a child process of this run")
        return c, None
    except Exception as ex:  # ncu missing, no permission for the counters, timeout ...
        return None, f"ncu pass failed: {ex}"


def ncu_child(args):
    """The process profiled by live_ncu_counts: one RK step (5 stages) of the workload."""
    import torch
    torch.cuda.set_device(0)
    from galaexi_b200 import dg
    wl = make_workload(args.config, args.scaling, 1, 0, degree=args.degree, elems=args.elems, curved=args.curved)
    s = dg.DGSolver(wl["c"], device=0)
    s.set_state(wl["U0"])
    s.set_keep_gradients(False)      # like the steps of the timed region: no analysis follows, the volume gradients are not stored
    dt0, _ = s.CalcTimeStep()
    s.TimeStepByLSERKW2(0.0, dt0)
    s.sync()
    s.FinalizeDG()


def multirank_parity(comm: Comm):
    """N ranks vs the single-rank CPU oracle, outside every timed region (the reference's MPI=1 vs MPI=2 criterion,
    regressioncheck/checks/parabolic/cavity_3D/command_line.ini): RHS and two RK steps of (a) TGV 8^3 N=7 split form and
    (b) a 4^3 channel with isothermal walls, N=5, gathered on rank 0 and compared by oracle/parity.py."""
    from galaexi_b200 import dg
    from galaexi_b200.host_standin import workloads as wl
    out = {}

    def tgv_disturbed(P, r):
        # a smooth 5 % disturbance on the low-Mach TGV field, the same function of x on every rank: the bare field's residual is
        # cancellation-dominated (p / (rho u^2) = 71, FP64 round-off floor ~9e-12 > 1e-12) and would test that floor instead of
        # the N-rank path (the bare field is in the single-rank GPU suite under the extended-precision criterion)
        c, U = wl.tgv((8, 8, 8), 7, nProcs=P, myRank=r)
        x = c.geo["Elem_xGP"]
        for v in range(5):
            U[..., v] *= 1.0 + 0.05 * np.sin((2.0 + v) * x[..., 0] + 0.3 * v) * np.cos(3.0 * x[..., 1] - 0.1 * v) * np.sin(2.0 * x[..., 2] + 0.5)
        return c, U

    for label, build in (("tgv_8x8x8_N7", tgv_disturbed),
                         ("channel_4x4x4_N5_walls", lambda P, r: wl.channel((4, 4, 4), 5, nProcs=P, myRank=r))):
        c, U0 = build(comm.world, comm.rank)
        s = dg.DGSolver(c, device=comm.local, nccl_id=comm.new_nccl_id())
        s.set_state(U0)
        s.DGTimeDerivative_weakForm(0.0)
        Ut = s.get_ut()
        dt, err = s.CalcTimeStep()
        t = 0.0
        for _ in range(2):
            s.TimeStepByLSERKW2(t, dt)
            t += dt
        U = s.get_state()
        s.sync()
        s.FinalizeDG()
        parts = [None] * comm.world
        comm.dist.gather_object((c.mesh.offsetElem, Ut, U, dt, err), parts if comm.rank == 0 else None, dst=0)
        if comm.rank == 0:
            from oracle import parity
            host_threads()
            parts.sort(key=lambda x: x[0])
            c1, U01 = build(1, 0)
            r = parity.compare(c1, U01, np.concatenate([p[1] for p in parts]), parts[0][3], np.concatenate([p[2] for p in parts]),
                               nsteps=2, label=label, nranks=comm.world)
            r["dt_equal_on_all_ranks"] = bool(all(p[3] == parts[0][3] and p[4] == 0 for p in parts))
            r["ok"] = bool(r["ok"] and r["dt_equal_on_all_ranks"])
            out[label] = r
        comm.dist.barrier()
    if comm.rank != 0:
        return None
    worst = lambda k: max(v[k] for v in out.values())
    return dict(ok=bool(all(v["ok"] for v in out.values())), ranks=comm.world, ut_rel_l2=worst("ut_rel_l2"), u_rel_l2=worst("u_rel_l2"),
                dt_rel=worst("dt_rel"), cases=out,
                criterion="Ut rel-L2 <= 1e-12 vs the single-rank FP64 oracle (or within 2x the FP64 oracle's own round-off of the 80-bit "
                          "oracle); a deviation above that only up to ut_geometry_roundoff_sensitivity = what 1e-14 relative noise on the "
                          "face normals does to the oracle's own Ut (the side masters, whose element provides the side's metric terms, "
                          "change with the partition; oracle/parity.py); U after 2 RK steps rel-L2 and Linf <= 1e-10, dt rel <= 1e-13, "
                          "identical dt on all ranks")


def e2e_leg(comm: Comm, s, wl, m, steps):
    """End to end through the C ABI with host buffers: H2D state + CalcTimeStep + RK step + D2H state, every step."""
    torch = comm.torch
    U0 = wl["U0"]
    pinned_in = torch.from_numpy(U0).pin_memory()
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    uin, uout = pinned_in.numpy(), pinned_out.numpy()
    steps = max(1, steps)
    s.set_state(uin)
    s.TimeStepByLSERKW2(0.0, m["dt0"])
    s.get_state(uout)
    comm.barrier(s)
    t0 = time.perf_counter()
    for _ in range(steps):
        s.set_state(uin)
        dt_e, _ = s.CalcTimeStep()
        s.TimeStepByLSERKW2(0.0, dt_e)
        s.get_state(uout)
    comm.barrier(s)
    sec = comm.max(time.perf_counter() - t0)
    nbytes = int(U0.nbytes)
    return dict(value=m["ndof_global"] * NSTAGES * steps / sec, unit="DOF*stage/s", h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes + 8,
                steps=steps, ms_per_step=1e3 * sec / steps,
                call="dgx_set_state(U_host) + dgx_calc_timestep + dgx_rk_step + dgx_get_state(U_host) per step, pinned host buffers")


def watchdog(seconds: float):
    """Last resort against a hang in a collective or in the teardown (another rank died, a communicator that does not come
    down): after `seconds` the process leaves without waiting for anybody."""
    def bite():
        time.sleep(seconds)
        sys.stderr.write(f"bench.py: watchdog after {seconds:.0f} s, leaving\n")
        sys.stderr.flush()
        os._exit(2)
    threading.Thread(target=bite, daemon=True).start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=(2, 3, 4, 5), help="BASELINE.json configuration (default 2, the headline)")
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"),
                    help="weak: --elems^3 elements per GPU; strong: fixed total (config 3/4: (2 elems)^3, config 5: the 652-element mesh)")
    ap.add_argument("--degree", type=int, default=None, help="polynomial degree N (default: the configuration's)")
    ap.add_argument("--elems", type=int, default=None, help="elements per direction and GPU (default 32)")
    ap.add_argument("--curved", action="store_true", help="TGV on the NGeo=2 mesh deformed by the reference's meshdeform sine (mesh.f90:224-235)")
    ap.add_argument("--pacing", default="auto", choices=("auto", "host", "device", "graph"),
                    help="dgx_run_steps: host = dt through the host every step (dgx_calc_timestep + dgx_rk_step); device = dt stays on the "
                         "device, CalcTimeStep fused into stage 1; graph = device + CUDA-graph replay of step pairs; bit-identical results. "
                         "auto (default): graph on one GPU, device on several (measured on 4 x B200: the replayed graph overlaps the "
                         "halo-dependent and the inner launches less well than the three eager streams, 18.15 vs 17.44 ms per step)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extras.config3_weak")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 parity leg")
    ap.add_argument("--no-ncu", action="store_true", help="skip the live ncu counter pass (N=1)")
    ap.add_argument("--ncu-timeout", type=int, default=240)
    ap.add_argument("--watchdog", type=float, default=1500.0, help="seconds after which a hung run gives up")
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.config == 5:
        args.scaling = "strong"
    if args.impl == "reference":
        return run_reference(args)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback for the product path)")
    if args.ncu_child:
        return ncu_child(args)
    comm = Comm()
    if comm.world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    watchdog(args.watchdog)
    comm.init()
    world, rank = comm.world, comm.rank
    if args.pacing == "auto":
        args.pacing = "graph" if world == 1 else "device"

    wl = make_workload(args.config, args.scaling, world, rank, degree=args.degree, elems=args.elems, curved=args.curved)
    s, m = measure(comm, wl, args.steps, args.warmup, pacing=args.pacing)
    e2e = e2e_leg(comm, s, wl, m, args.e2e_steps)
    s.FinalizeDG()
    del s

    extras = {}
    default_run = args.config == 2 and args.degree is None and args.elems is None and not args.curved
    if default_run and not args.no_extras:
        # BASELINE config #3, the configuration the 8-GPU target is quoted on (64^3 TGV, N=5): 32^3 elements per GPU
        wl3 = make_workload(3, "weak", world, rank)
        s3, m3 = measure(comm, wl3, 10, 4, sample_clocks=False, pacing=args.pacing)
        s3.FinalizeDG()
        del s3
        if rank == 0:
            r3 = roofline_of(wl3, m3)
            extras["config3_weak"] = dict(workload=wl3["desc"], value=m3["value"], unit="DOF*stage/s", pid_s=m3["pid_s"],
                                          ms_per_step=m3["ms_per_step"], steps=10, warmup=4, dof_global=m3["ndof_global"],
                                          gpu_launches=m3["launches"], stage_frac=r3["stage"]["frac"],
                                          kernels={k: dict(ms=v["ms"], hbm_frac=v["hbm_frac"], fp64_frac=v.get("fp64_frac"), bound=v["bound"])
                                                   for k, v in r3["kernels"].items()})
        del wl3

    parity = None
    if world > 1 and not args.no_parity:
        parity = multirank_parity(comm)

    if rank == 0:
        counts_live, ncu_note = None, None
        if world == 1 and not args.no_ncu:
            counts_live, ncu_note = live_ncu_counts(args, m["ndof_local"])
        roof = roofline_of(wl, m, counts_live)
        if ncu_note:
            roof["ncu_note"] = ncu_note
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            cpu = cpu_baseline(args.config, args.degree)
        n = wl["N"] + 1
        hbm_peak, _ = peaks()
        line = dict(metric="DOF-updates/s", value=m["value"], unit="DOF*stage/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=m["ms_per_step"], higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f64",
                    data="synthetic", pid_s=m["pid_s"], pid_floor_s=b_alg_stage_fused(n) / (hbm_peak * 1e9),
                    config=dict(workload=wl["desc"], rk_stages=NSTAGES, dof_global=m["ndof_global"],
                                l2_policy=l2_policy(40.0 * m["ndof_global"] / world)),
                    measurement=dict(
                        dof_per_gpu=m["ndof_local"],
                        timing="CUDA events on the launching stream around K steps, max over ranks",
                        parallelism=f"dd{world} (SFC element decomposition, NCCL face halos)", step_pacing=m["pacing"],
                        state_check="density > 0, pressure > 0 and a finite time step are checked on the device every step; the run fails otherwise"),
                    roofline=roof, cpu_baseline=cpu, clocks=m["clocks"], e2e=e2e, gpu_launches=m["launches"], setup_s=m["setup_s"])
        if extras:
            line["extras"] = extras
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    watchdog(90.0)     # the line is out: the teardown must not hold the launcher
    if world > 1:
        comm.dist.barrier()
        comm.dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:      # a failing rank must not leave the other ranks (and the launcher) waiting in a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
