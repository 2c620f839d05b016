// Host side of the C ABI (include/dgx.h): device memory ownership, layout conversion at the boundary,
// the RHS / Runge-Kutta orchestration on CUDA streams and the NCCL face-halo exchange.
//
// Orchestration replaces DGTimeDerivative_weakForm (dg/dg.f90:255-425) + TimeStepByLSERKW2
// (timedisc/timestep.f90:49-120) of the reference. Differences by design:
//   * 3 fused kernels per stage (k_lifting, k_sideflux, k_volsurf) instead of ~20, no device-wide syncs;
//   * face states are double buffered: the RK epilogue of stage s writes the faces stage s+1 reads;
//   * 2 halo phases per Navier-Stokes stage instead of 4: both sides of an MPI face exchange their face
//     state / face gradients, then lifting flux and numerical flux are evaluated redundantly (bit-identical)
//     on both ranks, which removes the lifting-flux and flux messages (mpi/mpi.f90:277-387, SURVEY 2.3 rows 2,4).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3 (no link dependency); ranges like globals/nvtx.f90, switched on with DGX_NVTX=1
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dgx.h"
#include "dgx_launch.h"

using namespace dgx;

#ifndef DGX_DEFAULT_FLAGS
#define DGX_DEFAULT_FLAGS 0  // see KParams::flags; chosen from the sweep in profiles/
#endif

namespace dgx {
const KernelTable* kernel_table(int N, int nodeType) {
    switch (N) {
#define DGX_CASE(NN) case NN: return kernel_table_N##NN(nodeType);
#ifdef DGX_HAVE_N1
        DGX_CASE(1)
#endif
#ifdef DGX_HAVE_N2
        DGX_CASE(2)
#endif
#ifdef DGX_HAVE_N3
        DGX_CASE(3)
#endif
#ifdef DGX_HAVE_N4
        DGX_CASE(4)
#endif
#ifdef DGX_HAVE_N5
        DGX_CASE(5)
#endif
#ifdef DGX_HAVE_N6
        DGX_CASE(6)
#endif
#ifdef DGX_HAVE_N7
        DGX_CASE(7)
#endif
#ifdef DGX_HAVE_N8
        DGX_CASE(8)
#endif
#ifdef DGX_HAVE_N9
        DGX_CASE(9)
#endif
        default: return nullptr;
    }
}
}  // namespace dgx

// ------------------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU use must not depend on libnccl being present
// ------------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclMin = 3 };
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define SYM(f) *(void**)(&f) = dlsym(lib, "nccl" #f); if (!f) { err = "missing symbol nccl" #f; return false; }
        SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
        return true;
    }
};
Nccl g_nccl;
}  // namespace

// ------------------------------------------------------------------------------------------------------
struct HaloMsg { int peer, isSend, slaveArray, side0, nSides; };
struct dgx_handle {
    dgx_config cfg;  // scalar copies only; pointers are not retained
    int n, n2, n3;
    const KernelTable* kt = nullptr;
    KParams P;
    std::string err;
    long long launches = 0;
    cudaStream_t s = nullptr, cs = nullptr, s2 = nullptr;  // compute, communication, halo-dependent compute (multi rank)
    cudaEvent_t evFaces = nullptr, evUhalo = nullptr, evGrad = nullptr, evGhalo = nullptr, evT0 = nullptr, evT1 = nullptr;
    cudaEvent_t evSide = nullptr, evBnd = nullptr, evDtPre = nullptr, evDt = nullptr, evNext = nullptr;
    // U-face halo of the NEXT stage: posted right behind the halo-dependent volume kernel of an RK stage so that it travels
    // while the inner elements are still being updated (dg.f90:335-350 starts it at the top of the next RHS instead)
    bool uHaloPosted = false, uHaloJoined = false;
    bool anyMortar = false;  // some rank of the job has non-conforming sides: every rank keeps the plain call order (see rhs)
    // device-paced stepping (dgx_run_steps flag 4): dt / bulk velocity stay on the device, one RK step pair is replayed as a CUDA graph
    double* dtHist = nullptr;
    int dtHistCap = 0, dtHistCount = 0;
    double bvVol = 0.0;  // Vol of the last dgx_calc_bulk_velocity (CalcForcing inside dgx_run_steps)
    cudaGraphExec_t graph = nullptr;
    int graphCur = -1, graphKey = -1;
    long long graphLaunches = 0;
    int graphFailed = 0;
    bool graphEndsPosted = false;      // the captured pair leaves the next U-face halo posted and joined into stream s
    std::vector<double> dtHistHost;    // dt of the steps of a host-paced dgx_run_steps
    bool histOnHost = false;
    // events used while capturing: an event recorded inside a stream capture cannot be waited on by eager work afterwards, so the
    // capture runs on its own set (same order as the eager ones in swap_event_sets)
    cudaEvent_t evCap[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool warm = false, capturing = false;  // warm: an RK stage has run outside any capture (first-use initialisation done)
    // device memory
    std::vector<void*> allocs;
    double *U = nullptr, *Ut = nullptr, *Ut_tmp = nullptr, *gradU = nullptr, *metrics = nullptr, *sJ = nullptr, *geo = nullptr;
    double *Uf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [buffer][master/slave]
    double *gm = nullptr, *gs = nullptr, *Flux = nullptr;
    double *stage = nullptr;  // staging for layout conversion, 5*nDOF doubles
    double *dtOut = nullptr;  // [3] device
    double *hPinned = nullptr;  // [4] pinned host scalars
    int *errFlag = nullptr;
    int cur = 0;
    // volume gradients: only analysis reads them (the viscous volume integral is formed inside k_lifting). keepGrad: the last
    // stage of every dgx_rk_step stores them, like the reference, whose d_gradUx/y/z then hold the last stage's gradients
    // (testcase.f90:361-364 reads them at analyze steps without a new RHS evaluation); gradValid: the last RHS stored them
    int keepGrad = 1;
    bool gradValid = false;
    bool bcChecked = false;
    // element / side lists for overlap
    int *innerList = nullptr, *bndList = nullptr;
    int nInner = 0, nBnd = 0;
    // RK
    std::vector<double> RKA, RKb, RKc;
    std::vector<double> RKdelta, RKg1, RKg2, RKg3;  // three-register schemes (empty: Williamson 2N)
    int rk3Stage = 0;                                // stage of the 3-register update rhs() applies (0: none)
    // MPI-like neighbour tables
    std::vector<int> NbProc, nMine, nYour, offMine, offYour;
    std::vector<HaloMsg> plan;
    ncclComm_t comm = nullptr;
    double *bvPart = nullptr, *bvW = nullptr;  // dgx_calc_bulk_velocity
    double *bfPart = nullptr, *bfW = nullptr;  // dgx_calc_body_forces
    int* bfBC = nullptr;
    int bfNBCs = 0;
    // TGV diagnostics (dgx_analyze_tgv)
    double *tgvV = nullptr, *tgvW = nullptr, *tgvPart = nullptr;
    int tgvNA1 = 0;
    // non-conforming interfaces: the two ranges of big mortar sides
    MortarParams mp;
    int nMortarInner = 0, nMortarMPI = 0;
    bool hasMortar() const { return nMortarInner + nMortarMPI > 0; }
    size_t nDOF() const { return (size_t)cfg.nElems * n3; }
    size_t nFace() const { return (size_t)cfg.nSides * n2; }
};

namespace {

int fail(dgx_handle* h, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return 1;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define NK(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != 0) return fail(h, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

template <class T>
int dalloc(dgx_handle* h, T** p, size_t count) {
    void* q = nullptr;
    CK(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
    CK(cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), h->s));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
template <class T>
int upload(dgx_handle* h, T** p, const T* host, size_t count) {
    if (dalloc(h, p, count)) return 1;
    if (count) CK(cudaMemcpyAsync(*p, host, count * sizeof(T), cudaMemcpyHostToDevice, h->s));
    return 0;
}

// NVTX range per phase of the RHS / time step (the reference brackets the same phases, globals/nvtx.f90 + dg.f90), for nsys timelines
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) : on(enabled()) { if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
    static bool enabled() { static const bool e = getenv("DGX_NVTX") != nullptr; return e; }
};

inline unsigned blocks_for(size_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

int check_launch(dgx_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, "kernel launch %s failed: %s", what, cudaGetErrorString(e));
    h->launches++;
    return 0;
}

// ---- halo exchange: both directions for every MPI side range -----------------------------------------
// The message plan is a pure function of the neighbour tables so that it can be executed (and tested) with any
// point-to-point layer that, like NCCL, matches the messages between one pair of ranks in issue order.
void halo_plan(const std::vector<int>& NbProc, const std::vector<int>& nMine, const std::vector<int>& nYour,
               const std::vector<int>& offMine, const std::vector<int>& offYour, std::vector<HaloMsg>& plan) {
    plan.clear();
    for (size_t ib = 0; ib < NbProc.size(); ib++) {
        const int peer = NbProc[ib];
        // Both ranks send their MINE range first and their YOUR range second, so the first message arriving from
        // the peer is ITS MINE range (= my YOUR range, master data), the second its YOUR range (= my MINE range).
        if (nMine[ib]) plan.push_back({peer, 1, 0, offMine[ib], nMine[ib]});  // I am master: my master data
        if (nYour[ib]) plan.push_back({peer, 1, 1, offYour[ib], nYour[ib]});  // I am slave: my slave data
        if (nYour[ib]) plan.push_back({peer, 0, 0, offYour[ib], nYour[ib]});  // the neighbour's master data
        if (nMine[ib]) plan.push_back({peer, 0, 1, offMine[ib], nMine[ib]});  // the neighbour's slave data
    }
}

int exchange(dgx_handle* h, double* am, double* as, int nvar) {
    const size_t per = (size_t)nvar * h->n2;
    NK(g_nccl.GroupStart());
    for (const HaloMsg& m : h->plan) {
        double* p = (m.slaveArray ? as : am) + (size_t)m.side0 * per;
        if (m.isSend) NK(g_nccl.Send(p, (size_t)m.nSides * per, ncclFloat64, m.peer, h->comm, h->cs));
        else NK(g_nccl.Recv(p, (size_t)m.nSides * per, ncclFloat64, m.peer, h->comm, h->cs));
    }
    NK(g_nccl.GroupEnd());
    return 0;
}

// ---- non-conforming interfaces: the same operation on the inner and the MPI range of big mortar sides --------
int mortar_u(dgx_handle* h, double* am, double* as, int nvar) {
    MortarParams mp = h->mp;
    mp.side0 = h->cfg.firstMortarMPISide - 1;
    h->kt->umortar(am, as, nvar, mp, h->nMortarMPI, h->s);
    if (h->nMortarMPI && check_launch(h, "k_umortar(mpi)")) return 1;
    mp.side0 = h->cfg.firstMortarInnerSide - 1;
    h->kt->umortar(am, as, nvar, mp, h->nMortarInner, h->s);
    if (h->nMortarInner && check_launch(h, "k_umortar")) return 1;
    return 0;
}
int mortar_flux(dgx_handle* h, double* F, int nvar, int weak) {
    MortarParams mp = h->mp;
    mp.side0 = h->cfg.firstMortarInnerSide - 1;
    h->kt->fluxmortar(F, nvar, weak, mp, h->nMortarInner, h->s);
    if (h->nMortarInner && check_launch(h, "k_fluxmortar")) return 1;
    mp.side0 = h->cfg.firstMortarMPISide - 1;
    h->kt->fluxmortar(F, nvar, weak, mp, h->nMortarMPI, h->s);
    if (h->nMortarMPI && check_launch(h, "k_fluxmortar(mpi)")) return 1;
    return 0;
}
int mortar_liftflux(dgx_handle* h, const KParams& P) {
    MortarParams mp = h->mp;
    mp.side0 = h->cfg.firstMortarInnerSide - 1;
    h->kt->mortar_liftflux(P, mp, h->nMortarInner, h->s);
    if (h->nMortarInner && check_launch(h, "k_mortar_liftflux")) return 1;
    mp.side0 = h->cfg.firstMortarMPISide - 1;
    h->kt->mortar_liftflux(P, mp, h->nMortarMPI, h->s);
    if (h->nMortarMPI && check_launch(h, "k_mortar_liftflux(mpi)")) return 1;
    return 0;
}

// ---- one RHS evaluation (mode 0: store Ut) or RK stage (mode 1) ---------------------------------------
// Mortar meshes (host FLEXI order, dg/lifting/lifting_br2.t90:94-181 and the U_Mortar / Flux_Mortar call pattern):
// the face states this stage reads already hold the small-side data (k_umortar runs right after every face
// extraction); the lifting flux of big sides is projected before k_lifting, the gradient traces are interpolated to the
// small sides after it, the numerical flux of the small sides is projected before k_volsurf.
struct StageTimes {
    cudaEvent_t ev[8];
    int nev = 0;
};
// per-stage options of rhs(): storeGrad: k_lifting also stores the volume gradients; devDt: the stage kernels read dt from
// dtOut[3] (b_dt carries RKb); fuseDt: this stage (the first of a step) evaluates CalcTimeStep on the way and finishes it in
// dtOut[3]; postHalo: the next stage's U-face halo may be posted behind this stage's halo-dependent volume kernel
struct StageOpt {
    int storeGrad = 0;
    bool devDt = false, fuseDt = false, postHalo = true;
};

// CalcTimeStep tail in device-paced stepping, on stream st: flag -> acc[2], min over the ranks, dt -> acc[3]
int dt_finish_on(dgx_handle* h, cudaStream_t st) {
    if (h->comm) {
        k_dt_pre<<<1, 1, 0, st>>>(h->dtOut, h->errFlag);
        if (check_launch(h, "k_dt_pre")) return 1;
        NK(g_nccl.AllReduce(h->dtOut, h->dtOut, 3, ncclFloat64, ncclMin, h->comm, st));  // calctimestep.f90:181
    }
    k_dt_finish<<<1, 1, 0, st>>>(h->dtOut, h->errFlag, h->dtHist, h->dtHistCap);
    return check_launch(h, "k_dt_finish");
}

int rhs(dgx_handle* h, int mode, double t, double mRKA, double b_dt, StageTimes* st = nullptr, const StageOpt& o = StageOpt()) {
    const dgx_config& c = h->cfg;
    const KernelTable* kt = h->kt;
    KParams P = h->P;
    P.storeGrad = (mode == 0 || o.storeGrad) ? 1 : 0;
    h->gradValid = c.parabolic && P.storeGrad;
    P.Um = h->Uf[h->cur][0];
    P.Us = h->Uf[h->cur][1];
    P.UmNext = h->Uf[h->cur ^ 1][0];
    P.UsNext = h->Uf[h->cur ^ 1][1];
    P.dtDev = o.devDt ? h->dtOut : nullptr;
    P.dtAcc = h->dtOut; P.dtCFL = c.CFLScale; P.dtDFL = c.DFLScale;
    // CalcTimeStep rides on stage 1's lifting kernel unless there is none (Euler) or the RHS filters U first (dg.f90:331: the
    // time step belongs to the unfiltered state the step starts from)
    P.dtFuse = (o.fuseDt && c.parabolic && !h->P.FilterMat) ? 1 : 0;
    if (!o.devDt) P.bulkDev = nullptr;
    const bool multi = c.nRanks > 1 && !h->NbProc.empty();
    // CalcSource (dg.f90:418): the volume kernels store Ut (MODE 0), k_source_rk adds the source and does the stage update
    if (mode == 1 && h->rk3Stage > 0) {  // timestep.f90:168-186
        const int i = h->rk3Stage - 1;
        P.rk3 = i == 0 ? 1 : 2;
        P.rk3Delta = h->RKdelta[i]; P.rk3G1 = h->RKg1[i]; P.rk3G2 = h->RKg2[i]; P.rk3G3 = h->RKg3[i];
    }
    // overintegration (step 14): the volume kernels leave Ut without the Jacobian, k_overint adds the sources of step 13 in
    // reference space, filters and applies the Jacobian; k_source_rk (sources switched off) then does the stage update
    const bool oint = P.overint != 0;
    P.noJac = oint ? 1 : 0;
    // the channel forcing alone does not need the extra pass: the volume kernels add it in their epilogue
    if (P.tcSource && !(P.iniExactFunc == 4 || P.spMat || P.rk3 || oint)) P.tcSource = 2;
    const bool src = P.iniExactFunc == 4 || P.tcSource == 1 || P.spMat || P.rk3 || oint;
    const int vmode = src ? 0 : mode;
    KParams Ps = P;   // parameters of k_source_rk
    if (oint) { Ps.iniExactFunc = 0; Ps.tcSource = 0; Ps.spMat = nullptr; }
    auto source_stage = [&](void) -> int {
        if (oint) { kt->overint(P, t, c.nElems, h->s); if (c.nElems > 0 && check_launch(h, "k_overint")) return 1; }
        if (mode == 0 && oint) return 0;  // Ut is complete (k_overint has added the sources)
        kt->source_rk(Ps, mode, t, mRKA, b_dt, c.nElems, h->s);
        return (c.nElems > 0 && check_launch(h, "k_source_rk")) ? 1 : 0;
    };
    auto mark = [&](void) { if (st) cudaEventRecord(st->ev[st->nev++], h->s); };
    NvtxRange nvtxRhs(mode == 1 ? "DGTimeDerivative_weakForm + RK stage" : "DGTimeDerivative_weakForm");
    mark();
    if (o.fuseDt && !P.dtFuse) {
        kt->timestep(P, c.CFLScale, c.DFLScale, h->dtOut, h->s);
        if (c.nElems && check_launch(h, "k_timestep")) return 1;
        if (dt_finish_on(h, h->s)) return 1;
    }
    if (P.FilterMat) {
        // 1. dg.f90:331: filter U in place (every RHS evaluation, like the reference) and re-extract the face states
        kt->filter(P, c.nElems, h->s);
        if (c.nElems > 0 && check_launch(h, "k_filter")) return 1;
        if (h->hasMortar() && mortar_u(h, P.Um, P.Us, 5)) return 1;
    }
    bool waitU = false;  // the halo-dependent launches of this stage have to wait for evUhalo
    if (multi) {
        CK(cudaEventRecord(h->evFaces, h->s));
        if (!h->uHaloPosted) {
            CK(cudaStreamWaitEvent(h->cs, h->evFaces, 0));
            if (exchange(h, P.Um, P.Us, 5)) return 1;
            CK(cudaEventRecord(h->evUhalo, h->cs));
            waitU = true;
        } else {
            waitU = !h->uHaloJoined;  // posted behind the previous stage's volume kernel; joined: already ordered before stream s
        }
        h->uHaloPosted = h->uHaloJoined = false;
    }
    KParams Pi = P, Pb = P;
    if (multi) {
        Pi.elemList = h->innerList; Pi.nList = h->nInner;
        Pb.elemList = h->bndList; Pb.nList = h->nBnd;
    }
    const bool mortar = h->hasMortar();
    // multi rank, conforming mesh: the halo-dependent launches (elements / sides touching an MPI side) run on a second,
    // high-priority stream concurrently with the inner-element kernels, so neither waits for the other's tail
    const bool split2 = multi && !mortar && !getenv("DGX_NO_SPLIT_STREAM");
    static const bool earlyHalo = !getenv("DGX_NO_EARLY_HALO");
    if (split2) {
        CK(cudaStreamWaitEvent(h->s2, h->evFaces, 0));
        if (waitU) CK(cudaStreamWaitEvent(h->s2, h->evUhalo, 0));
        if (c.parabolic) {
            kt->lifting(Pi, h->nInner, h->s);
            if (h->nInner > 0 && check_launch(h, "k_lifting")) return 1;
            if (h->nBnd) { kt->lifting(Pb, h->nBnd, h->s2); if (check_launch(h, "k_lifting(bnd)")) return 1; }
            CK(cudaEventRecord(h->evGrad, h->s2));
            CK(cudaStreamWaitEvent(h->cs, h->evGrad, 0));
            if (exchange(h, P.gm, P.gs, 12)) return 1;
            CK(cudaEventRecord(h->evGhalo, h->cs));
            CK(cudaStreamWaitEvent(h->s, h->evGrad, 0));  // inner sides may border halo-dependent elements
            if (P.dtFuse) {  // both lifting launches are complete on s: finish CalcTimeStep on the communication stream
                CK(cudaEventRecord(h->evDtPre, h->s));
                CK(cudaStreamWaitEvent(h->cs, h->evDtPre, 0));
                if (dt_finish_on(h, h->cs)) return 1;
                CK(cudaEventRecord(h->evDt, h->cs));
            }
        }
        mark();
        kt->sideflux(P, 0, c.lastInnerSide, h->s);
        if (c.lastInnerSide > 0 && check_launch(h, "k_sideflux")) return 1;
        CK(cudaEventRecord(h->evSide, h->s));
        mark();
        CK(cudaStreamWaitEvent(h->s2, h->evSide, 0));
        if (c.parabolic) CK(cudaStreamWaitEvent(h->s2, h->evGhalo, 0));
        const int nMPI = c.lastMPISide_YOUR - c.firstMPISide_MINE + 1;
        if (nMPI > 0) { kt->sideflux(P, c.firstMPISide_MINE - 1, nMPI, h->s2); if (check_launch(h, "k_sideflux(mpi)")) return 1; }
        if (P.dtFuse) { CK(cudaStreamWaitEvent(h->s2, h->evDt, 0)); CK(cudaStreamWaitEvent(h->s, h->evDt, 0)); }
        if (h->nBnd) { kt->volsurf(Pb, vmode, mRKA, b_dt, h->nBnd, h->s2); if (check_launch(h, "k_volsurf(bnd)")) return 1; }
        CK(cudaEventRecord(h->evBnd, h->s2));
        const bool post = mode == 1 && earlyHalo && o.postHalo && !P.FilterMat && !h->anyMortar;
        if (post && !src) {
            // every MPI side belongs to a halo-dependent element: their next-stage face states are complete here
            CK(cudaStreamWaitEvent(h->cs, h->evBnd, 0));
            if (exchange(h, P.UmNext, P.UsNext, 5)) return 1;
            CK(cudaEventRecord(h->evUhalo, h->cs));
            h->uHaloPosted = true;
        }
        if (h->nInner) { kt->volsurf(Pi, vmode, mRKA, b_dt, h->nInner, h->s); if (check_launch(h, "k_volsurf(inner)")) return 1; }
        CK(cudaStreamWaitEvent(h->s, h->evBnd, 0));
        if (src && source_stage()) return 1;
        if (post && src) {
            CK(cudaEventRecord(h->evNext, h->s));
            CK(cudaStreamWaitEvent(h->cs, h->evNext, 0));
            if (exchange(h, P.UmNext, P.UsNext, 5)) return 1;
            CK(cudaEventRecord(h->evUhalo, h->cs));
            h->uHaloPosted = true;
        }
        mark();
        if (mode == 1) h->cur ^= 1;
        if (!h->capturing) h->warm = true;
        return 0;
    }
    if (c.parabolic) {
        if (mortar && !multi && mortar_liftflux(h, P)) return 1;
        kt->lifting(multi ? Pi : P, multi ? h->nInner : c.nElems, h->s);
        if ((multi ? h->nInner : c.nElems) > 0 && check_launch(h, "k_lifting")) return 1;
        if (multi) {
            if (waitU) CK(cudaStreamWaitEvent(h->s, h->evUhalo, 0));
            if (mortar && mortar_liftflux(h, P)) return 1;
            if (h->nBnd) { kt->lifting(Pb, h->nBnd, h->s); if (check_launch(h, "k_lifting(bnd)")) return 1; }
            if (mortar && mortar_u(h, P.gm, P.gs, 12)) return 1;
            CK(cudaEventRecord(h->evGrad, h->s));
            CK(cudaStreamWaitEvent(h->cs, h->evGrad, 0));
            if (exchange(h, P.gm, P.gs, 12)) return 1;
            CK(cudaEventRecord(h->evGhalo, h->cs));
        } else if (mortar && mortar_u(h, P.gm, P.gs, 12)) return 1;
        if (P.dtFuse && dt_finish_on(h, h->s)) return 1;
    } else if (multi) {
        if (waitU) CK(cudaStreamWaitEvent(h->s, h->evUhalo, 0));
    }
    mark();
    // BC + inner sides
    kt->sideflux(P, 0, c.lastInnerSide, h->s);
    if (c.lastInnerSide > 0 && check_launch(h, "k_sideflux")) return 1;
    mark();
    if (!multi) {
        if (mortar && mortar_flux(h, P.Flux, 5, 1)) return 1;
        kt->volsurf(P, vmode, mRKA, b_dt, c.nElems, h->s);
        if (check_launch(h, "k_volsurf")) return 1;
    } else {
        if (h->nInner) { kt->volsurf(Pi, vmode, mRKA, b_dt, h->nInner, h->s); if (check_launch(h, "k_volsurf(inner)")) return 1; }
        if (c.parabolic) CK(cudaStreamWaitEvent(h->s, h->evGhalo, 0));
        const int nMPI = c.lastMPISide_YOUR - c.firstMPISide_MINE + 1;
        if (nMPI > 0) { kt->sideflux(P, c.firstMPISide_MINE - 1, nMPI, h->s); if (check_launch(h, "k_sideflux(mpi)")) return 1; }
        if (mortar && mortar_flux(h, P.Flux, 5, 1)) return 1;
        if (h->nBnd) { kt->volsurf(Pb, vmode, mRKA, b_dt, h->nBnd, h->s); if (check_launch(h, "k_volsurf(bnd)")) return 1; }
    }
    if (src && source_stage()) return 1;
    if (mode == 1 && mortar && mortar_u(h, P.UmNext, P.UsNext, 5)) return 1;
    mark();
    if (mode == 1) h->cur ^= 1;
    if (!h->capturing) h->warm = true;
    return 0;
}

int prolong_current(dgx_handle* h) {
    KParams P = h->P;
    P.Um = h->Uf[h->cur][0];
    P.Us = h->Uf[h->cur][1];
    h->kt->prolong(P, h->cfg.nElems, h->s);
    if (check_launch(h, "k_prolong")) return 1;
    return h->hasMortar() ? mortar_u(h, P.Um, P.Us, 5) : 0;
}

int check_err_flag(dgx_handle* h, const char* where, bool checkDt = false) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, h->errFlag, sizeof(int), cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    if (flag & 1) return fail(h, "%s: unsupported boundary condition type (supported: 2,3,4,9,91,23,24,25,27)", where);
    if ((flag & 2) && checkDt) return fail(h, "%s: timestep is NaN / state not admissible (density, convective / viscous timestep)", where);
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
extern "C" {

const char* dgx_last_error(const dgx_handle* h) { return h ? h->err.c_str() : "null handle"; }
long long dgx_launch_count(const dgx_handle* h) { return h ? h->launches : 0; }
unsigned long dgx_sizeof_config(void) { return (unsigned long)sizeof(dgx_config); }

int dgx_halo_plan(int nNbProcs, const int* NbProc, const int* nMine, const int* nYour, const int* offMine, const int* offYour,
                  int cap, int* out) {
    std::vector<int> nb(NbProc, NbProc + nNbProcs), m(nMine, nMine + nNbProcs), y(nYour, nYour + nNbProcs), om(offMine, offMine + nNbProcs),
        oy(offYour, offYour + nNbProcs);
    std::vector<HaloMsg> plan;
    halo_plan(nb, m, y, om, oy, plan);
    for (size_t i = 0; i < plan.size() && (int)i < cap; i++) {
        out[5 * i + 0] = plan[i].peer; out[5 * i + 1] = plan[i].isSend; out[5 * i + 2] = plan[i].slaveArray;
        out[5 * i + 3] = plan[i].side0; out[5 * i + 4] = plan[i].nSides;
    }
    return (int)plan.size();
}

int dgx_nccl_unique_id(char* out128) {
    std::string err;
    if (!g_nccl.load(err)) return 1;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return 2;
    memcpy(out128, id.internal, 128);
    return 0;
}

void dgx_destroy(dgx_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->s) cudaStreamSynchronize(h->s);
    if (h->cs) cudaStreamSynchronize(h->cs);
    if (h->s2) cudaStreamSynchronize(h->s2);
    // the step graph holds NCCL kernels of this communicator: it has to go first (ncclCommDestroy waits for captured work)
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    if (h->comm) g_nccl.CommDestroy(h->comm);
    for (void* p : h->allocs) cudaFree(p);
    if (h->hPinned) cudaFreeHost(h->hPinned);
    for (cudaEvent_t e : h->evCap) if (e) cudaEventDestroy(e);
    cudaEvent_t evs[] = {h->evFaces, h->evUhalo, h->evGrad, h->evGhalo, h->evT0, h->evT1, h->evSide, h->evBnd, h->evDtPre, h->evDt, h->evNext};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (h->s) cudaStreamDestroy(h->s);
    if (h->cs) cudaStreamDestroy(h->cs);
    if (h->s2) cudaStreamDestroy(h->s2);
    delete h;
}

int dgx_create(dgx_handle** out, const dgx_config* cfg) {
    if (!out || !cfg) return 1;
    dgx_handle* h = new dgx_handle();
    *out = h;  // returned even on failure so that dgx_last_error can be queried; caller destroys it
    h->cfg = *cfg;
    const dgx_config& c = h->cfg;
    if (c.N < 1 || c.N > 9) return fail(h, "polynomial degree N=%d not supported (1..9)", c.N);
    h->n = c.N + 1; h->n2 = h->n * h->n; h->n3 = h->n2 * h->n;
    h->kt = kernel_table(c.N, c.nodeType);
    if (!h->kt) return fail(h, "no kernels compiled for N=%d (rebuild with this degree enabled)", c.N);
    if (c.nodeType != 1 && c.nodeType != 2) return fail(h, "nodeType must be 1 (Gauss) or 2 (Gauss-Lobatto)");
    if (c.splitDG >= 0 && c.nodeType != 2) return fail(h, "Wrong Pointset: Gauss-Lobatto-Points are mandatory for using SplitDG !");
    if (c.splitDG < -1 || c.splitDG > 4) return fail(h, "SplitDG variant %d not available (SD=0, MO=1, DU=2, KG=3, PI=4; CH is not implemented for GPU, splitflux.f90:133-135)", c.splitDG);
    if (c.riemann < 0 || (c.riemann > 7 && c.riemann != 9)) return fail(h, "Riemann solver %d not available (LF=0, Roe=1, RoeL2=2, RoeEntropyFix=3, HLL=4, HLLC=5, HLLE=6, HLLEM=7, FluxAverage=9)", c.riemann);
    if (c.splitDG >= 0 && c.riemann >= 4 && c.riemann <= 7) return fail(h, "HLL-type Riemann solvers are not supported for SPLIT_DG=ON (as in the reference, src/CMakeLists.txt:108-127)");
    if (c.splitDG < 0 && c.riemann == 9) return fail(h, "the flux-average Riemann solver requires SplitDG (riemann.f90:1236-1252)");
    if (c.nRKStages < 1) return fail(h, "nRKStages < 1");
    if (c.lifting < 0 || c.lifting > 2) return fail(h, "lifting %d not available (1: BR1, 2: BR2)", c.lifting);
    if (c.nMortarSides < 0) return fail(h, "nMortarSides < 0");
    if (c.nMortarSides > 0 && (!c.MortarType || !c.MortarInfo || !c.M_0_1 || !c.M_0_2 || !c.M_1_0 || !c.M_2_0))
        return fail(h, "mortar mesh (nMortarSides=%d) needs MortarType, MortarInfo and the operators M_0_1, M_0_2, M_1_0, M_2_0", c.nMortarSides);
    // boundary conditions: types of the equation system (getboundaryflux.f90:262-838) and reference states that exist
    for (int sd = 0; sd < c.nBCSides; sd++) {
        const int bct = c.BCSides[2 * sd], bcs = c.BCSides[2 * sd + 1];
        const bool known = bct == 2 || bct == 3 || bct == 4 || bct == 9 || bct == 91 || bct == 23 || bct == 24 || bct == 25 || bct == 27;
        if (!known) return fail(h, "boundary side %d: boundary condition type %d not supported (supported: 2,3,4,9,91,23,24,25,27)", sd + 1, bct);
        if (bcs > c.nRefState) return fail(h, "boundary side %d: BC state %d exceeds nRefState = %d", sd + 1, bcs, c.nRefState);
        const bool needsRef = bct == 2 || bct == 4 || bct == 23 || bct == 24 || bct == 25 || bct == 27;
        if (needsRef && c.nRefState < 1) return fail(h, "boundary side %d: boundary condition type %d needs a reference state, nRefState = 0", sd + 1, bct);
    }
    CK(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major < 10) return fail(h, "device %s is sm_%d%d; this library is built for sm_100a (B200) only", prop.name, prop.major, prop.minor);
    int lo, hi;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->s, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&h->cs, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&h->s2, cudaStreamNonBlocking, hi));
    cudaEvent_t* evs[] = {&h->evFaces, &h->evUhalo, &h->evGrad, &h->evGhalo, &h->evSide, &h->evBnd, &h->evDtPre, &h->evDt, &h->evNext};
    for (cudaEvent_t* e : evs) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (cudaEvent_t& e : h->evCap) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreate(&h->evT0));
    CK(cudaEventCreate(&h->evT1));
    { int r = h->kt->setup(); if (r) return fail(h, "cudaFuncSetAttribute failed: %s", cudaGetErrorString((cudaError_t)r)); }

    const int n = h->n, n2 = h->n2, n3 = h->n3;
    const size_t nDOF = h->nDOF(), nFace = h->nFace();
    // ---- device arrays
    if (dalloc(h, &h->U, 5 * nDOF) || dalloc(h, &h->Ut, 5 * nDOF) || dalloc(h, &h->Ut_tmp, 5 * nDOF)) return 1;
    if (dalloc(h, &h->gradU, c.parabolic ? 12 * nDOF : 1)) return 1;
    if (dalloc(h, &h->metrics, 9 * nDOF) || dalloc(h, &h->sJ, nDOF) || dalloc(h, &h->geo, 10 * nFace)) return 1;
    for (int b = 0; b < 2; b++) for (int m = 0; m < 2; m++) if (dalloc(h, &h->Uf[b][m], 5 * nFace)) return 1;
    if (dalloc(h, &h->gm, c.parabolic ? 12 * nFace : 1) || dalloc(h, &h->gs, c.parabolic ? 12 * nFace : 1)) return 1;
    if (dalloc(h, &h->Flux, 5 * nFace)) return 1;
    if (dalloc(h, &h->stage, 5 * nDOF > 12 * nFace ? 5 * nDOF : 12 * nFace)) return 1;
    if (dalloc(h, &h->dtOut, 8) || dalloc(h, &h->errFlag, 1)) return 1;
    CK(cudaMallocHost((void**)&h->hPinned, 8 * sizeof(double)));
    // ---- geometry: upload in reference layout, repack on the device
    {
        double *mf, *mg, *mh, *nv, *t1, *t2, *se;
        if (upload(h, &mf, c.Metrics_fTilde, 3 * nDOF) || upload(h, &mg, c.Metrics_gTilde, 3 * nDOF) || upload(h, &mh, c.Metrics_hTilde, 3 * nDOF)) return 1;
        k_pack_metrics<<<blocks_for(9 * nDOF, 256), 256, 0, h->s>>>(mf, mg, mh, h->metrics, n3, 9 * nDOF);
        if (check_launch(h, "k_pack_metrics")) return 1;
        CK(cudaMemcpyAsync(h->sJ, c.sJ, nDOF * sizeof(double), cudaMemcpyHostToDevice, h->s));
        if (upload(h, &nv, c.NormVec, 3 * nFace) || upload(h, &t1, c.TangVec1, 3 * nFace) || upload(h, &t2, c.TangVec2, 3 * nFace) || upload(h, &se, c.SurfElem, nFace)) return 1;
        if (nFace) { k_pack_geo<<<blocks_for(10 * nFace, 256), 256, 0, h->s>>>(nv, t1, t2, se, h->geo, n2, 10 * nFace); if (check_launch(h, "k_pack_geo")) return 1; }
        CK(cudaStreamSynchronize(h->s));
        // free the temporaries (last 7 allocations)
        for (int x = 0; x < 7; x++) { cudaFree(h->allocs.back()); h->allocs.pop_back(); }
    }
    // ---- kernel parameter block
    KParams& P = h->P;
    memset(&P, 0, sizeof P);
    P.nElems = c.nElems; P.nSides = c.nSides; P.nBCSides = c.nBCSides;
    P.firstInner = c.firstInnerSide; P.lastInner = c.lastInnerSide;
    P.firstMINE = c.firstMPISide_MINE; P.lastMINE = c.lastMPISide_MINE; P.firstYOUR = c.firstMPISide_YOUR; P.lastYOUR = c.lastMPISide_YOUR;
    P.splitDG = c.splitDG; P.riemann = c.riemann; P.parabolic = c.parabolic;
    P.eos.kappa = c.EOS_Vars[0]; P.eos.R = c.EOS_Vars[1]; P.eos.Pr = c.EOS_Vars[2]; P.eos.mu0 = c.EOS_Vars[3];
    P.eos.Ts = c.EOS_Vars[4]; P.eos.Tref = c.EOS_Vars[5]; P.eos.ExpoSuth = c.EOS_Vars[6]; P.eos.cSuth = c.EOS_Vars[7];
    P.eos.viscLaw = c.viscLaw;
    for (int x = 0; x < n * n; x++) { P.D_T[x] = c.D_T[x]; P.D_Hat_T[x] = c.D_Hat_T[x]; P.DVolSurf[x] = c.DVolSurf ? c.DVolSurf[x] : 0.0; }
    for (int x = 0; x < n; x++) { P.L_Minus[x] = c.L_Minus[x]; P.L_Plus[x] = c.L_Plus[x]; P.L_HatMinus[x] = c.L_HatMinus[x]; P.L_HatPlus[x] = c.L_HatPlus[x]; }
    {
        int *e2s, *s2v2, *s2v2i, *bcs;
        double* ref;
        if (upload(h, &e2s, c.ElemToSide, (size_t)18 * c.nElems)) return 1;
        if (upload(h, &s2v2, c.S2V2, (size_t)2 * n2 * 30) || upload(h, &s2v2i, c.S2V2_inv, (size_t)2 * n2 * 30)) return 1;
        if (upload(h, &bcs, c.BCSides, (size_t)2 * c.nBCSides)) return 1;
        if (upload(h, &ref, c.RefStatePrim, (size_t)6 * (c.nRefState > 0 ? c.nRefState : 0))) return 1;
        P.E2S = e2s; P.S2V2 = s2v2; P.S2V2inv = s2v2i; P.BCSides = bcs; P.RefPrim = ref;
    }
    P.metrics = h->metrics; P.sJ = h->sJ; P.geo = h->geo;
    P.U = h->U; P.Ut = h->Ut; P.Ut_tmp = h->Ut_tmp; P.gradU = h->gradU;
    P.gm = h->gm; P.gs = h->gs; P.Flux = h->Flux; P.errFlag = h->errFlag;
    P.lifting = c.lifting == 2 ? 2 : 1; P.etaBR2 = c.etaBR2; P.etaBR2_wall = c.etaBR2_wall;
    P.liftWeak = (c.doWeakLifting && P.lifting != 2) ? 1 : 0;               // BR2 is always strong (lifting_vars.f90:77)
    P.liftCons = (P.liftWeak || c.doConservativeLifting) ? 1 : 0;          // lifting_br1.t90:95
    P.MortarType = nullptr;
    P.xGP = nullptr; P.advVel1 = c.AdvVel[0]; P.iniExactFunc = 0;
    P.tcSource = 0; P.tcDpdx = 0.0; P.tcBulkVel = 0.0;
    P.rk3 = 0; P.rk3Delta = P.rk3G1 = P.rk3G2 = P.rk3G3 = 0.0; P.rk3S2 = nullptr; P.rk3UPrev = nullptr;
    P.spMat = nullptr; P.spBase = nullptr;
    if (c.SpongeMat) {
        if (!c.SpBaseFlow) return fail(h, "SpongeMat given without SpBaseFlow");
        double *sm, *sb;
        if (upload(h, &sm, c.SpongeMat, nDOF) || dalloc(h, &sb, 5 * nDOF)) return 1;
        // base flow: reference layout -> tiles, through the staging buffer
        CK(cudaMemcpyAsync(h->stage, c.SpBaseFlow, 5 * nDOF * sizeof(double), cudaMemcpyHostToDevice, h->s));
        if (nDOF) { k_aos_to_soa<5><<<blocks_for(5 * nDOF, 256), 256, 0, h->s>>>(h->stage, sb, n3, 5 * nDOF); if (check_launch(h, "k_aos_to_soa(base flow)")) return 1; }
        CK(cudaStreamSynchronize(h->s));
        P.spMat = sm; P.spBase = sb;
    }
    if (c.IniExactFunc == 4) {
        if (!c.Elem_xGP) return fail(h, "IniExactFunc=4 (CalcSource) needs Elem_xGP");
        // reference layout (3,n^3,nElems) -> [elem][3][n^3]
        std::vector<double> xs((size_t)3 * nDOF);
        for (size_t e = 0; e < (size_t)c.nElems; e++)
            for (int node = 0; node < n3; node++)
                for (int d = 0; d < 3; d++) xs[(e * 3 + d) * n3 + node] = c.Elem_xGP[(e * n3 + node) * 3 + d];
        double* xd;
        if (upload(h, &xd, xs.data(), xs.size())) return 1;
        CK(cudaStreamSynchronize(h->s));
        P.xGP = xd; P.iniExactFunc = 4;
    }
    P.FilterMat = nullptr;
    if (c.FilterMat) {
        double* fm;
        if (upload(h, &fm, c.FilterMat, (size_t)n * n)) return 1;
        P.FilterMat = fm;
    }
    P.overint = 0; P.nUnder = c.N; P.noJac = 0;
    P.oiMat = P.oiDown = P.oiUp = P.sJNUnder = nullptr;
    if (c.OverintegrationType < 0 || c.OverintegrationType > 2) return fail(h, "Unknown OverintegrationType!");
    if (c.OverintegrationType == 1) {
        if (!c.OverintegrationMat) return fail(h, "OverintegrationType 1 (cut-off) needs OverintegrationMat");
        double* om;
        if (upload(h, &om, c.OverintegrationMat, (size_t)n * n)) return 1;
        P.oiMat = om; P.overint = 1; P.nUnder = c.NUnder;
    } else if (c.OverintegrationType == 2) {
        if (c.NUnder < 0 || c.NUnder >= c.N) return fail(h, "conservative cut-off overintegration needs 0 <= NUnder < N (overintegration.f90:139-160)");
        if (!c.Vdm_N_NUnder || !c.Vdm_NUnder_N || !c.sJNUnder) return fail(h, "OverintegrationType 2 needs Vdm_N_NUnder, Vdm_NUnder_N and sJNUnder");
        const size_t nu = (size_t)c.NUnder + 1;
        double *vd, *vu, *sj;
        if (upload(h, &vd, c.Vdm_N_NUnder, nu * n) || upload(h, &vu, c.Vdm_NUnder_N, nu * n) || upload(h, &sj, c.sJNUnder, nu * nu * nu * c.nElems)) return 1;
        P.oiDown = vd; P.oiUp = vu; P.sJNUnder = sj; P.overint = 2; P.nUnder = c.NUnder;
    }
    h->mp = MortarParams{nullptr, nullptr, nullptr, 0};
    if (c.nMortarSides > 0) {
        h->nMortarInner = c.lastMortarInnerSide - c.firstMortarInnerSide + 1;
        h->nMortarMPI = c.lastMortarMPISide - c.firstMortarMPISide + 1;
        if (h->nMortarInner < 0 || h->nMortarMPI < 0 || h->nMortarInner + h->nMortarMPI != c.nMortarSides)
            return fail(h, "mortar side ranges do not add up to nMortarSides");
        int *mt, *mi;
        double* M;
        if (upload(h, &mt, c.MortarType, (size_t)2 * c.nSides) || upload(h, &mi, c.MortarInfo, (size_t)8 * c.nMortarSides)) return 1;
        std::vector<double> Mh((size_t)4 * n * n);
        const double* src[4] = {c.M_0_1, c.M_0_2, c.M_1_0, c.M_2_0};
        for (int k = 0; k < 4; k++) for (int x = 0; x < n * n; x++) Mh[(size_t)k * n * n + x] = src[k][x];
        if (upload(h, &M, Mh.data(), Mh.size())) return 1;
        CK(cudaStreamSynchronize(h->s));
        P.MortarType = mt;
        h->mp = MortarParams{mt, mi, M, 0};
    }
    P.elemList = nullptr; P.nList = 0;
    P.flags = getenv("DGX_FLAGS") ? atoi(getenv("DGX_FLAGS")) : DGX_DEFAULT_FLAGS;
    // benign state on faces that are never written (slave side of BC sides), cf. dg.f90:118-121
    {
        std::vector<double> init(5 * nFace, 0.0);
        for (size_t s = 0; s < (size_t)c.nSides; s++)
            for (int pq = 0; pq < n2; pq++) { init[s * 5 * n2 + 0 * n2 + pq] = 1.0; init[s * 5 * n2 + 4 * n2 + pq] = 1.0; }
        for (int b = 0; b < 2; b++) for (int m = 0; m < 2; m++)
            if (nFace) CK(cudaMemcpyAsync(h->Uf[b][m], init.data(), 5 * nFace * sizeof(double), cudaMemcpyHostToDevice, h->s));
        CK(cudaStreamSynchronize(h->s));
    }
    h->RKA.assign(c.RKA, c.RKA + c.nRKStages);
    h->RKb.assign(c.RKb, c.RKb + c.nRKStages);
    h->RKc.assign(c.RKc, c.RKc + c.nRKStages);
    if (c.RKdelta || c.RKg1 || c.RKg2 || c.RKg3) {  // TimeDiscType LSERKK3 (timedisc_vars.f90:140-141)
        if (!(c.RKdelta && c.RKg1 && c.RKg2 && c.RKg3)) return fail(h, "three-register Runge-Kutta needs RKdelta, RKg1, RKg2 and RKg3");
        h->RKdelta.assign(c.RKdelta, c.RKdelta + c.nRKStages);
        h->RKg1.assign(c.RKg1, c.RKg1 + c.nRKStages);
        h->RKg2.assign(c.RKg2, c.RKg2 + c.nRKStages);
        h->RKg3.assign(c.RKg3, c.RKg3 + c.nRKStages);
        if (dalloc(h, &h->P.rk3S2, 5 * h->nDOF()) || dalloc(h, &h->P.rk3UPrev, 5 * h->nDOF())) return 1;
    }
    // ---- domain decomposition
    if (c.nRanks > 1) {
        for (int ib = 0; ib < c.nNbProcs; ib++) {
            h->NbProc.push_back(c.NbProc[ib]);
            h->nMine.push_back(c.nMPISides_MINE_Proc[ib]);
            h->nYour.push_back(c.nMPISides_YOUR_Proc[ib]);
            h->offMine.push_back(c.offsetMPISides_MINE[ib]);
            h->offYour.push_back(c.offsetMPISides_YOUR[ib]);
        }
        halo_plan(h->NbProc, h->nMine, h->nYour, h->offMine, h->offYour, h->plan);
        // element lists: elements touching an MPI side vs the rest
        std::vector<int> inner, bnd;
        for (int e = 0; e < c.nElems; e++) {
            bool mpi = false;
            for (int l = 0; l < 6; l++) {
                const int sid = c.ElemToSide[0 + 3 * (l + 6 * e)];
                if (sid >= c.firstMPISide_MINE) mpi = true;
                // big mortar sides depend on small sides that may be MPI sides: keep their elements behind the halo
                if (c.nMortarSides > 0 && c.MortarType[2 * (sid - 1)] > 0) mpi = true;
            }
            (mpi ? bnd : inner).push_back(e);
        }
        h->nInner = (int)inner.size(); h->nBnd = (int)bnd.size();
        if (upload(h, &h->innerList, inner.data(), inner.size()) || upload(h, &h->bndList, bnd.data(), bnd.size())) return 1;
        if (!c.ncclUniqueId) return fail(h, "nRanks>1 requires ncclUniqueId");
        if (!g_nccl.load(h->err)) return 1;
        ncclUniqueId id;
        memcpy(id.internal, c.ncclUniqueId, 128);
        NK(g_nccl.CommInitRank(&h->comm, c.nRanks, id, c.myRank));
        // orchestration choices that change the ORDER of the communication calls (U-face halo posted early, device-paced steps)
        // must be the same on every rank: a rank without mortar sides next to one with them would otherwise leave a point-to-point
        // exchange pending that its neighbour only answers after its next collective
        h->hPinned[5] = h->hasMortar() ? 1.0 : 0.0;
        CK(cudaMemcpyAsync(h->dtOut + 5, h->hPinned + 5, sizeof(double), cudaMemcpyHostToDevice, h->s));
        NK(g_nccl.AllReduce(h->dtOut + 5, h->dtOut + 5, 1, ncclFloat64, 2 /* ncclMax */, h->comm, h->s));
        CK(cudaMemcpyAsync(h->hPinned + 5, h->dtOut + 5, sizeof(double), cudaMemcpyDeviceToHost, h->s));
        CK(cudaStreamSynchronize(h->s));
        h->anyMortar = h->hPinned[5] > 0.0;
    }
    CK(cudaStreamSynchronize(h->s));
    h->cfg.RefStatePrim = nullptr;  // pointers are not retained
    h->cfg.MortarType = h->cfg.MortarInfo = nullptr;
    h->cfg.FilterMat = nullptr;
    h->cfg.Elem_xGP = nullptr;
    h->cfg.SpongeMat = h->cfg.SpBaseFlow = nullptr;
    h->cfg.M_0_1 = h->cfg.M_0_2 = h->cfg.M_1_0 = h->cfg.M_2_0 = nullptr;
    h->cfg.OverintegrationMat = h->cfg.Vdm_N_NUnder = h->cfg.Vdm_NUnder_N = h->cfg.sJNUnder = nullptr;
    return 0;
}

int dgx_sync(dgx_handle* h) {
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->s));
    CK(cudaStreamSynchronize(h->cs));
    CK(cudaStreamSynchronize(h->s2));
    return 0;
}

int dgx_set_state(dgx_handle* h, const double* U) {
    CK(cudaSetDevice(h->cfg.device));
    const size_t tot = 5 * h->nDOF();
    if (tot) {
        CK(cudaMemcpyAsync(h->stage, U, tot * sizeof(double), cudaMemcpyHostToDevice, h->s));
        k_aos_to_soa<5><<<blocks_for(tot, 256), 256, 0, h->s>>>(h->stage, h->U, h->n3, tot);
        if (check_launch(h, "k_aos_to_soa")) return 1;
        CK(cudaMemsetAsync(h->Ut_tmp, 0, tot * sizeof(double), h->s));
        h->gradValid = false;
        if (h->uHaloPosted) CK(cudaStreamWaitEvent(h->s, h->evUhalo, 0));  // a halo of the replaced state may still be travelling
        h->uHaloPosted = h->uHaloJoined = false;
        if (prolong_current(h)) return 1;
    }
    CK(cudaStreamSynchronize(h->s));
    return 0;
}

static int get_vol(dgx_handle* h, const double* soa, double* host, int soaStride, int soaOffset, int nvar) {
    CK(cudaSetDevice(h->cfg.device));
    const size_t tot = (size_t)nvar * h->nDOF();
    if (!tot) return 0;
    if (nvar == 5) k_soa_to_aos<5><<<blocks_for(tot, 256), 256, 0, h->s>>>(soa, h->stage, h->n3, tot, soaStride, soaOffset);
    else k_soa_to_aos<4><<<blocks_for(tot, 256), 256, 0, h->s>>>(soa, h->stage, h->n3, tot, soaStride, soaOffset);
    if (check_launch(h, "k_soa_to_aos")) return 1;
    CK(cudaMemcpyAsync(host, h->stage, tot * sizeof(double), cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    return 0;
}
int dgx_get_state(dgx_handle* h, double* U) { return get_vol(h, h->U, U, 5, 0, 5); }
int dgx_get_ut(dgx_handle* h, double* Ut) { return get_vol(h, h->Ut, Ut, 5, 0, 5); }
int dgx_get_gradients(dgx_handle* h, double* gx, double* gy, double* gz) {
    if (!h->cfg.parabolic) return fail(h, "gradients are only available for PARABOLIC runs");
    if (!h->gradValid) return fail(h, "dgx_get_gradients: the last RHS evaluation did not store the volume gradients (call dgx_time_derivative, or dgx_set_keep_gradients(h, 1) before the step)");
    if (get_vol(h, h->gradU, gx, 12, 0, 4)) return 1;
    if (get_vol(h, h->gradU, gy, 12, 4, 4)) return 1;
    return get_vol(h, h->gradU, gz, 12, 8, 4);
}

// Face arrays of the last RHS evaluation in the reference's layout (nVar,0:N,0:N,nSides), for parity checks of the
// orientation / sign conventions on the kernels themselves (unitTests/ProlongToFace.f90:70-114, SurfInt.f90:79-132):
// which = 0 U_master, 1 U_slave (5 variables), 2 Flux_master (5), 3..5 gradUx/y/z_master, 6..8 gradUx/y/z_slave (the 4 lifted
// variables u, v, w, T). A debugging path: device -> host copy, transposed on the host.
int dgx_get_face_array(dgx_handle* h, int which, double* out) {
    CK(cudaSetDevice(h->cfg.device));
    if (which < 0 || which > 8 || !out) return fail(h, "dgx_get_face_array: bad arguments");
    if (which >= 3 && !h->cfg.parabolic) return fail(h, "dgx_get_face_array: gradient traces exist only for PARABOLIC runs");
    const int n2 = h->n2, nvarDev = which < 3 ? 5 : 12, nvar = which < 3 ? 5 : 4, v0 = which < 3 ? 0 : 4 * ((which - 3) % 3);
    const double* src = which == 0 ? h->Uf[h->cur][0] : which == 1 ? h->Uf[h->cur][1] : which == 2 ? h->Flux : (which < 6 ? h->gm : h->gs);
    const size_t nS = (size_t)h->cfg.nSides;
    std::vector<double> tmp(nS * nvarDev * n2);
    CK(cudaStreamSynchronize(h->s));
    CK(cudaStreamSynchronize(h->cs));
    CK(cudaStreamSynchronize(h->s2));
    if (!tmp.empty()) CK(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t sd = 0; sd < nS; sd++)
        for (int pq = 0; pq < n2; pq++)
            for (int v = 0; v < nvar; v++) out[(sd * n2 + pq) * nvar + v] = tmp[(sd * nvarDev + v0 + v) * n2 + pq];
    return 0;
}

int dgx_time_derivative(dgx_handle* h, double t) {
    CK(cudaSetDevice(h->cfg.device));
    if (rhs(h, 0, t, 0.0, 0.0)) return 1;
    return check_err_flag(h, "dgx_time_derivative");
}

int dgx_rk_stage(dgx_handle* h, int iStage, double t, double dt) {
    CK(cudaSetDevice(h->cfg.device));
    if (iStage < 1 || iStage > h->cfg.nRKStages) return fail(h, "iStage out of range");
    StageOpt store;
    store.storeGrad = (iStage == h->cfg.nRKStages && h->keepGrad) ? 1 : 0;
    if (!h->RKg1.empty()) {  // TimeStepByLSERKK3 (timestep.f90:129-200)
        h->rk3Stage = iStage;
        const int rc = rhs(h, 1, t, 0.0, h->RKb[iStage - 1] * dt, nullptr, store);
        h->rk3Stage = 0;
        return rc;
    }
    const double mRKA = (iStage == 1) ? 0.0 : -1.0 * h->RKA[iStage - 1];
    return rhs(h, 1, t, mRKA, h->RKb[iStage - 1] * dt, nullptr, store);  // t: the stage time (timestep.f90:86-93)
}

int dgx_rk_step(dgx_handle* h, double t, double dt) {
    NvtxRange nvtxStep("TimeStepByLSERK");
    for (int st = 1; st <= h->cfg.nRKStages; st++)
        if (dgx_rk_stage(h, st, st == 1 ? t : t + h->RKc[st - 1] * dt, dt)) return 1;
    // an unsupported boundary condition type raises errFlag bit 1 in the first RHS (the reference Aborts in GetBoundaryFlux):
    // checked once, after the first step, so that the step loop stays free of host syncs
    if (!h->bcChecked) { h->bcChecked = true; return check_err_flag(h, "dgx_rk_step"); }
    return 0;
}

int dgx_set_keep_gradients(dgx_handle* h, int on) {
    if (!h) return 1;
    h->keepGrad = on ? 1 : 0;
    return 0;
}

int dgx_calc_timestep(dgx_handle* h, double* dt, int* errType) {
    NvtxRange nvtxDt("CalcTimeStep");
    CK(cudaSetDevice(h->cfg.device));
    const double big = 1.7976931348623157e308;
    h->hPinned[0] = big; h->hPinned[1] = big; h->hPinned[2] = 0.0;
    CK(cudaMemcpyAsync(h->dtOut, h->hPinned, 3 * sizeof(double), cudaMemcpyHostToDevice, h->s));
    k_clear_flag_bits<<<1, 1, 0, h->s>>>(h->errFlag, 2);  // bit 2 (dt) only: bit 1 (boundary condition) stays raised
    if (check_launch(h, "k_clear_flag_bits")) return 1;
    h->kt->timestep(h->P, h->cfg.CFLScale, h->cfg.DFLScale, h->dtOut, h->s);
    if (h->cfg.nElems && check_launch(h, "k_timestep")) return 1;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, h->errFlag, sizeof(int), cudaMemcpyDeviceToHost, h->s));
    if (h->comm) {
        CK(cudaStreamSynchronize(h->s));
        h->hPinned[4] = -(double)(flag != 0);
        CK(cudaMemcpyAsync(h->dtOut + 2, h->hPinned + 4, sizeof(double), cudaMemcpyHostToDevice, h->s));
        NK(g_nccl.AllReduce(h->dtOut, h->dtOut, 3, ncclFloat64, ncclMin, h->comm, h->s));  // calctimestep.f90:181
    }
    CK(cudaMemcpyAsync(h->hPinned, h->dtOut, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    if (h->comm) flag = (h->hPinned[2] < 0.0) ? 2 : 0;
    if (errType) *errType = flag ? 2 : 0;
    if (dt) *dt = h->hPinned[0] < h->hPinned[1] ? h->hPinned[0] : h->hPinned[1];
    return 0;
}

int dgx_temp_filter_time_deriv(dgx_handle* h, double dt, double tempFilterWidth) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->P.spBase) return fail(h, "dgx_temp_filter_time_deriv: no sponge base flow (SpongeMat / SpBaseFlow not set)");
    const size_t tot = 5 * h->nDOF();
    if (tot) { k_pruett<<<blocks_for(tot, 256), 256, 0, h->s>>>(h->U, h->P.spBase, dt / tempFilterWidth, tot); if (check_launch(h, "k_pruett")) return 1; }
    return 0;
}
int dgx_get_baseflow(dgx_handle* h, double* SpBaseFlow) {
    if (!h->P.spBase) return fail(h, "dgx_get_baseflow: no sponge base flow");
    return get_vol(h, h->P.spBase, SpBaseFlow, 5, 0, 5);
}

int dgx_set_channel_forcing(dgx_handle* h, int on, double dpdx, double BulkVel) {
    if (h->graph && ((on ? 1 : 0) != h->P.tcSource || dpdx != h->P.tcDpdx || BulkVel != h->P.tcBulkVel)) {
        cudaGraphExecDestroy(h->graph);   // the captured kernels carry the old forcing as launch parameters
        h->graph = nullptr;
    }
    h->P.tcSource = on ? 1 : 0;
    h->P.tcDpdx = dpdx;
    h->P.tcBulkVel = BulkVel;
    return 0;
}

namespace {
// CalcForcing on the device (testcase.f90:241-271): sum over this rank's nodes, sum over the ranks, / Vol -> tot[1]
int bulk_velocity_dev(dgx_handle* h) {
    h->kt->bulkvel(h->P, h->bvW, h->bvPart, h->s);
    if (h->cfg.nElems && check_launch(h, "k_bulkvel")) return 1;
    double* tot = h->bvPart + h->cfg.nElems;
    k_sum_partials<<<1, 256, 0, h->s>>>(h->bvPart, h->cfg.nElems, tot);
    if (check_launch(h, "k_sum_partials")) return 1;
    if (h->comm) NK(g_nccl.AllReduce(tot, tot, 1, ncclFloat64, 0 /* ncclSum */, h->comm, h->s));  // testcase.f90:266-268
    k_bulk_finish<<<1, 1, 0, h->s>>>(tot, h->bvVol);
    return check_launch(h, "k_bulk_finish");
}
}  // namespace

int dgx_calc_bulk_velocity(dgx_handle* h, const double* wGP, double Vol, double* BulkVel) {
    CK(cudaSetDevice(h->cfg.device));
    if (!wGP || !BulkVel) return fail(h, "dgx_calc_bulk_velocity: bad arguments");
    if (!h->bvPart) {
        if (dalloc(h, &h->bvPart, (size_t)h->cfg.nElems + 2) || dalloc(h, &h->bvW, (size_t)h->n)) return 1;
    }
    CK(cudaMemcpyAsync(h->bvW, wGP, h->n * sizeof(double), cudaMemcpyHostToDevice, h->s));
    h->bvVol = Vol;  // dgx_run_steps with CalcForcing every step reuses this quadrature
    if (bulk_velocity_dev(h)) return 1;
    double b = 0.0;
    CK(cudaMemcpyAsync(&b, h->bvPart + h->cfg.nElems, sizeof b, cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    *BulkVel = b / Vol;
    return 0;
}

namespace {
// wall diagnostics shared by dgx_calc_body_forces / dgx_calc_wall_velocity: r[x * nBCs + iBC], reduced over all ranks
int wall_diagnostics(dgx_handle* h, const char* who, const double* wGP, const int* BC, int nBCs, std::vector<double>& r) {
    CK(cudaSetDevice(h->cfg.device));
    const int nB = h->cfg.nBCSides;
    if (!wGP || nBCs < 1 || nBCs > 4096 || (nB > 0 && !BC)) return fail(h, "%s: bad arguments", who);
    for (int sd = 0; sd < nB; sd++)
        if (BC[sd] < 1 || BC[sd] > nBCs) return fail(h, "%s: BC(%d) = %d outside 1..nBCs = %d", who, sd + 1, BC[sd], nBCs);
    if (!h->bfPart || h->bfNBCs < nBCs) {
        if (dalloc(h, &h->bfPart, (size_t)WALL_NPART * (nB + nBCs)) || dalloc(h, &h->bfW, (size_t)h->n) || dalloc(h, &h->bfBC, (size_t)nB)) return 1;
        h->bfNBCs = nBCs;
    }
    CK(cudaMemcpyAsync(h->bfW, wGP, h->n * sizeof(double), cudaMemcpyHostToDevice, h->s));
    if (nB) CK(cudaMemcpyAsync(h->bfBC, BC, nB * sizeof(int), cudaMemcpyHostToDevice, h->s));
    double* tot = h->bfPart + (size_t)WALL_NPART * nB;
    if (nB) {
        k_wall_sides<<<nB, BF_THREADS, 0, h->s>>>(h->P, h->n, h->Uf[h->cur][0], h->bfW, h->bfPart);
        if (check_launch(h, "k_wall_sides")) return 1;
    }
    k_wall_reduce<<<nBCs, BF_THREADS, 0, h->s>>>(h->bfPart, h->bfBC, nB, nBCs, tot);
    if (check_launch(h, "k_wall_reduce")) return 1;
    if (h->comm) {  // calcbodyforces.f90:95-104, analyze_equation.f90:483-493
        NK(g_nccl.AllReduce(tot, tot, (size_t)7 * nBCs, ncclFloat64, 0 /* ncclSum */, h->comm, h->s));
        NK(g_nccl.AllReduce(tot + 7 * nBCs, tot + 7 * nBCs, (size_t)nBCs, ncclFloat64, 2 /* ncclMax */, h->comm, h->s));
        NK(g_nccl.AllReduce(tot + 8 * nBCs, tot + 8 * nBCs, (size_t)nBCs, ncclFloat64, 3 /* ncclMin */, h->comm, h->s));
    }
    r.resize((size_t)WALL_NPART * nBCs);
    CK(cudaMemcpyAsync(r.data(), tot, r.size() * sizeof(double), cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    return 0;
}
}  // namespace

int dgx_calc_body_forces(dgx_handle* h, const double* wGP, const int* BC, int nBCs, double* Fp, double* Fv) {
    if (!Fp || !Fv) return fail(h, "dgx_calc_body_forces: bad arguments");
    std::vector<double> r;
    if (wall_diagnostics(h, "dgx_calc_body_forces", wGP, BC, nBCs, r)) return 1;
    for (int b = 0; b < nBCs; b++)
        for (int d = 0; d < 3; d++) {
            Fp[3 * b + d] = r[(size_t)d * nBCs + b];
            Fv[3 * b + d] = r[(size_t)(3 + d) * nBCs + b];
        }
    return 0;
}

int dgx_calc_wall_velocity(dgx_handle* h, const double* wGP, const int* BC, int nBCs, const double* Surf, double* maxV, double* minV,
                           double* meanV) {
    if (!Surf || !maxV || !minV || !meanV) return fail(h, "dgx_calc_wall_velocity: bad arguments");
    std::vector<double> r;
    if (wall_diagnostics(h, "dgx_calc_wall_velocity", wGP, BC, nBCs, r)) return 1;
    for (int b = 0; b < nBCs; b++) {
        maxV[b] = r[(size_t)7 * nBCs + b];
        minV[b] = r[(size_t)8 * nBCs + b];
        // a boundary condition without wall sides keeps the reference's initial values (-1e14 / 1e14 / 0)
        meanV[b] = (maxV[b] > -WALL_HUGE && Surf[b] > 0.0) ? r[(size_t)6 * nBCs + b] / Surf[b] : 0.0;
    }
    return 0;
}

int dgx_analyze_tgv(dgx_handle* h, int NAnalyze, const double* Vdm, const double* wAnalyze, double Vol, double rho0, double* out15) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->cfg.parabolic) return fail(h, "dgx_analyze_tgv needs PARABOLIC (lifted gradients)");
    if (NAnalyze < 1 || NAnalyze > 32 || !Vdm || !wAnalyze || !out15) return fail(h, "dgx_analyze_tgv: bad arguments");
    if (!h->gradValid) return fail(h, "dgx_analyze_tgv: the last RHS evaluation did not store the volume gradients (call dgx_time_derivative, or dgx_set_keep_gradients(h, 1) before the step)");
    const int NA1 = NAnalyze + 1;
    if (h->tgvNA1 != NA1) {  // (re)upload the analysis basis
        if (upload(h, &h->tgvV, Vdm, (size_t)NA1 * h->n) || upload(h, &h->tgvW, wAnalyze, (size_t)NA1)) return 1;
        if (!h->tgvPart && dalloc(h, &h->tgvPart, (size_t)TGV_NPART * (h->cfg.nElems + 1) + 2 * TGV_NPART)) return 1;
        h->tgvNA1 = NA1;
    }
    double* tot = h->tgvPart + (size_t)TGV_NPART * h->cfg.nElems;
    int r = h->kt->tgv_analyze(h->P, NA1, h->tgvV, h->tgvW, h->tgvPart, h->s);
    if (r) return fail(h, "k_tgv_analyze: %s", cudaGetErrorString((cudaError_t)r));
    if (h->cfg.nElems && check_launch(h, "k_tgv_analyze")) return 1;
    k_tgv_reduce<<<1, TGV_THREADS, 0, h->s>>>(h->tgvPart, h->cfg.nElems, tot);
    if (check_launch(h, "k_tgv_reduce")) return 1;
    if (h->comm) {  // testcase.f90:416-446 MPI_REDUCE (sum x 11, max x 1); every rank gets the result here
        NK(g_nccl.AllReduce(tot, tot, TGV_NPART - 1, ncclFloat64, 0 /* ncclSum */, h->comm, h->s));
        NK(g_nccl.AllReduce(tot + TGV_NPART - 1, tot + TGV_NPART - 1, 1, ncclFloat64, 2 /* ncclMax */, h->comm, h->s));
    }
    double p[TGV_NPART];
    CK(cudaMemcpyAsync(p, tot, sizeof p, cudaMemcpyDeviceToHost, h->s));
    CK(cudaStreamSynchronize(h->s));
    const double mu0 = h->P.eos.mu0;
    const double T_mean = p[0] / Vol, Entropy = p[1] / Vol, Ekin = p[2] / Vol, Ekin_comp = p[3] / Vol / rho0;
    const double Enstr = p[4] / (rho0 * Vol), DR_u = p[5] * mu0 / Vol, DR_S = p[6] * 2. * mu0 / (rho0 * Vol);
    const double DR_Sd = p[7] * 2. * mu0 / (rho0 * Vol), DR_p = -p[8] / (rho0 * Vol), ED_S = p[9] / Vol, ED_D = p[10] * 4. / 3. / Vol;
    const double uPrime = sqrt(2. / 3. * Ekin);
    const double o[15] = {DR_S, DR_Sd + DR_p, Ekin, Ekin_comp, Enstr, DR_u, DR_S, DR_Sd, DR_p, p[11], T_mean, uPrime, Entropy, ED_S, ED_D};
    memcpy(out15, o, sizeof o);
    return 0;
}

namespace {
// one RK step of device-paced stepping: CalcForcing (optional), then the stages; stage 1 evaluates CalcTimeStep on the way
int dev_step(dgx_handle* h, bool forcing, bool storeLast, bool postLast) {
    NvtxRange nvtxStep("TimeStepByLSERK (device-paced)");
    if (forcing && bulk_velocity_dev(h)) return 1;
    const int nst = h->cfg.nRKStages;
    for (int st = 1; st <= nst; st++) {
        StageOpt o;
        o.devDt = true;
        o.fuseDt = st == 1;
        o.storeGrad = (st == nst && storeLast) ? 1 : 0;
        o.postHalo = st < nst || postLast;
        const double mRKA = (st == 1) ? 0.0 : -1.0 * h->RKA[st - 1];
        if (rhs(h, 1, 0.0, mRKA, h->RKb[st - 1], nullptr, o)) return 1;
    }
    return 0;
}

// A pair of RK steps captured once as a CUDA graph and replayed (the face double buffer is back in place after 2 nRKStages
// flips): one launch per 2 x (3 nRKStages + 1) kernels. Multi rank: the capture starts from "U-face halo complete and ordered
// before stream s" and ends in the same state (the halo the last stage posted is joined into s), which dgx_run_steps
// establishes before the first replay. Nothing is executed here.
void swap_event_sets(dgx_handle* h) {
    cudaEvent_t* evs[9] = {&h->evFaces, &h->evUhalo, &h->evGrad, &h->evGhalo, &h->evSide, &h->evBnd, &h->evDtPre, &h->evDt, &h->evNext};
    for (int i = 0; i < 9; i++) { cudaEvent_t t = *evs[i]; *evs[i] = h->evCap[i]; h->evCap[i] = t; }
}

int build_step_graph(dgx_handle* h, bool forcing, int key) {
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    swap_event_sets(h);   // the eager events keep their last recorded state
    struct Restore { dgx_handle* h; ~Restore() { swap_event_sets(h); } } restore{h};
    const long long l0 = h->launches;
    const int cur0 = h->cur;
    const bool posted0 = h->uHaloPosted, joined0 = h->uHaloJoined;
    const bool multi = h->cfg.nRanks > 1 && !h->NbProc.empty();
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(h->s, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return 1; }
    h->capturing = true;
    if (multi) h->uHaloPosted = h->uHaloJoined = true;
    int rc = dev_step(h, forcing, false, true);
    if (!rc) rc = dev_step(h, forcing, false, true);
    const bool endsPosted = multi && h->uHaloPosted;
    if (!rc && endsPosted && cudaStreamWaitEvent(h->s, h->evUhalo, 0) != cudaSuccess) rc = 1;  // join the communication stream
    cudaError_t e = cudaStreamEndCapture(h->s, &g);
    h->capturing = false;
    h->cur = cur0;
    h->uHaloPosted = posted0; h->uHaloJoined = joined0;
    h->graphLaunches = h->launches - l0;
    h->launches = l0;  // nothing ran
    if (rc || e != cudaSuccess || !g) {
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        return 1;
    }
    e = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); h->graph = nullptr; return 1; }
    h->graphCur = cur0;
    h->graphKey = key;
    h->graphEndsPosted = endsPosted;
    return 0;
}
}  // namespace

int dgx_run_steps(dgx_handle* h, int nSteps, double t, double dt, int flags, float* ms, long long* launches) {
    CK(cudaSetDevice(h->cfg.device));
    const bool adaptive = flags & 1, forcing = flags & 2;
    // non-conforming meshes cut by a rank boundary keep the host-paced sequence (same results; their projection kernels sit
    // between the halo phases on one stream, which the device-paced orchestration has not been validated for)
    const bool dev = (flags & 4) && !(h->cfg.nRanks > 1 && h->anyMortar);
    static const bool noGraph = getenv("DGX_NO_GRAPH") != nullptr;
    const bool graph = (flags & 8) && !noGraph && !h->graphFailed;
    if (forcing && (!h->bvPart || !(h->bvVol > 0.0))) return fail(h, "dgx_run_steps: CalcForcing every step needs a previous dgx_calc_bulk_velocity (weights, volume)");
    if ((flags & 4) && (!adaptive || h->P.iniExactFunc == 4 || !h->RKg1.empty()))
        return fail(h, "dgx_run_steps: device-paced stepping needs adaptive dt, a time-independent source and a 2-register Runge-Kutta scheme");
    const long long l0 = h->launches;
    const int keep = h->keepGrad;
    const bool multi = h->cfg.nRanks > 1 && !h->NbProc.empty();
    if (dev && h->dtHistCap < nSteps) {
        if (dalloc(h, &h->dtHist, (size_t)nSteps + 64)) return 1;
        h->dtHistCap = nSteps + 64;
        if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }  // the history buffer is a kernel argument
    }
    CK(cudaStreamSynchronize(h->s));
    CK(cudaEventRecord(h->evT0, h->s));
    if (dev) {
        // arm the accumulators: minima, flag, dt, step counter
        h->hPinned[0] = DT_HUGE; h->hPinned[1] = DT_HUGE; h->hPinned[2] = 0.0; h->hPinned[3] = 0.0; h->hPinned[4] = 0.0;
        CK(cudaMemcpyAsync(h->dtOut, h->hPinned, 5 * sizeof(double), cudaMemcpyHostToDevice, h->s));
        k_clear_flag_bits<<<1, 1, 0, h->s>>>(h->errFlag, 2);
        if (check_launch(h, "k_clear_flag_bits")) return 1;
        if (forcing) h->P.bulkDev = h->bvPart + h->cfg.nElems + 1;
        const int key = (forcing ? 1 : 0) | (h->P.tcSource ? 2 : 0) | (h->P.spMat ? 4 : 0) | (h->P.FilterMat ? 8 : 0);
        int it = 0;
        if (graph) {
            if (!h->warm && nSteps > 0) {  // first-use initialisation (occupancy queries, communicator channels) outside any capture
                if (dev_step(h, forcing, nSteps == 1 && keep, true)) return 1;
                it = 1;
            }
            const int nPairs = (nSteps - it - 1) / 2;
            if (nPairs > 0) {
                if ((!h->graph || h->graphCur != h->cur || h->graphKey != key) && build_step_graph(h, forcing, key)) h->graphFailed = 1;
                if (h->graph && h->graphCur == h->cur && h->graphKey == key) {
                    for (int p = 0; p < nPairs; p++) {
                        if (multi) {
                            // the graph's first stage expects its U-face halo complete and ordered before stream s
                            if (!h->uHaloPosted) {
                                CK(cudaEventRecord(h->evFaces, h->s));
                                CK(cudaStreamWaitEvent(h->cs, h->evFaces, 0));
                                if (exchange(h, h->Uf[h->cur][0], h->Uf[h->cur][1], 5)) return 1;
                                CK(cudaEventRecord(h->evUhalo, h->cs));
                            }
                            if (!h->uHaloPosted || !h->uHaloJoined) CK(cudaStreamWaitEvent(h->s, h->evUhalo, 0));
                        }
                        CK(cudaGraphLaunch(h->graph, h->s));
                        h->launches += h->graphLaunches;
                        if (multi) h->uHaloPosted = h->uHaloJoined = h->graphEndsPosted;  // the state the replay ends in
                    }
                    it += 2 * nPairs;
                }
            }
        }
        for (; it < nSteps; it++)
            if (dev_step(h, forcing, it == nSteps - 1 && keep, true)) return 1;
        h->dtHistCount = nSteps;
        // capture for the next call, outside its timed region
        if (graph && h->warm && (!h->graph || h->graphCur != h->cur || h->graphKey != key) && build_step_graph(h, forcing, key)) h->graphFailed = 1;
        h->P.bulkDev = nullptr;
        h->histOnHost = false;
    } else {
        h->dtHistHost.clear();
        h->histOnHost = true;
        for (int it = 0; it < nSteps; it++) {
            h->keepGrad = (it == nSteps - 1) ? keep : 0;  // nothing can read the gradients between the steps of this call
            if (forcing) {
                double b = 0.0;
                if (bulk_velocity_dev(h)) { h->keepGrad = keep; return 1; }
                CK(cudaMemcpyAsync(&b, h->bvPart + h->cfg.nElems + 1, sizeof b, cudaMemcpyDeviceToHost, h->s));
                CK(cudaStreamSynchronize(h->s));
                h->P.tcBulkVel = b;
            }
            if (adaptive) {
                int et = 0;
                if (dgx_calc_timestep(h, &dt, &et)) { h->keepGrad = keep; return 1; }
                if (et) { h->keepGrad = keep; return fail(h, "timestep is NaN / state not admissible at t=%g", t); }
            }
            if (dgx_rk_step(h, t, dt)) { h->keepGrad = keep; return 1; }
            h->dtHistHost.push_back(dt);
            t += dt;
        }
        h->keepGrad = keep;
    }
    CK(cudaEventRecord(h->evT1, h->s));
    CK(cudaEventSynchronize(h->evT1));
    CK(cudaStreamSynchronize(h->cs));
    CK(cudaStreamSynchronize(h->s2));
    float m = 0.f;
    CK(cudaEventElapsedTime(&m, h->evT0, h->evT1));
    if (ms) *ms = m;
    if (launches) *launches = h->launches - l0;
    return check_err_flag(h, "dgx_run_steps", dev);
}

int dgx_get_dt_history(dgx_handle* h, int cap, double* dts, int* count) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->histOnHost) {
        const int nh = (int)h->dtHistHost.size();
        if (count) *count = nh;
        for (int i = 0; i < nh && i < cap && dts; i++) dts[i] = h->dtHistHost[i];
        return 0;
    }
    const int n = h->dtHistCount < cap ? h->dtHistCount : cap;
    if (count) *count = h->dtHistCount;
    if (n > 0 && dts) {
        CK(cudaMemcpyAsync(dts, h->dtHist, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->s));
        CK(cudaStreamSynchronize(h->s));
    }
    return 0;
}

int dgx_step_graph_active(const dgx_handle* h) { return h && h->graph ? 1 : 0; }

int dgx_profile_stage(dgx_handle* h, double t, double dt, int cap, const char** names, float* ms, int* count) {
    if (count) *count = 0;
    CK(cudaSetDevice(h->cfg.device));
    StageTimes st;
    for (int i = 0; i < 8; i++) CK(cudaEventCreate(&st.ev[i]));
    const int stage = h->cfg.nRKStages > 1 ? 2 : 1;
    const double mRKA = (stage == 1) ? 0.0 : -1.0 * h->RKA[stage - 1];
    h->rk3Stage = h->RKg1.empty() ? 0 : stage;
    const int rc = rhs(h, 1, t, mRKA, h->RKb[stage - 1] * dt, &st);
    h->rk3Stage = 0;
    if (rc) return 1;
    CK(cudaStreamSynchronize(h->s));
    static const char* nm[] = {"halo+lifting", "sideflux", "volsurf_rk"};
    int cnt = 0;
    for (int i = 0; i + 1 < st.nev && cnt < cap && i < 3; i++) {
        float m = 0.f;
        cudaEventElapsedTime(&m, st.ev[i], st.ev[i + 1]);
        names[cnt] = nm[i];
        ms[cnt] = m;
        cnt++;
    }
    for (int i = 0; i < 8; i++) cudaEventDestroy(st.ev[i]);
    if (count) *count = cnt;
    return 0;
}

}  // extern "C"
