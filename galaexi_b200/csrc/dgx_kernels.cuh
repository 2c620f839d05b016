// Element- and side-centric CUDA kernels of the DGSEM operator for sm_100a (FP64 CUDA cores, HBM-bound).
//
// Data layout in HBM (all FP64, per-variable contiguous tiles so every warp access is a coalesced 8-byte
// stream; the reference's AoS layout U(nVar,i,j,k,iElem) is only used at the C ABI):
//   volume   U, Ut, Ut_tmp : [elem][5][n^3]        gradU : [elem][3*4][n^3] (d*4+v; d=x,y,z; v=u,v,w,T)
//            metrics       : [elem][9][n^3]        (f1 f2 f3 g1 g2 g3 h1 h2 h3)   sJ : [elem][n^3]
//   sides    Um, Us, Flux  : [side][5][n^2]        gm, gs : [side][12][n^2]       geo: [side][10][n^2]
//            geo = nv(3), t1(3), t2(3), SurfElem   (side-local (p,q) index, p fastest)
// One CTA works on one element (n^3 threads); its tile is staged in shared memory, the three 1-D
// tensor-product sweeps read lines from shared memory, face data are staged through shared memory in the
// element's own face-node order so that the surface integral is a conflict-free gather.
//
// Reference kernels replaced per launch (paths relative to /root/reference/src):
//   k_prolong   interpolation/prolongtoface.t90:168-344
//   k_lifting   equations/.../eos.f90:328 (ConsToPrim), dg/lifting/lifting_fillflux.t90:39-256,
//               idealgas/getboundaryflux.f90:945-1019, dg/lifting/lifting_volint.t90:262-328,
//               dg/surfint.t90:591-725 (x3, with sJ), interpolation/prolongtoface.t90 (x3 lifting)
//   k_sideflux  dg/fillflux.f90:45-172, equations/navierstokes/riemann.f90:401,638, getboundaryflux.f90:542-838
//   k_volsurf   eos.f90:328, flux.f90:304-384, dg/applydmatrix.t90:19-75, dg/volint.f90:265-353,
//               dg/surfint.t90:352-586, globals/vector.f90:210-226 (x -1), interpolation/applyjacobian.t90:196,
//               globals/vector.f90:163-183 (RK 2N update) and prolongtoface of the next stage
//   k_timestep  equations/navierstokes/calctimestep.f90:192-296
#pragma once
#include "dgx_physics.cuh"

namespace dgx {

constexpr int MAXN1 = 10;  // n = N+1 <= 10

struct KParams {
    int nElems, nSides, nBCSides;
    int firstInner, lastInner, firstMINE, lastMINE, firstYOUR, lastYOUR;  // 1-based inclusive
    int splitDG, riemann, parabolic;
    Eos eos;
    // small operator tables (copied to shared memory by the kernels), Fortran (a,b) at [a + n*b]
    double D_T[MAXN1 * MAXN1], D_Hat_T[MAXN1 * MAXN1], DVolSurf[MAXN1 * MAXN1];
    double L_Minus[MAXN1], L_Plus[MAXN1], L_HatMinus[MAXN1], L_HatPlus[MAXN1];
    const int* E2S;      // (3,6,nElems)
    const int* S2V2;     // (2,n,n,5,6)
    const int* S2V2inv;
    const int* BCSides;  // (2,nBCSides)
    const double* RefPrim;
    const double* metrics;
    const double* sJ;
    const double* geo;
    double* U;
    double* Ut;
    double* Ut_tmp;
    double* gradU;
    double* Um;  // face states read by this stage
    double* Us;
    double* UmNext;  // face states written by the RK epilogue (double buffered)
    double* UsNext;
    double* gm;
    double* gs;
    double* Flux;
    const int* elemList;  // optional indirection (MPI-boundary / inner element lists), or nullptr
    int nList;
    int* errFlag;
    // general lifting path (k_lifting<..., GEN=1>): BR2 (host FLEXI, dg/lifting/lifting_br2.t90) and/or big mortar faces
    int lifting;               // 1 BR1, 2 BR2
    int liftWeak, liftCons;    // non-default lifting forms (lifting.f90:81-85): weak form; conservative volume form
    double etaBR2, etaBR2_wall;
    const int* MortarType;     // (2,nSides) or nullptr when the mesh has no mortars
    const double* FilterMat;   // device (0:N,0:N) Fortran layout, or nullptr (FilterType 0)
    // manufactured-solution source (exactfunc.f90:946-1113): coordinates [elem][3][n^3] or nullptr, AdvVel(1), selector
    const double* xGP;
    double advVel1;
    int iniExactFunc;
    // sponge (sponge/sponge.f90:529-588): SpongeMat [elem][n^3] (= damping sigma / sJ) and base flow [elem][5][n^3], or nullptr
    const double* spMat;
    double* spBase;
    // channel forcing (testcase/channel/testcase.f90:277-296 TestcaseSource)
    int tcSource;   // 0 off, 1 added by k_source_rk / k_overint, 2 folded into the volume kernel's epilogue (no other source active)
    double tcDpdx, tcBulkVel;
    // three-register low-storage Runge-Kutta (timestep.f90:129-200 TimeStepByLSERKK3): 0 off, 1 first stage, 2 later stage;
    // registers S2 and UPrev [elem][5][n^3]
    int rk3;
    double rk3Delta, rk3G1, rk3G2, rk3G3;
    double *rk3S2, *rk3UPrev;
    int storeGrad;  // k_lifting also writes the volume gradients (only analysis / dgx_get_gradients read them; the viscous volume
                    // integral is formed inside k_lifting and travels as 4 doubles per node in Ut)
    // device-paced time stepping (dgx_run_steps flag 4): dtDev[3] holds the step's dt and the stage kernels form b_dt = RKb * dt
    // themselves; dtFuse: this k_lifting launch (stage 1 of a step) also evaluates CalcTimeStep of the state it reads and leaves
    // the minima in dtDev[0..1] (calctimestep.f90:98-186); bulkDev: CalcForcing's bulk velocity kept on the device
    const double* dtDev;
    double* dtAcc;
    int dtFuse;
    double dtCFL, dtDFL;
    const double* bulkDev;
    // overintegration of JU_t (dg/overintegration.f90:179-340, step 14 of the RHS): 0 none, 1 cut-off filter with oiMat (n x n),
    // 2 conservative cut-off: oiDown (0:NUnder,0:N), oiUp (0:N,0:NUnder), sJNUnder [elem][(NUnder+1)^3]; noJac: the volume
    // kernels leave Ut = -(DG operator) without the Jacobian, k_overint filters it and applies the Jacobian
    int overint, nUnder, noJac;
    const double *oiMat, *oiDown, *oiUp, *sJNUnder;
    int flags;  // tuning switches (DGX_FLAGS): 1 lifting: L2 prefetch of own later-phase data; 2 lifting: L2 prefetch of the element
                // a resident wave ahead; 4 / 8: the same two for k_volsurf2
};

enum { ZETA_MINUS = 1, ETA_MINUS = 2, XI_PLUS = 3, ETA_PLUS = 4, XI_MINUS = 5, ZETA_PLUS = 6 };

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// prefetch `bytes` contiguous bytes (128-byte lines) cooperatively with `nthr` threads
__device__ __forceinline__ void prefetch_block(const void* base, size_t bytes, int tid, int nthr) {
    const char* p = reinterpret_cast<const char*>(base);
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthr * 128) prefetch_l2(p + off);
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) with mbarrier completion: one thread moves a contiguous element tile from HBM
// to shared memory without holding registers; the consumers wait on the mbarrier's phase parity.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned phase) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(mbar), "r"(phase)
                     : "memory");
    }
}

template <int n>
__device__ __forceinline__ int s2v2(const int* __restrict__ tab, int c, int p, int q, int flip, int loc) {
    return __ldg(&tab[c + 2 * (p + n * (q + n * (flip + 5 * (loc - 1))))]);
}

// volume index of face node (a,b) at depth l for local side loc
template <int n>
__device__ __forceinline__ int face_vol_index(int loc, int a, int b, int l) {
    switch (loc) {
        case XI_MINUS:
        case XI_PLUS: return l + n * (a + n * b);
        case ETA_MINUS:
        case ETA_PLUS: return a + n * (l + n * b);
        default: return a + n * (b + n * l);
    }
}
__device__ __forceinline__ bool is_minus(int loc) { return loc == XI_MINUS || loc == ETA_MINUS || loc == ZETA_MINUS; }

// Extract the 6 faces of a shared-memory element tile [NVAR][n^3] into the side arrays (master if flip==0,
// else slave), in side-local orientation. GL: copy of the boundary layer; Gauss: contraction with L_Minus/L_Plus.
template <int n, int NT, int NVAR>
__device__ __forceinline__ void extract_faces(const double* __restrict__ tile, double* __restrict__ dstM, double* __restrict__ dstS,
                                              const int* __restrict__ e2s, const int* __restrict__ S2V2, const double* __restrict__ sLm,
                                              const double* __restrict__ sLp) {
    constexpr int n2 = n * n, n3 = n2 * n;
    for (int f = threadIdx.x; f < 6 * n2; f += blockDim.x) {
        const int loc = f / n2 + 1;
        const int pq = f - (loc - 1) * n2;
        const int q = pq / n, p = pq - q * n;
        const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
        const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
        const int a = s2v2<n>(S2V2, 0, p, q, flip, loc);
        const int b = s2v2<n>(S2V2, 1, p, q, flip, loc);
        double* dst = (flip == 0 ? dstM : dstS) + (size_t)side * NVAR * n2 + pq;
        if (NT == 2) {
            const int node = face_vol_index<n>(loc, a, b, is_minus(loc) ? 0 : n - 1);
#pragma unroll
            for (int v = 0; v < NVAR; v++) dst[v * n2] = tile[v * n3 + node];
        } else {
            const double* L = is_minus(loc) ? sLm : sLp;
            const int base = face_vol_index<n>(loc, a, b, 0);
            const int stride = face_vol_index<n>(loc, a, b, 1) - base;
#pragma unroll
            for (int v = 0; v < NVAR; v++) {
                double acc = tile[v * n3 + base] * L[0];
#pragma unroll
                for (int l = 1; l < n; l++) acc += tile[v * n3 + base + l * stride] * L[l];
                dst[v * n2] = acc;
            }
        }
    }
}

// Shared-memory element tile layout [slot][n^3].
template <int n>
struct Tile {
    // n == 8: bank-conflict-free 64-bit accesses for lines of all three directions WITHOUT padding. Position inside a
    // 16-double row pair = (i + 8 (j&1)) xor (k + 8 (k&1)); 8-byte bank (of 16 per half warp) = (i^k) + 8 ((j^k)&1).
    // Lane mappings that make a half warp hit 16 distinct banks: point-wise / zeta / eta lines: first coordinate fast;
    // xi lines: k fast, two adjacent j.
    static constexpr bool swz = (n == 8);
    static constexpr int SLOT = n * n * n;
    __device__ __forceinline__ static int idx(int i, int j, int k) {
        return swz ? (((i + 8 * (j & 1)) ^ (k + 8 * (k & 1))) + 16 * (j >> 1) + 64 * k) : (i + n * j + n * n * k);
    }
};

// Side-local node (p,q) handled by lane x of a face: the assignment is transposed when needed so that the wanted tile
// coordinate (b if b_fast, else a; (a,b) = S2V2(p,q)) varies fastest over the lanes -- conflict-free tile access.
template <int n>
__device__ __forceinline__ void face_lane(const int* __restrict__ S2V2, int x, int flip, int loc, bool b_fast, int& p, int& qq) {
    const bool a_on_p = s2v2<n>(S2V2, 0, 1, 0, flip, loc) != s2v2<n>(S2V2, 0, 0, 0, flip, loc);
    const bool tr = (a_on_p == b_fast);
    p = tr ? x / n : x % n;
    qq = tr ? x % n : x / n;
}

// extract_faces for a tile in Tile<n> layout, with conflict-free lane assignment (see face_lane)
template <int n, int NT, int NVAR>
__device__ __forceinline__ void extract_faces_tile(const double* __restrict__ tile, double* __restrict__ dstM, double* __restrict__ dstS,
                                                   const int* __restrict__ e2s, const int* __restrict__ S2V2, const double* __restrict__ sLm,
                                                   const double* __restrict__ sLp) {
    constexpr int n2 = n * n, SL = Tile<n>::SLOT;
    for (int f = threadIdx.x; f < 6 * n2; f += blockDim.x) {
        const int loc = f / n2 + 1;
        const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
        const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
        const bool xi = (loc == XI_MINUS || loc == XI_PLUS), eta = (loc == ETA_MINUS || loc == ETA_PLUS);
        int p, q;
        face_lane<n>(S2V2, f - (loc - 1) * n2, flip, loc, xi, p, q);
        const int a = s2v2<n>(S2V2, 0, p, q, flip, loc);
        const int b = s2v2<n>(S2V2, 1, p, q, flip, loc);
        double* dst = (flip == 0 ? dstM : dstS) + (size_t)side * NVAR * n2 + (p + n * q);
        if (NT == 2) {
            const int l = is_minus(loc) ? 0 : n - 1;
            const int id = xi ? Tile<n>::idx(l, a, b) : (eta ? Tile<n>::idx(a, l, b) : Tile<n>::idx(a, b, l));
#pragma unroll
            for (int v = 0; v < NVAR; v++) dst[v * n2] = tile[v * SL + id];
        } else {
            const double* L = is_minus(loc) ? sLm : sLp;
#pragma unroll
            for (int v = 0; v < NVAR; v++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < n; l++) {
                    const int id = xi ? Tile<n>::idx(l, a, b) : (eta ? Tile<n>::idx(a, l, b) : Tile<n>::idx(a, b, l));
                    acc = (l == 0) ? tile[v * SL + id] * L[0] : acc + tile[v * SL + id] * L[l];
                }
                dst[v * n2] = acc;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
template <int n, int NT>
__global__ void __launch_bounds__(n* n* n) k_prolong(const KParams P) {
    constexpr int n2 = n * n, n3 = n2 * n;
    __shared__ double tile[5 * n3];
    __shared__ double sLm[n], sLp[n];
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int t = threadIdx.x;
    if (t < n) { sLm[t] = P.L_Minus[t]; sLp[t] = P.L_Plus[t]; }
    const double* U = P.U + (size_t)e * 5 * n3;
#pragma unroll
    for (int v = 0; v < 5; v++) tile[v * n3 + t] = U[v * n3 + t];
    __syncthreads();
    extract_faces<n, NT, 5>(tile, P.Um, P.Us, P.E2S + 18 * e, P.S2V2, sLm, sLp);
}

// ---------------------------------------------------------------------------------------------------------
// Step 1 of the RHS when FilterType > 0 (dg/dg.f90:331, filter/filter.f90:272-306 -> changeBasis.t90:287-360): U <- FilterMat
// applied along xi, eta, zeta, in place; the face states of the filtered solution are extracted from the tile in the same
// kernel (they replace the ones the previous stage's epilogue wrote from the unfiltered state).
template <int n, int NT>
__global__ void __launch_bounds__(n* n* n) k_filter(const KParams P) {
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ double smem[];
    double *b0 = smem, *b1 = smem + 5 * n3, *sM = smem + 10 * n3, *sLm = sM + n * n, *sLp = sLm + n;
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int t = threadIdx.x;
    for (int x = t; x < n * n; x += n3) sM[x] = P.FilterMat[x];
    if (t < n) { sLm[t] = P.L_Minus[t]; sLp[t] = P.L_Plus[t]; }
    double* U = P.U + (size_t)e * 5 * n3;
#pragma unroll
    for (int v = 0; v < 5; v++) b0[v * n3 + t] = U[v * n3 + t];
    __syncthreads();
    const int k = t / n2, j = (t - k * n2) / n, i = t - k * n2 - j * n;
#pragma unroll
    for (int v = 0; v < 5; v++) {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < n; l++) a += sM[i + n * l] * b0[v * n3 + l + n * j + n2 * k];
        b1[v * n3 + t] = a;
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 5; v++) {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < n; l++) a += sM[j + n * l] * b1[v * n3 + i + n * l + n2 * k];
        b0[v * n3 + t] = a;
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 5; v++) {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < n; l++) a += sM[k + n * l] * b0[v * n3 + i + n * j + n2 * l];
        b1[v * n3 + t] = a;
        U[v * n3 + t] = a;
    }
    __syncthreads();
    extract_faces<n, NT, 5>(b1, P.Um, P.Us, P.E2S + 18 * e, P.S2V2, sLm, sLp);
}

// Source terms of one node in physical space (added to sJ * (-R); before the Jacobian they enter as src / sJ): manufactured
// solution (exactfunc.f90:946-1113), channel forcing (testcase/channel/testcase.f90:277-296), sponge (sponge.f90:529-588)
template <int n>
__device__ __forceinline__ void source_terms(const KParams& P, int e, int tt, double t, double (&src)[5]) {
    constexpr int n3 = n * n * n;
#pragma unroll
    for (int v = 0; v < 5; v++) src[v] = 0.0;
    if (P.iniExactFunc == 4) {
        const double* X = P.xGP + (size_t)e * 3 * n3 + tt;
        const double Kappa = P.eos.kappa, PP_Pi = acos(-1.0), Amplitude = 0.1;
        const double Omega = PP_Pi * 1.0, a = P.advVel1 * 2. * PP_Pi;
        double tmp[6];
        tmp[0] = -a + 3. * Omega;
        tmp[1] = -a + 0.5 * Omega * (1. + Kappa * 5.);
        tmp[2] = Amplitude * Omega * (Kappa - 1.);
        tmp[3] = 0.5 * ((9. + Kappa * 15.) * Omega - 8. * a);
        tmp[4] = Amplitude * (3. * Omega * Kappa - a);
        tmp[5] = P.parabolic ? 3. * P.eos.mu0 * Kappa * Omega * Omega / P.eos.Pr : 0.;
#pragma unroll
        for (int x = 0; x < 6; x++) tmp[x] = tmp[x] * Amplitude;
        const double arg = Omega * (X[0] + X[n3] + X[2 * n3]) - a * t;
        const double cosX = cos(arg), sinX = sin(arg), sin2 = 2. * sinX * cosX;
        src[0] = tmp[0] * cosX;
        src[1] = src[2] = src[3] = tmp[1] * cosX + tmp[2] * sin2;
        src[4] = tmp[3] * cosX + tmp[4] * sin2 + tmp[5] * sinX;
    }
    if (P.tcSource) { src[1] = __dadd_rn(src[1], -P.tcDpdx); src[4] = __dadd_rn(src[4], -__dmul_rn(P.tcDpdx, P.bulkDev ? __ldg(P.bulkDev) : P.tcBulkVel)); }
    if (P.spMat) {
        // Ut = Ut - SpongeMat (U - SpBaseFlow) before the Jacobian (sponge.f90:574-579): times sJ here
        const double sm = P.spMat[(size_t)e * n3 + tt] * P.sJ[(size_t)e * n3 + tt];
#pragma unroll
        for (int v = 0; v < 5; v++) src[v] -= sm * (P.U[(size_t)e * 5 * n3 + v * n3 + tt] - P.spBase[(size_t)e * 5 * n3 + v * n3 + tt]);
    }
}

// Source term + (MODE 1) Williamson 2N or Ketcheson 3-register update + next-stage face states, for runs with CalcSource
// (dg.f90:418) or a 3-register scheme: the volume
// kernels then run in MODE 0 and leave Ut = -sJ * (DG operator); here Ut += Ut_src (the reference adds Ut_src / sJ before
// the Jacobian is applied, exactfunc.f90:1109), then vector.f90:163-183 and the face extraction of the fused epilogue.
template <int n, int NT, int MODE>
__global__ void __launch_bounds__(n* n* n) k_source_rk(const KParams P, double t, double mRKA, double b_dt_in) {
    const double b_dt = P.dtDev ? b_dt_in * __ldg(P.dtDev + 3) : b_dt_in;  // device-paced stepping: b_dt_in carries RKb, dt lives on the device
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ double smem[];
    double *tile = smem, *sLm = smem + 5 * n3, *sLp = sLm + n;
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int tt = threadIdx.x;
    if (tt < n) { sLm[tt] = P.L_Minus[tt]; sLp[tt] = P.L_Plus[tt]; }
    double src[5];
    source_terms<n>(P, e, tt, t, src);
    double* Utg = P.Ut + (size_t)e * 5 * n3 + tt;
#pragma unroll
    for (int v = 0; v < 5; v++) {
        // Ut holds sJ * (-R); the reference forms (-R + src / sJ) * sJ = sJ * (-R) + src up to one rounding
        const double ut = Utg[v * n3] + src[v];
        if (MODE == 0) {
            Utg[v * n3] = ut;
        } else {
            double* ou = P.U + (size_t)e * 5 * n3 + tt;
            double un;
            if (P.rk3) {
                // S1 == U, S2, S3 == UPrev (timestep.f90:176-185)
                double* s2 = P.rk3S2 + (size_t)e * 5 * n3 + tt;
                double* up = P.rk3UPrev + (size_t)e * 5 * n3 + tt;
                const double u = ou[v * n3];
                if (P.rk3 == 1) {
                    up[v * n3] = u;
                    s2[v * n3] = u;
                    un = u;
                } else {
                    const double sn = s2[v * n3] + u * P.rk3Delta;
                    s2[v * n3] = sn;
                    un = u * P.rk3G1 + sn * P.rk3G2;
                    un = un + up[v * n3] * P.rk3G3;
                }
                un = un + ut * b_dt;
            } else {
                double* ot = P.Ut_tmp + (size_t)e * 5 * n3 + tt;
                const double r = (mRKA == 0.0) ? ut : ot[v * n3] * mRKA + ut;
                ot[v * n3] = r;
                un = ou[v * n3] + r * b_dt;
            }
            ou[v * n3] = un;
            tile[v * n3 + tt] = un;
        }
    }
    if (MODE == 1) {
        __syncthreads();
        extract_faces<n, NT, 5>(tile, P.UmNext, P.UsNext, P.E2S + 18 * e, P.S2V2, sLm, sLp);
    }
}
template <int n>
constexpr size_t source_smem_bytes() { return sizeof(double) * (5 * n * n * n + 2 * n); }

// Step 14 of the RHS with overintegration (dg/dg.f90:311-328 list, dg/overintegration.f90:179-340; host FLEXI code -- GALAEXI's
// GPU build stops in InitOverintegration): Ut arrives as -(DG operator) WITHOUT the Jacobian (KParams::noJac); the source terms of
// step 13 are added in reference space (src / sJ), then
//   type 1: Filter(Ut, OverintegrationMat) along xi, eta, zeta (changeBasis.t90:287-360) and ApplyJacobian;
//   type 2: FilterConservative (:212-339): J U_t projected to NUnder, times sJNUnder, interpolated back to N.
// The stage update and the face extraction follow in k_source_rk (launched with the sources switched off).
template <int n>
__global__ void __launch_bounds__(n* n* n) k_overint(const KParams P, double t) {
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ double smem[];
    double *b0 = smem, *b1 = smem + 5 * n3, *sA = smem + 10 * n3, *sB = sA + n * n;
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int tt = threadIdx.x;
    const int nu = P.nUnder + 1;
    if (P.overint == 1) {
        for (int x = tt; x < n * n; x += n3) sA[x] = P.oiMat[x];
    } else {
        for (int x = tt; x < nu * n; x += n3) { sA[x] = P.oiDown[x]; sB[x] = P.oiUp[x]; }
    }
    double* Ut = P.Ut + (size_t)e * 5 * n3;
    const double sJ = P.sJ[(size_t)e * n3 + tt];
    {
        double src[5];
        source_terms<n>(P, e, tt, t, src);
#pragma unroll
        for (int v = 0; v < 5; v++) b0[v * n3 + tt] = Ut[v * n3 + tt] + src[v] / sJ;
    }
    __syncthreads();
    const int k = tt / n2, j = (tt - k * n2) / n, i = tt - k * n2 - j * n;
    if (P.overint == 1) {
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = 0.0;
            for (int l = 0; l < n; l++) a += sA[i + n * l] * b0[v * n3 + l + n * j + n2 * k];
            b1[v * n3 + tt] = a;
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = 0.0;
            for (int l = 0; l < n; l++) a += sA[j + n * l] * b1[v * n3 + i + n * l + n2 * k];
            b0[v * n3 + tt] = a;
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = 0.0;
            for (int l = 0; l < n; l++) a += sA[k + n * l] * b0[v * n3 + i + n * j + n2 * l];
            Ut[v * n3 + tt] = a * sJ;   // applyjacobian.t90:196
        }
        return;
    }
    // conservative cut-off. Down: oiDown(iU,i) at [iU + nu*i]; buffers indexed [v][a + na*(b + nb*c)] with the current extents
    if (tt < nu * n2) {  // xi: (iU, j, k)
        const int kk = tt / (nu * n), jj = (tt - kk * nu * n) / nu, iU = tt - kk * nu * n - jj * nu;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = sA[iU] * b0[v * n3 + n * jj + n2 * kk];
            for (int l = 1; l < n; l++) a += sA[iU + nu * l] * b0[v * n3 + l + n * jj + n2 * kk];
            b1[v * n3 + iU + nu * (jj + n * kk)] = a;
        }
    }
    __syncthreads();
    if (tt < nu * nu * n) {  // eta: (iU, jU, k)
        const int kk = tt / (nu * nu), jU = (tt - kk * nu * nu) / nu, iU = tt - kk * nu * nu - jU * nu;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = sA[jU] * b1[v * n3 + iU + nu * (0 + n * kk)];
            for (int l = 1; l < n; l++) a += sA[jU + nu * l] * b1[v * n3 + iU + nu * (l + n * kk)];
            b0[v * n3 + iU + nu * (jU + nu * kk)] = a;
        }
    }
    __syncthreads();
    if (tt < nu * nu * nu) {  // zeta: (iU, jU, kU), then the Jacobian of NUnder
        const int kU = tt / (nu * nu), jU = (tt - kU * nu * nu) / nu, iU = tt - kU * nu * nu - jU * nu;
        const double sJu = P.sJNUnder[(size_t)e * nu * nu * nu + tt];
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = 0.0;
            for (int l = 0; l < n; l++) a += sA[kU + nu * l] * b0[v * n3 + iU + nu * (jU + nu * l)];
            b1[v * n3 + iU + nu * (jU + nu * kU)] = a * sJu;
        }
    }
    __syncthreads();
    // Up: oiUp(i,iU) at [i + n*iU]
    if (tt < n * nu * nu) {  // xi: (i, jU, kU)
        const int kU = tt / (n * nu), jU = (tt - kU * n * nu) / n, ii = tt - kU * n * nu - jU * n;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = sB[ii] * b1[v * n3 + 0 + nu * (jU + nu * kU)];
            for (int l = 1; l < nu; l++) a += sB[ii + n * l] * b1[v * n3 + l + nu * (jU + nu * kU)];
            b0[v * n3 + ii + n * (jU + nu * kU)] = a;
        }
    }
    __syncthreads();
    if (tt < n2 * nu) {  // eta: (i, j, kU)
        const int kU = tt / n2, jj = (tt - kU * n2) / n, ii = tt - kU * n2 - jj * n;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double a = sB[jj] * b0[v * n3 + ii + n * (0 + nu * kU)];
            for (int l = 1; l < nu; l++) a += sB[jj + n * l] * b0[v * n3 + ii + n * (l + nu * kU)];
            b1[v * n3 + ii + n * (jj + n * kU)] = a;
        }
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 5; v++) {  // zeta: (i, j, k)
        double a = 0.0;
        for (int l = 0; l < nu; l++) a += sB[k + n * l] * b1[v * n3 + i + n * (j + n * l)];
        Ut[v * n3 + tt] = a;
    }
}
template <int n>
constexpr size_t overint_smem_bytes() { return sizeof(double) * (10 * n * n * n + 2 * n * n); }

template <int n>
constexpr size_t filter_smem_bytes() { return sizeof(double) * (10 * n * n * n + n * n + 2 * n); }


// ---------------------------------------------------------------------------------------------------------
// CalcTimeStep at one node (equations/navierstokes/calctimestep.f90:192-296): the three convective and the three viscous
// eigenvalue bounds lam[0..5]; returns true when the state is not admissible. M: the nine metric terms of the node.
__device__ __forceinline__ bool timestep_node(const Eos& eos, bool parabolic, const double (&Uc)[5], const double (&M)[9], double sJ, double (&lam)[6]) {
    double Pr[6];
    cons_to_prim(Pr, Uc, eos);
    const double c = sqrt(eos.kappa * Pr[PRES] / Uc[DENS]);
    const bool bad = !(Uc[DENS] > 0.0) || !(Pr[PRES] > 0.0) || !isfinite(Uc[ENER]);
    const double kmax = fmax(4.0 / 3.0, eos.kappa / eos.Pr);
    const double mu = parabolic ? viscosity(eos, Pr[TEMP]) : 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double m0 = M[3 * d], m1 = M[3 * d + 1], m2 = M[3 * d + 2];
        const double nrm2 = m0 * m0 + m1 * m1 + m2 * m2;
        lam[d] = fabs(m0 * (Pr[VEL1] * sJ) + m1 * (Pr[VEL2] * sJ) + m2 * (Pr[VEL3] * sJ)) + c * (sJ * sqrt(nrm2));
        lam[3 + d] = mu / Uc[DENS] * (kmax * (nrm2 * sJ * sJ));
    }
    return bad;
}
// positive doubles order like their bit patterns
__device__ __forceinline__ void atomic_min_pos(double* addr, double v) {
    atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
// Element maxima of lam[0..5] over the NTHR threads of the block (max is exact in any order), then the element's time step
// bounds into the global minima out[0] (convective) / out[1] (viscous); errFlag bit 2 for inadmissible states
// (calctimestep.f90:134-186). Every thread of the block must call it; the last warp may be partial.
template <int NTHR>
__device__ __forceinline__ void timestep_block_min(const double (&lam)[6], bool bad, double (*red)[32], const KParams& P, double CFL, double DFL,
                                                   double* out) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nAct = (NTHR - 32 * w) < 32 ? (NTHR - 32 * w) : 32;
    const unsigned mask = nAct == 32 ? 0xffffffffu : ((1u << nAct) - 1u);
#pragma unroll
    for (int x = 0; x < 6; x++) {
        double v = lam[x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_down_sync(mask, v, o);
            if (lane + o < nAct) v = fmax(v, u);
        }
        if (lane == 0) red[x][w] = v;
    }
    const int anyBad = __syncthreads_or(bad ? 1 : 0);
    if (t == 0) {
        constexpr int nw = (NTHR + 31) / 32;
        double mx[6];
#pragma unroll
        for (int x = 0; x < 6; x++) {
            double v = red[x][0];
            for (int ww = 1; ww < nw; ww++) v = fmax(v, red[x][ww]);
            mx[x] = v;
        }
        const double lc = mx[0] + mx[1] + mx[2];
        const double dtc = CFL * 2.0 / lc;
        if (anyBad || !(dtc > 0.0) || !isfinite(dtc)) atomicOr(P.errFlag, 2);
        else atomic_min_pos(&out[0], dtc);
        if (P.parabolic) {
            const double lv = mx[3] + mx[4] + mx[5];
            const double dtv = DFL * 4.0 / lv;
            if (dtv > 0.0 && isfinite(dtv)) atomic_min_pos(&out[1], dtv);
        }
    }
}
// End of CalcTimeStep in device-paced stepping: dt = min(convective, viscous) into acc[3] and the history, accumulators
// re-armed for the next step; acc[2] < 0: another rank found an inadmissible state (after the min-reduction over the ranks)
constexpr double DT_HUGE = 1.7976931348623157e308;
static __global__ void k_dt_pre(double* acc, const int* errFlag) { acc[2] = (*errFlag & 2) ? -1.0 : 0.0; }
static __global__ void k_dt_finish(double* acc, int* errFlag, double* hist, int cap) {
    if (acc[2] < 0.0) atomicOr(errFlag, 2);
    const double dt = acc[0] < acc[1] ? acc[0] : acc[1];
    acc[3] = dt;
    const int slot = (int)acc[4];  // steps finished since dgx_run_steps armed the accumulators
    if (hist && slot < cap) hist[slot] = dt;
    acc[4] = (double)(slot + 1);
    acc[0] = DT_HUGE; acc[1] = DT_HUGE; acc[2] = 0.0;
}
static __global__ void k_bulk_finish(double* tot, double Vol) { tot[1] = tot[0] / Vol; }

// ---------------------------------------------------------------------------------------------------------
// BR1 lifting (strong form, non-conservative volume integral): gradU = sJ * ( M . D U + sum_faces F n Lhat )
// n == 8 (N=7): the three derivative sweeps of the lifting run on the FP64 tensor-core path (mma.sync m8n8k4, DMMA): the
// same 36 TFLOP/s as DFMA on a B200 (tools/microbench/fp64_rates.cu) but 8x fewer issue slots and 10x fewer shared-memory
// loads per flop -- the sweeps of the thread-per-node form were bound by LDS issue (profiles/r02a: mio_throttle).
template <int n>
constexpr bool lifting_uses_dmma() { return n == 8; }
// even n without the DMMA path: metrics / Jacobian staged by TMA (the DMMA path needs the shared memory for its 12 output
// slots and prefetches them to L2 instead)
template <int n>
constexpr bool lifting_uses_tma() { return n % 2 == 0 && !lifting_uses_dmma<n>(); }
template <int n>
constexpr int lifting_tile_slots() { return lifting_uses_dmma<n>() ? 16 : 12; }

// Tile layout of the DMMA path (n = 8): conflict-free 64-bit accesses for the point-wise lane mapping and for the B
// fragments of all three sweep directions (4 consecutive l x 4 consecutive lines per half warp)
__device__ __forceinline__ int idx_m8(int i, int j, int k) { return ((i ^ (4 * (((j >> 1) ^ k) & 1))) + 8 * (j ^ ((k >> 1) & 1))) + 64 * k; }
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// GEN=0: BR1 on conforming meshes (the GALAEXI configuration, hot path). GEN=1: BR2 and/or elements with a big mortar
// face, whose projected lifting flux (times normal) k_mortar_liftflux has left in gm[bigSide].
// resident CTAs per SM the register allocation aims at: about 1024 threads per SM (64 registers per thread)
template <int n>
constexpr int lifting_min_blocks() {
    constexpr int warps = (n * n * n + 31) / 32;
    return warps >= 32 ? 1 : (32 / warps > 16 ? 16 : 32 / warps);
}
template <int n, int NT, int GEN = 0>
__global__ void __launch_bounds__(n* n* n, lifting_min_blocks<n>()) k_lifting(const KParams P, int lookahead) {
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ __align__(16) double smem[];
    __shared__ int sMort[6];  // GEN: 0-based big mortar side of local side loc, or -1
    __shared__ double sSig[6];  // GEN: sign of the face in the surface integral (weak form: -1 on slave faces, surfint.t90:640-644)
    __shared__ __align__(8) unsigned long long sBar;  // mbarrier of the metric / Jacobian bulk copy
    __shared__ double sRed[6][32];  // dtFuse: element maxima of the time step eigenvalue bounds
    const bool br2 = GEN && P.lifting == 2;
    // even n: the element's metrics (9 n^3) and Jacobian (n^3) are fetched by TMA bulk copies issued at kernel entry into a
    // staging area behind the operator tables, so this read is in flight from the first cycle of the CTA instead of
    // starting after the sweeps (16-byte granularity of cp.async.bulk: n^3 * 8 B must be a multiple of 16)
    constexpr bool TMA = lifting_uses_tma<n>();
    constexpr bool DMMA = lifting_uses_dmma<n>();
    double* sT = smem;                 // [4][n3] lifting variables; later aliased by the gradient tile [12][n3]
    double* sG = smem;                 // alias (used after the sweeps are done)
    double* sO = smem + 4 * n3;        // DMMA path: [12][n3] derivatives (dir*4+v) in reference space
    double* sF = smem + lifting_tile_slots<n>() * n3;  // [6][7][n2] face lifting flux (4) + normal (3), element face order
    double* sD = sF + 6 * 7 * n2;      // D_T [n*n]
    double* sDx = sD + n * n;          // D_T transposed
    double* sLhm = sDx + n * n;
    double* sLhp = sLhm + n;
    double* sLm = sLhp + n;
    double* sLp = sLm + n;
    double* sDh = sLp + n;             // D_Hat_T [l + n o] (viscous volume integral)
    double* sDhx = sDh + n * n;        // its transpose [o + n l] (lane-varying o)
    double* stM = sDhx + n * n;        // TMA staging: metrics [9][n3], then sJ [n3]
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int t = threadIdx.x;
    if (TMA && t == 0) {
        const unsigned bar = smem_u32(&sBar);
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (unsigned)(10 * n3 * sizeof(double)));
        tma_load_1d(smem_u32(stM), P.metrics + (size_t)e * 9 * n3, (unsigned)(9 * n3 * sizeof(double)), bar);
        tma_load_1d(smem_u32(stM + 9 * n3), P.sJ + (size_t)e * n3, (unsigned)(n3 * sizeof(double)), bar);
    }
    for (int x = t; x < n * n; x += n3) {
        sD[x] = P.D_T[x]; sDx[(x / n) + n * (x % n)] = P.D_T[x];
        sDh[x] = P.D_Hat_T[x]; sDhx[(x / n) + n * (x % n)] = P.D_Hat_T[x];
    }
    if (t < n) { sLhm[t] = P.L_HatMinus[t]; sLhp[t] = P.L_HatPlus[t]; sLm[t] = P.L_Minus[t]; sLp[t] = P.L_Plus[t]; }
    const Eos eos = P.eos;
    const int* e2s = P.E2S + 18 * e;
    const int k = t / n2, j = (t - k * n2) / n, i = t - k * n2 - j * n;
    const int tid_ = Tile<n>::idx(i, j, k);
    // L2 prefetches (fire and forget, no registers): the metrics / Jacobian this CTA reads two barriers from now, and the
    // volume data of the element that takes this CTA's place once it retires (one resident wave ahead in the grid)
    if (!TMA && (DMMA || !(P.flags & 16))) {  // on by default (N=4 Gauss: 0.749 -> 0.730 ms); DGX_FLAGS bit 16 switches it off
        prefetch_block(P.metrics + (size_t)e * 9 * n3, sizeof(double) * 9 * n3, t, n3);
        prefetch_block(P.sJ + (size_t)e * n3, sizeof(double) * n3, t, n3);
    }
    if ((P.flags & 2) && (int)blockIdx.x + lookahead < (P.elemList ? P.nList : P.nElems)) {
        const int en = P.elemList ? P.elemList[blockIdx.x + lookahead] : (int)blockIdx.x + lookahead;
        prefetch_block(P.U + (size_t)en * 5 * n3, sizeof(double) * 5 * n3, t, n3);
        prefetch_block(P.metrics + (size_t)en * 9 * n3, sizeof(double) * 9 * n3, t, n3);
        prefetch_block(P.sJ + (size_t)en * n3, sizeof(double) * n3, t, n3);
    }
    // 1. node: primitive lifting variables into the tile
    {
        const double* U = P.U + (size_t)e * 5 * n3;
        double Uc[5], Pr[6];
#pragma unroll
        for (int v = 0; v < 5; v++) Uc[v] = U[v * n3 + t];
        cons_to_prim(Pr, Uc, eos);
        const int tin = DMMA ? idx_m8(i, j, k) : tid_;
        sT[0 * n3 + tin] = Pr[VEL1];
        sT[1 * n3 + tin] = Pr[VEL2];
        sT[2 * n3 + tin] = Pr[VEL3];
        sT[3 * n3 + tin] = Pr[TEMP];
        if (P.dtFuse) {
            // CalcTimeStep of the state this stage reads (= the state at the start of the time step, calctimestep.f90:98-186)
            // while it is in registers: the node's metrics and Jacobian are fetched here already (the same DRAM round trip as
            // the state; step 3 then finds them in L1 / L2 or in the TMA staging area), which saves k_timestep's separate pass
            // over U, metrics and sJ (device-paced stepping, dgx_run_steps)
            if (TMA) { __syncthreads(); mbar_wait(smem_u32(&sBar), 0); }
            double M9[9], lam[6];
            const double* M = TMA ? stM + t : P.metrics + (size_t)e * 9 * n3 + t;
#pragma unroll
            for (int x = 0; x < 9; x++) M9[x] = M[x * n3];
            const double sJ = TMA ? stM[9 * n3 + t] : P.sJ[(size_t)e * n3 + t];
            const bool bad = timestep_node(eos, true, Uc, M9, sJ, lam);
            timestep_block_min<n3>(lam, bad, sRed, P, P.dtCFL, P.dtDFL, P.dtAcc);
        }
    }
    // 2. faces: lifting flux F = 1/2 (U_s - U_m) SurfElem in side orientation -> stored in element face order
    for (int f = t; f < 6 * n2; f += n3) {
        const int loc = f / n2 + 1;
        const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
        const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
        int p, q;
        face_lane<n>(P.S2V2, f - (loc - 1) * n2, flip, loc, false, p, q);  // a fastest: contiguous writes to sF
        const int pq = p + n * q;
        const int a = s2v2<n>(P.S2V2, 0, p, q, flip, loc);
        const int b = s2v2<n>(P.S2V2, 1, p, q, flip, loc);
        if (GEN) {
            const bool big = P.MortarType && __ldg(&P.MortarType[2 * side]) > 0;
            if (f == (loc - 1) * n2) { sMort[loc - 1] = big ? side : -1; sSig[loc - 1] = (P.liftWeak && flip != 0) ? -1.0 : 1.0; }
            if (big) continue;  // flux comes projected from the small sides
        }
        const double* g = P.geo + (size_t)side * 10 * n2 + pq;
        const double nv[3] = {g[0 * n2], g[1 * n2], g[2 * n2]};
        const double se = g[9 * n2];
        double Um[5], Pm[6], Fl[4];
#pragma unroll
        for (int v = 0; v < 5; v++) Um[v] = P.Um[(size_t)side * 5 * n2 + v * n2 + pq];
        cons_to_prim(Pm, Um, eos);
        if (side < P.nBCSides) {
            const int bct = __ldg(&P.BCSides[2 * side]), bcs = __ldg(&P.BCSides[2 * side + 1]);
            const double t1[3] = {g[3 * n2], g[4 * n2], g[5 * n2]};
            const double t2[3] = {g[6 * n2], g[7 * n2], g[8 * n2]};
            double Pb[6];
            if (boundary_state(bct, eos, Pb, Pm, P.RefPrim + 6 * (bcs > 0 ? bcs - 1 : 0), nv, t1, t2)) atomicOr(P.errFlag, 1);
            if (is_riemann_bc(bct)) {
                Fl[0] = 0.5 * (Pm[VEL1] + Pb[VEL1]); Fl[1] = 0.5 * (Pm[VEL2] + Pb[VEL2]);
                Fl[2] = 0.5 * (Pm[VEL3] + Pb[VEL3]); Fl[3] = 0.5 * (Pm[TEMP] + Pb[TEMP]);
            } else if (bct == 3 || bct == 4) {
                Fl[0] = Fl[1] = Fl[2] = 0.0; Fl[3] = Pb[TEMP];
            } else {
                Fl[0] = Pb[VEL1]; Fl[1] = Pb[VEL2]; Fl[2] = Pb[VEL3]; Fl[3] = Pm[TEMP];
            }
            if (GEN && P.liftWeak) {  // weak form: the inner state is not subtracted (getboundaryflux.f90:1014)
                Fl[0] *= se; Fl[1] *= se; Fl[2] *= se; Fl[3] *= se;
            } else {
                Fl[0] = (Fl[0] - Pm[VEL1]) * se; Fl[1] = (Fl[1] - Pm[VEL2]) * se;
                Fl[2] = (Fl[2] - Pm[VEL3]) * se; Fl[3] = (Fl[3] - Pm[TEMP]) * se;
            }
        } else {
            double Us[5], Ps[6];
#pragma unroll
            for (int v = 0; v < 5; v++) Us[v] = P.Us[(size_t)side * 5 * n2 + v * n2 + pq];
            cons_to_prim(Ps, Us, eos);
            const double sig = (GEN && P.liftWeak) ? 1.0 : -1.0;  // lifting_fillflux.t90:78
            Fl[0] = 0.5 * se * (sig * Pm[VEL1] + Ps[VEL1]); Fl[1] = 0.5 * se * (sig * Pm[VEL2] + Ps[VEL2]);
            Fl[2] = 0.5 * se * (sig * Pm[VEL3] + Ps[VEL3]); Fl[3] = 0.5 * se * (sig * Pm[TEMP] + Ps[TEMP]);
        }
        double* d = sF + (loc - 1) * 7 * n2 + (b * n + a);
        d[0 * n2] = Fl[0]; d[1 * n2] = Fl[1]; d[2 * n2] = Fl[2]; d[3 * n2] = Fl[3];
        d[4 * n2] = nv[0]; d[5 * n2] = nv[1]; d[6 * n2] = nv[2];
    }
    __syncthreads();
    // 3. node: volume derivative + surface lifting, Jacobian
    double G[12], fv[12];
    {
        double gxi[4] = {0, 0, 0, 0}, get[4] = {0, 0, 0, 0}, gze[4] = {0, 0, 0, 0};
        if constexpr (DMMA) {
            // out(o, line) = sum_l D_T(l,o) u(l, line) as C[8x8] = A[8x8] B[8x8] in two k-steps of m8n8k4:
            // A[row o][l] = D_T(l,o) (the same fragment for the three directions), B[l][col] = u on 8 lines, one task =
            // (direction, variable, group of 8 lines); 96 tasks over the 16 warps. Fragments: A: row = lane/4, k = lane%4;
            // B: k = lane%4, col = lane/4; C: row = lane/4, cols 2 (lane%4) + {0,1}.
            // 96 tasks over the 16 warps: task = warp + 16 it, so that the direction is a compile-time value of the unrolled
            // loop and the line group g = warp & 7 is fixed per warp (all tile indices are loop invariants)
            const int lane = t & 31, r = lane >> 2, c4 = lane & 3;
            const int g = (t >> 5) & 7, vh = t >> 8;
            const double a0 = sD[c4 + n * r], a1 = sD[c4 + 4 + n * r];
#pragma unroll
            for (int it = 0; it < 6; it++) {
                const int dir = it >> 1, v = vh + 2 * (it & 1);
                int ib0, ib1, ic0, ic1;
                if (dir == 0) {         // xi: lines (j = col, k = g)
                    ib0 = idx_m8(c4, r, g); ib1 = idx_m8(c4 + 4, r, g);
                    ic0 = idx_m8(r, 2 * c4, g); ic1 = idx_m8(r, 2 * c4 + 1, g);
                } else if (dir == 1) {  // eta: lines (i = col, k = g)
                    ib0 = idx_m8(r, c4, g); ib1 = idx_m8(r, c4 + 4, g);
                    ic0 = idx_m8(2 * c4, r, g); ic1 = idx_m8(2 * c4 + 1, r, g);
                } else {                // zeta: lines (i = col, j = g)
                    ib0 = idx_m8(r, g, c4); ib1 = idx_m8(r, g, c4 + 4);
                    ic0 = idx_m8(2 * c4, g, r); ic1 = idx_m8(2 * c4 + 1, g, r);
                }
                const double b0 = sT[v * n3 + ib0], b1 = sT[v * n3 + ib1];
                double c0 = 0.0, c1 = 0.0;
                dmma_m8n8k4(c0, c1, a0, b0);
                dmma_m8n8k4(c0, c1, a1, b1);
                sO[(dir * 4 + v) * n3 + ic0] = c0;
                sO[(dir * 4 + v) * n3 + ic1] = c1;
            }
            __syncthreads();
            const int tm = idx_m8(i, j, k);
#pragma unroll
            for (int v = 0; v < 4; v++) { gxi[v] = sO[v * n3 + tm]; get[v] = sO[(4 + v) * n3 + tm]; gze[v] = sO[(8 + v) * n3 + tm]; }
        } else {
#pragma unroll
            for (int l = 0; l < n; l++) {
                const double dx = sDx[i + n * l], dy = sD[l + n * j], dz = sD[l + n * k];  // sDx: transposed copy, conflict-free over i
                const int ix = Tile<n>::idx(l, j, k), iy = Tile<n>::idx(i, l, k), iz = Tile<n>::idx(i, j, l);
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    gxi[v] += dx * sT[v * n3 + ix];
                    get[v] += dy * sT[v * n3 + iy];
                    gze[v] += dz * sT[v * n3 + iz];
                }
            }
        }
        if (TMA) mbar_wait(smem_u32(&sBar), 0);  // (the barrier after step 2 ordered the init before this wait)
        const double* M = TMA ? stM + t : P.metrics + (size_t)e * 9 * n3 + t;
        double S[12];
#pragma unroll
        for (int x = 0; x < 12; x++) S[x] = 0.0;
        // surface contributions in local-side order 1..6 (surfint.t90:624-716)
#pragma unroll
        for (int loc = 1; loc <= 6; loc++) {
            int a, b, l;
            if (loc == XI_MINUS || loc == XI_PLUS) { a = j; b = k; l = i; }
            else if (loc == ETA_MINUS || loc == ETA_PLUS) { a = i; b = k; l = j; }
            else { a = i; b = j; l = k; }
            const bool minus = is_minus(loc);
            double Lh;
            if (NT == 2) {
                if (minus ? (l != 0) : (l != n - 1)) continue;
                Lh = minus ? sLhm[0] : sLhp[n - 1];
            } else {
                Lh = minus ? sLhm[l] : sLhp[l];
            }
            const double sJn = GEN ? (TMA ? stM[9 * n3 + t] : P.sJ[(size_t)e * n3 + t]) : 1.0;  // BR2: F_loc = sJ * Flux * L_hat (lifting_br2.t90:252,293)
            if (GEN && sMort[loc - 1] >= 0) {
                // big mortar face: projected flux*normal in side-local (flip 0) node order
                const int p = s2v2<n>(P.S2V2inv, 0, a, b, 0, loc), q = s2v2<n>(P.S2V2inv, 1, a, b, 0, loc);
                const double* gF = P.gm + (size_t)sMort[loc - 1] * 12 * n2 + (p + n * q);
#pragma unroll
                for (int x = 0; x < 12; x++) S[x] += br2 ? (sJn * gF[x * n2]) * Lh : gF[x * n2] * Lh;
                continue;
            }
            const double* s = sF + (loc - 1) * 7 * n2 + (b * n + a);
            double nx = s[4 * n2], ny = s[5 * n2], nz = s[6 * n2];
            if (GEN) { const double fsig = sSig[loc - 1]; nx *= fsig; ny *= fsig; nz *= fsig; }
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double F = s[v * n2];
                if (br2) {
                    S[0 * 4 + v] += (sJn * (F * nx)) * Lh;
                    S[1 * 4 + v] += (sJn * (F * ny)) * Lh;
                    S[2 * 4 + v] += (sJn * (F * nz)) * Lh;
                } else {
                    S[0 * 4 + v] += (F * nx) * Lh;
                    S[1 * 4 + v] += (F * ny) * Lh;
                    S[2 * 4 + v] += (F * nz) * Lh;
                }
            }
        }
        const double sJ = TMA ? stM[9 * n3 + t] : P.sJ[(size_t)e * n3 + t];
        double V[12];
        if (GEN && P.liftCons) {
            // conservative volume form (lifting_volint.t90:125-200): the metrics sit inside the derivative; DMat = D_Hat_T
            // for the weak form, D_T otherwise. Metrics of the line nodes come from the TMA staging area or from L2.
            const double* __restrict__ Dm = P.liftWeak ? P.D_Hat_T : P.D_T;
            const double* __restrict__ Mb = TMA ? stM : P.metrics + (size_t)e * 9 * n3;
#pragma unroll
            for (int x = 0; x < 12; x++) V[x] = 0.0;
            for (int l = 0; l < n; l++) {
                const double dx = Dm[l + n * i], dy = Dm[l + n * j], dz = Dm[l + n * k];
                const int nxn = l + n * j + n2 * k, nyn = i + n * l + n2 * k, nzn = i + n * j + n2 * l;
                const int ix = DMMA ? idx_m8(l, j, k) : Tile<n>::idx(l, j, k), iy = DMMA ? idx_m8(i, l, k) : Tile<n>::idx(i, l, k),
                          iz = DMMA ? idx_m8(i, j, l) : Tile<n>::idx(i, j, l);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double mfx = Mb[(0 + d) * n3 + nxn], mgy = Mb[(3 + d) * n3 + nyn], mhz = Mb[(6 + d) * n3 + nzn];
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        V[d * 4 + v] += dx * (mfx * sT[v * n3 + ix]) + dz * (mhz * sT[v * n3 + iz]) + dy * (mgy * sT[v * n3 + iy]);
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double mf = M[(0 + d) * n3], mg = M[(3 + d) * n3], mh = M[(6 + d) * n3];
#pragma unroll
                for (int v = 0; v < 4; v++) V[d * 4 + v] = mf * gxi[v] + mg * get[v] + mh * gze[v];
            }
        }
#pragma unroll
        for (int x = 0; x < 12; x++) {
            // BR2: volume part times sJ (ApplyJacobianLifting, lifting_br2.t90:118-124), S holds the surface part
            G[x] = br2 ? V[x] * sJ : sJ * (V[x] + S[x]);
        }
        // 3b. transformed viscous fluxes of the node (VolInt_weakForm_Visc, dg/volint.f90:60-119 -> flux.f90:351-384, 619-697)
        // from the volume gradients, while the metrics are at hand; they are swept with D_Hat_T in step 5
        {
            const int tin = DMMA ? idx_m8(i, j, k) : tid_;
            double Pn[6];
            Pn[VEL1] = sT[0 * n3 + tin]; Pn[VEL2] = sT[1 * n3 + tin]; Pn[VEL3] = sT[2 * n3 + tin]; Pn[TEMP] = sT[3 * n3 + tin];
            double gr[12];
#pragma unroll
            for (int x = 0; x < 12; x++) gr[x] = br2 ? G[x] + S[x] : G[x];
            const double mu = viscosity(eos, Pn[TEMP]);
            Tau ta;
            stress(ta, Pn, gr, mu, conductivity(eos, mu));
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double Md[3] = {M[(3 * d + 0) * n3], M[(3 * d + 1) * n3], M[(3 * d + 2) * n3]};
                visc_flux_dir(ta, Md, fv + 4 * d);
            }
        }
        if (br2) {
            __syncthreads();  // all reads of sT done before it is overwritten by the gradient tile
#pragma unroll
            for (int x = 0; x < 12; x++) sG[x * n3 + tid_] = G[x];
            if (P.storeGrad) {
                double* gU = P.gradU + (size_t)e * 12 * n3 + t;
#pragma unroll
                for (int x = 0; x < 12; x++) gU[x * n3] = G[x] + S[x];
            }
        }
    }
    if (!br2) {
        __syncthreads();  // all reads of sT done before it is overwritten by the gradient tile
#pragma unroll
        for (int x = 0; x < 12; x++) sG[x * n3 + tid_] = G[x];
        if (P.storeGrad) {
            double* gU = P.gradU + (size_t)e * 12 * n3 + t;
#pragma unroll
            for (int x = 0; x < 12; x++) gU[x * n3] = G[x];
        }
    }
    __syncthreads();
    // 4. gradients on the faces (ProlongToFaceLifting)
    if (!br2) extract_faces_tile<n, NT, 12>(sG, P.gm, P.gs, e2s, P.S2V2, sLm, sLp);
    // BR2 (lifting_br2.t90:126-137, 193-313): trace of the volume part + eta * sum_l L_Minus(l) F_loc(l), F_loc(l) =
    // sJ(depth l) * Flux * L_HatMinus(l); l = depth from the face (Gauss-Lobatto: l = 0 only)
    if (GEN && br2) {
        constexpr int SL = Tile<n>::SLOT;
        for (int f = t; f < 6 * n2; f += n3) {
            const int loc = f / n2 + 1;
            const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
            const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
            const bool xi = (loc == XI_MINUS || loc == XI_PLUS), eta_ = (loc == ETA_MINUS || loc == ETA_PLUS);
            const bool minus = is_minus(loc);
            int p, q;
            face_lane<n>(P.S2V2, f - (loc - 1) * n2, flip, loc, xi, p, q);
            const int a = s2v2<n>(P.S2V2, 0, p, q, flip, loc);
            const int b = s2v2<n>(P.S2V2, 1, p, q, flip, loc);
            double eta = P.etaBR2;
            if (side < P.nBCSides) {
                const int bct = __ldg(&P.BCSides[2 * side]);
                if (bct == 3 || bct == 4) eta = P.etaBR2_wall;
            }
            // flux times normal at this face node
            double Fd[12];
            if (sMort[loc - 1] >= 0) {
                const double* gF = P.gm + (size_t)side * 12 * n2 + (p + n * q);  // flip == 0 on a big mortar side
#pragma unroll
                for (int x = 0; x < 12; x++) Fd[x] = gF[x * n2];
            } else {
                const double* s = sF + (loc - 1) * 7 * n2 + (b * n + a);
#pragma unroll
                for (int d = 0; d < 3; d++)
#pragma unroll
                    for (int v = 0; v < 4; v++) Fd[d * 4 + v] = s[v * n2] * s[(4 + d) * n2];
            }
            double acc[12];
            // trace of the volume part
            if (NT == 2) {
                const int l = minus ? 0 : n - 1;
                const int id = xi ? Tile<n>::idx(l, a, b) : (eta_ ? Tile<n>::idx(a, l, b) : Tile<n>::idx(a, b, l));
#pragma unroll
                for (int x = 0; x < 12; x++) acc[x] = sG[x * SL + id];
            } else {
                const double* L = minus ? sLm : sLp;
#pragma unroll
                for (int x = 0; x < 12; x++) {
                    double r = 0.0;
#pragma unroll
                    for (int l = 0; l < n; l++) {
                        const int id = xi ? Tile<n>::idx(l, a, b) : (eta_ ? Tile<n>::idx(a, l, b) : Tile<n>::idx(a, b, l));
                        r = (l == 0) ? sG[x * SL + id] * L[0] : r + sG[x * SL + id] * L[l];
                    }
                    acc[x] = r;
                }
            }
            // penalised surface part
            const int lmax = (NT == 2) ? 1 : n;
            for (int l = 0; l < lmax; l++) {
                const int ln = minus ? l : n - 1 - l;  // node index along the face normal at depth l
                const int node = xi ? (ln + n * (a + n * b)) : (eta_ ? (a + n * (ln + n * b)) : (a + n * (b + n * ln)));
                const double sJl = P.sJ[(size_t)e * n3 + node];
                const double w = eta * sLm[l];
#pragma unroll
                for (int x = 0; x < 12; x++) acc[x] += w * ((sJl * Fd[x]) * sLhm[l]);
            }
            double* dst = (flip == 0 ? P.gm : P.gs) + (size_t)side * 12 * n2 + (p + n * q);
#pragma unroll
            for (int x = 0; x < 12; x++) dst[x * n2] = acc[x];
        }
    }
    // 5. viscous volume integral (applydmatrix.t90:19-75 with D_Hat_T): Ut_visc(v) = sum_l D_Hat_T(l,i) f_v(l,j,k) +
    // D_Hat_T(l,k) h_v(i,j,l) + D_Hat_T(l,j) g_v(i,l,k) for the four momentum / energy rows (the density row of the viscous
    // flux is zero). The result goes to Ut(1..4) of the element; the volume kernel takes it as its initial accumulator
    // (volint.f90:238-243: the viscous integral overwrites Ut, the advective one adds), so the 12 gradient components per
    // node never make the round trip through HBM.
    __syncthreads();  // face extraction done with the gradient tile
    double* UtV = P.Ut + (size_t)e * 5 * n3 + n3;
    if constexpr (DMMA) {
        const int tm = idx_m8(i, j, k);
#pragma unroll
        for (int x = 0; x < 12; x++) smem[x * n3 + tm] = fv[x];
        __syncthreads();
        // FP64 tensor-core form (fragments as in step 3). zeta first, as C[i][o=k] = sum_l h(i,g,l) D_Hat_T(l,o) on the planes
        // j = g (A = flux tile, B = operator), parked in slots 12..15; then on the planes k = g the xi product
        // C[o=i][j] += sum_l D_Hat_T(l,o) f(l,j,g) (A = operator) and the eta product C[i][o=j] += sum_l g(i,l,g) D_Hat_T(l,o)
        // (A = flux tile) accumulate on top of it in the same fragment, which is stored straight to global memory.
        const int lane = t & 31, r = lane >> 2, c4 = lane & 3;
        const int g = (t >> 5) & 7, vh = t >> 8;  // 32 tasks per phase over the 16 warps: v = vh + 2 it, plane g = warp & 7
        const double a0 = sDh[c4 + n * r], a1 = sDh[c4 + 4 + n * r];
        {
            const int ih0 = idx_m8(r, g, c4), ih1 = idx_m8(r, g, c4 + 4), io0 = idx_m8(r, g, 2 * c4), io1 = idx_m8(r, g, 2 * c4 + 1);
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int v = vh + 2 * it;
                const double h0 = smem[(8 + v) * n3 + ih0], h1 = smem[(8 + v) * n3 + ih1];
                double c0 = 0.0, c1 = 0.0;
                dmma_m8n8k4(c0, c1, h0, a0);
                dmma_m8n8k4(c0, c1, h1, a1);
                smem[(12 + v) * n3 + io0] = c0;
                smem[(12 + v) * n3 + io1] = c1;
            }
        }
        __syncthreads();
        {
            const int ic0 = idx_m8(r, 2 * c4, g), ic1 = idx_m8(r, 2 * c4 + 1, g);
            const int if0 = idx_m8(c4, r, g), if1 = idx_m8(c4 + 4, r, g), ig0 = idx_m8(r, c4, g), ig1 = idx_m8(r, c4 + 4, g);
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int v = vh + 2 * it;
                double c0 = smem[(12 + v) * n3 + ic0], c1 = smem[(12 + v) * n3 + ic1];
                const double f0 = smem[v * n3 + if0], f1 = smem[v * n3 + if1];
                const double g0 = smem[(4 + v) * n3 + ig0], g1 = smem[(4 + v) * n3 + ig1];
                dmma_m8n8k4(c0, c1, a0, f0);
                dmma_m8n8k4(c0, c1, a1, f1);
                dmma_m8n8k4(c0, c1, g0, a0);
                dmma_m8n8k4(c0, c1, g1, a1);
                double* o = UtV + v * n3 + (r + n * (2 * c4) + n2 * g);
                o[0] = c0;
                o[n] = c1;
            }
        }
    } else {
#pragma unroll
        for (int x = 0; x < 12; x++) smem[x * n3 + tid_] = fv[x];
        __syncthreads();
        double a[4] = {0, 0, 0, 0};
#pragma unroll
        for (int l = 0; l < n; l++) {
            const double dx = sDhx[i + n * l], dy = sDh[l + n * j], dz = sDh[l + n * k];
            const int ix = Tile<n>::idx(l, j, k), iy = Tile<n>::idx(i, l, k), iz = Tile<n>::idx(i, j, l);
#pragma unroll
            for (int v = 0; v < 4; v++) a[v] += dx * smem[v * n3 + ix] + dz * smem[(8 + v) * n3 + iz] + dy * smem[(4 + v) * n3 + iy];
        }
#pragma unroll
        for (int v = 0; v < 4; v++) UtV[v * n3 + t] = a[v];
    }
}

template <int n>
constexpr size_t lifting_smem_bytes() {
    return sizeof(double) * (lifting_tile_slots<n>() * n * n * n + 6 * 7 * n * n + 4 * n * n + 4 * n + (lifting_uses_tma<n>() ? 10 * n * n * n : 0));
}

// ---------------------------------------------------------------------------------------------------------
// Numerical flux on a range of sides [side0, side0+nS): BC flux or Riemann + 1/2(Fv_L+Fv_R).n, times SurfElem
template <int n>
__global__ void __launch_bounds__(128, 4) k_sideflux(const KParams P, int side0, int nS) {
    constexpr int n2 = n * n;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nS * n2) return;
    const int side = side0 + gid / n2;
    const int pq = gid % n2;
    if (side >= P.nBCSides && side < P.firstInner - 1) return;  // big mortar sides: filled by k_fluxmortar
    const Eos eos = P.eos;
    const double* g = P.geo + (size_t)side * 10 * n2 + pq;
    const double nv[3] = {g[0 * n2], g[1 * n2], g[2 * n2]};
    const double t1[3] = {g[3 * n2], g[4 * n2], g[5 * n2]};
    const double t2[3] = {g[6 * n2], g[7 * n2], g[8 * n2]};
    const double se = g[9 * n2];
    double Um[5], Pm[6], F[5];
#pragma unroll
    for (int v = 0; v < 5; v++) Um[v] = P.Um[(size_t)side * 5 * n2 + v * n2 + pq];
    cons_to_prim(Pm, Um, eos);
    double gm[12];
    if (P.parabolic) {
#pragma unroll
        for (int x = 0; x < 12; x++) gm[x] = P.gm[(size_t)side * 12 * n2 + x * n2 + pq];
    }
    if (side < P.nBCSides) {
        const int bct = __ldg(&P.BCSides[2 * side]), bcs = __ldg(&P.BCSides[2 * side + 1]);
        double Pb[6];
        if (boundary_state(bct, eos, Pb, Pm, P.RefPrim + 6 * (bcs > 0 ? bcs - 1 : 0), nv, t1, t2)) atomicOr(P.errFlag, 1);
        if (is_riemann_bc(bct)) {
            double Umc[5], Ubc[5];
            prim_to_cons(Pm, Umc, eos);
            prim_to_cons(Pb, Ubc, eos);
            riemann(P.riemann, P.splitDG, eos.kappa, F, Umc, Ubc, Pm, Pb, nv, t1, t2);
            if (P.parabolic) {
                Tau tL, tR;
                double mu = viscosity(eos, Pm[TEMP]);
                stress(tL, Pm, gm, mu, conductivity(eos, mu));
                mu = viscosity(eos, Pb[TEMP]);
                stress(tR, Pb, gm, mu, conductivity(eos, mu));
                double fL[4], fR[4];
                visc_flux_dir(tL, nv, fL);
                visc_flux_dir(tR, nv, fR);
#pragma unroll
                for (int v = 0; v < 4; v++) F[1 + v] += 0.5 * (fL[v] + fR[v]);
            }
        } else {
            F[DENS] = 0.0;
            F[MOM1] = Pb[PRES] * nv[0]; F[MOM2] = Pb[PRES] * nv[1]; F[MOM3] = Pb[PRES] * nv[2];
            F[ENER] = 0.0;
            if (P.parabolic) {
                const double mu = viscosity(eos, Pb[TEMP]);
                const double la = conductivity(eos, mu);
                Tau tb;
                double fd[4];
                if (bct == 9) {
                    double gf[12];
                    const double B[3][3] = {{1.0 - nv[0] * nv[0], -nv[0] * nv[1], -nv[0] * nv[2]},
                                            {-nv[0] * nv[1], 1.0 - nv[1] * nv[1], -nv[2] * nv[1]},
                                            {-nv[0] * nv[2], -nv[2] * nv[1], 1.0 - nv[2] * nv[2]}};
#pragma unroll
                    for (int d = 0; d < 3; d++)
#pragma unroll
                        for (int v = 0; v < 4; v++) gf[d * 4 + v] = B[d][0] * gm[0 * 4 + v] + B[d][1] * gm[1 * 4 + v] + B[d][2] * gm[2 * 4 + v];
                    stress(tb, Pb, gf, mu, la);
                } else if (bct == 91) {
                    double gf[12];
                    slip_wall_gradients_91(gm, nv, t1, t2, gf);
                    stress(tb, Pb, gf, mu, la);
                } else {
                    stress(tb, Pb, gm, mu, la);
                }
                visc_flux_dir(tb, nv, fd);
                if (bct == 3) {  // adiabatic wall: no energy flux (getboundaryflux.f90:664-668)
                    fd[3] = 0.0;
                }
#pragma unroll
                for (int v = 0; v < 4; v++) F[1 + v] += fd[v];
            }
        }
    } else {
        double Us[5], Ps[6];
#pragma unroll
        for (int v = 0; v < 5; v++) Us[v] = P.Us[(size_t)side * 5 * n2 + v * n2 + pq];
        cons_to_prim(Ps, Us, eos);
        riemann(P.riemann, P.splitDG, eos.kappa, F, Um, Us, Pm, Ps, nv, t1, t2);
        if (P.parabolic) {
            double gs[12];
#pragma unroll
            for (int x = 0; x < 12; x++) gs[x] = P.gs[(size_t)side * 12 * n2 + x * n2 + pq];
            Tau tL, tR;
            double mu = viscosity(eos, Pm[TEMP]);
            stress(tL, Pm, gm, mu, conductivity(eos, mu));
            mu = viscosity(eos, Ps[TEMP]);
            stress(tR, Ps, gs, mu, conductivity(eos, mu));
            double fL[4], fR[4];
            visc_flux_dir(tL, nv, fL);
            visc_flux_dir(tR, nv, fR);
#pragma unroll
            for (int v = 0; v < 4; v++) { F[1 + v] += 0.5 * fL[v]; F[1 + v] += 0.5 * fR[v]; }
        }
    }
    double* out = P.Flux + (size_t)side * 5 * n2 + pq;
#pragma unroll
    for (int v = 0; v < 5; v++) out[v * n2] = F[v] * se;
}

// ---------------------------------------------------------------------------------------------------------
// Volume integral (weak or split form, + viscous weak form), surface integral, sign, Jacobian,
// optional RK 2N update and prolongation of the updated state to the faces.
//   MODE 0: store Ut only (DGTimeDerivative_weakForm);  MODE 1: RK stage update (+ face extraction)
// thread-per-node kernel of the weak form / Gauss paths: two resident CTAs at n = 6 (without the bound ptxas takes 158
// registers and a single 216-thread CTA fits an SM); smaller n already fit 4+ CTAs, larger n would spill
template <int n>
constexpr int volsurf_min_blocks() { return n == 6 ? 2 : 1; }
template <int n, int NT, int MODE>
__global__ void __launch_bounds__(n* n* n, volsurf_min_blocks<n>()) k_volsurf(const KParams P, double mRKA, double b_dt_in) {
    const double b_dt = P.dtDev ? b_dt_in * __ldg(P.dtDev + 3) : b_dt_in;  // device-paced stepping: b_dt_in carries RKb, dt lives on the device
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ double smem[];
    double* sA = smem;               // [15][n3] multipurpose: fluxes f,g,h (15) | node record (6) + metrics (9) | U tile (5)
    double* sFl = smem + 15 * n3;    // [6][5][n2] signed face fluxes in element face order
    double* sDh = sFl + 30 * n2;     // D_Hat_T
    double* sDv = sDh + n * n;       // DVolSurf
    double* sLhm = sDv + n * n;
    double* sLhp = sLhm + n;
    double* sLm = sLhp + n;
    double* sLp = sLm + n;
    const int e = P.elemList ? P.elemList[blockIdx.x] : blockIdx.x;
    const int t = threadIdx.x;
    for (int x = t; x < n * n; x += n3) { sDh[x] = P.D_Hat_T[x]; sDv[x] = P.DVolSurf[x]; }
    if (t < n) { sLhm[t] = P.L_HatMinus[t]; sLhp[t] = P.L_HatPlus[t]; sLm[t] = P.L_Minus[t]; sLp[t] = P.L_Plus[t]; }
    const Eos eos = P.eos;
    const int* e2s = P.E2S + 18 * e;
    const int k = t / n2, j = (t - k * n2) / n, i = t - k * n2 - j * n;
    const bool split = P.splitDG >= 0;

    // face fluxes -> shared memory (element face order, sign: +master / -slave)
    for (int f = t; f < 6 * n2; f += n3) {
        const int loc = f / n2 + 1;
        const int pq = f - (loc - 1) * n2;
        const int q = pq / n, p = pq - q * n;
        const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
        const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
        const int a = s2v2<n>(P.S2V2, 0, p, q, flip, loc);
        const int b = s2v2<n>(P.S2V2, 1, p, q, flip, loc);
        const double sg = (flip == 0) ? 1.0 : -1.0;
        const double* F = P.Flux + (size_t)side * 5 * n2 + pq;
        double* d = sFl + (loc - 1) * 5 * n2 + (b * n + a);
#pragma unroll
        for (int v = 0; v < 5; v++) d[v * n2] = sg * F[v * n2];
    }

    // node state
    double Uc[5], Pr[6], M[9];
    {
        const double* U = P.U + (size_t)e * 5 * n3 + t;
#pragma unroll
        for (int v = 0; v < 5; v++) Uc[v] = U[v * n3];
        const double* Mg = P.metrics + (size_t)e * 9 * n3 + t;
#pragma unroll
        for (int x = 0; x < 9; x++) M[x] = Mg[x * n3];
    }
    cons_to_prim(Pr, Uc, eos);
    double Ut[5] = {0, 0, 0, 0, 0};

    // ---- viscous volume integral: formed by k_lifting (its step 5), taken over as the initial value (volint.f90:238-243)
    if (P.parabolic) {
        const double* uv = P.Ut + (size_t)e * 5 * n3 + n3 + t;
#pragma unroll
        for (int v = 0; v < 4; v++) Ut[1 + v] = uv[v * n3];
    }
    // ---- weak-form part: Euler fluxes (non-split)
    if (!split) {
        double f[5], g[5], h[5];
        const double Ep = (Uc[ENER] + Pr[PRES]) / Uc[DENS];
        double* Fs[3] = {f, g, h};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double* Md = M + 3 * d;
            const double Mmom = Md[0] * Uc[MOM1] + Md[1] * Uc[MOM2] + Md[2] * Uc[MOM3];
            Fs[d][DENS] = Mmom;
            Fs[d][MOM1] = Mmom * Pr[VEL1] + Md[0] * Pr[PRES];
            Fs[d][MOM2] = Mmom * Pr[VEL2] + Md[1] * Pr[PRES];
            Fs[d][MOM3] = Mmom * Pr[VEL3] + Md[2] * Pr[PRES];
            Fs[d][ENER] = Mmom * Ep;
        }
#pragma unroll
        for (int v = 0; v < 5; v++) { sA[v * n3 + t] = f[v]; sA[(5 + v) * n3 + t] = g[v]; sA[(10 + v) * n3 + t] = h[v]; }
        __syncthreads();
        // D_Hat sweep (applydmatrix.t90:60-67)
        double A[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int l = 0; l < n; l++) {
            const double dx = sDh[l + n * i], dy = sDh[l + n * j], dz = sDh[l + n * k];
#pragma unroll
            for (int v = 0; v < 5; v++)
                A[v] += dx * sA[v * n3 + l + n * (j + n * k)] + dz * sA[(10 + v) * n3 + i + n * (j + n * l)] +
                        dy * sA[(5 + v) * n3 + i + n * (l + n * k)];
        }
#pragma unroll
        for (int v = 0; v < 5; v++) Ut[v] += A[v];
        __syncthreads();
    }

    // ---- split-form flux differencing (volint.f90:306-347)
    if (split) {
        const int var = P.splitDG;
        double me[6] = {Pr[DENS], Pr[VEL1], Pr[VEL2], Pr[VEL3], Pr[PRES], split_sixth(var, Uc, Pr)};
#pragma unroll
        for (int v = 0; v < 6; v++) sA[v * n3 + t] = me[v];
#pragma unroll
        for (int x = 0; x < 9; x++) sA[(6 + x) * n3 + t] = M[x];
        __syncthreads();
        double acc[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int idx = (d == 0) ? i : (d == 1 ? j : k);
            const int stride = (d == 0) ? 1 : (d == 1 ? n : n2);
            const int base = t - idx * stride;
#pragma unroll
            for (int l = 0; l < n; l++) {
                const int o = base + l * stride;
                double ot[6], Ms[3], F[5];
#pragma unroll
                for (int v = 0; v < 6; v++) ot[v] = sA[v * n3 + o];
#pragma unroll
                for (int c = 0; c < 3; c++) Ms[c] = M[3 * d + c] + sA[(6 + 3 * d + c) * n3 + o];
                split_volume_flux(var, me, ot, Ms, F);
                const double w = sDv[l + n * idx];
#pragma unroll
                for (int v = 0; v < 5; v++) acc[v] += w * F[v];
            }
        }
#pragma unroll
        for (int v = 0; v < 5; v++) Ut[v] += acc[v];
    }
    __syncthreads();  // sFl complete (written before the first barrier in any path), sA free for reuse

    // ---- surface integral (surfint.t90:463-584), local sides in order 1..6
    {
        double S[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int loc = 1; loc <= 6; loc++) {
            int a, b, l;
            if (loc == XI_MINUS || loc == XI_PLUS) { a = j; b = k; l = i; }
            else if (loc == ETA_MINUS || loc == ETA_PLUS) { a = i; b = k; l = j; }
            else { a = i; b = j; l = k; }
            const bool minus = is_minus(loc);
            double Lh;
            if (NT == 2) {
                if (minus ? (l != 0) : (l != n - 1)) continue;
                Lh = minus ? sLhm[0] : sLhp[n - 1];
            } else {
                Lh = minus ? sLhm[l] : sLhp[l];
            }
            const double* s = sFl + (loc - 1) * 5 * n2 + (b * n + a);
#pragma unroll
            for (int v = 0; v < 5; v++) S[v] += s[v * n2] * Lh;
        }
#pragma unroll
        for (int v = 0; v < 5; v++) Ut[v] += S[v];
    }
    // ---- sign and Jacobian (dg.f90:413,423)
    const double msJ = P.noJac ? -1.0 : -P.sJ[(size_t)e * n3 + t];  // noJac: k_overint applies the Jacobian after its filter
#pragma unroll
    for (int v = 0; v < 5; v++) Ut[v] *= msJ;
    if (P.tcSource == 2) {  // channel forcing folded into this epilogue (TestcaseSource, testcase/channel/testcase.f90:277-296)
        const double bulk = P.bulkDev ? __ldg(P.bulkDev) : P.tcBulkVel;
        Ut[MOM1] = __dadd_rn(Ut[MOM1], -P.tcDpdx);   // two roundings, like Ut + src in k_source_rk
        Ut[ENER] = __dadd_rn(Ut[ENER], -__dmul_rn(P.tcDpdx, bulk));
    }

    if (MODE == 0) {
        double* o = P.Ut + (size_t)e * 5 * n3 + t;
#pragma unroll
        for (int v = 0; v < 5; v++) o[v * n3] = Ut[v];
    } else {
        // Williamson 2N update (vector.f90:163-183): Ut_tmp = Ut_tmp*mRKA + Ut ; U = U + Ut_tmp*b_dt
        double* ot = P.Ut_tmp + (size_t)e * 5 * n3 + t;
        double* ou = P.U + (size_t)e * 5 * n3 + t;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            const double r = (mRKA == 0.0) ? Ut[v] : ot[v * n3] * mRKA + Ut[v];
            ot[v * n3] = r;
            const double un = Uc[v] + r * b_dt;
            ou[v * n3] = un;
            sA[v * n3 + t] = un;
        }
        __syncthreads();
        extract_faces<n, NT, 5>(sA, P.UmNext, P.UsNext, e2s, P.S2V2, sLm, sLp);
    }
}

template <int n>
constexpr size_t volsurf_smem_bytes() { return sizeof(double) * (15 * n * n * n + 30 * n * n + 2 * n * n + 4 * n); }

// ---------------------------------------------------------------------------------------------------------
// per element max eigenvalues -> global min of CFL*2/lam_c and DFL*4/lam_v  (calctimestep.f90:98-296)
template <int n>
constexpr int timestep_threads() { return ((n * n * n + 31) / 32) * 32; }

template <int n>
__global__ void __launch_bounds__(timestep_threads<n>()) k_timestep(const KParams P, double CFL, double DFL, double* out /*[2]*/) {
    constexpr int n3 = n * n * n;
    __shared__ double red[6][32];
    const int e = blockIdx.x, t = threadIdx.x;
    const bool active = t < n3;  // block is padded to full warps
    double lam[6] = {0, 0, 0, 0, 0, 0};
    bool bad = false;
    if (active) {
        double Uc[5], M[9];
        const double* U = P.U + (size_t)e * 5 * n3 + t;
#pragma unroll
        for (int v = 0; v < 5; v++) Uc[v] = U[v * n3];
        const double* Mg = P.metrics + (size_t)e * 9 * n3 + t;
#pragma unroll
        for (int x = 0; x < 9; x++) M[x] = Mg[x * n3];
        bad = timestep_node(P.eos, P.parabolic != 0, Uc, M, P.sJ[(size_t)e * n3 + t], lam);
    }
    timestep_block_min<timestep_threads<n>()>(lam, bad, red, P, CFL, DFL, out);
}

// ---------------------------------------------------------------------------------------------------------
// CalcForcing of the channel testcase (testcase/channel/testcase.f90:241-271): per-element sum of u wGPVol / sJ; the
// partials are reduced in a fixed order by k_sum_partials
template <int n>
__global__ void __launch_bounds__(timestep_threads<n>()) k_bulkvel(const KParams P, const double* __restrict__ wGP, double* __restrict__ partials) {
    constexpr int n2 = n * n, n3 = n2 * n;
    __shared__ double red[32];
    const int e = blockIdx.x, t = threadIdx.x;
    double v = 0.0;
    if (t < n3) {
        const int k = t / n2, j = (t - k * n2) / n, i = t - k * n2 - j * n;
        const double* U = P.U + (size_t)e * 5 * n3 + t;
        v = U[n3] / U[0] * (wGP[i] * wGP[j] * wGP[k]) / P.sJ[(size_t)e * n3 + t];
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
        double s = red[0];
        for (int w = 1; w < timestep_threads<n>() / 32; w++) s += red[w];
        partials[e] = s;
    }
}
static __global__ void k_clear_flag_bits(int* flag, int bits) { *flag &= ~bits; }
// sponge/pruettdamping.f90:69-92 TempFilterTimeDeriv: SpBaseFlow += (U - SpBaseFlow) dt / tempFilterWidth
static __global__ void k_pruett(const double* __restrict__ U, double* __restrict__ base, double fac, size_t total) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) base[gid] = base[gid] + (U[gid] - base[gid]) * fac;
}
static __global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partials, int count, double* __restrict__ out) {
    __shared__ double red[256];
    double v = 0.0;
    for (int e = threadIdx.x; e < count; e += 256) v += partials[e];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ---------------------------------------------------------------------------------------------------------
// layout conversion at the C ABI: reference AoS (nv, n^3, nElems) <-> SoA tiles [elem][nv][n^3]
template <int NVAR>
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int n3, size_t total) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over elem*n3*NVAR in SoA order
    if (gid >= total) return;
    const size_t e = gid / ((size_t)NVAR * n3);
    const int r = (int)(gid - e * NVAR * n3);
    const int v = r / n3, node = r - v * n3;
    soa[gid] = aos[(e * n3 + node) * NVAR + v];
}
template <int NVAR>
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int n3, size_t total, int soaStride, int soaOffset) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over elem*n3*NVAR in AoS order
    if (gid >= total) return;
    const size_t dof = gid / NVAR;
    const int v = (int)(gid - dof * NVAR);
    const size_t e = dof / n3;
    const int node = (int)(dof - e * n3);
    aos[gid] = soa[(e * soaStride + soaOffset + v) * n3 + node];
}
// side geometry (3,n,n,S)x3 + (n,n,S) -> [side][10][n2]
static __global__ void k_pack_geo(const double* nv, const double* t1, const double* t2, const double* se, double* geo, int n2, size_t total) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const size_t s = gid / (10 * n2);
    const int r = (int)(gid - s * 10 * n2);
    const int c = r / n2, pq = r - c * n2;
    double v;
    if (c < 3) v = nv[(s * n2 + pq) * 3 + c];
    else if (c < 6) v = t1[(s * n2 + pq) * 3 + c - 3];
    else if (c < 9) v = t2[(s * n2 + pq) * 3 + c - 6];
    else v = se[s * n2 + pq];
    geo[gid] = v;
}
// metrics (3,n3,E)x3 -> [e][9][n3]
static __global__ void k_pack_metrics(const double* mf, const double* mg, const double* mh, double* out, int n3, size_t total) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const size_t e = gid / (9 * (size_t)n3);
    const int r = (int)(gid - e * 9 * n3);
    const int c = r / n3, node = r - c * n3;
    const double* src = c < 3 ? mf : (c < 6 ? mg : mh);
    out[gid] = src[(e * n3 + node) * 3 + (c % 3)];
}

}  // namespace dgx
