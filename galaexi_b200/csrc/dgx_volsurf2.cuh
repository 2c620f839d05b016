// k_volsurf2: split-form (Gauss-Lobatto) volume + surface integral + Jacobian + RK stage, register-blocked.
//
// Same operator as k_volsurf (dgx_kernels.cuh; reference: dg/volint.f90:60-119 + :211-353, dg/surfint.t90:352-586,
// globals/vector.f90:163-226, interpolation/applyjacobian.t90:196, interpolation/prolongtoface.t90:168-344), different
// work decomposition, chosen from the ncu profile of k_volsurf (profiles/r01a_volsurf_full.md: shared-memory
// wavefronts and a single 512-thread CTA per SM were the limiters):
//   * 2*n^2 threads per element instead of n^3. In the three flux-differencing sweeps a thread owns HALF A LINE
//     (SEG = ceil(n/2) output nodes, held in registers with their accumulators): pairs inside the own segment are
//     evaluated once (the two-point flux is symmetric), the other half of the line streams through registers.
//     Shared-memory loads per node drop from 24*9 to about 2*9 per direction.
//   * In the point-wise phases the same thread owns the nodes (i,j,k in its half of the zeta column), so global
//     accesses stay coalesced and the zeta sweep accumulates directly into the registers of the epilogue.
//   * The viscous volume integral (volint.f90:60-119) is NOT evaluated here: k_lifting, which has the element's gradients in
//     registers, forms it and hands over 4 doubles per node in Ut(1..4); this kernel starts its accumulators from them.
//   * 17 tile slots of shared memory per element (5 Ut partials, 6 node record, 2 metric triples) -> 3 CTAs per SM at N=7,
//     which overlaps the load / sweep / store phases of different elements.
//   * Shared-memory tile layout for n=8: (i xor k) + 8 j + 72 k, conflict-free for 64-bit accesses of all three
//     line directions with the lane mappings used below.
//   * Gauss-Lobatto only (split DG requires it, splitflux.f90:116-119): surface integral and next-stage face
//     extraction touch only the boundary layer of nodes and go directly to global memory.
#pragma once
#include "dgx_kernels.cuh"

namespace dgx {

// Tile layout of k_volsurf2: the conflict-free swizzle for n = 8 (Tile<8>), and for the other sizes the plain layout with
// a padded zeta stride chosen by exhaustive search over the kernel's access patterns (point-wise and the three line
// directions, both line halves, 2 elements per CTA): n = 6: 1.84 -> 1.39 wavefronts per ideal wavefront, n = 4: 2.71 ->
// 1.29; n = 5, 7 gain nothing (odd strides) and stay unpadded.
template <int n>
struct TileV {
    static constexpr int PK = (n == 6) ? 41 : ((n == 4) ? 19 : n * n);
    static constexpr int SLOT = (n == 8) ? Tile<n>::SLOT : PK * n;
    __device__ __forceinline__ static int idx(int i, int j, int k) {
        if constexpr (n == 8) return Tile<n>::idx(i, j, k);
        else return i + n * j + PK * k;
    }
};

#ifndef VS2_EPB6
#define VS2_EPB6 1   // elements per CTA at n = 6 (tuning knob; 1 element / 5 CTAs per SM: 0.93 -> 0.79 ms at 32^3, profiles/r03f_*)
#endif
template <int n>
constexpr int vs2_epb() { return n == 6 ? VS2_EPB6 : ((128 + n * n) / (2 * n * n) > 0 ? (128 + n * n) / (2 * n * n) : 1); }
#ifndef VS2_PARTS8
#define VS2_PARTS8 2   // threads per line at n = 8 (2: half line per thread, 4: quarter line per thread)
#endif
template <int n>
constexpr int vs2_parts() { return n == 8 ? VS2_PARTS8 : 2; }
template <int n>
constexpr int vs2_seg() { return (n + vs2_parts<n>() - 1) / vs2_parts<n>(); }
template <int n>
constexpr int vs2_threads() { return vs2_epb<n>() * vs2_parts<n>() * n * n; }
// shared-memory slots of one element: Ut partials (momentum, energy), density partial, node record (6), metric triple of the
// xi (later zeta) sweep, metric triple of the eta sweep
constexpr int S_UT = 0, S_RHO = 4, S_REC = 5, S_MX = 11, S_ME = 14;
constexpr int VS2_SLOTS = 17;
#ifndef VS2_CROSS_UNROLL
#define VS2_CROSS_UNROLL 1
#endif
// 1: the consistent (l == i) terms of the flux differencing are skipped: on Gauss-Lobatto nodes DVolSurf is zero on its main
// diagonal (volint.f90:257-259 "Attention 5"; the reference's CPU loop starts at l = i+1, its GPU kernel multiplies the
// consistent flux by that zero). In floating point the table holds O(1e-15) there, so the switch moves Ut by ~1e-15 relative.
// the three sweep directions as three copies of the sweep code (direction a compile-time value: no index selects) for n in
// [VS2_DIRU_MINN, VS2_DIRU_MAXN]. Round 1's kernel (with the viscous sweep inside) was instruction-cache bound in this form
// (profiles/r01e_volsurf2_full.md); without it the three copies fit: n = 8 1.599 -> 1.516 ms, n = 6 0.836 -> 0.786 ms (r03)
#ifndef VS2_DIRU_MAXN
#define VS2_DIRU_MAXN 8
#endif
#ifndef VS2_DIRU_MINN
#define VS2_DIRU_MINN 6   // measured at n = 6 only (0.84 -> 0.79 ms); n = 5 spills with three copies at its register cap
#endif
#ifndef VS2_SKIP_DIAG
#define VS2_SKIP_DIAG 0
#endif
constexpr int VS2_CU = VS2_CROSS_UNROLL;
#ifndef VS2_MIN_BLOCKS
#ifndef VS2_MINB6
#define VS2_MINB6 5
#endif
#ifndef VS2_MINB8
#define VS2_MINB8 3
#endif
#define VS2_MIN_BLOCKS(n) ((n) == 6 ? VS2_MINB6 : ((n) == 8 ? VS2_MINB8 : ((n) >= 6 ? 3 : 4)))
#endif
template <int n>
constexpr size_t vs2_smem_bytes() { return sizeof(double) * ((size_t)vs2_epb<n>() * VS2_SLOTS * TileV<n>::SLOT); }

// Paired operand layout (n = 8, knob VS2_PAIRED8): the node record is kept as three double2 ([0,1] [2,3] [4,5]) and a metric
// triple as double2 + double, so that a partner node costs 5 shared-memory load instructions (4 LDS.128 + 1 LDS.64) instead
// of 9 LDS.64. Tile<8> puts (i xor k) into the low three bits of the node index: the eight lanes of a quarter warp hit eight
// different 16-byte bank groups in the point-wise phases and in all three sweep directions.
#ifndef VS2_PAIRED8
#define VS2_PAIRED8 0
#endif
template <int n>
constexpr bool vs2_paired() { return n == 8 && VS2_PAIRED8 != 0; }
template <int n>
__device__ __forceinline__ void vs2_rec_load(const double* __restrict__ R, int id, double* __restrict__ o) {
    constexpr int SL = TileV<n>::SLOT;
    if constexpr (vs2_paired<n>()) {
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const double2 x = reinterpret_cast<const double2*>(R + 2 * p * SL)[id];
            o[2 * p] = x.x; o[2 * p + 1] = x.y;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 6; v++) o[v] = R[v * SL + id];
    }
}
template <int n>
__device__ __forceinline__ void vs2_rec_store(double* __restrict__ R, int id, const double* __restrict__ r) {
    constexpr int SL = TileV<n>::SLOT;
    if constexpr (vs2_paired<n>()) {
#pragma unroll
        for (int p = 0; p < 3; p++) reinterpret_cast<double2*>(R + 2 * p * SL)[id] = make_double2(r[2 * p], r[2 * p + 1]);
    } else {
#pragma unroll
        for (int v = 0; v < 6; v++) R[v * SL + id] = r[v];
    }
}
// address of component c of the metric triple stored at Mx for node id
template <int n>
__device__ __forceinline__ double* vs2_met_ptr(double* Mx, int id, int c) {
    constexpr int SL = TileV<n>::SLOT;
    if constexpr (vs2_paired<n>()) return c < 2 ? Mx + 2 * id + c : Mx + 2 * SL + id;
    else return Mx + c * SL + id;
}
template <int n>
__device__ __forceinline__ void vs2_met_load(const double* __restrict__ Mx, int id, double* __restrict__ o) {
    constexpr int SL = TileV<n>::SLOT;
    if constexpr (vs2_paired<n>()) {
        const double2 x = reinterpret_cast<const double2*>(Mx)[id];
        o[0] = x.x; o[1] = x.y; o[2] = Mx[2 * SL + id];
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) o[c] = Mx[c * SL + id];
    }
}

// tile index of position l on the line (c1,c2) of direction d (0 xi: (j,k), 1 eta: (i,k), 2 zeta: (i,j)); d is a run-time
// value on purpose: one copy of the sweep code serves the three directions (the fully unrolled variant was
// instruction-cache bound, profiles/r01e_volsurf2_full.md)
template <int n>
__device__ __forceinline__ int line_idx(int d, int l, int c1, int c2) {
    const int i = (d == 0) ? l : c1;
    const int j = (d == 0) ? c1 : ((d == 1) ? l : c2);
    const int k = (d == 2) ? l : c2;
    return TileV<n>::idx(i, j, k);
}

// Two-point flux of one node pair for split variant VAR (compile-time: no variant branches in the sweeps, and the dead
// variants stay out of the instruction cache). a, b: node record (6) followed by the metric triple of the direction (3).
// Record layouts (vs2_record): PI/KG: {rho/8, u, v, w, p/2, H or e}  -- the constant factors of splitflux.f90:437-501 /
// :669-730 folded into the operands (exact: powers of two); SD/MO/DU: {rho, u, v, w, p, rhoE}.
template <int VAR>
__device__ __forceinline__ void vs2_record(double* rec, const double* Uc, const double* Pr) {
    if (VAR >= 3) {
        rec[0] = 0.125 * Pr[DENS]; rec[1] = Pr[VEL1]; rec[2] = Pr[VEL2]; rec[3] = Pr[VEL3]; rec[4] = 0.5 * Pr[PRES];
        rec[5] = split_sixth(VAR, Uc, Pr);
    } else {
        rec[0] = Pr[DENS]; rec[1] = Pr[VEL1]; rec[2] = Pr[VEL2]; rec[3] = Pr[VEL3]; rec[4] = Pr[PRES]; rec[5] = Uc[ENER];
    }
}
template <int VAR>
__device__ __forceinline__ void vs2_pair_flux(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ F) {
    double Ms[3];
#pragma unroll
    for (int c = 0; c < 3; c++) Ms[c] = a[6 + c] + b[6 + c];
    if (VAR >= 3) {
        const double us = a[1] + b[1], vs = a[2] + b[2], ws = a[3] + b[3];
        const double dot = Ms[0] * us + Ms[1] * vs + Ms[2] * ws;  // 2 x contravariant velocity sum
        const double q = (a[0] + b[0]) * dot;                    // 1/4 {rho}_s * 1/2 dot
        const double ph = a[4] + b[4];                            // 1/2 p_s
        F[DENS] = q + q;
        F[MOM1] = q * us + Ms[0] * ph;
        F[MOM2] = q * vs + Ms[1] * ph;
        F[MOM3] = q * ws + Ms[2] * ph;
        if (VAR == 3) F[ENER] = q * (a[5] + b[5]) + ph * (0.5 * dot);
        else F[ENER] = q * (a[5] + b[5]);
    } else {
        split_volume_flux(VAR, a, b, Ms, F);
    }
}
// acc += w * F#(a,b) for a pair used by one end point only (the other end lives in another thread)
template <int VAR>
__device__ __forceinline__ void vs2_pair_acc(const double* __restrict__ a, const double* __restrict__ b, double w, double* __restrict__ acc) {
    if (VAR >= 3) {
        double Ms[3];
#pragma unroll
        for (int c = 0; c < 3; c++) Ms[c] = a[6 + c] + b[6 + c];
        const double us = a[1] + b[1], vs = a[2] + b[2], ws = a[3] + b[3];
        const double dot = Ms[0] * us + Ms[1] * vs + Ms[2] * ws;
        const double wq = w * ((a[0] + b[0]) * dot);
        const double wp = w * (a[4] + b[4]);
        acc[DENS] = fma(wq, 2.0, acc[DENS]);
        acc[MOM1] = fma(wp, Ms[0], fma(wq, us, acc[MOM1]));
        acc[MOM2] = fma(wp, Ms[1], fma(wq, vs, acc[MOM2]));
        acc[MOM3] = fma(wp, Ms[2], fma(wq, ws, acc[MOM3]));
        acc[ENER] = fma(wq, a[5] + b[5], acc[ENER]);
        if (VAR == 3) acc[ENER] = fma(wp, 0.5 * dot, acc[ENER]);
    } else {
        double F[5];
        vs2_pair_flux<VAR>(a, b, F);
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v] += w * F[v];
    }
}

// One flux-differencing sweep of direction d for the half line (c1,c2,h): acc[m][v] = sum_l DVolSurf(l,a_m) F#(a_m,l)
template <int n, int VAR>
__device__ __forceinline__ void vs2_sweep(const double* __restrict__ R, const double* __restrict__ Mx, const double* __restrict__ Dv, int d,
                                          int c1, int c2, int h, double (&acc)[vs2_seg<n>()][5]) {
    constexpr int SEG = vs2_seg<n>(), SL = TileV<n>::SLOT;
    const int a0 = h * SEG, cnt = (vs2_parts<n>() == 2) ? (h ? n - SEG : SEG) : ((n - a0 < SEG) ? n - a0 : SEG);
    const int f0 = h ? 0 : SEG, fcnt = n - cnt;
    double own[SEG][9];
#pragma unroll
    for (int m = 0; m < SEG; m++) {
        const int id = line_idx<n>(d, m < cnt ? a0 + m : a0, c1, c2);
        vs2_rec_load<n>(R, id, own[m]);
        vs2_met_load<n>(Mx, id, own[m] + 6);
#pragma unroll
        for (int v = 0; v < 5; v++) acc[m][v] = 0.0;
    }
    // pairs inside the own segment: evaluated once, used for both end points
#pragma unroll
    for (int m = 0; m < SEG; m++) {
        if (m < cnt) {
            if (!VS2_SKIP_DIAG) vs2_pair_acc<VAR>(own[m], own[m], Dv[(a0 + m) + n * (a0 + m)], acc[m]);
#pragma unroll
            for (int m2 = m + 1; m2 < SEG; m2++) {
                if (m2 < cnt) {
                    double F[5];
                    vs2_pair_flux<VAR>(own[m], own[m2], F);
                    const double w1 = Dv[(a0 + m2) + n * (a0 + m)], w2 = Dv[(a0 + m) + n * (a0 + m2)];
#pragma unroll
                    for (int v = 0; v < 5; v++) { acc[m][v] += w1 * F[v]; acc[m2][v] += w2 * F[v]; }
                }
            }
        }
    }
    // pairs with the rest of the line (rolled loop: code size)
#pragma unroll(VS2_CU)
    for (int mf = 0; mf < fcnt; mf++) {
        const int b = (vs2_parts<n>() == 2) ? f0 + mf : ((mf < a0) ? mf : mf + cnt);
        const int id = line_idx<n>(d, b, c1, c2);
        double ot[9];
        vs2_rec_load<n>(R, id, ot);
        vs2_met_load<n>(Mx, id, ot + 6);
#pragma unroll
        for (int m = 0; m < SEG; m++)
            if (m < cnt) vs2_pair_acc<VAR>(own[m], ot, Dv[b + n * (a0 + m)], acc[m]);
    }
}

template <int n, int MODE, int VAR>
__global__ void __launch_bounds__(vs2_threads<n>(), VS2_MIN_BLOCKS(n)) k_volsurf2(const __grid_constant__ KParams P, int nWork, double mRKA, double b_dt_in, int lookahead) {
    const double b_dt = P.dtDev ? b_dt_in * __ldg(P.dtDev + 3) : b_dt_in;  // device-paced stepping: b_dt_in carries RKb, dt lives on the device
    constexpr int n2 = n * n, n3 = n2 * n, SEG = vs2_seg<n>(), SL = TileV<n>::SLOT, T = vs2_parts<n>() * n2, EPB = vs2_epb<n>();
    extern __shared__ double smem[];
    const int le = threadIdx.x / T, tid = threadIdx.x - le * T;
    const int we = blockIdx.x * EPB + le;
    const bool live = we < nWork;
    const int e = live ? (P.elemList ? P.elemList[we] : we) : 0;
    double* S = smem + (size_t)le * VS2_SLOTS * SL;
    const int h = tid / n2, q = tid - h * n2;
    const int c1 = q % n, c2 = q / n;          // point-wise phases: (i,j) = (c1,c2), k in the own half of the zeta column
    const int a0 = h * SEG, cnt = (vs2_parts<n>() == 2) ? (h ? n - SEG : SEG) : ((n - a0 < SEG) ? n - a0 : SEG);
    const bool facer = (vs2_parts<n>() == 2) || h < 2;                  // P4: the threads of the first two parts take the - / + face of each axis
    const Eos eos = P.eos;
    const bool par = P.parabolic != 0;
    const double* __restrict__ Dv = P.DVolSurf;
    const double* __restrict__ gU_e = P.U + (size_t)e * 5 * n3;
    const double* __restrict__ gM_e = P.metrics + (size_t)e * 9 * n3;

    // L2 prefetches (no registers): data this CTA reads several barriers from now (zeta metrics, Jacobian, Ut_tmp, face
    // fluxes), and the P0 data of the element that replaces this CTA when it retires (one resident wave ahead)
    if (live && (P.flags & 4)) {
        prefetch_block(gM_e + (size_t)6 * n3, sizeof(double) * 3 * n3, tid, T);
        prefetch_block(P.sJ + (size_t)e * n3, sizeof(double) * n3, tid, T);
        if (MODE == 1) prefetch_block(P.Ut_tmp + (size_t)e * 5 * n3, sizeof(double) * 5 * n3, tid, T);
        constexpr int LPS = (5 * n2 * 8 + 127) / 128;  // 128-byte lines per side of the flux array
        if (tid < 6 * LPS) {
            const int side = __ldg(&P.E2S[18 * e + 3 * (tid / LPS)]) - 1;
            prefetch_l2(reinterpret_cast<const char*>(P.Flux + (size_t)side * 5 * n2) + (tid % LPS) * 128);
        }
    }
    if ((P.flags & 8) && we + lookahead * EPB < nWork) {
        const int en = P.elemList ? P.elemList[we + lookahead * EPB] : we + lookahead * EPB;
        prefetch_block(P.U + (size_t)en * 5 * n3, sizeof(double) * 5 * n3, tid, T);
        prefetch_block(P.metrics + (size_t)en * 9 * n3, sizeof(double) * 6 * n3, tid, T);
        if (par) prefetch_block(P.Ut + (size_t)en * 5 * n3 + n3, sizeof(double) * 4 * n3, tid, T);
    }
    // ---- P0: point-wise. The viscous volume integral of the element (4 doubles per node, formed by k_lifting's step 5)
    // is the initial value of the Ut partials (volint.f90:238-243): 8-byte cp.async straight into slots S_UT.. at the
    // thread's own nodes (no registers held); state and metrics into registers; node record and the xi / eta metric triples
    // into their slots.
    if (live) {
        if (par) {
            const double* gV = P.Ut + (size_t)e * 5 * n3 + n3;
#pragma unroll
            for (int m = 0; m < SEG; m++) {
                if (m < cnt) {
                    const int node = c1 + n * c2 + n2 * (a0 + m);
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(S + S_UT * SL + TileV<n>::idx(c1, c2, a0 + m));
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (unsigned)(v * SL * 8)), "l"(gV + v * n3 + node) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        double Uc[SEG][5], M[SEG][6];
#pragma unroll
        for (int m = 0; m < SEG; m++) {
            const int node = c1 + n * c2 + n2 * (m < cnt ? a0 + m : a0);
#pragma unroll
            for (int v = 0; v < 5; v++) Uc[m][v] = gU_e[v * n3 + node];
#pragma unroll
            for (int x = 0; x < 6; x++) M[m][x] = gM_e[x * n3 + node];
        }
#pragma unroll
        for (int m = 0; m < SEG; m++) {
            if (m < cnt) {
                const int id = TileV<n>::idx(c1, c2, a0 + m);
                double Pr[6], Rec[6];
                cons_to_prim(Pr, Uc[m], eos);
                vs2_record<VAR>(Rec, Uc[m], Pr);
                vs2_rec_store<n>(S + S_REC * SL, id, Rec);
#pragma unroll
                for (int c = 0; c < 6; c++) *vs2_met_ptr<n>(S + (c < 3 ? S_MX : S_ME) * SL, id, c % 3) = M[m][c];
                S[S_RHO * SL + id] = 0.0;
                if (!par) {
#pragma unroll
                    for (int v = 0; v < 4; v++) S[(S_UT + v) * SL + id] = 0.0;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ---- P3: the three flux-differencing sweeps (volint.f90:306-347). Metric triples: xi in slots S_MX.. and eta in
    // S_ME.. (from P0); zeta is copied by cp.async into S_MX.. while the eta sweep runs.
    constexpr int DIRU = (n <= VS2_DIRU_MAXN && n >= VS2_DIRU_MINN) ? 3 : 1;
#pragma unroll(DIRU)
    for (int d = 0; d < 3; d++) {
        if (d == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // node record / previous partials / metric triple visible, previous triple consumed
        if (live) {
            if (d == 1) {
#pragma unroll
                for (int m = 0; m < SEG; m++) {
                    if (m < cnt) {
                        const double* Mg = gM_e + (c1 + n * c2 + n2 * (a0 + m)) + (size_t)6 * n3;
                        const int idm = TileV<n>::idx(c1, c2, a0 + m);
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(vs2_met_ptr<n>(S + S_MX * SL, idm, c));
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(Mg + c * n3) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // lane -> line mapping: xi lines take (j,k) = (q/n, q%n), eta (i,k) and zeta (i,j) = (q%n, q/n): conflict-free
            const int l1 = (d == 0) ? c2 : c1, l2 = (d == 0) ? c1 : c2;
            double acc[SEG][5];
            vs2_sweep<n, VAR>(S + S_REC * SL, S + (d == 1 ? S_ME : S_MX) * SL, Dv, d, l1, l2, h, acc);
#pragma unroll
            for (int m = 0; m < SEG; m++) {
                if (m < cnt) {
                    const int id = line_idx<n>(d, a0 + m, l1, l2);
                    S[S_RHO * SL + id] += acc[m][0];
#pragma unroll
                    for (int v = 0; v < 4; v++) S[(S_UT + v) * SL + id] += acc[m][1 + v];
                }
            }
        }
    }
    // ---- P4: surface integral (surfint.t90:519-573; GL: only the boundary layer). One face node per thread and round;
    // the two faces of a round (-/+ side of one axis) touch disjoint nodes. All global reads of P4 and P5 (face fluxes,
    // sJ, Ut_tmp, U) are issued here, before the first barrier of the rounds: one DRAM round trip for the rest of the CTA.
    const int* e2s = P.E2S + 18 * e;
    double Ff[3][5], wf[3];
    int idf[3];
    double sJv[SEG], Uo[SEG][5], Uto[SEG][5];
    if (live) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            if (!facer) break;
            const int loc = (r == 0) ? (h ? XI_PLUS : XI_MINUS) : ((r == 1) ? (h ? ETA_PLUS : ETA_MINUS) : (h ? ZETA_PLUS : ZETA_MINUS));
            const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
            const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
            int p, qq;
            face_lane<n>(P.S2V2, q, flip, loc, (loc == XI_MINUS || loc == XI_PLUS), p, qq);
            const int a = s2v2<n>(P.S2V2, 0, p, qq, flip, loc);
            const int b = s2v2<n>(P.S2V2, 1, p, qq, flip, loc);
            const int l = h ? n - 1 : 0;
            idf[r] = (r == 0) ? TileV<n>::idx(l, a, b) : ((r == 1) ? TileV<n>::idx(a, l, b) : TileV<n>::idx(a, b, l));
            wf[r] = ((flip == 0) ? 1.0 : -1.0) * (h ? P.L_HatPlus[n - 1] : P.L_HatMinus[0]);
            const double* F = P.Flux + (size_t)side * 5 * n2 + (p + n * qq);
#pragma unroll
            for (int v = 0; v < 5; v++) Ff[r][v] = F[v * n2];
        }
#pragma unroll
        for (int m = 0; m < SEG; m++) {
            const int node = c1 + n * c2 + n2 * (m < cnt ? a0 + m : a0);
            sJv[m] = P.sJ[(size_t)e * n3 + node];
            if (MODE == 1) {
#pragma unroll
                for (int v = 0; v < 5; v++) { Uo[m][v] = gU_e[v * n3 + node]; Uto[m][v] = P.Ut_tmp[(size_t)e * 5 * n3 + v * n3 + node]; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        __syncthreads();
        if (live && facer) {
            S[S_RHO * SL + idf[r]] += Ff[r][0] * wf[r];
#pragma unroll
            for (int v = 0; v < 4; v++) S[(S_UT + v) * SL + idf[r]] += Ff[r][1 + v] * wf[r];
        }
    }
    __syncthreads();  // Ut complete in slots 10, 0..3; node record consumed: slots 4..8 take the updated state
    // ---- P5: -sJ (dg.f90:413,423) and the Williamson 2N update (vector.f90:163-183)
    if (live) {
#pragma unroll
        for (int m = 0; m < SEG; m++) {
            if (m < cnt) {
                const int k = a0 + m;
                const int node = c1 + n * c2 + n2 * k;
                const int id = TileV<n>::idx(c1, c2, k);
                const double msJ = P.noJac ? -1.0 : -sJv[m];  // noJac: k_overint applies the Jacobian after its filter
                double Ut[5];
                Ut[0] = S[S_RHO * SL + id] * msJ;
#pragma unroll
                for (int v = 0; v < 4; v++) Ut[1 + v] = S[(S_UT + v) * SL + id] * msJ;
                if (P.tcSource == 2) {  // channel forcing folded into this epilogue (TestcaseSource, testcase/channel/testcase.f90:277-296)
                    const double bulk = P.bulkDev ? __ldg(P.bulkDev) : P.tcBulkVel;
                    Ut[MOM1] = __dadd_rn(Ut[MOM1], -P.tcDpdx);   // two roundings, like Ut + src in k_source_rk
                    Ut[ENER] = __dadd_rn(Ut[ENER], -__dmul_rn(P.tcDpdx, bulk));
                }
                if (MODE == 0) {
                    double* o = P.Ut + (size_t)e * 5 * n3 + node;
#pragma unroll
                    for (int v = 0; v < 5; v++) o[v * n3] = Ut[v];
                } else {
                    double* ot = P.Ut_tmp + (size_t)e * 5 * n3 + node;
                    double* ou = P.U + (size_t)e * 5 * n3 + node;
#pragma unroll
                    for (int v = 0; v < 5; v++) {
                        const double r = (mRKA == 0.0) ? Ut[v] : Uto[m][v] * mRKA + Ut[v];
                        ot[v * n3] = r;
                        const double un = Uo[m][v] + r * b_dt;
                        ou[v * n3] = un;
                        S[(S_REC + v) * SL + id] = un;
                    }
                }
            }
        }
    }
    if (MODE == 1) {
        __syncthreads();
        // next-stage face states (GL: copy of the boundary layer), side-local orientation, coalesced in (p,q)
        if (live) {
#pragma unroll 1
            for (int f = tid; f < 6 * n2; f += T) {
                const int loc = f / n2 + 1;
                const int side = __ldg(&e2s[0 + 3 * (loc - 1)]) - 1;
                const int flip = __ldg(&e2s[1 + 3 * (loc - 1)]);
                int p, qq;
                face_lane<n>(P.S2V2, f - (loc - 1) * n2, flip, loc, (loc == XI_MINUS || loc == XI_PLUS), p, qq);
                const int pq = p + n * qq;
                const int a = s2v2<n>(P.S2V2, 0, p, qq, flip, loc);
                const int b = s2v2<n>(P.S2V2, 1, p, qq, flip, loc);
                const int l = is_minus(loc) ? 0 : n - 1;
                int id;
                if (loc == XI_MINUS || loc == XI_PLUS) id = TileV<n>::idx(l, a, b);
                else if (loc == ETA_MINUS || loc == ETA_PLUS) id = TileV<n>::idx(a, l, b);
                else id = TileV<n>::idx(a, b, l);
                double* dst = (flip == 0 ? P.UmNext : P.UsNext) + (size_t)side * 5 * n2 + pq;
#pragma unroll
                for (int v = 0; v < 5; v++) dst[v * n2] = S[(S_REC + v) * SL + id];
            }
        }
    }
}

}  // namespace dgx
