// Taylor-Green vortex diagnostics on the device: AnalyzeTestcase of testcase/taylorgreenvortex/testcase.f90:283-515
// (SURVEY.md 8f rank 2). The reference copies U and the three gradient arrays to the host at every analyze step
// (testcase.f90:361-364) and integrates there; here one CTA per element interpolates the 15 fields it needs (U 5, velocity
// gradients 9, sJ 1) to the (NAnalyze+1)^3 Gauss-Lobatto analysis nodes plane by plane (sum factorisation in shared
// memory, nothing written to HBM), evaluates the integrands and leaves 12 partial results per element; a second,
// single-CTA kernel reduces them in a fixed order (deterministic for a given partition).
#pragma once
#include "dgx_kernels.cuh"

namespace dgx {

constexpr int TGV_NPART = 12;  // T_mean, Entropy, Ekin, Ekin_comp, Enstr, DR_u, DR_S, DR_Sd, DR_p, ED_S, ED_D (sums), max vorticity
constexpr int TGV_THREADS = 256;

template <int n>
size_t tgv_smem_bytes(int NA1) { return sizeof(double) * (size_t)(15 * n * n * n + 15 * n * n + 15 * n * NA1 + NA1 * n + NA1 + TGV_NPART * (TGV_THREADS / 32)); }

template <int n>
__global__ void __launch_bounds__(TGV_THREADS) k_tgv_analyze(const KParams P, int NA1, const double* __restrict__ Vdm /* (0:NA,0:N) Fortran: [I + NA1*i] */,
                                                             const double* __restrict__ wA, double* __restrict__ partials) {
    constexpr int n2 = n * n, n3 = n2 * n;
    extern __shared__ double smem[];
    double* f = smem;                 // [15][n3]: U 0..4, du_i/dx_j at 5 + 3*j + i, sJ at 14
    double* g = f + 15 * n3;          // [15][n2]   zeta-interpolated plane
    double* hh = g + 15 * n2;         // [15][n][NA1]
    double* sV = hh + 15 * n * NA1;   // [NA1][n] as V[I*n + i]
    double* sW = sV + NA1 * n;        // [NA1]
    double* red = sW + NA1;           // [TGV_NPART][warps]
    const int e = blockIdx.x, t = threadIdx.x;
    const Eos eos = P.eos;
    for (int x = t; x < NA1 * n; x += TGV_THREADS) { const int I = x / n, i = x - I * n; sV[x] = Vdm[I + NA1 * i]; }
    for (int x = t; x < NA1; x += TGV_THREADS) sW[x] = wA[x];
    for (int x = t; x < 5 * n3; x += TGV_THREADS) f[x] = P.U[(size_t)e * 5 * n3 + x];
    for (int x = t; x < 9 * n3; x += TGV_THREADS) {
        const int c = x / n3, node = x - c * n3, j = c / 3, i = c - 3 * j;  // field 5 + 3 j + i = d u_i / d x_j = gradU[j*4 + i]
        f[5 * n3 + x] = P.gradU[((size_t)e * 12 + j * 4 + i) * n3 + node];
    }
    for (int x = t; x < n3; x += TGV_THREADS) f[14 * n3 + x] = P.sJ[(size_t)e * n3 + x];
    double acc[TGV_NPART];
#pragma unroll
    for (int x = 0; x < TGV_NPART; x++) acc[x] = 0.0;
    __syncthreads();
    for (int K = 0; K < NA1; K++) {
        for (int x = t; x < 15 * n2; x += TGV_THREADS) {
            const int c = x / n2, ij = x - c * n2;
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < n; k++) a += sV[K * n + k] * f[c * n3 + k * n2 + ij];
            g[x] = a;
        }
        __syncthreads();
        for (int x = t; x < 15 * n * NA1; x += TGV_THREADS) {
            const int c = x / (n * NA1), r = x - c * n * NA1, j = r / NA1, I = r - j * NA1;
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) a += sV[I * n + i] * g[c * n2 + j * n + i];
            hh[x] = a;
        }
        __syncthreads();
        for (int pt = t; pt < NA1 * NA1; pt += TGV_THREADS) {
            const int J = pt / NA1, I = pt - J * NA1;
            double v[15];
#pragma unroll
            for (int c = 0; c < 15; c++) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < n; j++) a += sV[J * n + j] * hh[(c * n + j) * NA1 + I];
                v[c] = a;
            }
            double Pr[6];
            cons_to_prim(Pr, v, eos);
            const double* G = v + 5;  // G[3*j + i] = d u_i / d x_j
            const double divU = G[0] + G[4] + G[8];
            double uu = 0.0, SS = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double gij = G[3 * j + i], s = 0.5 * (gij + G[3 * i + j]);
                    uu += gij * gij;
                    SS += s * s;
                }
            // Sd = S - divU/3 I  ->  Sd:Sd = S:S - 2/3 divU tr(S) + 3 (divU/3)^2, evaluated term by term like the reference
            double SdSd = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double s = 0.5 * (G[3 * j + i] + G[3 * i + j]);
                    if (i == j) s -= 1.0 / 3.0 * divU;
                    SdSd += s * s;
                }
            const double w1 = G[3 * 1 + 2] - G[3 * 2 + 1];  // du_3/dx_2 - du_2/dx_3
            const double w2 = G[3 * 2 + 0] - G[3 * 0 + 2];
            const double w3 = G[3 * 0 + 1] - G[3 * 1 + 0];
            const double ww = w1 * w1 + w2 * w2 + w3 * w3;
            const double F = sW[I] * sW[J] * sW[K] / v[14];
            const double v2 = Pr[VEL1] * Pr[VEL1] + Pr[VEL2] * Pr[VEL2] + Pr[VEL3] * Pr[VEL3];
            const double mu = viscosity(eos, Pr[TEMP]);
            acc[0] += F * Pr[TEMP];
            acc[1] += F * (-1.0 / (eos.kappa - 1.0)) * Pr[DENS] * (log(Pr[PRES]) - eos.kappa * log(Pr[DENS]));
            acc[2] += F * 0.5 * v2;
            acc[3] += F * 0.5 * Pr[DENS] * v2;
            acc[4] += F * 0.5 * Pr[DENS] * ww;
            acc[5] += F * uu;
            acc[6] += F * SS;
            acc[7] += F * SdSd;
            acc[8] += F * Pr[PRES] * divU;
            acc[9] += F * mu * ww;
            acc[10] += F * mu * divU * divU;
            acc[11] = fmax(acc[11], sqrt(ww));
        }
        __syncthreads();
    }
    // block reduction (fixed order: lanes by xor-shuffle, then warps in order)
#pragma unroll
    for (int x = 0; x < TGV_NPART; x++) {
        double v = acc[x];
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = (x == TGV_NPART - 1) ? fmax(v, u) : v + u;
        }
        if ((t & 31) == 0) red[x * (TGV_THREADS / 32) + (t >> 5)] = v;
    }
    __syncthreads();
    if (t < TGV_NPART) {
        double v = red[t * (TGV_THREADS / 32)];
        for (int w = 1; w < TGV_THREADS / 32; w++) v = (t == TGV_NPART - 1) ? fmax(v, red[t * (TGV_THREADS / 32) + w]) : v + red[t * (TGV_THREADS / 32) + w];
        partials[(size_t)e * TGV_NPART + t] = v;
    }
}

// out[x] = reduce over elements of partials[e][x], one CTA, fixed order
static __global__ void __launch_bounds__(TGV_THREADS) k_tgv_reduce(const double* __restrict__ partials, int nElems, double* __restrict__ out) {
    __shared__ double red[TGV_THREADS];
    for (int x = 0; x < TGV_NPART; x++) {
        const bool mx = (x == TGV_NPART - 1);
        double v = 0.0;
        for (int e = threadIdx.x; e < nElems; e += TGV_THREADS) v = mx ? fmax(v, partials[(size_t)e * TGV_NPART + x]) : v + partials[(size_t)e * TGV_NPART + x];
        red[threadIdx.x] = v;
        __syncthreads();
        for (int s = TGV_THREADS / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] = mx ? fmax(red[threadIdx.x], red[threadIdx.x + s]) : red[threadIdx.x] + red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[x] = red[0];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// CalcBodyForces (equations/navierstokes/calcbodyforces.f90:41-205): pressure force sum p n wGPSurf SurfElem and friction
// force -sum tau n wGPSurf SurfElem on the wall boundary sides (BC types 3, 4, 9: analyze_equation.f90:113-121), and
// CalcWallVelocity (analyze_equation.f90:435-499): sum |v| wGPSurf SurfElem, max |v|, min |v| on the same sides -- from the
// face states and the lifted gradient traces of the last DGTimeDerivative_weakForm. One CTA per boundary side writes the
// WALL_NPART values of that side (neutral elements for a non-wall side); k_wall_reduce combines them per boundary condition
// in a fixed order: out[x * nBCs + iBC], x = 0..2 Fp, 3..5 Fv, 6 sum |v| dA, 7 max |v|, 8 min |v|.
constexpr int BF_THREADS = 128;
constexpr int WALL_NPART = 9;
constexpr double WALL_HUGE = 1.e14;  // the reference's initial values of minV / maxV
__device__ __forceinline__ double wall_combine(int x, double a, double b) { return x < 7 ? a + b : (x == 7 ? fmax(a, b) : fmin(a, b)); }
static __global__ void __launch_bounds__(BF_THREADS) k_wall_sides(const KParams P, int n, const double* __restrict__ Um, const double* __restrict__ wGP,
                                                                  double* __restrict__ partials) {
    __shared__ double red[WALL_NPART * (BF_THREADS / 32)];
    const int side = blockIdx.x, t = threadIdx.x, n2 = n * n;
    const int type = P.BCSides[2 * side];
    double acc[WALL_NPART] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -WALL_HUGE, WALL_HUGE};
    if (type == 3 || type == 4 || type == 9) {
        const Eos eos = P.eos;
        for (int pq = t; pq < n2; pq += BF_THREADS) {
            const int q = pq / n, p = pq - q * n;
            const double* g = P.geo + (size_t)side * 10 * n2 + pq;
            const double nv[3] = {g[0], g[n2], g[2 * n2]};
            const double dA = wGP[p] * wGP[q] * g[9 * n2];
            double U[5], Pr[6];
#pragma unroll
            for (int c = 0; c < 5; c++) U[c] = Um[((size_t)side * 5 + c) * n2 + pq];
            cons_to_prim(Pr, U, eos);
#pragma unroll
            for (int d = 0; d < 3; d++) acc[d] += Pr[PRES] * nv[d] * dA;
            const double locV = sqrt(Pr[VEL1] * Pr[VEL1] + Pr[VEL2] * Pr[VEL2] + Pr[VEL3] * Pr[VEL3]);
            acc[6] += locV * dA;
            acc[7] = fmax(acc[7], locV);
            acc[8] = fmin(acc[8], locV);
            if (P.parabolic) {
                const double mu = viscosity(eos, Pr[TEMP]);
                double G[3][3];  // G[i][d] = d v_i / d x_d
#pragma unroll
                for (int d = 0; d < 3; d++)
#pragma unroll
                    for (int i = 0; i < 3; i++) G[i][d] = P.gm[((size_t)side * 12 + d * 4 + i) * n2 + pq];
                const double div = G[0][0] + G[1][1] + G[2][2];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    double f = 0.0;
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        double tau = mu * (G[i][d] + G[d][i]);
                        if (i == d) tau -= 2.0 / 3.0 * mu * div;
                        f += tau * nv[d];
                    }
                    acc[3 + i] -= f * dA;  // force acting on the wall
                }
            }
        }
    }
#pragma unroll
    for (int x = 0; x < WALL_NPART; x++) {
        double v = acc[x];
        for (int o = 16; o > 0; o >>= 1) v = wall_combine(x, v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((t & 31) == 0) red[x * (BF_THREADS / 32) + (t >> 5)] = v;
    }
    __syncthreads();
    if (t < WALL_NPART) {
        double v = red[t * (BF_THREADS / 32)];
        for (int w = 1; w < BF_THREADS / 32; w++) v = wall_combine(t, v, red[t * (BF_THREADS / 32) + w]);
        partials[(size_t)side * WALL_NPART + t] = v;
    }
}
// one CTA per boundary condition: combine the sides with BC(side) == iBC+1
static __global__ void __launch_bounds__(BF_THREADS) k_wall_reduce(const double* __restrict__ partials, const int* __restrict__ BC, int nBCSides,
                                                                   int nBCs, double* __restrict__ out) {
    __shared__ double red[BF_THREADS];
    const int iBC = blockIdx.x + 1;
    for (int x = 0; x < WALL_NPART; x++) {
        double v = x < 7 ? 0.0 : (x == 7 ? -WALL_HUGE : WALL_HUGE);
        for (int sd = threadIdx.x; sd < nBCSides; sd += BF_THREADS)
            if (BC[sd] == iBC) v = wall_combine(x, v, partials[(size_t)sd * WALL_NPART + x]);
        red[threadIdx.x] = v;
        __syncthreads();
        for (int s = BF_THREADS / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] = wall_combine(x, red[threadIdx.x], red[threadIdx.x + s]);
            __syncthreads();
        }
        if (threadIdx.x == 0) out[(size_t)x * nBCs + blockIdx.x] = red[0];
        __syncthreads();
    }
}

}  // namespace dgx
