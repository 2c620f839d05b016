// Non-conforming (mortar) interfaces on the device. GALAEXI has no GPU mortar path (it aborts on mortar meshes,
// mesh/mesh.f90:140-143); these kernels follow the inherited host FLEXI routines (paths relative to /root/reference/src):
//   k_umortar          mortar/fillmortar.t90:34-195   U_Mortar: big-side data -> its 4 / 2 small sides (M_0_1, M_0_2)
//   k_fluxmortar       mortar/fillmortar.t90:215-376  Flux_Mortar: small-side fluxes -> big side (M_1_0, M_2_0)
//   k_mortar_liftflux  the same projection applied to the BR1/BR2 lifting flux times the normal vector, which the fused
//                      lifting kernel never materialises on the sides (dg/lifting/lifting_br2.t90:102-106 pattern)
// One CTA per (big side, variable), n^2 threads: mortar sides are a tiny fraction of a mesh, so these kernels are
// launch-latency bound and only run when the mesh has mortars. Summation order as in the reference (l = 0 first).
#pragma once
#include "dgx_kernels.cuh"

namespace dgx {

struct MortarParams {
    const int* MortarType;  // (2,nSides): type 1..3 for big sides, index into MortarInfo
    const int* MortarInfo;  // (2,4,nMortarSides): SideID (1-based), flip of the small sides
    const double* M;        // [4][n*n]: M_0_1, M_0_2, M_1_0, M_2_0, each Fortran (l,p) at [l + n*p] (stored transposed)
    int side0;              // first big side of the range (0-based)
};

// FS2M(1:2,p,q,flip), mesh/mappings.f90:237-268
template <int n>
__device__ __forceinline__ void fs2m(int p, int q, int flip, int& pm, int& qm) {
    switch (flip) {
        case 1: pm = q; qm = p; break;
        case 2: pm = n - 1 - p; qm = q; break;
        case 3: pm = n - 1 - q; qm = n - 1 - p; break;
        case 4: pm = p; qm = n - 1 - q; break;
        default: pm = p; qm = q;
    }
}

// out(p) = sum_l M(l,p) in(l) along the first (dir 0) or second (dir 1) side index
template <int n>
__device__ __forceinline__ double mortar_line(const double* __restrict__ M, const double* __restrict__ in, int dir, int p, int q) {
    const int o = dir == 0 ? p : q;
    double a = M[0 + n * o] * in[dir == 0 ? (0 + n * q) : (p + n * 0)];
#pragma unroll
    for (int l = 1; l < n; l++) a = a + M[l + n * o] * in[dir == 0 ? (l + n * q) : (p + n * l)];
    return a;
}
template <int n>
__device__ __forceinline__ double mortar_line_pair(const double* __restrict__ M1, const double* __restrict__ M2, const double* __restrict__ a,
                                                   const double* __restrict__ b, int dir, int p, int q) {
    const int o = dir == 0 ? p : q;
    int i0 = dir == 0 ? (0 + n * q) : (p + n * 0);
    double t = M1[0 + n * o] * a[i0] + M2[0 + n * o] * b[i0];
#pragma unroll
    for (int l = 1; l < n; l++) {
        const int il = dir == 0 ? (l + n * q) : (p + n * l);
        t = t + M1[l + n * o] * a[il] + M2[l + n * o] * b[il];
    }
    return t;
}

// projection of the four / two small-side fields held in sIn[im][n2] onto the big side (value for node (p,q))
template <int n>
__device__ __forceinline__ double mortar_project(int type, const double* sM10, const double* sM20, double (*sIn)[n * n], double (*sT2)[n * n], int p,
                                                 int q) {
    const int t = p + n * q;
    if (type == 1) {
        sT2[0][t] = mortar_line_pair<n>(sM10, sM20, sIn[0], sIn[1], 0, p, q);
        sT2[1][t] = mortar_line_pair<n>(sM10, sM20, sIn[2], sIn[3], 0, p, q);
        __syncthreads();
        return mortar_line_pair<n>(sM10, sM20, sT2[0], sT2[1], 1, p, q);
    }
    return mortar_line_pair<n>(sM10, sM20, sIn[0], sIn[1], type == 2 ? 1 : 0, p, q);
}

template <int n>
__global__ void __launch_bounds__(n* n) k_umortar(double* __restrict__ am, double* __restrict__ as, int nvar, const MortarParams mp) {
    constexpr int n2 = n * n;
    __shared__ double sBig[n2], sT2[2][n2], sOut[4][n2], sM1[n2], sM2[n2];
    const int sd = mp.side0 + blockIdx.x, v = blockIdx.y, t = threadIdx.x;
    const int type = mp.MortarType[2 * sd], iSide = mp.MortarType[2 * sd + 1];
    sM1[t] = mp.M[0 * n2 + t];
    sM2[t] = mp.M[1 * n2 + t];
    sBig[t] = am[((size_t)sd * nvar + v) * n2 + t];
    __syncthreads();
    const int q = t / n, p = t - q * n;
    const int nMortars = type == 1 ? 4 : 2;
    if (type == 1) {
        sT2[0][t] = mortar_line<n>(sM1, sBig, 1, p, q);
        sT2[1][t] = mortar_line<n>(sM2, sBig, 1, p, q);
        __syncthreads();
        sOut[0][t] = mortar_line<n>(sM1, sT2[0], 0, p, q);
        sOut[1][t] = mortar_line<n>(sM2, sT2[0], 0, p, q);
        sOut[2][t] = mortar_line<n>(sM1, sT2[1], 0, p, q);
        sOut[3][t] = mortar_line<n>(sM2, sT2[1], 0, p, q);
    } else {
        const int dir = type == 2 ? 1 : 0;
        sOut[0][t] = mortar_line<n>(sM1, sBig, dir, p, q);
        sOut[1][t] = mortar_line<n>(sM2, sBig, dir, p, q);
    }
    __syncthreads();
    for (int im = 0; im < nMortars; im++) {
        const int SideID = mp.MortarInfo[0 + 2 * (im + 4 * (iSide - 1))], flip = mp.MortarInfo[1 + 2 * (im + 4 * (iSide - 1))];
        int pm, qm;
        fs2m<n>(p, q, flip, pm, qm);
        double* dst = (flip == 0 ? am : as) + ((size_t)(SideID - 1) * nvar + v) * n2;
        dst[t] = sOut[im][pm + n * qm];
    }
}

// Fm (in place): big side <- projection of its small sides; small slave sides (flip>0) through FS2M, negated if weak
template <int n>
__global__ void __launch_bounds__(n* n) k_fluxmortar(double* __restrict__ F, int nvar, int weak, const MortarParams mp) {
    constexpr int n2 = n * n;
    __shared__ double sIn[4][n2], sT2[2][n2], sM1[n2], sM2[n2];
    const int sd = mp.side0 + blockIdx.x, v = blockIdx.y, t = threadIdx.x;
    const int type = mp.MortarType[2 * sd], iSide = mp.MortarType[2 * sd + 1];
    sM1[t] = mp.M[2 * n2 + t];
    sM2[t] = mp.M[3 * n2 + t];
    const int q = t / n, p = t - q * n;
    const int nMortars = type == 1 ? 4 : 2;
    for (int im = 0; im < nMortars; im++) {
        const int SideID = mp.MortarInfo[0 + 2 * (im + 4 * (iSide - 1))], flip = mp.MortarInfo[1 + 2 * (im + 4 * (iSide - 1))];
        int pm, qm;
        fs2m<n>(p, q, flip, pm, qm);
        const double f = F[((size_t)(SideID - 1) * nvar + v) * n2 + pm + n * qm];
        sIn[im][t] = (flip != 0 && weak) ? -f : f;
    }
    __syncthreads();
    F[((size_t)sd * nvar + v) * n2 + t] = mortar_project<n>(type, sM1, sM2, sIn, sT2, p, q);
}

// Lifting flux times normal on the small sides, 1/2 (U_s - U_m) SurfElem n_d for the lifted variables (u,v,w,T),
// projected onto the big side and stored in gm[bigSide][d*4+v] (read back by the big element's k_lifting CTA, which
// afterwards overwrites it with the gradient trace). blockIdx.y = d*4+v.
template <int n>
__global__ void __launch_bounds__(n* n) k_mortar_liftflux(const KParams P, const MortarParams mp) {
    constexpr int n2 = n * n;
    __shared__ double sIn[4][n2], sT2[2][n2], sM1[n2], sM2[n2];
    const int sd = mp.side0 + blockIdx.x, x = blockIdx.y, t = threadIdx.x;
    const int d = x / 4, v = x - 4 * d;
    const int type = mp.MortarType[2 * sd], iSide = mp.MortarType[2 * sd + 1];
    sM1[t] = mp.M[2 * n2 + t];
    sM2[t] = mp.M[3 * n2 + t];
    const int q = t / n, p = t - q * n;
    const int nMortars = type == 1 ? 4 : 2;
    const Eos eos = P.eos;
    for (int im = 0; im < nMortars; im++) {
        const int side = mp.MortarInfo[0 + 2 * (im + 4 * (iSide - 1))] - 1, flip = mp.MortarInfo[1 + 2 * (im + 4 * (iSide - 1))];
        int pm, qm;
        fs2m<n>(p, q, flip, pm, qm);
        const int pq = pm + n * qm;
        const double* g = P.geo + (size_t)side * 10 * n2 + pq;
        double Um[5], Us[5], Pm[6], Ps[6];
#pragma unroll
        for (int c = 0; c < 5; c++) {
            Um[c] = P.Um[(size_t)side * 5 * n2 + c * n2 + pq];
            Us[c] = P.Us[(size_t)side * 5 * n2 + c * n2 + pq];
        }
        cons_to_prim(Pm, Um, eos);
        cons_to_prim(Ps, Us, eos);
        const int iv = (v < 3) ? VEL1 + v : TEMP;
        const double Fl = 0.5 * g[9 * n2] * ((P.liftWeak ? 1.0 : -1.0) * Pm[iv] + Ps[iv]);
        const double val = Fl * g[d * n2];
        sIn[im][t] = (P.liftWeak && flip != 0) ? -val : val;  // Flux_MortarLifting(weak=doWeakLifting)
    }
    __syncthreads();
    P.gm[((size_t)sd * 12 + x) * n2 + t] = mortar_project<n>(type, sM1, sM2, sIn, sT2, p, q);
}

}  // namespace dgx
