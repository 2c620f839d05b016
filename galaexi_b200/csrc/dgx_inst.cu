// Kernel instantiations for one polynomial degree (compile with -DDGX_N=<N>); see dgx_kernels.cuh.
#include "dgx_launch.h"
#include "dgx_volsurf2.cuh"
#include <stdlib.h>
#ifndef DGX_N
#error "compile with -DDGX_N=<polynomial degree>"
#endif
namespace dgx {
namespace {
constexpr int n = DGX_N + 1;
constexpr int n3 = n * n * n;

// k_volsurf2 is instantiated per split variant (SPLIT_DG 0..4), like the reference compiles one variant in
template <int VAR>
cudaError_t setup_vs2_one() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_volsurf2<n, 0, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs2_smem_bytes<n>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_volsurf2<n, 1, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs2_smem_bytes<n>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_volsurf2<n, 0, VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_volsurf2<n, 1, VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}
template <int VAR>
cudaError_t setup_vs2() {
    cudaError_t e = setup_vs2_one<VAR>();
    if (e != cudaSuccess) return e;
    if constexpr (VAR < 4) return setup_vs2<VAR + 1>();
    return e;
}
template <int MODE, int VAR>
void launch_vs2(const KParams& P, double mRKA, double b_dt, int nb, cudaStream_t s) {
    if (P.splitDG == VAR) {
        static int resident = 0;  // CTAs of one resident wave (prefetch distance)
        if (!resident) {
            int dev = 0, sms = 148, per = 1;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_volsurf2<n, MODE, VAR>, vs2_threads<n>(), vs2_smem_bytes<n>());
            resident = sms * (per > 0 ? per : 1);
        }
        const int grid = (nb + vs2_epb<n>() - 1) / vs2_epb<n>();
        k_volsurf2<n, MODE, VAR><<<grid, vs2_threads<n>(), vs2_smem_bytes<n>(), s>>>(P, nb, mRKA, b_dt, resident);
        return;
    }
    if constexpr (VAR < 4) launch_vs2<MODE, VAR + 1>(P, mRKA, b_dt, nb, s);
}

template <int NT>
struct L {
    static int setup() {
        cudaError_t e;
        e = cudaFuncSetAttribute(k_lifting<n, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lifting_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_lifting<n, NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lifting_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_volsurf<n, NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)volsurf_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_volsurf<n, NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)volsurf_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_source_rk<n, NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)source_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_filter<n, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_overint<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)overint_smem_bytes<n>());
        if (e != cudaSuccess) return (int)e;
        if (NT == 2) e = setup_vs2<0>();
        return (int)e;
    }
    static void prolong(const KParams& P, int nb, cudaStream_t s) {
        if (nb > 0) k_prolong<n, NT><<<nb, n3, 0, s>>>(P);
    }
    static void lifting(const KParams& P, int nb, cudaStream_t s) {
        if (nb <= 0) return;
        static int resident = 0;
        if (!resident) {
            int dev = 0, sms = 148, per = 1;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lifting<n, NT>, n3, lifting_smem_bytes<n>());
            resident = sms * (per > 0 ? per : 1);
        }
        if (P.lifting == 2 || P.MortarType || P.liftWeak || P.liftCons) k_lifting<n, NT, 1><<<nb, n3, lifting_smem_bytes<n>(), s>>>(P, resident);
        else k_lifting<n, NT><<<nb, n3, lifting_smem_bytes<n>(), s>>>(P, resident);
    }
    static void umortar(double* am, double* as, int nvar, const MortarParams& mp, int nBig, cudaStream_t s) {
        if (nBig > 0) k_umortar<n><<<dim3(nBig, nvar), n * n, 0, s>>>(am, as, nvar, mp);
    }
    static void fluxmortar(double* F, int nvar, int weak, const MortarParams& mp, int nBig, cudaStream_t s) {
        if (nBig > 0) k_fluxmortar<n><<<dim3(nBig, nvar), n * n, 0, s>>>(F, nvar, weak, mp);
    }
    static void mortar_liftflux(const KParams& P, const MortarParams& mp, int nBig, cudaStream_t s) {
        if (nBig > 0) k_mortar_liftflux<n><<<dim3(nBig, 12), n * n, 0, s>>>(P, mp);
    }
    static void sideflux(const KParams& P, int side0, int nS, cudaStream_t s) {
        if (nS <= 0) return;
        const long long tot = (long long)nS * n * n;
        k_sideflux<n><<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(P, side0, nS);
    }
    static void volsurf(const KParams& P, int mode, double mRKA, double b_dt, int nb, cudaStream_t s) {
        if (nb <= 0) return;
        // split form on Gauss-Lobatto nodes: register-blocked kernel (DGX_VOLSURF=1 selects the thread-per-node one)
        static const bool v1 = getenv("DGX_VOLSURF") && atoi(getenv("DGX_VOLSURF")) == 1;
        if (NT == 2 && P.splitDG >= 0 && !v1) {
            if (mode == 0) launch_vs2<0, 0>(P, mRKA, b_dt, nb, s);
            else launch_vs2<1, 0>(P, mRKA, b_dt, nb, s);
            return;
        }
        if (mode == 0) k_volsurf<n, NT, 0><<<nb, n3, volsurf_smem_bytes<n>(), s>>>(P, mRKA, b_dt);
        else k_volsurf<n, NT, 1><<<nb, n3, volsurf_smem_bytes<n>(), s>>>(P, mRKA, b_dt);
    }
    static void filter(const KParams& P, int nb, cudaStream_t s) {
        if (nb > 0) k_filter<n, NT><<<nb, n3, filter_smem_bytes<n>(), s>>>(P);
    }
    static void source_rk(const KParams& P, int mode, double t, double mRKA, double b_dt, int nb, cudaStream_t s) {
        if (nb <= 0) return;
        if (mode == 0) k_source_rk<n, NT, 0><<<nb, n3, source_smem_bytes<n>(), s>>>(P, t, mRKA, b_dt);
        else k_source_rk<n, NT, 1><<<nb, n3, source_smem_bytes<n>(), s>>>(P, t, mRKA, b_dt);
    }
    static void overint(const KParams& P, double t, int nb, cudaStream_t s) {
        if (nb > 0) k_overint<n><<<nb, n3, overint_smem_bytes<n>(), s>>>(P, t);
    }
    static void bulkvel(const KParams& P, const double* wGP, double* partials, cudaStream_t s) {
        if (P.nElems > 0) k_bulkvel<n><<<P.nElems, timestep_threads<n>(), 0, s>>>(P, wGP, partials);
    }
    static int tgv_analyze(const KParams& P, int NA1, const double* Vdm, const double* wA, double* partials, cudaStream_t s) {
        if (P.nElems <= 0) return 0;
        const size_t sm = tgv_smem_bytes<n>(NA1);
        cudaError_t e = cudaFuncSetAttribute(k_tgv_analyze<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        k_tgv_analyze<n><<<P.nElems, TGV_THREADS, sm, s>>>(P, NA1, Vdm, wA, partials);
        return 0;
    }
    static void timestep(const KParams& P, double CFL, double DFL, double* out, cudaStream_t s) {
        if (P.nElems > 0) k_timestep<n><<<P.nElems, timestep_threads<n>(), 0, s>>>(P, CFL, DFL, out);
    }
};
const KernelTable tabG = {L<1>::setup, L<1>::prolong, L<1>::lifting, L<1>::sideflux, L<1>::volsurf, L<1>::timestep,
                          L<1>::umortar, L<1>::fluxmortar, L<1>::mortar_liftflux, L<1>::filter, L<1>::source_rk, L<1>::overint, L<1>::bulkvel, L<1>::tgv_analyze};
const KernelTable tabGL = {L<2>::setup, L<2>::prolong, L<2>::lifting, L<2>::sideflux, L<2>::volsurf, L<2>::timestep,
                           L<2>::umortar, L<2>::fluxmortar, L<2>::mortar_liftflux, L<2>::filter, L<2>::source_rk, L<2>::overint, L<2>::bulkvel, L<2>::tgv_analyze};
}  // namespace

#define DGX_CAT2(a, b) a##b
#define DGX_CAT(a, b) DGX_CAT2(a, b)
const KernelTable* DGX_CAT(kernel_table_N, DGX_N)(int nodeType) { return nodeType == 2 ? &tabGL : &tabG; }
}  // namespace dgx
