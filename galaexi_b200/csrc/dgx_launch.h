// Launch table: one entry per (N, node type), instantiated in dgx_inst.cu (compiled once per N with -DDGX_N=<N>).
#pragma once
#include <cuda_runtime.h>
#include "dgx_kernels.cuh"
#include "dgx_mortar.cuh"
#include "dgx_analyze.cuh"

namespace dgx {
struct KernelTable {
    int (*setup)();  // opt-in to large dynamic shared memory; returns cudaError_t
    void (*prolong)(const KParams&, int nBlocks, cudaStream_t);
    void (*lifting)(const KParams&, int nBlocks, cudaStream_t);
    void (*sideflux)(const KParams&, int side0, int nS, cudaStream_t);
    void (*volsurf)(const KParams&, int mode, double mRKA, double b_dt, int nBlocks, cudaStream_t);
    void (*timestep)(const KParams&, double CFL, double DFL, double* out, cudaStream_t);
    // non-conforming interfaces (nBig big mortar sides starting at mp.side0)
    void (*umortar)(double* am, double* as, int nvar, const MortarParams& mp, int nBig, cudaStream_t);
    void (*fluxmortar)(double* F, int nvar, int weak, const MortarParams& mp, int nBig, cudaStream_t);
    void (*mortar_liftflux)(const KParams&, const MortarParams& mp, int nBig, cudaStream_t);
    void (*filter)(const KParams&, int nBlocks, cudaStream_t);  // FilterType > 0: U <- FilterMat U, faces of the filtered state
    void (*source_rk)(const KParams&, int mode, double t, double mRKA, double b_dt, int nBlocks, cudaStream_t);  // CalcSource path
    void (*overint)(const KParams&, double t, int nBlocks, cudaStream_t);  // step 14 with overintegration
    void (*bulkvel)(const KParams&, const double* wGP, double* partials, cudaStream_t);  // channel CalcForcing
    // TGV diagnostics: per-element partials [nElems][TGV_NPART]; returns a cudaError_t
    int (*tgv_analyze)(const KParams&, int NA1, const double* Vdm, const double* wA, double* partials, cudaStream_t);
};
const KernelTable* kernel_table(int N, int nodeType);  // nullptr if this N was not compiled in
}  // namespace dgx

#define DGX_DECLARE_TABLE(NN) namespace dgx { const KernelTable* kernel_table_N##NN(int nodeType); }
DGX_DECLARE_TABLE(1) DGX_DECLARE_TABLE(2) DGX_DECLARE_TABLE(3) DGX_DECLARE_TABLE(4) DGX_DECLARE_TABLE(5)
DGX_DECLARE_TABLE(6) DGX_DECLARE_TABLE(7) DGX_DECLARE_TABLE(8) DGX_DECLARE_TABLE(9)
