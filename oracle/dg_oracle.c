/*
 * dg_oracle.c -- CPU restatement of GALAEXI's DGSEM right-hand side + low-storage RK stage.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity oracle for the CUDA library in galaexi_b200/csrc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it. The product path never routes through it.
 *
 * The reference (flexi-framework/galaexi @1119a65, CUDA Fortran) cannot be compiled in this
 * environment (needs nvfortran + HDF5 + MPI), so this is a plain-C restatement of its arithmetic in the
 * reference's own array layout (variable index fastest, U(nVar,i,j,k,iElem)) and operation order.
 * Every routine cites the reference file:line it follows (paths relative to /root/reference/src).
 *
 * Pinned against (tests/test_oracle_goldens.py): unitTests/ProlongToFace_{G,GL}3D.bin,
 * unitTests/SurfInt_{G,GL}3D.bin (100 eps), regressioncheck tgv/split CSV rows 1-2 (IC gradients,
 * first dt, kinetic energy after one RK step), parabolic/cavity_3D reference state.
 *
 * Scope of the restatement: 3D, conforming meshes, PP_nVar=5, PP_nVarPrim=6, PP_nVarLifting=5
 * (PP_OPTLIFT=0), BR1 lifting in strong non-conservative form (the GALAEXI defaults, lifting.f90:81-85),
 * node types Gauss / Gauss-Lobatto, weak form or split form (SD, MO, DU, KG, PI), Riemann solvers LF, Roe, RoeL2,
 * RoeEntropyFix, HLL, HLLC, HLLE, HLLEM, FluxAverage, constant or Sutherland viscosity, BC types 2, 3, 4, 9, 91,
 * 23, 24, 25, 27.
 *
 * Two builds of this one source (oracle/Makefile): libdgoracle.so (FP64, no FMA contraction: the reference's
 * arithmetic) and libdgoracle_ld.so (-DDGO_EXTENDED: the same formulas and the same FP64 constants
 * evaluated in 80-bit extended precision). The second one measures how far FP64 round-off moves a result, so
 * that ill-conditioned comparisons (e.g. the low-Mach TGV initial residual, dominated by cancellation) can be
 * judged against the exact value of the reference's formulas instead of against one particular rounding order.
 *
 * OpenMP over elements / sides is only used to time the CPU baseline on all host cores; results do not
 * depend on the thread count (every loop iteration owns its outputs, as in the reference's gather kernels).
 */
#include <tgmath.h> /* type-generic sqrt/pow/fabs/fmax: the same source also builds the extended-precision variant */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef DGO_EXTENDED
#define double long double /* after the system headers: every FP64 quantity below becomes 80-bit extended */
#endif

#define NV 5   /* PP_nVar        */
#define NP 6   /* PP_nVarPrim    */
#define NL 5   /* PP_nVarLifting */

/* eos.h index macros, 0-based */
enum { DENS = 0, MOM1, MOM2, MOM3, ENER };
enum { VEL1 = 1, VEL2, VEL3, PRES, TEMP };
static const int PRIM_LIFT[NL] = {0, 1, 2, 3, 5}; /* eos.h:143 */
enum { LIFT_DENS = 0, LIFT_VEL1, LIFT_VEL2, LIFT_VEL3, LIFT_TEMP };
/* extended state U_LL/U_RR (eos.h:116-131): cons 0..4, then sRho, vel1..3, pres, temp */
enum { E_DENS = 0, E_MOM1, E_MOM2, E_MOM3, E_ENER, E_SRHO, E_VEL1, E_VEL2, E_VEL3, E_PRES, E_TEMP, NEXT };
enum { EOS_KAPPA = 0, EOS_R, EOS_PR, EOS_MU0, EOS_TS, EOS_TREF, EOS_EXPOSUTH, EOS_CSUTH };

/* local sides, flexi.h:97-102 */
enum { ZETA_MINUS = 1, ETA_MINUS = 2, XI_PLUS = 3, ETA_PLUS = 4, XI_MINUS = 5, ZETA_PLUS = 6 };

typedef struct {
    int N, nElems, nSides;
    int nBCSides, firstInnerSide, lastInnerSide;
    int firstMPISide_MINE, lastMPISide_MINE, firstMPISide_YOUR, lastMPISide_YOUR;
    int nodeType;  /* 1 Gauss, 2 Gauss-Lobatto (PP_NodeType) */
    int splitDG;   /* -1 off; 0 SD, 1 MO, 2 DU, 3 KG, 4 PI (SPLIT_DG) */
    int riemann;   /* 0 LF, 1 Roe, 2 RoeL2, 3 RoeEntropyFix, 4 HLL, 5 HLLC, 6 HLLE, 7 HLLEM (RIEMANN); 9 FluxAverage */
    int parabolic; /* PARABOLIC */
    int viscLaw;   /* PP_VISC: 0 constant, 1 Sutherland */
    int nRefState;
    double EOS[8];
    /* operators, Fortran column-major (0:N,0:N): M(a,b) at [a + n*b] */
    const double *D_T, *D_Hat_T, *DVolSurf, *L_Minus, *L_Plus, *L_HatMinus, *L_HatPlus;
    /* mesh tables (Fortran memory) */
    const int *ElemToSide;  /* (3,6,nElems) */
    const int *S2V2;        /* (2,0:N,0:N,0:4,1:6) */
    const int *S2V2_inv;
    const int *BCSides;     /* (2,nBCSides): BC_TYPE, BC_STATE */
    const double *Metrics_fTilde, *Metrics_gTilde, *Metrics_hTilde; /* (3,n,n,n,nElems) */
    const double *sJ;       /* (n,n,n,nElems) */
    const double *NormVec, *TangVec1, *TangVec2; /* (3,n,n,nSides) */
    const double *SurfElem; /* (n,n,nSides) */
    const double *RefStatePrim; /* (6,nRefState) */
    /* lifting variant: 1 BR1 (GALAEXI), 2 BR2 (host FLEXI code, dg/lifting/lifting_br2.t90); penalties lifting.f90:86-91 */
    int lifting;
    double etaBR2, etaBR2_wall;
    /* non-conforming interfaces (host FLEXI code, mortar/): side ranges mesh.f90:271-283, MortarType(2,nSides),
     * MortarInfo(2,4,nMortarSides), FS2M(2,0:N,0:N,0:4) mappings.f90:180-189, 1-D operators mortar.f90 (stored transposed) */
    int firstMortarInnerSide, lastMortarInnerSide, firstMortarMPISide, lastMortarMPISide;
    const int *MortarType, *MortarInfo, *FS2M;
    const int *SideToElem; /* (5,nSides) */
    const double *M_0_1, *M_0_2, *M_1_0, *M_2_0;
    /* modal filter applied to U at the start of the RHS (dg.f90:331): FilterMat(0:N,0:N) or NULL (FilterType 0) */
    const double *FilterMat;
    /* source term of the manufactured solution (exactfunc.f90:665-926 CalcSource, GPU kernel :946-1113): IniExactFunc 4 */
    int iniExactFunc;
    double AdvVel[3];
    const double *Elem_xGP; /* (3,n,n,n,nElems) */
    /* channel testcase forcing (testcase/channel/testcase.f90:277-296 TestcaseSource): active if tcSource != 0 */
    int tcSource;
    double dpdx, BulkVel;
    /* non-default lifting forms (lifting.f90:81-85, 139-141): weak form (surface flux 1/2 (U_m + U_s), D_Hat_T, signed
     * surface integral) and conservative volume form (metrics inside the derivative); BR2 is always strong */
    int doWeakLifting, doConservativeLifting;
    /* sponge (sponge/sponge.f90:529-588): SpongeMat(0:N,0:N,0:N,nElems) = damping sigma / sJ, zero outside the sponge zone,
     * or NULL; the base flow lives in the oracle (array "SpBaseFlow") */
    const double *SpongeMat;
    /* overintegration of JU_t, step 14 of the RHS (dg/overintegration.f90:179-340; host FLEXI: GALAEXI stops in InitOverintegration,
     * :108-114): 0 none; 1 cut-off: Filter(Ut, OverintegrationMat) then ApplyJacobian; 2 conservative cut-off (FilterConservative):
     * JU_t projected to NUnder with Vdm_N_NUnder(0:NUnder,0:N), times sJNUnder(0:NUnder,0:NUnder,0:NUnder,nElems), interpolated
     * back with Vdm_NUnder_N(0:N,0:NUnder); matrices in Fortran layout */
    int OverintegrationType, NUnder;
    const double *OverintegrationMat, *Vdm_N_NUnder, *Vdm_NUnder_N, *sJNUnder;
} dgo_config;

typedef struct {
    dgo_config c;
    int n, n2, n3;
    size_t nDOF, nFace;
    double *U, *Ut, *UPrim, *Ut_tmp;
    double *U_master, *U_slave, *UPrim_master, *UPrim_slave, *Flux_master, *Flux_slave;
    double *gradUx, *gradUy, *gradUz;
    double *gradUx_master, *gradUy_master, *gradUz_master, *gradUx_slave, *gradUy_slave, *gradUz_slave;
    double *f, *g, *h;
    double *MetricsAdv, *MetricsVisc;
    double *FluxX, *FluxY, *FluxZ; /* BR2 lifting fluxes (lifting_br2.t90) */
    double *SpBaseFlow;            /* sponge base flow (PP_nVar,n,n,n,nElems) */
} dgo;

#define IDX_VOL(s, nv, v, i, j, k, e) ((size_t)(v) + (size_t)(nv) * ((size_t)(i) + (s)->n * ((size_t)(j) + (s)->n * ((size_t)(k) + (size_t)(s)->n * (size_t)(e)))))
#define IDX_FACE(s, nv, v, p, q, sd) ((size_t)(v) + (size_t)(nv) * ((size_t)(p) + (s)->n * ((size_t)(q) + (size_t)(s)->n * (size_t)(sd))))

static inline int s2v2(const int *tab, int n, int c, int p, int q, int flip, int loc)
{ /* S2V2(c,p,q,flip,loc) with c in 1..2, loc in 1..6 */
    return tab[(c - 1) + 2 * (p + n * (q + n * (flip + 5 * (loc - 1))))];
}

/* ------------------------------------------------------------------------------------------------ */
/* equations/navierstokes/idealgas/eos.f90:212-239 ConsToPrim */
static inline void cons_to_prim(double *prim, const double *cons, double kappa, double R)
{
    double sRho = 1. / cons[DENS];
    prim[DENS] = cons[DENS];
    prim[VEL1] = cons[MOM1] * sRho;
    prim[VEL2] = cons[MOM2] * sRho;
    prim[VEL3] = cons[MOM3] * sRho;
    prim[PRES] = (kappa - 1.) * (cons[ENER] - 0.5 * (cons[MOM1] * prim[VEL1] + cons[MOM2] * prim[VEL2] + cons[MOM3] * prim[VEL3]));
    prim[TEMP] = prim[PRES] * sRho / R;
}
/* eos.f90:467-489 PrimToCons */
static inline void prim_to_cons(const double *prim, double *cons, double kappa)
{
    cons[DENS] = prim[DENS];
    cons[MOM1] = prim[VEL1] * prim[DENS];
    cons[MOM2] = prim[VEL2] * prim[DENS];
    cons[MOM3] = prim[VEL3] * prim[DENS];
    cons[ENER] = prim[PRES] / (kappa - 1.) + 0.5 * (cons[MOM1] * prim[VEL1] + cons[MOM2] * prim[VEL2] + cons[MOM3] * prim[VEL3]);
}
/* idealgas/viscosity.f90 muSuth + eos.h:89-110 VISCOSITY_PRIM_EOS / THERMAL_CONDUCTIVITY_EOS */
static inline double viscosity(const dgo_config *c, const double *prim)
{
    if (c->viscLaw == 0) return c->EOS[EOS_MU0];
    /* viscosity.f90:105: TnoDim=T*Tref; IF(TnoDim >= Ts) mu0*TnoDim**ExpoSuth*(1+Ts)/(TnoDim+Ts) ELSE mu0*TnoDim*cSuth */
    double Tn = prim[TEMP] * c->EOS[EOS_TREF];
    double Ts = c->EOS[EOS_TS];
    if (Tn >= Ts) return c->EOS[EOS_MU0] * pow(Tn, c->EOS[EOS_EXPOSUTH]) * (1. + Ts) / (Tn + Ts);
    return c->EOS[EOS_MU0] * Tn * c->EOS[EOS_CSUTH];
}
static inline double conductivity(const dgo_config *c, double mu)
{
    return mu * c->EOS[EOS_R] * c->EOS[EOS_KAPPA] / ((c->EOS[EOS_KAPPA] - 1.) * c->EOS[EOS_PR]);
}

/* ------------------------------------------------------------------------------------------------ */
/* interpolation/prolongtoface.t90:168-344 EvalElemFaceG_Device / EvalElemFaceGL_Device
 * (one thread per (elem, locSide, p, q); routes to master (flip==0) or slave) */
static void prolong_to_face(const dgo *s, int nVar, const double *Uvol, double *Um, double *Us)
{
    const dgo_config *c = &s->c;
    const int n = s->n, N = c->N;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++)
        for (int loc = 1; loc <= 6; loc++) {
            int SideID = c->ElemToSide[0 + 3 * ((loc - 1) + 6 * e)];
            int flip = c->ElemToSide[1 + 3 * ((loc - 1) + 6 * e)];
            double *dst = (flip == 0) ? Um : Us;
            for (int q = 0; q < n; q++)
                for (int p = 0; p < n; p++) {
                    int i = s2v2(c->S2V2, n, 1, p, q, flip, loc);
                    int j = s2v2(c->S2V2, n, 2, p, q, flip, loc);
                    double Uface[8];
                    if (c->nodeType == 2) {
                        for (int v = 0; v < nVar; v++) {
                            switch (loc) {
                            case XI_MINUS:   Uface[v] = Uvol[IDX_VOL(s, nVar, v, 0, i, j, e)]; break;
                            case ETA_MINUS:  Uface[v] = Uvol[IDX_VOL(s, nVar, v, i, 0, j, e)]; break;
                            case ZETA_MINUS: Uface[v] = Uvol[IDX_VOL(s, nVar, v, i, j, 0, e)]; break;
                            case XI_PLUS:    Uface[v] = Uvol[IDX_VOL(s, nVar, v, N, i, j, e)]; break;
                            case ETA_PLUS:   Uface[v] = Uvol[IDX_VOL(s, nVar, v, i, N, j, e)]; break;
                            default:         Uface[v] = Uvol[IDX_VOL(s, nVar, v, i, j, N, e)]; break;
                            }
                        }
                    } else {
                        const double *L = (loc == XI_MINUS || loc == ETA_MINUS || loc == ZETA_MINUS) ? c->L_Minus : c->L_Plus;
                        for (int v = 0; v < nVar; v++) {
                            double a = 0.;
                            for (int l = 0; l < n; l++) {
                                double u;
                                switch (loc) {
                                case XI_MINUS: case XI_PLUS:   u = Uvol[IDX_VOL(s, nVar, v, l, i, j, e)]; break;
                                case ETA_MINUS: case ETA_PLUS: u = Uvol[IDX_VOL(s, nVar, v, i, l, j, e)]; break;
                                default:                       u = Uvol[IDX_VOL(s, nVar, v, i, j, l, e)]; break;
                                }
                                a = (l == 0) ? u * L[0] : a + u * L[l];
                            }
                            Uface[v] = a;
                        }
                    }
                    for (int v = 0; v < nVar; v++) dst[IDX_FACE(s, nVar, v, p, q, SideID - 1)] = Uface[v];
                }
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* dg/surfint.t90:352-586 SurfInt_Device (two fluxes, slave gets -Flux_slave) and :591-725
 * SurfInt_Device_Single_Flux (lifting; strong form: sig=+1, weak: sig = 2*isMaster-1), incl. optional sJ */
static void surf_int(const dgo *s, int nVar, const double *Fm, const double *Fs, double *Ut, int single, int weak, int applyJac)
{
    const dgo_config *c = &s->c;
    const int n = s->n, N = c->N;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++)
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            double loc_[8];
            for (int v = 0; v < nVar; v++) loc_[v] = 0.;
            for (int loc = 1; loc <= 6; loc++) {
                int SideID = c->ElemToSide[0 + 3 * ((loc - 1) + 6 * e)];
                int flip = c->ElemToSide[1 + 3 * ((loc - 1) + 6 * e)];
                int isM = c->ElemToSide[2 + 3 * ((loc - 1) + 6 * e)];
                int a, b, l; double Lh; int minus;
                switch (loc) {
                case XI_MINUS:   a = j; b = k; l = i; minus = 1; break;
                case ETA_MINUS:  a = i; b = k; l = j; minus = 1; break;
                case ZETA_MINUS: a = i; b = j; l = k; minus = 1; break;
                case XI_PLUS:    a = j; b = k; l = i; minus = 0; break;
                case ETA_PLUS:   a = i; b = k; l = j; minus = 0; break;
                default:         a = i; b = j; l = k; minus = 0; break;
                }
                if (c->nodeType == 2) { /* collocation: only the boundary layer contributes */
                    if (minus ? (l != 0) : (l != N)) continue;
                    Lh = minus ? c->L_HatMinus[0] : c->L_HatPlus[N];
                } else {
                    Lh = minus ? c->L_HatMinus[l] : c->L_HatPlus[l];
                }
                int p = s2v2(c->S2V2_inv, n, 1, a, b, flip, loc);
                int q = s2v2(c->S2V2_inv, n, 2, a, b, flip, loc);
                if (single) {
                    double sig = weak ? (2. * (double)isM - 1.) : 1.;
                    for (int v = 0; v < nVar; v++) loc_[v] = loc_[v] + sig * Fm[IDX_FACE(s, nVar, v, p, q, SideID - 1)] * Lh;
                } else if (flip == 0) {
                    for (int v = 0; v < nVar; v++) loc_[v] = loc_[v] + Fm[IDX_FACE(s, nVar, v, p, q, SideID - 1)] * Lh;
                } else {
                    for (int v = 0; v < nVar; v++) loc_[v] = loc_[v] - Fs[IDX_FACE(s, nVar, v, p, q, SideID - 1)] * Lh;
                }
            }
            for (int v = 0; v < nVar; v++) {
                size_t id = IDX_VOL(s, nVar, v, i, j, k, e);
                if (applyJac) Ut[id] = c->sJ[i + n * (j + n * (k + (size_t)n * e))] * (Ut[id] + loc_[v]);
                else Ut[id] = Ut[id] + loc_[v];
            }
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* BCs: idealgas/eos.f90:605-628 PRESSURE_RIEMANN_PURE */
static double pressure_riemann(const double *P, double kappa)
{
    double P_RP;
    if (P[VEL1] <= 0.) {
        double kf = 2. * kappa / (kappa - 1.);
        P_RP = P[PRES] * pow(fmax(0.0001, (1. + 0.5 * (kappa - 1.) * P[VEL1] / sqrt(kappa * P[PRES] / P[DENS]))), kf);
    } else {
        double ar = 2. / ((kappa + 1.) * P[DENS]);
        double br = (kappa - 1.) / (kappa + 1.) * P[PRES];
        P_RP = P[PRES] + P[VEL1] / ar * 0.5 * (P[VEL1] + sqrt(P[VEL1] * P[VEL1] + 4. * ar * (P[PRES] + br)));
    }
    return P_RP;
}
/* idealgas/getboundaryflux.f90:262-487 GetBoundaryState, BC types 2,3,4,9,91,23,24,25,27 */
static int is_riemann_bc(int t) { return t == 2 || t == 23 || t == 24 || t == 25 || t == 27; }
static int is_wall_bc(int t) { return t == 3 || t == 4 || t == 9 || t == 91; }
static int get_boundary_state(const dgo_config *c, int BCType, double *out, const double *Pm, const double *Ref,
                              const double *nv, const double *t1, const double *t2)
{
    double kappa = c->EOS[EOS_KAPPA], R = c->EOS[EOS_R];
    if (BCType == 2) { for (int v = 0; v < NP; v++) out[v] = Ref[v]; return 0; }
    if (is_wall_bc(BCType) || (is_riemann_bc(BCType) && BCType != 2)) {
        double b[NP];
        b[DENS] = Pm[DENS];
        b[VEL1] = Pm[VEL1] * nv[0] + Pm[VEL2] * nv[1] + Pm[VEL3] * nv[2];
        b[VEL2] = Pm[VEL1] * t1[0] + Pm[VEL2] * t1[1] + Pm[VEL3] * t1[2];
        b[VEL3] = Pm[VEL1] * t2[0] + Pm[VEL2] * t2[1] + Pm[VEL3] * t2[2];
        b[PRES] = Pm[PRES];
        b[TEMP] = Pm[TEMP];
        if (BCType == 3) {
            b[PRES] = pressure_riemann(b, kappa);
            b[VEL1] = b[VEL2] = b[VEL3] = 0.;
            b[TEMP] = Pm[TEMP];
            b[DENS] = b[PRES] / (b[TEMP] * R);
        } else if (BCType == 4) {
            b[PRES] = pressure_riemann(b, kappa);
            b[VEL1] = b[VEL2] = b[VEL3] = 0.;
            b[TEMP] = Ref[TEMP];
            b[DENS] = b[PRES] / (b[TEMP] * R);
        } else if (BCType == 9 || BCType == 91) {
            b[PRES] = pressure_riemann(b, kappa);
            b[VEL1] = 0.;
            b[DENS] = Pm[DENS];
            b[TEMP] = b[PRES] / (b[DENS] * R);
        } else if (BCType == 23) { /* :361-383 outflow Mach number, RefState = (x,MaOut,x,x,x) */
            double MaOut = Ref[1];
            double cs = sqrt(kappa * b[PRES] / b[DENS]);
            double Ma = b[VEL1] / cs, pb;
            if (Ma < 1) {
                double pt = b[PRES] * pow(1 + 0.5 * (kappa - 1) * Ma * Ma, kappa / (kappa - 1.));
                pb = pt * pow(1 + 0.5 * (kappa - 1) * MaOut * MaOut, -kappa / (kappa - 1.));
            } else {
                pb = b[PRES] + 0.5 * b[DENS] * (b[VEL1] * b[VEL1] + b[VEL2] * b[VEL2] + b[VEL3] * b[VEL3]);
            }
            b[DENS] = kappa * pb / (cs * cs);
            b[PRES] = pb;
            b[TEMP] = b[PRES] / (R * b[DENS]);
        } else if (BCType == 24) { /* :385-401 pressure outflow, RefState = (x,x,x,x,p) */
            double cs = sqrt(kappa * b[PRES] / b[DENS]);
            double Ma = b[VEL1] / cs;
            if (Ma < 1) {
                double pb = Ref[4];
                b[DENS] = kappa * pb / (cs * cs);
                b[PRES] = pb;
                b[TEMP] = b[PRES] / (R * b[DENS]);
            }
        } else if (BCType == 25) { /* :403-426 subsonic outflow */
            double cs = sqrt(kappa * b[PRES] / b[DENS]);
            double Ma = b[VEL1] / cs, pb;
            if (Ma < 1) pb = Ref[4];
            else pb = b[PRES] + 0.5 * b[DENS] * (b[VEL1] * b[VEL1] + b[VEL2] * b[VEL2] + b[VEL3] * b[VEL3]);
            if (b[VEL1] < 0.) { b[VEL1] = fabs(b[VEL1]); b[VEL2] = 0.; b[VEL3] = 0.; }
            b[DENS] = kappa * pb / (cs * cs);
            b[PRES] = Ref[4];
            b[TEMP] = b[PRES] / (R * b[DENS]);
        } else { /* 27, :428-474 subsonic inflow; RefState = (Tt, a(1:3) precomputed at init :201-210, pt) */
            double Tt = Ref[0], pt = Ref[4];
            const double *a = &Ref[1];
            double A = -1. * (a[0] * nv[0] + a[1] * nv[1] + a[2] * nv[2]);
            double cs = sqrt(kappa * b[PRES] / b[DENS]);
            double Rplus = -b[VEL1] - 2. * cs / (kappa - 1.);
            double tmp1 = A * A + 2. / (kappa - 1.);
            double tmp2 = 2 * Rplus;
            double tmp3 = (kappa - 1.) / 2. * (Rplus * Rplus) - kappa * R * Tt * (A * A);
            double cb = fmax((-tmp2 + sqrt(tmp2 * tmp2 - 4 * tmp1 * tmp3)) / (2 * tmp1), (-tmp2 - sqrt(tmp2 * tmp2 - 4 * tmp1 * tmp3)) / (2 * tmp1));
            double Tb = cb * cb / (kappa * R);
            double Ma = sqrt(2. / (kappa - 1.) * (Tt / Tb - 1.));
            double pb = pt * pow(1. + 0.5 * (kappa - 1.) * (Ma * Ma), -kappa / (kappa - 1.));
            double Um = Ma * sqrt(kappa * R * Tb);
            b[DENS] = pb / (R * Tb);
            b[VEL1] = Um * (a[0] * nv[0] + a[1] * nv[1] + a[2] * nv[2]);
            b[VEL2] = Um * (a[0] * t1[0] + a[1] * t1[1] + a[2] * t1[2]);
            b[VEL3] = Um * (a[0] * t2[0] + a[1] * t2[1] + a[2] * t2[2]);
            b[PRES] = pb;
            b[TEMP] = Tb;
        }
        out[DENS] = b[DENS];
        for (int d = 0; d < 3; d++) out[VEL1 + d] = b[VEL1] * nv[d] + b[VEL2] * t1[d] + b[VEL3] * t2[d];
        out[PRES] = b[PRES];
        out[TEMP] = b[TEMP];
        return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* equations/navierstokes/splitflux.f90: surface fluxes :315 (SD) :627 (MO) :407 (DU) :506 (KG) :735 (PI) */
static void split_surface_flux(int variant, const double *L, const double *R, double *F)
{
    if (variant == 1) { /* MO :627-660 */
        double rhoep_LL = L[E_ENER] - 0.5 * L[E_DENS] * (L[E_VEL1] * L[E_VEL1] + L[E_VEL2] * L[E_VEL2] + L[E_VEL3] * L[E_VEL3]) + L[E_PRES];
        double rhoep_RR = R[E_ENER] - 0.5 * R[E_DENS] * (R[E_VEL1] * R[E_VEL1] + R[E_VEL2] * R[E_VEL2] + R[E_VEL3] * R[E_VEL3]) + R[E_PRES];
        F[DENS] = 0.5 * (L[E_MOM1] + R[E_MOM1]);
        F[MOM1] = 0.25 * (L[E_MOM1] + R[E_MOM1]) * (L[E_VEL1] + R[E_VEL1]) + 0.5 * (L[E_PRES] + R[E_PRES]);
        F[MOM2] = 0.25 * (L[E_MOM1] + R[E_MOM1]) * (L[E_VEL2] + R[E_VEL2]);
        F[MOM3] = 0.25 * (L[E_MOM1] + R[E_MOM1]) * (L[E_VEL3] + R[E_VEL3]);
        F[ENER] = 0.5 * (rhoep_LL * L[E_VEL1] + rhoep_RR * R[E_VEL1]) +
                  0.25 * (L[E_MOM1] * L[E_VEL1] + R[E_MOM1] * R[E_VEL1]) * (L[E_VEL1] + R[E_VEL1]) +
                  0.25 * (L[E_MOM1] * L[E_VEL2] + R[E_MOM1] * R[E_VEL2]) * (L[E_VEL2] + R[E_VEL2]) +
                  0.25 * (L[E_MOM1] * L[E_VEL3] + R[E_MOM1] * R[E_VEL3]) * (L[E_VEL3] + R[E_VEL3]) -
                  0.25 * (L[E_MOM1] * L[E_VEL1] * L[E_VEL1] + R[E_MOM1] * R[E_VEL1] * R[E_VEL1]) -
                  0.25 * (L[E_MOM1] * L[E_VEL2] * L[E_VEL2] + R[E_MOM1] * R[E_VEL2] * R[E_VEL2]) -
                  0.25 * (L[E_MOM1] * L[E_VEL3] * L[E_VEL3] + R[E_MOM1] * R[E_VEL3] * R[E_VEL3]);
    } else if (variant == 2) { /* DU :407-428 */
        F[DENS] = 0.25 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]);
        F[MOM1] = 0.25 * (L[E_MOM1] + R[E_MOM1]) * (L[E_VEL1] + R[E_VEL1]) + 0.5 * (L[E_PRES] + R[E_PRES]);
        F[MOM2] = 0.25 * (L[E_MOM2] + R[E_MOM2]) * (L[E_VEL1] + R[E_VEL1]);
        F[MOM3] = 0.25 * (L[E_MOM3] + R[E_MOM3]) * (L[E_VEL1] + R[E_VEL1]);
        F[ENER] = 0.25 * (L[E_ENER] + R[E_ENER] + L[E_PRES] + R[E_PRES]) * (L[E_VEL1] + R[E_VEL1]);
    } else if (variant == 0) {
        F[DENS] = 0.5 * (L[E_MOM1] + R[E_MOM1]);
        F[MOM1] = 0.5 * (L[E_MOM1] * L[E_VEL1] + L[E_PRES] + R[E_MOM1] * R[E_VEL1] + R[E_PRES]);
        F[MOM2] = 0.5 * (L[E_MOM1] * L[E_VEL2] + R[E_MOM1] * R[E_VEL2]);
        F[MOM3] = 0.5 * (L[E_MOM1] * L[E_VEL3] + R[E_MOM1] * R[E_VEL3]);
        F[ENER] = 0.5 * ((L[E_ENER] + L[E_PRES]) * L[E_VEL1] + (R[E_ENER] + R[E_PRES]) * R[E_VEL1]);
    } else if (variant == 3) {
        double E_LL = L[E_ENER] / L[E_DENS], E_RR = R[E_ENER] / R[E_DENS];
        F[DENS] = 0.25 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]);
        F[MOM1] = 0.125 * (L[E_DENS] + R[E_DENS]) * ((L[E_VEL1] + R[E_VEL1]) * (L[E_VEL1] + R[E_VEL1])) + 0.5 * (L[E_PRES] + R[E_PRES]);
        F[MOM2] = 0.125 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]) * (L[E_VEL2] + R[E_VEL2]);
        F[MOM3] = 0.125 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]) * (L[E_VEL3] + R[E_VEL3]);
        F[ENER] = 0.125 * (L[E_DENS] + R[E_DENS]) * (E_LL + E_RR) * (L[E_VEL1] + R[E_VEL1]) +
                  0.25 * (L[E_PRES] + R[E_PRES]) * (L[E_VEL1] + R[E_VEL1]);
    } else { /* PI */
        double H_LL = (L[E_ENER] + L[E_PRES]) / L[E_DENS], H_RR = (R[E_ENER] + R[E_PRES]) / R[E_DENS];
        F[DENS] = 0.25 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]);
        F[MOM1] = 0.125 * (L[E_DENS] + R[E_DENS]) * ((L[E_VEL1] + R[E_VEL1]) * (L[E_VEL1] + R[E_VEL1])) + 0.5 * (L[E_PRES] + R[E_PRES]);
        F[MOM2] = 0.125 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]) * (L[E_VEL2] + R[E_VEL2]);
        F[MOM3] = 0.125 * (L[E_DENS] + R[E_DENS]) * (L[E_VEL1] + R[E_VEL1]) * (L[E_VEL3] + R[E_VEL3]);
        F[ENER] = 0.125 * (L[E_DENS] + R[E_DENS]) * (H_LL + H_RR) * (L[E_VEL1] + R[E_VEL1]);
    }
}

/* splitflux.f90: volume two-point fluxes :145 (SD) :543 (MO) :346 (DU) :437 (KG) :669 (PI) */
static void split_volume_flux(int variant, const double *URef, const double *PRef, const double *U, const double *P,
                              const double *MRef, const double *M, double *Flux)
{
    double f[NV], g[NV], h[NV];
    if (variant == 1) { /* MO :543-622 */
        double rhoepRef = URef[ENER] - 0.5 * URef[DENS] * (PRef[VEL1] * PRef[VEL1] + PRef[VEL2] * PRef[VEL2] + PRef[VEL3] * PRef[VEL3]) + PRef[PRES];
        double rhoep = U[ENER] - 0.5 * U[DENS] * (P[VEL1] * P[VEL1] + P[VEL2] * P[VEL2] + P[VEL3] * P[VEL3]) + P[PRES];
        double *F3[3] = {f, g, h};
        for (int d = 0; d < 3; d++) {
            double *F = F3[d];
            double mR = URef[MOM1 + d], m = U[MOM1 + d];
            F[DENS] = (mR + m);
            F[MOM1] = 0.5 * (mR + m) * (PRef[VEL1] + P[VEL1]);
            F[MOM2] = 0.5 * (mR + m) * (PRef[VEL2] + P[VEL2]);
            F[MOM3] = 0.5 * (mR + m) * (PRef[VEL3] + P[VEL3]);
            F[MOM1 + d] = F[MOM1 + d] + (PRef[PRES] + P[PRES]);
            F[ENER] = (rhoepRef * PRef[VEL1 + d] + rhoep * P[VEL1 + d]) +
                      0.5 * (mR * PRef[VEL1] + m * P[VEL1]) * (PRef[VEL1] + P[VEL1]) +
                      0.5 * (mR * PRef[VEL2] + m * P[VEL2]) * (PRef[VEL2] + P[VEL2]) +
                      0.5 * (mR * PRef[VEL3] + m * P[VEL3]) * (PRef[VEL3] + P[VEL3]) -
                      0.5 * (mR * PRef[VEL1] * PRef[VEL1] + m * P[VEL1] * P[VEL1]) -
                      0.5 * (mR * PRef[VEL2] * PRef[VEL2] + m * P[VEL2] * P[VEL2]) -
                      0.5 * (mR * PRef[VEL3] * PRef[VEL3] + m * P[VEL3] * P[VEL3]);
        }
    } else if (variant == 2) { /* DU :346-402 */
        double *F3[3] = {f, g, h};
        for (int d = 0; d < 3; d++) {
            double *F = F3[d];
            double vs = PRef[VEL1 + d] + P[VEL1 + d];
            F[DENS] = 0.5 * (URef[DENS] + U[DENS]) * vs;
            F[MOM1] = 0.5 * (URef[MOM1] + U[MOM1]) * vs;
            F[MOM2] = 0.5 * (URef[MOM2] + U[MOM2]) * vs;
            F[MOM3] = 0.5 * (URef[MOM3] + U[MOM3]) * vs;
            F[MOM1 + d] = F[MOM1 + d] + (PRef[PRES] + P[PRES]);
            F[ENER] = 0.5 * (URef[ENER] + U[ENER] + PRef[PRES] + P[PRES]) * vs;
        }
    } else if (variant == 0) {
        double rhoEpRef = URef[ENER] + PRef[PRES], rhoEp = U[ENER] + P[PRES];
        f[DENS] = (URef[MOM1] + U[MOM1]);
        f[MOM1] = (URef[MOM1] * PRef[VEL1] + PRef[PRES] + U[MOM1] * P[VEL1] + P[PRES]);
        f[MOM2] = (URef[MOM1] * PRef[VEL2] + U[MOM1] * P[VEL2]);
        f[MOM3] = (URef[MOM1] * PRef[VEL3] + U[MOM1] * P[VEL3]);
        f[ENER] = (rhoEpRef * PRef[VEL1] + rhoEp * P[VEL1]);
        g[DENS] = (URef[MOM2] + U[MOM2]);
        g[MOM1] = (URef[MOM1] * PRef[VEL2] + U[MOM1] * P[VEL2]);
        g[MOM2] = (URef[MOM2] * PRef[VEL2] + PRef[PRES] + U[MOM2] * P[VEL2] + P[PRES]);
        g[MOM3] = (URef[MOM2] * PRef[VEL3] + U[MOM2] * P[VEL3]);
        g[ENER] = (rhoEpRef * PRef[VEL2] + rhoEp * P[VEL2]);
        h[DENS] = (URef[MOM3] + U[MOM3]);
        h[MOM1] = (URef[MOM1] * PRef[VEL3] + U[MOM1] * P[VEL3]);
        h[MOM2] = (URef[MOM2] * PRef[VEL3] + U[MOM2] * P[VEL3]);
        h[MOM3] = (URef[MOM3] * PRef[VEL3] + PRef[PRES] + U[MOM3] * P[VEL3] + P[PRES]);
        h[ENER] = (rhoEpRef * PRef[VEL3] + rhoEp * P[VEL3]);
    } else {
        double rs = URef[DENS] + U[DENS];
        double us = PRef[VEL1] + P[VEL1], vs = PRef[VEL2] + P[VEL2], ws = PRef[VEL3] + P[VEL3];
        double ps = PRef[PRES] + P[PRES];
        f[DENS] = 0.5 * rs * us;
        f[MOM1] = 0.25 * rs * (us * us) + ps;
        f[MOM2] = 0.25 * rs * us * vs;
        f[MOM3] = 0.25 * rs * us * ws;
        g[DENS] = 0.5 * rs * vs;
        g[MOM1] = f[MOM2];
        g[MOM2] = 0.25 * rs * (vs * vs) + ps;
        g[MOM3] = 0.25 * rs * vs * ws;
        h[DENS] = 0.5 * rs * ws;
        h[MOM1] = f[MOM3];
        h[MOM2] = g[MOM3];
        h[MOM3] = 0.25 * rs * (ws * ws) + ps;
        if (variant == 3) { /* KG: {rho}{E}{u} + {p}{u} */
            double eRef = URef[ENER] / URef[DENS], e = U[ENER] / U[DENS];
            f[ENER] = 0.25 * rs * us * (eRef + e) + 0.5 * ps * us;
            g[ENER] = 0.25 * rs * vs * (eRef + e) + 0.5 * ps * vs;
            h[ENER] = 0.25 * rs * ws * (eRef + e) + 0.5 * ps * ws;
        } else { /* PI: {rho}{H}{u} */
            double HRef = (URef[ENER] + PRef[PRES]) / URef[DENS], H = (U[ENER] + P[PRES]) / U[DENS];
            f[ENER] = 0.25 * rs * us * (HRef + H);
            g[ENER] = 0.25 * rs * vs * (HRef + H);
            h[ENER] = 0.25 * rs * ws * (HRef + H);
        }
    }
    for (int v = 0; v < NV; v++)
        Flux[v] = 0.5 * (MRef[0] + M[0]) * f[v] + 0.5 * (MRef[2] + M[2]) * h[v] + 0.5 * (MRef[1] + M[1]) * g[v];
}

/* ------------------------------------------------------------------------------------------------ */
/* riemann.f90: solvers :707 LF, :738 HLLC, :809 Roe, :885 RoeEntropyFix, :994 RoeL2, :1075 HLL, :1126 HLLE, :1175 HLLEM, :1239 FluxAverage */
static void roe_averages(const double *L, const double *R, double kappa, double *RoeVel, double *RoeH, double *Roec, double *absVel)
{
    double H_L = (L[E_ENER] + L[E_PRES]) * L[E_SRHO];
    double H_R = (R[E_ENER] + R[E_PRES]) * R[E_SRHO];
    double sl = sqrt(L[E_DENS]), sr = sqrt(R[E_DENS]);
    double ss = 1. / (sl + sr);
    for (int d = 0; d < 3; d++) RoeVel[d] = (sr * R[E_VEL1 + d] + sl * L[E_VEL1 + d]) * ss;
    *RoeH = (sr * H_R + sl * H_L) * ss;
    *absVel = RoeVel[0] * RoeVel[0] + RoeVel[1] * RoeVel[1] + RoeVel[2] * RoeVel[2];
    *Roec = sqrt((kappa - 1.) * (*RoeH - 0.5 * (*absVel)));
}

static void riemann_solver(const dgo_config *c, double *F, const double *F_L, const double *F_R, const double *L, const double *R)
{
    const double kappa = c->EOS[EOS_KAPPA];
    const int split = c->splitDG >= 0;
    if (c->riemann == 0) { /* LF */
        double cL = sqrt(kappa * L[E_PRES] * L[E_SRHO]), cR = sqrt(kappa * R[E_PRES] * R[E_SRHO]);
        double LambdaMax = fmax(fabs(R[E_VEL1]), fabs(L[E_VEL1])) + fmax(cL, cR);
        if (!split) {
            for (int v = 0; v < NV; v++) F[v] = 0.5 * ((F_L[v] + F_R[v]) - LambdaMax * (R[v] - L[v]));
        } else {
            split_surface_flux(c->splitDG, L, R, F);
            for (int v = 0; v < NV; v++) F[v] = F[v] - 0.5 * LambdaMax * (R[v] - L[v]);
        }
        return;
    }
    if (c->riemann == 9) { /* Riemann_FluxAverage :1239 (SPLIT_DG only) */
        split_surface_flux(c->splitDG, L, R, F);
        return;
    }
    if (c->riemann == 4 || c->riemann == 6 || c->riemann == 7) { /* HLL :1075, HLLE :1126, HLLEM :1175 (non-split builds only) */
        double RoeVel[3], RoeH, Roec, absVel;
        roe_averages(L, R, kappa, RoeVel, &RoeH, &Roec, &absVel);
        double Ssl, Ssr;
        if (c->riemann == 4) { Ssl = RoeVel[0] - Roec; Ssr = RoeVel[0] + Roec; }
        else {
            double beta = sqrt(0.5 * (kappa - 1.) / kappa);
            double cL = sqrt(kappa * L[E_PRES] * L[E_SRHO]), cR = sqrt(kappa * R[E_PRES] * R[E_SRHO]);
            Ssl = fmin(fmin(RoeVel[0] - Roec, L[E_VEL1] - beta * cL), 0.);
            Ssr = fmax(fmax(RoeVel[0] + Roec, R[E_VEL1] + beta * cR), 0.);
        }
        if (Ssl >= 0.) { for (int v = 0; v < NV; v++) F[v] = F_L[v]; }
        else if (Ssr <= 0.) { for (int v = 0; v < NV; v++) F[v] = F_R[v]; }
        else if (c->riemann != 7) {
            for (int v = 0; v < NV; v++) F[v] = (Ssr * F_L[v] - Ssl * F_R[v] + Ssl * Ssr * (R[v] - L[v])) / (Ssr - Ssl);
        } else {
            double RoeDens = sqrt(L[E_DENS] * R[E_DENS]);
            double delta = Roec / (Roec + fabs(0.5 * (Ssl + Ssr)));
            double A2 = (R[E_DENS] - L[E_DENS]) - (R[E_PRES] - L[E_PRES]) / (Roec * Roec);
            double A3 = RoeDens * (R[E_VEL2] - L[E_VEL2]), A4 = RoeDens * (R[E_VEL3] - L[E_VEL3]);
            double r2[5] = {1., RoeVel[0], RoeVel[1], RoeVel[2], 0.5 * absVel};
            double r3[5] = {0., 0., 1., 0., RoeVel[1]};
            double r4[5] = {0., 0., 0., 1., RoeVel[2]};
            for (int v = 0; v < NV; v++)
                F[v] = (Ssr * F_L[v] - Ssl * F_R[v] + Ssl * Ssr * (R[v] - L[v] - delta * (r2[v] * A2 + r3[v] * A3 + r4[v] * A4))) / (Ssr - Ssl);
        }
        return;
    }
    if (c->riemann == 5) { /* HLLC (non-split builds only, CMakeLists.txt:113-117) */
        double RoeVel[3], RoeH, Roec, absVel;
        roe_averages(L, R, kappa, RoeVel, &RoeH, &Roec, &absVel);
        double Ssl = RoeVel[0] - Roec, Ssr = RoeVel[0] + Roec;
        if (Ssl >= 0.) { for (int v = 0; v < NV; v++) F[v] = F_L[v]; }
        else if (Ssr <= 0.) { for (int v = 0; v < NV; v++) F[v] = F_R[v]; }
        else {
            double sMu_L = Ssl - L[E_VEL1], sMu_R = Ssr - R[E_VEL1];
            double SStar = (R[E_PRES] - L[E_PRES] + L[E_MOM1] * sMu_L - R[E_MOM1] * sMu_R) / (L[E_DENS] * sMu_L - R[E_DENS] * sMu_R);
            double Us[NV];
            if (Ssl <= 0. && SStar >= 0.) {
                double EStar = L[E_ENER] * L[E_SRHO] + (SStar - L[E_VEL1]) * (SStar + L[E_PRES] * L[E_SRHO] / sMu_L);
                double fac = L[E_DENS] * sMu_L / (Ssl - SStar);
                Us[0] = fac * 1.; Us[1] = fac * SStar; Us[2] = fac * L[E_VEL2]; Us[3] = fac * L[E_VEL3]; Us[4] = fac * EStar;
                for (int v = 0; v < NV; v++) F[v] = F_L[v] + Ssl * (Us[v] - L[v]);
            } else {
                double EStar = R[E_ENER] * R[E_SRHO] + (SStar - R[E_VEL1]) * (SStar + R[E_PRES] * R[E_SRHO] / sMu_R);
                double fac = R[E_DENS] * sMu_R / (Ssr - SStar);
                Us[0] = fac * 1.; Us[1] = fac * SStar; Us[2] = fac * R[E_VEL2]; Us[3] = fac * R[E_VEL3]; Us[4] = fac * EStar;
                for (int v = 0; v < NV; v++) F[v] = F_R[v] + Ssr * (Us[v] - R[v]);
            }
        }
        return;
    }
    /* Roe family */
    double RoeVel[3], RoeH, Roec, absVel;
    roe_averages(L, R, kappa, RoeVel, &RoeH, &Roec, &absVel);
    double a[5] = {RoeVel[0] - Roec, RoeVel[0], RoeVel[0], RoeVel[0], RoeVel[0] + Roec};
    double r1[5] = {1., a[0], RoeVel[1], RoeVel[2], RoeH - RoeVel[0] * Roec};
    double r2[5] = {1., RoeVel[0], RoeVel[1], RoeVel[2], 0.5 * absVel};
    double r3[5] = {0., 0., 1., 0., RoeVel[1]};
    double r4[5] = {0., 0., 0., 1., RoeVel[2]};
    double r5[5] = {1., a[4], RoeVel[1], RoeVel[2], RoeH + RoeVel[0] * Roec};
    double Alpha[5];
    if (c->riemann == 1 || c->riemann == 2) { /* Roe :809-877, RoeL2 :994-1070 */
        double dU[6];
        for (int v = 0; v < NV; v++) dU[v] = R[v] - L[v];
        dU[5] = dU[4] - (dU[2] - RoeVel[1] * dU[0]) * RoeVel[1] - (dU[3] - RoeVel[2] * dU[0]) * RoeVel[2];
        if (c->riemann == 2) { /* low Mach number fix :1044-1046 */
            double Ma_loc = sqrt(absVel) / (Roec * sqrt(kappa));
            dU[1] = dU[1] * Ma_loc; dU[2] = dU[2] * Ma_loc; dU[3] = dU[3] * Ma_loc;
        }
        Alpha[2] = dU[2] - RoeVel[1] * dU[0];
        Alpha[3] = dU[3] - RoeVel[2] * dU[0];
        Alpha[1] = (kappa - 1.) / (Roec * Roec) * (dU[0] * (RoeH - RoeVel[0] * RoeVel[0]) - dU[5] + RoeVel[0] * dU[1]);
        Alpha[0] = 0.5 / Roec * (dU[0] * (RoeVel[0] + Roec) - dU[1] - Roec * Alpha[1]);
        Alpha[4] = dU[0] - Alpha[0] - Alpha[1];
        for (int i = 0; i < 5; i++) a[i] = fabs(a[i]);
    } else { /* RoeEntropyFix :885-986 */
        double c_L = sqrt(kappa * L[E_PRES] * L[E_SRHO]), c_R = sqrt(kappa * R[E_PRES] * R[E_SRHO]);
        double RoeDens = sqrt(L[E_DENS] * R[E_DENS]);
        double dU[5];
        dU[0] = R[E_DENS] - L[E_DENS];
        dU[1] = R[E_VEL1] - L[E_VEL1]; dU[2] = R[E_VEL2] - L[E_VEL2]; dU[3] = R[E_VEL3] - L[E_VEL3];
        dU[4] = R[E_PRES] - L[E_PRES];
        double tmp = 0.5 / (Roec * Roec);
        Alpha[0] = tmp * (dU[4] - RoeDens * Roec * dU[1]);
        Alpha[1] = dU[0] - dU[4] * 2. * tmp;
        Alpha[2] = RoeDens * dU[2];
        Alpha[3] = RoeDens * dU[3];
        Alpha[4] = tmp * (dU[4] + RoeDens * Roec * dU[1]);
        double al[5] = {L[E_VEL1] - c_L, L[E_VEL1], L[E_VEL1], L[E_VEL1], L[E_VEL1] + c_L};
        double ar[5] = {R[E_VEL1] - c_R, R[E_VEL1], R[E_VEL1], R[E_VEL1], R[E_VEL1] + c_R};
        for (int i = 0; i < 5; i++) {
            double da = fmax(0., fmax(a[i] - al[i], ar[i] - a[i]));
            if (fabs(a[i]) < da) a[i] = 0.5 * (a[i] * a[i] / da + da);
            else a[i] = fabs(a[i]);
        }
    }
    if (!split) {
        for (int v = 0; v < NV; v++)
            F[v] = 0.5 * ((F_L[v] + F_R[v]) - Alpha[0] * a[0] * r1[v] - Alpha[1] * a[1] * r2[v] - Alpha[2] * a[2] * r3[v]
                          - Alpha[3] * a[3] * r4[v] - Alpha[4] * a[4] * r5[v]);
    } else {
        split_surface_flux(c->splitDG, L, R, F);
        for (int v = 0; v < NV; v++)
            F[v] = F[v] - 0.5 * (Alpha[0] * a[0] * r1[v] + Alpha[1] * a[1] * r2[v] + Alpha[2] * a[2] * r3[v]
                                 + Alpha[3] * a[3] * r4[v] + Alpha[4] * a[4] * r5[v]);
    }
}

/* riemann.f90:216-289 Riemann (rotate, solve, rotate back); flux.f90:1123 EvalEulerFlux1D_fast */
static void riemann(const dgo_config *c, double *FOut, const double *U_L, const double *U_R, const double *P_L, const double *P_R,
                    const double *nv, const double *t1, const double *t2)
{
    double LL[NEXT], RR[NEXT], F_L[NV], F_R[NV], F[NV];
    LL[E_DENS] = U_L[DENS]; LL[E_SRHO] = 1. / LL[E_DENS]; LL[E_ENER] = U_L[ENER]; LL[E_PRES] = P_L[PRES];
    LL[E_VEL1] = P_L[VEL1] * nv[0] + P_L[VEL2] * nv[1] + P_L[VEL3] * nv[2];
    LL[E_VEL2] = P_L[VEL1] * t1[0] + P_L[VEL2] * t1[1] + P_L[VEL3] * t1[2];
    LL[E_MOM1] = LL[E_DENS] * LL[E_VEL1]; LL[E_MOM2] = LL[E_DENS] * LL[E_VEL2];
    LL[E_VEL3] = P_L[VEL1] * t2[0] + P_L[VEL2] * t2[1] + P_L[VEL3] * t2[2];
    LL[E_MOM3] = LL[E_DENS] * LL[E_VEL3]; LL[E_TEMP] = 0.;
    RR[E_DENS] = U_R[DENS]; RR[E_SRHO] = 1. / RR[E_DENS]; RR[E_ENER] = U_R[ENER]; RR[E_PRES] = P_R[PRES];
    RR[E_VEL1] = P_R[VEL1] * nv[0] + P_R[VEL2] * nv[1] + P_R[VEL3] * nv[2];
    RR[E_VEL2] = P_R[VEL1] * t1[0] + P_R[VEL2] * t1[1] + P_R[VEL3] * t1[2];
    RR[E_MOM1] = RR[E_DENS] * RR[E_VEL1]; RR[E_MOM2] = RR[E_DENS] * RR[E_VEL2];
    RR[E_VEL3] = P_R[VEL1] * t2[0] + P_R[VEL2] * t2[1] + P_R[VEL3] * t2[2];
    RR[E_MOM3] = RR[E_DENS] * RR[E_VEL3]; RR[E_TEMP] = 0.;
    if (c->splitDG < 0) {
        F_L[DENS] = LL[E_MOM1]; F_L[MOM1] = LL[E_MOM1] * LL[E_VEL1] + LL[E_PRES]; F_L[MOM2] = LL[E_MOM1] * LL[E_VEL2];
        F_L[MOM3] = LL[E_MOM1] * LL[E_VEL3]; F_L[ENER] = (LL[E_ENER] + LL[E_PRES]) * LL[E_VEL1];
        F_R[DENS] = RR[E_MOM1]; F_R[MOM1] = RR[E_MOM1] * RR[E_VEL1] + RR[E_PRES]; F_R[MOM2] = RR[E_MOM1] * RR[E_VEL2];
        F_R[MOM3] = RR[E_MOM1] * RR[E_VEL3]; F_R[ENER] = (RR[E_ENER] + RR[E_PRES]) * RR[E_VEL1];
    } else {
        for (int v = 0; v < NV; v++) F_L[v] = F_R[v] = 0.;
    }
    riemann_solver(c, F, F_L, F_R, LL, RR);
    FOut[DENS] = F[DENS];
    for (int d = 0; d < 3; d++) FOut[MOM1 + d] = nv[d] * F[MOM1] + t1[d] * F[MOM2] + t2[d] * F[MOM3];
    FOut[ENER] = F[ENER];
}

/* ------------------------------------------------------------------------------------------------ */
/* flux.f90:700-745 EvalDiffFlux3D (3 Cartesian viscous fluxes) */
static void eval_diff_flux3d(const double *P, const double *gx, const double *gy, const double *gz, double *f, double *g, double *h,
                             double mu, double lambda)
{
    const double s23 = 2. / 3., s43 = 4. / 3.;
    double tau_xx = mu * (s43 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
    double tau_yy = mu * (-s23 * gx[LIFT_VEL1] + s43 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
    double tau_zz = mu * (-s23 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] + s43 * gz[LIFT_VEL3]);
    double tau_xy = mu * (gy[LIFT_VEL1] + gx[LIFT_VEL2]);
    double tau_xz = mu * (gz[LIFT_VEL1] + gx[LIFT_VEL3]);
    double tau_yz = mu * (gz[LIFT_VEL2] + gy[LIFT_VEL3]);
    f[DENS] = 0.; f[MOM1] = -tau_xx; f[MOM2] = -tau_xy; f[MOM3] = -tau_xz;
    f[ENER] = -tau_xx * P[VEL1] - tau_xy * P[VEL2] - tau_xz * P[VEL3] - lambda * gx[LIFT_TEMP];
    g[DENS] = 0.; g[MOM1] = -tau_xy; g[MOM2] = -tau_yy; g[MOM3] = -tau_yz;
    g[ENER] = -tau_xy * P[VEL1] - tau_yy * P[VEL2] - tau_yz * P[VEL3] - lambda * gy[LIFT_TEMP];
    h[DENS] = 0.; h[MOM1] = -tau_xz; h[MOM2] = -tau_yz; h[MOM3] = -tau_zz;
    h[ENER] = -tau_xz * P[VEL1] - tau_yz * P[VEL2] - tau_zz * P[VEL3] - lambda * gz[LIFT_TEMP];
}
/* flux.f90:779-822 EvalNormalDiffFlux3D */
static void eval_normal_diff_flux3d(const double *P, const double *gx, const double *gy, const double *gz, double *f, const double *nv,
                                    double mu, double lambda)
{
    const double s23 = 2. / 3., s43 = 4. / 3.;
    double tau_xx = mu * (s43 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
    double tau_yy = mu * (-s23 * gx[LIFT_VEL1] + s43 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
    double tau_zz = mu * (-s23 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] + s43 * gz[LIFT_VEL3]);
    double tau_xy = mu * (gy[LIFT_VEL1] + gx[LIFT_VEL2]);
    double tau_xz = mu * (gz[LIFT_VEL1] + gx[LIFT_VEL3]);
    double tau_yz = mu * (gz[LIFT_VEL2] + gy[LIFT_VEL3]);
    f[DENS] = 0.;
    f[MOM1] = -nv[0] * tau_xx - nv[1] * tau_xy - nv[2] * tau_xz;
    f[MOM2] = -nv[0] * tau_xy - nv[1] * tau_yy - nv[2] * tau_yz;
    f[MOM3] = -nv[0] * tau_xz - nv[1] * tau_yz - nv[2] * tau_zz;
    f[ENER] = -nv[0] * (tau_xx * P[VEL1] + tau_xy * P[VEL2] + tau_xz * P[VEL3] + lambda * gx[LIFT_TEMP])
              - nv[1] * (tau_xy * P[VEL1] + tau_yy * P[VEL2] + tau_yz * P[VEL3] + lambda * gy[LIFT_TEMP])
              - nv[2] * (tau_xz * P[VEL1] + tau_yz * P[VEL2] + tau_zz * P[VEL3] + lambda * gz[LIFT_TEMP]);
}

/* flux.f90:137-186 (Euler, _fast), :509-614 (Euler+diff), :619-697 (diff only): transformed fluxes at one node */
static void eval_transformed_flux(int euler, int visc, const double *U, const double *P, const double *gx, const double *gy, const double *gz,
                                  double *f, double *g, double *h, const double *Mf, const double *Mg, const double *Mh,
                                  double mu, double lambda)
{
    double tau_xx = 0, tau_yy = 0, tau_zz = 0, tau_xy = 0, tau_xz = 0, tau_yz = 0, tX = 0, tY = 0, tZ = 0;
    if (visc) {
        const double s23 = 2. / 3., s43 = 4. / 3.;
        tau_xx = mu * (s43 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
        tau_yy = mu * (-s23 * gx[LIFT_VEL1] + s43 * gy[LIFT_VEL2] - s23 * gz[LIFT_VEL3]);
        tau_zz = mu * (-s23 * gx[LIFT_VEL1] - s23 * gy[LIFT_VEL2] + s43 * gz[LIFT_VEL3]);
        tau_xy = mu * (gy[LIFT_VEL1] + gx[LIFT_VEL2]);
        tau_xz = mu * (gz[LIFT_VEL1] + gx[LIFT_VEL3]);
        tau_yz = mu * (gz[LIFT_VEL2] + gy[LIFT_VEL3]);
        tX = tau_xx * P[VEL1] + tau_xy * P[VEL2] + tau_xz * P[VEL3] + lambda * gx[LIFT_TEMP];
        tY = tau_xy * P[VEL1] + tau_yy * P[VEL2] + tau_yz * P[VEL3] + lambda * gy[LIFT_TEMP];
        tZ = tau_xz * P[VEL1] + tau_yz * P[VEL2] + tau_zz * P[VEL3] + lambda * gz[LIFT_TEMP];
    }
    const double *Ms[3] = {Mf, Mg, Mh};
    double *Fs[3] = {f, g, h};
    double Ep = euler ? (U[ENER] + P[PRES]) / U[DENS] : 0.;
    for (int d = 0; d < 3; d++) {
        const double *M = Ms[d];
        double *F = Fs[d];
        if (euler && visc) {
            double Mmom = M[0] * U[MOM1] + M[1] * U[MOM2] + M[2] * U[MOM3];
            F[DENS] = Mmom;
            F[MOM1] = Mmom * P[VEL1] + M[0] * P[PRES] - M[0] * tau_xx - M[1] * tau_xy - M[2] * tau_xz;
            F[MOM2] = Mmom * P[VEL2] + M[1] * P[PRES] - M[0] * tau_xy - M[1] * tau_yy - M[2] * tau_yz;
            F[MOM3] = Mmom * P[VEL3] + M[2] * P[PRES] - M[0] * tau_xz - M[1] * tau_yz - M[2] * tau_zz;
            F[ENER] = Mmom * Ep - M[0] * tX - M[1] * tY - M[2] * tZ;
        } else if (euler) {
            double Mmom = M[0] * U[MOM1] + M[1] * U[MOM2] + M[2] * U[MOM3];
            F[DENS] = Mmom;
            F[MOM1] = Mmom * P[VEL1] + M[0] * P[PRES];
            F[MOM2] = Mmom * P[VEL2] + M[1] * P[PRES];
            F[MOM3] = Mmom * P[VEL3] + M[2] * P[PRES];
            F[ENER] = Mmom * Ep;
        } else {
            F[DENS] = 0.;
            F[MOM1] = -M[0] * tau_xx - M[1] * tau_xy - M[2] * tau_xz;
            F[MOM2] = -M[0] * tau_xy - M[1] * tau_yy - M[2] * tau_yz;
            F[MOM3] = -M[0] * tau_xz - M[1] * tau_yz - M[2] * tau_zz;
            F[ENER] = -M[0] * tX - M[1] * tY - M[2] * tZ;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Non-conforming interfaces, mortar/fillmortar.t90 (host FLEXI code; GALAEXI has no GPU mortar path).
 * 1-D operator application along the first (dir 0) or second (dir 1) side index with the reference's summation
 * order: the l=0 term first, then l=1..N. M is stored transposed: out(p) = sum_l M(l,p) in(l). */
static void mortar_1d(int n, int nVar, int dir, const double *M, const double *in, double *out)
{
    for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
        const int o = dir == 0 ? p : q;
        for (int v = 0; v < nVar; v++) {
            double a = M[0 + n * o] * in[v + nVar * (dir == 0 ? (0 + n * q) : (p + n * 0))];
            for (int l = 1; l < n; l++) a = a + M[l + n * o] * in[v + nVar * (dir == 0 ? (l + n * q) : (p + n * l))];
            out[v + nVar * (p + n * q)] = a;
        }
    }
}
/* the same for the sum of two operators on two inputs: out(p) = sum_l ( M1(l,p) a(l) + M2(l,p) b(l) ) */
static void mortar_1d_pair(int n, int nVar, int dir, const double *M1, const double *M2, const double *a, const double *b, double *out)
{
    for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
        const int o = dir == 0 ? p : q;
        for (int v = 0; v < nVar; v++) {
            int i0 = v + nVar * (dir == 0 ? (0 + n * q) : (p + n * 0));
            double t = M1[0 + n * o] * a[i0] + M2[0 + n * o] * b[i0];
            for (int l = 1; l < n; l++) {
                int il = v + nVar * (dir == 0 ? (l + n * q) : (p + n * l));
                t = t + M1[l + n * o] * a[il] + M2[l + n * o] * b[il];
            }
            out[v + nVar * (p + n * q)] = t;
        }
    }
}
#define MAXFACE (8 * 10 * 10)
/* fillmortar.t90:34-195 U_Mortar: big-side master data -> the 4 (type 1) or 2 (types 2: eta split, 3: xi split) small
 * sides; small sides that are slave sides here (flip>0, MPI) are filled through FS2M */
static void u_mortar(const dgo *s, int nVar, double *Um, double *Us, int firstSide, int lastSide)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    for (int sd = firstSide - 1; sd < lastSide; sd++) {
        const double *big = &Um[IDX_FACE(s, nVar, 0, 0, 0, sd)];
        const int type = c->MortarType[0 + 2 * sd], iSide = c->MortarType[1 + 2 * sd];
        double tmp[4][MAXFACE], tmp2[2][MAXFACE];
        int nMortars = 2;
        switch (type) {
        case 1:
            nMortars = 4;
            mortar_1d(n, nVar, 1, c->M_0_1, big, tmp2[0]);
            mortar_1d(n, nVar, 1, c->M_0_2, big, tmp2[1]);
            mortar_1d(n, nVar, 0, c->M_0_1, tmp2[0], tmp[0]);
            mortar_1d(n, nVar, 0, c->M_0_2, tmp2[0], tmp[1]);
            mortar_1d(n, nVar, 0, c->M_0_1, tmp2[1], tmp[2]);
            mortar_1d(n, nVar, 0, c->M_0_2, tmp2[1], tmp[3]);
            break;
        case 2:
            mortar_1d(n, nVar, 1, c->M_0_1, big, tmp[0]);
            mortar_1d(n, nVar, 1, c->M_0_2, big, tmp[1]);
            break;
        default:
            mortar_1d(n, nVar, 0, c->M_0_1, big, tmp[0]);
            mortar_1d(n, nVar, 0, c->M_0_2, big, tmp[1]);
        }
        for (int im = 0; im < nMortars; im++) {
            const int SideID = c->MortarInfo[0 + 2 * (im + 4 * (iSide - 1))], flip = c->MortarInfo[1 + 2 * (im + 4 * (iSide - 1))];
            for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
                if (flip == 0) {
                    for (int v = 0; v < nVar; v++) Um[IDX_FACE(s, nVar, v, p, q, SideID - 1)] = tmp[im][v + nVar * (p + n * q)];
                } else {
                    const int pm = c->FS2M[0 + 2 * (p + n * (q + n * flip))], qm = c->FS2M[1 + 2 * (p + n * (q + n * flip))];
                    for (int v = 0; v < nVar; v++) Us[IDX_FACE(s, nVar, v, p, q, SideID - 1)] = tmp[im][v + nVar * (pm + n * qm)];
                }
            }
        }
    }
}
/* fillmortar.t90:215-376 Flux_Mortar: small-side fluxes -> big side with the projection operators M_1_0, M_2_0;
 * small slave sides (MPI) enter through FS2M with the sign of the weak form */
static void flux_mortar(const dgo *s, int nVar, double *Fm, const double *Fs, int firstSide, int lastSide, int weak)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    for (int sd = firstSide - 1; sd < lastSide; sd++) {
        const int type = c->MortarType[0 + 2 * sd], iSide = c->MortarType[1 + 2 * sd];
        const int nMortars = type == 1 ? 4 : 2;
        double tmp[4][MAXFACE], tmp2[2][MAXFACE];
        for (int im = 0; im < nMortars; im++) {
            const int SideID = c->MortarInfo[0 + 2 * (im + 4 * (iSide - 1))], flip = c->MortarInfo[1 + 2 * (im + 4 * (iSide - 1))];
            for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
                if (flip == 0) {
                    for (int v = 0; v < nVar; v++) tmp[im][v + nVar * (p + n * q)] = Fm[IDX_FACE(s, nVar, v, p, q, SideID - 1)];
                } else {
                    const int pm = c->FS2M[0 + 2 * (p + n * (q + n * flip))], qm = c->FS2M[1 + 2 * (p + n * (q + n * flip))];
                    for (int v = 0; v < nVar; v++) {
                        double f = Fs[IDX_FACE(s, nVar, v, pm, qm, SideID - 1)];
                        tmp[im][v + nVar * (p + n * q)] = weak ? -f : f;
                    }
                }
            }
        }
        double *big = &Fm[IDX_FACE(s, nVar, 0, 0, 0, sd)];
        switch (type) {
        case 1:
            mortar_1d_pair(n, nVar, 0, c->M_1_0, c->M_2_0, tmp[0], tmp[1], tmp2[0]);
            mortar_1d_pair(n, nVar, 0, c->M_1_0, c->M_2_0, tmp[2], tmp[3], tmp2[1]);
            mortar_1d_pair(n, nVar, 1, c->M_1_0, c->M_2_0, tmp2[0], tmp2[1], big);
            break;
        case 2: mortar_1d_pair(n, nVar, 1, c->M_1_0, c->M_2_0, tmp[0], tmp[1], big); break;
        default: mortar_1d_pair(n, nVar, 0, c->M_1_0, c->M_2_0, tmp[0], tmp[1], big);
        }
    }
}
/* both ranges of big mortar sides (inner and MPI); on one rank the second range is empty */
static void u_mortar_all(const dgo *s, int nVar, double *Um, double *Us)
{
    u_mortar(s, nVar, Um, Us, s->c.firstMortarMPISide, s->c.lastMortarMPISide);
    u_mortar(s, nVar, Um, Us, s->c.firstMortarInnerSide, s->c.lastMortarInnerSide);
}
static void flux_mortar_all(const dgo *s, int nVar, double *Fm, const double *Fs, int weak)
{
    flux_mortar(s, nVar, Fm, Fs, s->c.firstMortarInnerSide, s->c.lastMortarInnerSide, weak);
    flux_mortar(s, nVar, Fm, Fs, s->c.firstMortarMPISide, s->c.lastMortarMPISide, weak);
}

/* lifting.f90:139-141: BR2 is always strong; the conservative form is only read for strong lifting */
static inline int lift_weak(const dgo_config *c) { return c->doWeakLifting && c->lifting != 2; }
static inline int lift_cons(const dgo_config *c) { return lift_weak(c) || c->doConservativeLifting; }

/* ------------------------------------------------------------------------------------------------ */
/* lifting: dg/lifting/lifting_br1.t90:47-167 (strong form, non-conservative volume integral) */
static void lifting_br1_fillflux(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    double *Flux = s->gradUz_slave; /* the reference reuses this buffer for the untransformed flux, lifting_br1.t90:76-80 */
    /* lifting_fillflux.t90:109-140 + getboundaryflux.f90:945-1019 Lifting_GetBoundaryFlux_Kernel (strong form) */
#pragma omp parallel for schedule(static)
    for (int sd = 0; sd < c->nBCSides; sd++)
        for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
            int bct = c->BCSides[0 + 2 * sd], bcs = c->BCSides[1 + 2 * sd];
            const double *Pm = &s->UPrim_master[IDX_FACE(s, NP, 0, p, q, sd)];
            const double *nv = &c->NormVec[IDX_FACE(s, 3, 0, p, q, sd)];
            const double *t1 = &c->TangVec1[IDX_FACE(s, 3, 0, p, q, sd)];
            const double *t2 = &c->TangVec2[IDX_FACE(s, 3, 0, p, q, sd)];
            double Pb[NP], Fl[NL];
            get_boundary_state(c, bct, Pb, Pm, &c->RefStatePrim[NP * (bcs > 0 ? bcs - 1 : 0)], nv, t1, t2);
            if (is_riemann_bc(bct)) {
                for (int v = 0; v < NL; v++) Fl[v] = 0.5 * (Pm[PRIM_LIFT[v]] + Pb[PRIM_LIFT[v]]);
            } else if (bct == 3 || bct == 4) {
                Fl[LIFT_DENS] = Pb[DENS]; Fl[LIFT_VEL1] = Fl[LIFT_VEL2] = Fl[LIFT_VEL3] = 0.; Fl[LIFT_TEMP] = Pb[TEMP];
            } else {
                Fl[LIFT_DENS] = Pm[DENS]; Fl[LIFT_VEL1] = Pb[VEL1]; Fl[LIFT_VEL2] = Pb[VEL2]; Fl[LIFT_VEL3] = Pb[VEL3]; Fl[LIFT_TEMP] = Pm[TEMP];
            }
            if (!lift_weak(c)) for (int v = 0; v < NL; v++) Fl[v] = Fl[v] - Pm[PRIM_LIFT[v]]; /* getboundaryflux.f90:1014 */
            double se = c->SurfElem[p + n * (q + (size_t)n * sd)];
            for (int v = 0; v < NL; v++) Flux[IDX_FACE(s, NL, v, p, q, sd)] = Fl[v] * se;
        }
    /* lifting_fillflux.t90:39-93 Lifting_FillFlux: inner + MPI MINE sides, sig = -1 (strong) or +1 (weak) */
    const double sig = lift_weak(c) ? 1. : -1.;
#pragma omp parallel for schedule(static)
    for (int sd = c->firstInnerSide - 1; sd < c->lastMPISide_MINE; sd++)
        for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
            double se = c->SurfElem[p + n * (q + (size_t)n * sd)];
            for (int v = 0; v < NL; v++)
                Flux[IDX_FACE(s, NL, v, p, q, sd)] = 0.5 * se * (sig * s->UPrim_master[IDX_FACE(s, NP, PRIM_LIFT[v], p, q, sd)]
                                                                 + s->UPrim_slave[IDX_FACE(s, NP, PRIM_LIFT[v], p, q, sd)]);
        }
}

static void lifting_volint(dgo *s);
/* (MPI) between the two parts the lifting flux on YOUR sides arrives from the master rank, lifting_br1.t90:93-112 */
static void lifting_br1_finish(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    double *Flux = s->gradUz_slave;
    /* lifting_fillflux.t90:212-256 Lifting_FillFlux_NormVec */
#pragma omp parallel for schedule(static)
    for (int sd = 0; sd < c->nSides; sd++)
        for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
            const double *nv = &c->NormVec[IDX_FACE(s, 3, 0, p, q, sd)];
            for (int v = 0; v < NL; v++) {
                double fl = Flux[IDX_FACE(s, NL, v, p, q, sd)];
                s->gradUx_master[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[0];
                s->gradUy_master[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[1];
                s->gradUz_master[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[2];
            }
        }
    /* big mortar sides: project the small-side lifting fluxes (strong form: no sign change), as the host code does
     * with Flux_MortarLifting on FluxX/Y/Z (lifting_br2.t90:102-106) */
    const int weak = lift_weak(c);
    flux_mortar_all(s, NL, s->gradUx_master, s->gradUx_master, weak);
    flux_mortar_all(s, NL, s->gradUy_master, s->gradUy_master, weak);
    flux_mortar_all(s, NL, s->gradUz_master, s->gradUz_master, weak);
    lifting_volint(s);
    /* lifting_br1.t90:118-124 SurfIntLifting x3 (single flux, strong or weak sign, with sJ) */
    surf_int(s, NL, s->gradUx_master, NULL, s->gradUx, 1, weak, 1);
    surf_int(s, NL, s->gradUy_master, NULL, s->gradUy, 1, weak, 1);
    surf_int(s, NL, s->gradUz_master, NULL, s->gradUz, 1, weak, 1);
    /* lifting_br1.t90:152-156 ProlongToFaceLifting x3 */
    prolong_to_face(s, NL, s->gradUx, s->gradUx_master, s->gradUx_slave);
    prolong_to_face(s, NL, s->gradUy, s->gradUy_master, s->gradUy_slave);
    prolong_to_face(s, NL, s->gradUz, s->gradUz_master, s->gradUz_slave);
    /* gradients of the big side to its small sides (U_MortarLifting, lifting_br2.t90:175-181) */
    u_mortar_all(s, NL, s->gradUx_master, s->gradUx_slave);
    u_mortar_all(s, NL, s->gradUy_master, s->gradUy_slave);
    u_mortar_all(s, NL, s->gradUz_master, s->gradUz_slave);
}

/* lifting_volint.t90:125-200 Lifting_VolInt_Conservative (x3 directions): DMat = D_Hat_T (weak) or D_T (strong) */
static void lifting_volint_conservative(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const double *DMat = lift_weak(c) ? c->D_Hat_T : c->D_T;
    double *grad[3] = {s->gradUx, s->gradUy, s->gradUz};
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++)
        for (int dir = 0; dir < 3; dir++)
            for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++)
                for (int v = 0; v < NL; v++) {
                    double a = 0.;
                    for (int l = 0; l < n; l++) {
                        double f = c->Metrics_fTilde[IDX_VOL(s, 3, dir, l, j, k, e)] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], l, j, k, e)];
                        double h = c->Metrics_hTilde[IDX_VOL(s, 3, dir, i, j, l, e)] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], i, j, l, e)];
                        double g = c->Metrics_gTilde[IDX_VOL(s, 3, dir, i, l, k, e)] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], i, l, k, e)];
                        if (l == 0) a = DMat[0 + n * i] * f + DMat[0 + n * k] * h + DMat[0 + n * j] * g;
                        else a = a + DMat[l + n * i] * f + DMat[l + n * k] * h + DMat[l + n * j] * g;
                    }
                    grad[dir][IDX_VOL(s, NL, v, i, j, k, e)] = a;
                }
}

/* lifting_volint.t90:262-328 Lifting_VolInt_Nonconservative_GPU_Kernel */
static void lifting_volint(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    if (lift_cons(c)) { lifting_volint_conservative(s); return; }
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++)
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            double gxi[NL], get[NL], gze[NL];
            for (int v = 0; v < NL; v++) gxi[v] = get[v] = gze[v] = 0.;
            for (int l = 0; l < n; l++)
                for (int v = 0; v < NL; v++) {
                    gxi[v] = gxi[v] + c->D_T[l + n * i] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], l, j, k, e)];
                    get[v] = get[v] + c->D_T[l + n * j] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], i, l, k, e)];
                    gze[v] = gze[v] + c->D_T[l + n * k] * s->UPrim[IDX_VOL(s, NP, PRIM_LIFT[v], i, j, l, e)];
                }
            const double *Mf = &c->Metrics_fTilde[IDX_VOL(s, 3, 0, i, j, k, e)];
            const double *Mg = &c->Metrics_gTilde[IDX_VOL(s, 3, 0, i, j, k, e)];
            const double *Mh = &c->Metrics_hTilde[IDX_VOL(s, 3, 0, i, j, k, e)];
            for (int v = 0; v < NL; v++) {
                s->gradUx[IDX_VOL(s, NL, v, i, j, k, e)] = Mf[0] * gxi[v] + Mg[0] * get[v] + Mh[0] * gze[v];
                s->gradUy[IDX_VOL(s, NL, v, i, j, k, e)] = Mf[1] * gxi[v] + Mg[1] * get[v] + Mh[1] * gze[v];
                s->gradUz[IDX_VOL(s, NL, v, i, j, k, e)] = Mf[2] * gxi[v] + Mg[2] * get[v] + Mh[2] * gze[v];
            }
        }
}

/* volume index of side node (p,q) at depth l: S2V(1:3,l,p,q,flip,locSide), mappings.f90:337-377 after the flip */
static inline void s2v3(const dgo *s, int l, int p, int q, int flip, int loc, int *ijk)
{
    const int n = s->n, N = s->c.N;
    const int a = s2v2(s->c.S2V2, n, 1, p, q, flip, loc), b = s2v2(s->c.S2V2, n, 2, p, q, flip, loc);
    switch (loc) {
    case XI_MINUS:   ijk[0] = l;     ijk[1] = a; ijk[2] = b; break;
    case XI_PLUS:    ijk[0] = N - l; ijk[1] = a; ijk[2] = b; break;
    case ETA_MINUS:  ijk[0] = a; ijk[1] = l;     ijk[2] = b; break;
    case ETA_PLUS:   ijk[0] = a; ijk[1] = N - l; ijk[2] = b; break;
    case ZETA_MINUS: ijk[0] = a; ijk[1] = b; ijk[2] = l;     break;
    default:         ijk[0] = a; ijk[1] = b; ijk[2] = N - l; break;
    }
}

/* lifting_br2.t90:193-311 Lifting_SurfInt_BR2 for one gradient direction: slave sides, then master sides */
static void lifting_surfint_br2(dgo *s, const double *Flux, double *gradU, double *gm, double *gs)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const int lmax = (c->nodeType == 2) ? 1 : n; /* Gauss-Lobatto: only the boundary node (l=0) */
    /* slave sides: inner sides and MPI YOUR sides */
    const int sr[2][2] = {{c->firstInnerSide, c->lastInnerSide}, {c->firstMPISide_YOUR, c->lastMPISide_YOUR}};
    for (int r = 0; r < 2; r++)
        for (int sd = sr[r][0] - 1; sd < sr[r][1]; sd++) {
            const int nbElemID = c->SideToElem[1 + 5 * sd];
            if (nbElemID <= 0) continue;
            const int nbloc = c->SideToElem[3 + 5 * sd], flip = c->SideToElem[4 + 5 * sd];
            const double eta = c->etaBR2;
            for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) for (int l = 0; l < lmax; l++) {
                int ijk[3];
                s2v3(s, l, p, q, flip, nbloc, ijk);
                const double sJ = c->sJ[ijk[0] + n * (ijk[1] + n * (ijk[2] + (size_t)n * (nbElemID - 1)))];
                for (int v = 0; v < NL; v++) {
                    double F_loc = sJ * Flux[IDX_FACE(s, NL, v, p, q, sd)] * c->L_HatMinus[l];
                    size_t id = IDX_VOL(s, NL, v, ijk[0], ijk[1], ijk[2], nbElemID - 1);
                    gradU[id] = gradU[id] + F_loc;
                    gs[IDX_FACE(s, NL, v, p, q, sd)] = gs[IDX_FACE(s, NL, v, p, q, sd)] + eta * c->L_Minus[l] * F_loc;
                }
            }
        }
    /* master sides: everything but YOUR, plus the MPI mortars */
    const int mr[2][2] = {{1, c->lastMPISide_MINE}, {c->firstMortarMPISide, c->lastMortarMPISide}};
    for (int r = 0; r < 2; r++)
        for (int sd = mr[r][0] - 1; sd < mr[r][1]; sd++) {
            double eta = c->etaBR2;
            if (sd < c->nBCSides && (c->BCSides[0 + 2 * sd] == 4 || c->BCSides[0 + 2 * sd] == 3)) eta = c->etaBR2_wall;
            const int ElemID = c->SideToElem[0 + 5 * sd];
            if (ElemID <= 0) continue;
            const int loc = c->SideToElem[2 + 5 * sd];
            for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) for (int l = 0; l < lmax; l++) {
                int ijk[3];
                s2v3(s, l, p, q, 0, loc, ijk);
                const double sJ = c->sJ[ijk[0] + n * (ijk[1] + n * (ijk[2] + (size_t)n * (ElemID - 1)))];
                for (int v = 0; v < NL; v++) {
                    double F_loc = sJ * Flux[IDX_FACE(s, NL, v, p, q, sd)] * c->L_HatMinus[l];
                    size_t id = IDX_VOL(s, NL, v, ijk[0], ijk[1], ijk[2], ElemID - 1);
                    gradU[id] = gradU[id] + F_loc;
                    gm[IDX_FACE(s, NL, v, p, q, sd)] = gm[IDX_FACE(s, NL, v, p, q, sd)] + eta * c->L_Minus[l] * F_loc;
                }
            }
        }
}

/* dg/lifting/lifting_br2.t90:43-184 Lifting_BR2 (host FLEXI code; non-conservative volume part). The untransformed
 * strong-form flux 1/2 (U_s - U_m) SurfElem (BC sides: lifting_fillflux.t90:109-140) is the BR1 one; FluxX/Y/Z are
 * that flux times the normal components. */
static void lifting_br2_finish(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const double *Flux = s->gradUz_slave;
#pragma omp parallel for schedule(static)
    for (int sd = 0; sd < c->nSides; sd++)
        for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
            const double *nv = &c->NormVec[IDX_FACE(s, 3, 0, p, q, sd)];
            for (int v = 0; v < NL; v++) {
                double fl = Flux[IDX_FACE(s, NL, v, p, q, sd)];
                s->FluxX[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[0];
                s->FluxY[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[1];
                s->FluxZ[IDX_FACE(s, NL, v, p, q, sd)] = fl * nv[2];
            }
        }
    flux_mortar_all(s, NL, s->FluxX, s->FluxX, 0);
    flux_mortar_all(s, NL, s->FluxY, s->FluxY, 0);
    flux_mortar_all(s, NL, s->FluxZ, s->FluxZ, 0);
    lifting_volint(s);
    /* ApplyJacobianLifting(toPhysical), lifting_br2.t90:118-124 */
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nDOF; d++)
        for (int v = 0; v < NL; v++) {
            s->gradUx[NL * d + v] = s->gradUx[NL * d + v] * c->sJ[d];
            s->gradUy[NL * d + v] = s->gradUy[NL * d + v] * c->sJ[d];
            s->gradUz[NL * d + v] = s->gradUz[NL * d + v] * c->sJ[d];
        }
    prolong_to_face(s, NL, s->gradUx, s->gradUx_master, s->gradUx_slave);
    prolong_to_face(s, NL, s->gradUy, s->gradUy_master, s->gradUy_slave);
    prolong_to_face(s, NL, s->gradUz, s->gradUz_master, s->gradUz_slave);
    lifting_surfint_br2(s, s->FluxX, s->gradUx, s->gradUx_master, s->gradUx_slave);
    lifting_surfint_br2(s, s->FluxY, s->gradUy, s->gradUy_master, s->gradUy_slave);
    lifting_surfint_br2(s, s->FluxZ, s->gradUz, s->gradUz_master, s->gradUz_slave);
    u_mortar_all(s, NL, s->gradUx_master, s->gradUx_slave);
    u_mortar_all(s, NL, s->gradUy_master, s->gradUy_slave);
    u_mortar_all(s, NL, s->gradUz_master, s->gradUz_slave);
}

static void lifting_finish(dgo *s) { if (s->c.lifting == 2) lifting_br2_finish(s); else lifting_br1_finish(s); }
static void lifting_br1(dgo *s) { lifting_br1_fillflux(s); lifting_finish(s); }

/* ------------------------------------------------------------------------------------------------ */
/* dg/applydmatrix.t90:19-75 ApplyDMatrix_Kernel */
static void apply_dmatrix(const dgo *s, double *Ut, int overwrite)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++)
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++)
            for (int v = 0; v < NV; v++) {
                double a = 0.;
                for (int l = 0; l < n; l++)
                    a = a + c->D_Hat_T[l + n * i] * s->f[IDX_VOL(s, NV, v, l, j, k, e)]
                          + c->D_Hat_T[l + n * k] * s->h[IDX_VOL(s, NV, v, i, j, l, e)]
                          + c->D_Hat_T[l + n * j] * s->g[IDX_VOL(s, NV, v, i, l, k, e)];
                size_t id = IDX_VOL(s, NV, v, i, j, k, e);
                Ut[id] = overwrite ? a : Ut[id] + a;
            }
}

/* dg/volint.f90:60-119 (visc), :129-188 (weak), :211-353 (split) */
static void vol_int(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const int split = c->splitDG >= 0;
    if (!split || c->parabolic) {
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++) {
            const double *P = &s->UPrim[NP * d];
            double mu = 0., lambda = 0.;
            if (c->parabolic) { mu = viscosity(c, P); lambda = conductivity(c, mu); }
            eval_transformed_flux(!split, c->parabolic, &s->U[NV * d], P,
                                  c->parabolic ? &s->gradUx[NL * d] : NULL, c->parabolic ? &s->gradUy[NL * d] : NULL,
                                  c->parabolic ? &s->gradUz[NL * d] : NULL, &s->f[NV * d], &s->g[NV * d], &s->h[NV * d],
                                  &c->Metrics_fTilde[3 * d], &c->Metrics_gTilde[3 * d], &c->Metrics_hTilde[3 * d], mu, lambda);
        }
        apply_dmatrix(s, s->Ut, 1);
    }
    if (split) { /* volint.f90:265-353 VolInt_splitForm_Kernel */
#pragma omp parallel for schedule(static)
        for (int e = 0; e < c->nElems; e++)
            for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
                double acc[NV], Fl[NV];
                for (int v = 0; v < NV; v++) acc[v] = 0.;
                size_t d0 = IDX_VOL(s, 1, 0, i, j, k, e);
                for (int l = 0; l < n; l++) {
                    size_t d1 = IDX_VOL(s, 1, 0, l, j, k, e);
                    split_volume_flux(c->splitDG, &s->U[NV * d0], &s->UPrim[NP * d0], &s->U[NV * d1], &s->UPrim[NP * d1],
                                      &c->Metrics_fTilde[3 * d0], &c->Metrics_fTilde[3 * d1], Fl);
                    for (int v = 0; v < NV; v++) acc[v] = acc[v] + c->DVolSurf[l + n * i] * Fl[v];
                }
                for (int l = 0; l < n; l++) {
                    size_t d1 = IDX_VOL(s, 1, 0, i, l, k, e);
                    split_volume_flux(c->splitDG, &s->U[NV * d0], &s->UPrim[NP * d0], &s->U[NV * d1], &s->UPrim[NP * d1],
                                      &c->Metrics_gTilde[3 * d0], &c->Metrics_gTilde[3 * d1], Fl);
                    for (int v = 0; v < NV; v++) acc[v] = acc[v] + c->DVolSurf[l + n * j] * Fl[v];
                }
                for (int l = 0; l < n; l++) {
                    size_t d1 = IDX_VOL(s, 1, 0, i, j, l, e);
                    split_volume_flux(c->splitDG, &s->U[NV * d0], &s->UPrim[NP * d0], &s->U[NV * d1], &s->UPrim[NP * d1],
                                      &c->Metrics_hTilde[3 * d0], &c->Metrics_hTilde[3 * d1], Fl);
                    for (int v = 0; v < NV; v++) acc[v] = acc[v] + c->DVolSurf[l + n * k] * Fl[v];
                }
                for (int v = 0; v < NV; v++) {
                    if (c->parabolic) s->Ut[NV * d0 + v] = s->Ut[NV * d0 + v] + acc[v];
                    else s->Ut[NV * d0 + v] = acc[v];
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* dg/fillflux.f90:45-172 FillFlux (BC sides, inner + MPI-MINE sides), getboundaryflux.f90:542-838 */
static int fill_flux(dgo *s)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const double kappa = c->EOS[EOS_KAPPA];
    int err = 0;
#pragma omp parallel for schedule(static) reduction(|:err)
    for (int sd = 0; sd < c->lastMPISide_MINE; sd++)
        for (int q = 0; q < n; q++) for (int p = 0; p < n; p++) {
            if (sd >= c->nBCSides && sd < c->firstInnerSide - 1) continue; /* big mortar sides: filled by Flux_Mortar */
            const double *nv = &c->NormVec[IDX_FACE(s, 3, 0, p, q, sd)];
            const double *t1 = &c->TangVec1[IDX_FACE(s, 3, 0, p, q, sd)];
            const double *t2 = &c->TangVec2[IDX_FACE(s, 3, 0, p, q, sd)];
            const double *Pm = &s->UPrim_master[IDX_FACE(s, NP, 0, p, q, sd)];
            double *F = &s->Flux_master[IDX_FACE(s, NV, 0, p, q, sd)];
            const double *gxm = &s->gradUx_master[IDX_FACE(s, NL, 0, p, q, sd)];
            const double *gym = &s->gradUy_master[IDX_FACE(s, NL, 0, p, q, sd)];
            const double *gzm = &s->gradUz_master[IDX_FACE(s, NL, 0, p, q, sd)];
            if (sd < c->nBCSides) {
                int bct = c->BCSides[0 + 2 * sd], bcs = c->BCSides[1 + 2 * sd];
                double Pb[NP];
                err |= get_boundary_state(c, bct, Pb, Pm, &c->RefStatePrim[NP * (bcs > 0 ? bcs - 1 : 0)], nv, t1, t2);
                if (is_riemann_bc(bct)) {
                    double Um[NV], Ub[NV];
                    prim_to_cons(Pm, Um, kappa);
                    prim_to_cons(Pb, Ub, kappa);
                    riemann(c, F, Um, Ub, Pm, Pb, nv, t1, t2);
                    if (c->parabolic) { /* riemann.f90:485-528 ViscousFlux (point), master gradients on both sides */
                        double fL[NV], gL[NV], hL[NV], fR[NV], gR[NV], hR[NV];
                        double mu = viscosity(c, Pm), la = conductivity(c, mu);
                        eval_diff_flux3d(Pm, gxm, gym, gzm, fL, gL, hL, mu, la);
                        mu = viscosity(c, Pb); la = conductivity(c, mu);
                        eval_diff_flux3d(Pb, gxm, gym, gzm, fR, gR, hR, mu, la);
                        for (int v = 0; v < NV; v++)
                            F[v] = F[v] + 0.5 * (nv[0] * (fL[v] + fR[v]) + nv[1] * (gL[v] + gR[v]) + nv[2] * (hL[v] + hR[v]));
                    }
                } else if (is_wall_bc(bct)) {
                    F[DENS] = 0.;
                    for (int d = 0; d < 3; d++) F[MOM1 + d] = Pb[PRES] * nv[d];
                    F[ENER] = 0.;
                    if (c->parabolic) {
                        double mu = viscosity(c, Pb), la = conductivity(c, mu);
                        double fd[NV], gd[NV], hd[NV];
                        if (bct == 9) {
                            double B[3][3], gxf[NL], gyf[NL], gzf[NL];
                            B[0][0] = 1. - nv[0] * nv[0]; B[1][1] = 1. - nv[1] * nv[1]; B[2][2] = 1. - nv[2] * nv[2];
                            B[0][1] = -nv[0] * nv[1]; B[0][2] = -nv[0] * nv[2]; B[2][1] = -nv[2] * nv[1];
                            B[1][0] = B[0][1]; B[2][0] = B[0][2]; B[1][2] = B[2][1];
                            for (int v = 0; v < NL; v++) {
                                gxf[v] = B[0][0] * gxm[v] + B[0][1] * gym[v] + B[0][2] * gzm[v];
                                gyf[v] = B[1][0] * gxm[v] + B[1][1] * gym[v] + B[1][2] * gzm[v];
                                gzf[v] = B[2][0] * gxm[v] + B[2][1] * gym[v] + B[2][2] * gzm[v];
                            }
                            eval_diff_flux3d(Pb, gxf, gyf, gzf, fd, gd, hd, mu, la);
                        } else if (bct == 91) { /* :716-781 slip wall, version 2 */
                            double B[3][3], gf[3][NL];
                            const double *gm_[3] = {gxm, gym, gzm};
                            const double *tv[3] = {nv, t1, t2};
                            B[0][0] = 1. - nv[0] * nv[0]; B[1][1] = 1. - nv[1] * nv[1]; B[2][2] = 1. - nv[2] * nv[2];
                            B[0][1] = -nv[0] * nv[1]; B[0][2] = -nv[0] * nv[2]; B[2][1] = -nv[2] * nv[1];
                            B[1][0] = B[0][1]; B[2][0] = B[0][2]; B[1][2] = B[2][1];
                            for (int d = 0; d < 3; d++) {
                                gf[d][LIFT_DENS] = 0.; /* not set by the reference; unused by the viscous flux */
                                gf[d][LIFT_TEMP] = B[d][0] * gxm[LIFT_TEMP] + B[d][1] * gym[LIFT_TEMP] + B[d][2] * gzm[LIFT_TEMP];
                            }
                            /* first: gradients (x,y,z) of the wall-aligned velocities w = (normal, tang1, tang2) */
                            double gw[3][3]; /* [dir x,y,z][w] */
                            for (int d = 0; d < 3; d++)
                                for (int w = 0; w < 3; w++)
                                    gw[d][w] = tv[w][0] * gm_[d][LIFT_VEL1] + tv[w][1] * gm_[d][LIFT_VEL2] + tv[w][2] * gm_[d][LIFT_VEL3];
                            /* second: derivatives along the wall-aligned directions a = (n,t1,t2); boundary conditions */
                            double ga[3][3]; /* [a][w] */
                            for (int a = 0; a < 3; a++)
                                for (int w = 0; w < 3; w++)
                                    ga[a][w] = tv[a][0] * gw[0][w] + tv[a][1] * gw[1][w] + tv[a][2] * gw[2][w];
                            ga[0][1] = 0.; ga[0][2] = 0.; ga[1][0] = 0.; ga[2][0] = 0.;
                            /* third: back to x/y/z derivatives */
                            for (int d = 0; d < 3; d++)
                                for (int w = 0; w < 3; w++)
                                    gw[d][w] = nv[d] * ga[0][w] + t1[d] * ga[1][w] + t2[d] * ga[2][w];
                            /* fourth: back to the Cartesian velocity components */
                            for (int d = 0; d < 3; d++)
                                for (int x = 0; x < 3; x++)
                                    gf[d][LIFT_VEL1 + x] = nv[x] * gw[d][0] + t1[x] * gw[d][1] + t2[x] * gw[d][2];
                            eval_diff_flux3d(Pb, gf[0], gf[1], gf[2], fd, gd, hd, mu, la);
                        } else {
                            eval_diff_flux3d(Pb, gxm, gym, gzm, fd, gd, hd, mu, la);
                            if (bct == 3) fd[ENER] = gd[ENER] = hd[ENER] = 0.;
                        }
                        for (int v = 0; v < NV; v++) F[v] = F[v] + nv[0] * fd[v] + nv[1] * gd[v] + nv[2] * hd[v];
                    }
                } else err |= 1;
            } else {
                const double *Ps = &s->UPrim_slave[IDX_FACE(s, NP, 0, p, q, sd)];
                riemann(c, F, &s->U_master[IDX_FACE(s, NV, 0, p, q, sd)], &s->U_slave[IDX_FACE(s, NV, 0, p, q, sd)], Pm, Ps, nv, t1, t2);
                if (c->parabolic) { /* riemann.f90:638-700 ViscousFlux_Kernel_CUDA */
                    double nd[NV];
                    double mu = viscosity(c, Pm), la = conductivity(c, mu);
                    eval_normal_diff_flux3d(Pm, gxm, gym, gzm, nd, nv, mu, la);
                    for (int v = 0; v < NV; v++) F[v] = F[v] + 0.5 * nd[v];
                    mu = viscosity(c, Ps); la = conductivity(c, mu);
                    eval_normal_diff_flux3d(Ps, &s->gradUx_slave[IDX_FACE(s, NL, 0, p, q, sd)], &s->gradUy_slave[IDX_FACE(s, NL, 0, p, q, sd)],
                                            &s->gradUz_slave[IDX_FACE(s, NL, 0, p, q, sd)], nd, nv, mu, la);
                    for (int v = 0; v < NV; v++) F[v] = F[v] + 0.5 * nd[v];
                }
            }
            double se = c->SurfElem[p + n * (q + (size_t)n * sd)];
            for (int v = 0; v < NV; v++) {
                F[v] = F[v] * se;
                s->Flux_slave[IDX_FACE(s, NV, v, p, q, sd)] = F[v];
            }
        }
    return err;
}

/* ------------------------------------------------------------------------------------------------ */
/* idealgas/exactfunc.f90:946-1113 CalcSource_CUDA_Kernel, CASE(4) (3D): Ut += Ut_src / sJ, called between the sign change
 * and the Jacobian (dg.f90:413-423) */
static void calc_source(dgo *s, double t)
{
    const dgo_config *c = &s->c;
    if (c->iniExactFunc != 4) return;
    const double Kappa = c->EOS[EOS_KAPPA], PP_Pi = acos(-1.0);
    const double Frequency = 1., Amplitude = 0.1;
    const double Omega = PP_Pi * Frequency, a = c->AdvVel[0] * 2. * PP_Pi;
    double tmp[6];
    tmp[0] = -a + 3. * Omega;
    tmp[1] = -a + 0.5 * Omega * (1. + Kappa * 5.);
    tmp[2] = Amplitude * Omega * (Kappa - 1.);
    tmp[3] = 0.5 * ((9. + Kappa * 15.) * Omega - 8. * a);
    tmp[4] = Amplitude * (3. * Omega * Kappa - a);
    tmp[5] = c->parabolic ? 3. * c->EOS[EOS_MU0] * Kappa * Omega * Omega / c->EOS[EOS_PR] : 0.;
    for (int x = 0; x < 6; x++) tmp[x] = tmp[x] * Amplitude;
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nDOF; d++) {
        const double *X = &c->Elem_xGP[3 * d];
        const double arg = Omega * (X[0] + X[1] + X[2]) - a * t;
        const double cosX = cos(arg), sinX = sin(arg), sin2 = 2. * sinX * cosX;
        double src[NV];
        src[DENS] = tmp[0] * cosX;
        src[MOM1] = src[MOM2] = src[MOM3] = tmp[1] * cosX + tmp[2] * sin2;
        src[ENER] = tmp[3] * cosX + tmp[4] * sin2 + tmp[5] * sinX;
        for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] + src[v] / c->sJ[d];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* filter/filter.f90:272-306 Filter -> interpolation/changeBasis.t90:287-360 ChangeBasis3D_GPU with X_Out absent: U is
 * replaced by FilterMat applied along xi, then eta, then zeta */
static void filter_array(dgo *s, double *A, const double *Mat)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++) {
        double b1[NV * 4096], b2[NV * 4096]; /* n <= 16 */
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++)
            for (int v = 0; v < NV; v++) {
                double a = 0.;
                for (int l = 0; l < n; l++) a = a + Mat[i + n * l] * A[IDX_VOL(s, NV, v, l, j, k, e)];
                b1[v + NV * (i + n * (j + n * k))] = a;
            }
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++)
            for (int v = 0; v < NV; v++) {
                double a = 0.;
                for (int l = 0; l < n; l++) a = a + Mat[j + n * l] * b1[v + NV * (i + n * (l + n * k))];
                b2[v + NV * (i + n * (j + n * k))] = a;
            }
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++)
            for (int v = 0; v < NV; v++) {
                double a = 0.;
                for (int l = 0; l < n; l++) a = a + Mat[k + n * l] * b2[v + NV * (i + n * (j + n * l))];
                A[IDX_VOL(s, NV, v, i, j, k, e)] = a;
            }
    }
}
static void filter_u(dgo *s)
{
    if (s->c.FilterMat) filter_array(s, s->U, s->c.FilterMat);
}

/* dg/overintegration.f90:212-339 FilterConservative (3-D branch): JU_t on N -> NUnder (xi, eta, zeta in this order, each sum
 * started with l = 0), times sJNUnder, back to N (xi, eta, zeta); the zeta passes accumulate into a zeroed array like the
 * reference (VNullify + "U = U + ..."). Input JU_t, output U_t on N. */
static void filter_conservative(dgo *s, double *A)
{
    const dgo_config *c = &s->c;
    const int n = s->n, nu = c->NUnder + 1;
    const double *VD = c->Vdm_N_NUnder;  /* (0:NUnder,0:N): VD[iU + nu*i] */
    const double *VU = c->Vdm_NUnder_N;  /* (0:N,0:NUnder): VU[i + n*iU] */
#pragma omp parallel for schedule(static)
    for (int e = 0; e < c->nElems; e++) {
        double *B1 = (double *)malloc(sizeof(double) * NV * (size_t)nu * n * n);    /* (nVar,0:NUnder,0:N,0:N) */
        double *B2 = (double *)malloc(sizeof(double) * NV * (size_t)nu * nu * n);   /* (nVar,0:NUnder,0:NUnder,0:N) */
        double *UL = (double *)calloc((size_t)NV * nu * nu * nu, sizeof(double));   /* U_loc (nVar,0:NUnder,0:NUnder,0:NUnder) */
        double *B3 = (double *)malloc(sizeof(double) * NV * (size_t)n * nu * nu);   /* (nVar,0:N,0:NUnder,0:NUnder) */
        double *B4 = (double *)malloc(sizeof(double) * NV * (size_t)n * n * nu);    /* (nVar,0:N,0:N,0:NUnder) */
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++)
            for (int iU = 0; iU < nu; iU++) for (int v = 0; v < NV; v++) {
                double a = VD[iU + nu * 0] * A[IDX_VOL(s, NV, v, 0, j, k, e)];
                for (int i = 1; i < n; i++) a = a + VD[iU + nu * i] * A[IDX_VOL(s, NV, v, i, j, k, e)];
                B1[v + NV * (iU + nu * (j + n * k))] = a;
            }
        for (int k = 0; k < n; k++)
            for (int jU = 0; jU < nu; jU++) for (int iU = 0; iU < nu; iU++) for (int v = 0; v < NV; v++) {
                double a = VD[jU + nu * 0] * B1[v + NV * (iU + nu * (0 + n * k))];
                for (int j = 1; j < n; j++) a = a + VD[jU + nu * j] * B1[v + NV * (iU + nu * (j + n * k))];
                B2[v + NV * (iU + nu * (jU + nu * k))] = a;
            }
        for (int k = 0; k < n; k++)
            for (int kU = 0; kU < nu; kU++) for (int jU = 0; jU < nu; jU++) for (int iU = 0; iU < nu; iU++) for (int v = 0; v < NV; v++)
                UL[v + NV * (iU + nu * (jU + nu * kU))] = UL[v + NV * (iU + nu * (jU + nu * kU))] + VD[kU + nu * k] * B2[v + NV * (iU + nu * (jU + nu * k))];
        for (int kU = 0; kU < nu; kU++) for (int jU = 0; jU < nu; jU++) for (int iU = 0; iU < nu; iU++) for (int v = 0; v < NV; v++)
            UL[v + NV * (iU + nu * (jU + nu * kU))] = UL[v + NV * (iU + nu * (jU + nu * kU))] * c->sJNUnder[iU + nu * (jU + nu * (kU + (size_t)nu * e))];
        for (int kU = 0; kU < nu; kU++) for (int jU = 0; jU < nu; jU++)
            for (int i = 0; i < n; i++) for (int v = 0; v < NV; v++) {
                double a = VU[i + n * 0] * UL[v + NV * (0 + nu * (jU + nu * kU))];
                for (int iU = 1; iU < nu; iU++) a = a + VU[i + n * iU] * UL[v + NV * (iU + nu * (jU + nu * kU))];
                B3[v + NV * (i + n * (jU + nu * kU))] = a;
            }
        for (int kU = 0; kU < nu; kU++)
            for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) for (int v = 0; v < NV; v++) {
                double a = VU[j + n * 0] * B3[v + NV * (i + n * (0 + nu * kU))];
                for (int jU = 1; jU < nu; jU++) a = a + VU[j + n * jU] * B3[v + NV * (i + n * (jU + nu * kU))];
                B4[v + NV * (i + n * (j + n * kU))] = a;
            }
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) for (int v = 0; v < NV; v++) {
            double a = 0.;
            for (int kU = 0; kU < nu; kU++) a = a + VU[k + n * kU] * B4[v + NV * (i + n * (j + n * kU))];
            A[IDX_VOL(s, NV, v, i, j, k, e)] = a;
        }
        free(B1); free(B2); free(UL); free(B3); free(B4);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* dg/dg.f90:255-425 DGTimeDerivative_weakForm (single rank: no halo phases) */
int dgo_time_derivative(dgo *s, double t)
{
    const dgo_config *c = &s->c;
    const double kappa = c->EOS[EOS_KAPPA], R = c->EOS[EOS_R];
    /* 1. dg.f90:331 */ filter_u(s);
    /* 2. */ prolong_to_face(s, NV, s->U, s->U_master, s->U_slave);
    /* 2b. (host FLEXI) big mortar sides -> small sides */ u_mortar_all(s, NV, s->U_master, s->U_slave);
    /* 3. eos.f90:328 ConsToPrim volume */
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nDOF; d++) cons_to_prim(&s->UPrim[NP * d], &s->U[NV * d], kappa, R);
    /* 5. equation.f90:198-261 GetPrimitiveStateSurface */
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nFace; d++) {
        cons_to_prim(&s->UPrim_master[NP * d], &s->U_master[NV * d], kappa, R);
        cons_to_prim(&s->UPrim_slave[NP * d], &s->U_slave[NV * d], kappa, R);
    }
    /* 6. */ if (c->parabolic) lifting_br1(s);
    /* 8. */ vol_int(s);
    /* 11. */ int err = fill_flux(s);
    /* 11b. (host FLEXI) small-side fluxes -> big mortar sides, weak form */ flux_mortar_all(s, NV, s->Flux_master, s->Flux_slave, 1);
    /* 11.5 */ surf_int(s, NV, s->Flux_master, s->Flux_slave, s->Ut, 0, 0, 0);
    /* 12. vector.f90:210 VAX_GPU(-1), 14. applyjacobian.t90:196 */
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nDOF; d++)
        for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] * (-1.);
    /* 13. */ calc_source(s, t);
    if (c->SpongeMat) { /* sponge.f90:529-588 Sponge */
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++)
            for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] - c->SpongeMat[d] * (s->U[NV * d + v] - s->SpBaseFlow[NV * d + v]);
    }
    if (c->tcSource) { /* dg.f90:419 TestcaseSource (commented out in GALAEXI; host FLEXI testcase/channel/testcase.f90:277-296) */
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++) {
            s->Ut[NV * d + MOM1] = s->Ut[NV * d + MOM1] - c->dpdx / c->sJ[d];
            s->Ut[NV * d + ENER] = s->Ut[NV * d + ENER] - c->dpdx / c->sJ[d] * c->BulkVel;
        }
    }
    /* 14. overintegration.f90:179-201 Overintegration(Ut): cut-off filter on JU_t, or the conservative variant which applies
     * the Jacobian (of NUnder) itself; otherwise applyjacobian.t90:196 */
    if (c->OverintegrationType == 1) filter_array(s, s->Ut, c->OverintegrationMat);
    if (c->OverintegrationType == 2) {
        filter_conservative(s, s->Ut);
        return err;
    }
#pragma omp parallel for schedule(static)
    for (size_t d = 0; d < s->nDOF; d++)
        for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] * c->sJ[d];
    return err;
}

/* The same routine cut at the reference's four halo phases (dg.f90:290-300 U_slave YOUR->MINE; lifting_br1.t90:93-112
 * lifting flux MINE->YOUR; lifting_br1.t90:158-165 / dg.f90:352-366 gradU*_slave YOUR->MINE; dg.f90:392-401 Flux_slave
 * MINE->YOUR). A multi-rank test driver calls phase 0..4 and performs the exchange named after each phase with any
 * message layer; calling the five phases back to back on one rank equals dgo_time_derivative. */
int dgo_rhs_phase(dgo *s, int phase)
{
    const dgo_config *c = &s->c;
    const double kappa = c->EOS[EOS_KAPPA], R = c->EOS[EOS_R];
    int err = 0;
    switch (phase) {
    case 0:
        filter_u(s);
        prolong_to_face(s, NV, s->U, s->U_master, s->U_slave);
        u_mortar_all(s, NV, s->U_master, s->U_slave);
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++) cons_to_prim(&s->UPrim[NP * d], &s->U[NV * d], kappa, R);
        break; /* -> exchange U_slave, YOUR -> MINE */
    case 1:
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nFace; d++) {
            cons_to_prim(&s->UPrim_master[NP * d], &s->U_master[NV * d], kappa, R);
            cons_to_prim(&s->UPrim_slave[NP * d], &s->U_slave[NV * d], kappa, R);
        }
        if (c->parabolic) lifting_br1_fillflux(s);
        break; /* -> exchange lifting flux (buffer gradUz_slave), MINE -> YOUR */
    case 2:
        if (c->parabolic) lifting_finish(s);
        break; /* -> exchange gradUx/y/z_slave, YOUR -> MINE */
    case 3:
        vol_int(s);
        err = fill_flux(s);
        break; /* -> exchange Flux_slave, MINE -> YOUR */
    case 4:
        flux_mortar_all(s, NV, s->Flux_master, s->Flux_slave, 1);
        surf_int(s, NV, s->Flux_master, s->Flux_slave, s->Ut, 0, 0, 0);
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++)
            for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] * (-1.);
        /* 14. overintegration (element-local: no halo involved), then / or the Jacobian */
        if (c->OverintegrationType == 1) filter_array(s, s->Ut, c->OverintegrationMat);
        if (c->OverintegrationType == 2) { filter_conservative(s, s->Ut); break; }
#pragma omp parallel for schedule(static)
        for (size_t d = 0; d < s->nDOF; d++)
            for (int v = 0; v < NV; v++) s->Ut[NV * d + v] = s->Ut[NV * d + v] * c->sJ[d];
        break;
    default: err = 4;
    }
    return err;
}

/* vector.f90:163-183 VAXPB_OUT_VAXPB_IN alone (the stage update after an externally driven RHS) */
void dgo_rk_update(dgo *s, double mRKA, double b_dt)
{
    size_t nTot = NV * s->nDOF;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nTot; i++) {
        s->Ut_tmp[i] = s->Ut_tmp[i] * mRKA + s->Ut[i];
        s->U[i] = s->U[i] + s->Ut_tmp[i] * b_dt;
    }
}

/* timedisc/timestep.f90:49-120 one stage of TimeStepByLSERKW2 + globals/vector.f90:163-183 VAXPB_OUT_VAXPB_IN */
int dgo_rk_stage(dgo *s, double tStage, double mRKA, double b_dt)
{
    int err = dgo_time_derivative(s, tStage);
    size_t nTot = NV * s->nDOF;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nTot; i++) {
        s->Ut_tmp[i] = s->Ut_tmp[i] * mRKA + s->Ut[i];
        s->U[i] = s->U[i] + s->Ut_tmp[i] * b_dt;
    }
    return err;
}

int dgo_rk_step(dgo *s, double t, double dt, int nStages, const double *RKA, const double *RKb, const double *RKc)
{
    int err = 0;
    for (int st = 0; st < nStages; st++) {
        double tStage = (st == 0) ? t : t + RKc[st] * dt;
        double mRKA = (st == 0) ? 0. : -1. * RKA[st];
        err |= dgo_rk_stage(s, tStage, mRKA, RKb[st] * dt);
    }
    return err;
}

/* calctimestep.f90:47-90 InitCalctimestep, :98-296 CalcTimeStep / CalcMaxEigenvalue.
 * The min over elements is taken as min_e(CFL*2/lambda_e): the value the reference's CUF reduction produces
 * when every thread owns one element (calctimestep.f90:157 has a misplaced parenthesis otherwise). */
double dgo_calc_timestep(dgo *s, double CFLScale, double DFLScale, double *dt_conv, double *dt_visc)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    const double kappa = c->EOS[EOS_KAPPA], R = c->EOS[EOS_R];
    double tc = HUGE_VAL, tv = HUGE_VAL;
    double KappasPr_max = fmax(4. / 3., kappa / c->EOS[EOS_PR]);
#pragma omp parallel for schedule(static) reduction(min:tc) reduction(min:tv)
    for (int e = 0; e < c->nElems; e++) {
        double mx[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            size_t d = IDX_VOL(s, 1, 0, i, j, k, e);
            const double *U = &s->U[NV * d];
            double sJ = c->sJ[d];
            const double *Mf = &c->Metrics_fTilde[3 * d], *Mg = &c->Metrics_gTilde[3 * d], *Mh = &c->Metrics_hTilde[3 * d];
            double sRho = 1. / U[DENS];
            double vel[3] = {U[MOM1] * sRho, U[MOM2] * sRho, U[MOM3] * sRho};
            double pres = (kappa - 1.) * (U[ENER] - 0.5 * (vel[0] * U[MOM1] + vel[1] * U[MOM2] + vel[2] * U[MOM3]));
            double prim[NP] = {U[DENS], vel[0], vel[1], vel[2], pres, pres * sRho / R};
            double cs = sqrt(kappa * pres * sRho);
            double vsJ[3] = {vel[0] * sJ, vel[1] * sJ, vel[2] * sJ};
            double adv[3] = {sJ * sqrt(Mf[0] * Mf[0] + Mf[1] * Mf[1] + Mf[2] * Mf[2]),
                             sJ * sqrt(Mg[0] * Mg[0] + Mg[1] * Mg[1] + Mg[2] * Mg[2]),
                             sJ * sqrt(Mh[0] * Mh[0] + Mh[1] * Mh[1] + Mh[2] * Mh[2])};
            mx[0] = fmax(mx[0], fabs(Mf[0] * vsJ[0] + Mf[1] * vsJ[1] + Mf[2] * vsJ[2]) + cs * adv[0]);
            mx[1] = fmax(mx[1], fabs(Mg[0] * vsJ[0] + Mg[1] * vsJ[1] + Mg[2] * vsJ[2]) + cs * adv[1]);
            mx[2] = fmax(mx[2], fabs(Mh[0] * vsJ[0] + Mh[1] * vsJ[1] + Mh[2] * vsJ[2]) + cs * adv[2]);
            if (c->parabolic) {
                double mu = viscosity(c, prim);
                double vis[3] = {KappasPr_max * ((Mf[0] * sJ) * (Mf[0] * sJ) + (Mf[1] * sJ) * (Mf[1] * sJ) + (Mf[2] * sJ) * (Mf[2] * sJ)),
                                 KappasPr_max * ((Mg[0] * sJ) * (Mg[0] * sJ) + (Mg[1] * sJ) * (Mg[1] * sJ) + (Mg[2] * sJ) * (Mg[2] * sJ)),
                                 KappasPr_max * ((Mh[0] * sJ) * (Mh[0] * sJ) + (Mh[1] * sJ) * (Mh[1] * sJ) + (Mh[2] * sJ) * (Mh[2] * sJ))};
                for (int d3 = 0; d3 < 3; d3++) mx[3 + d3] = fmax(mx[3 + d3], mu * sRho * vis[d3]);
            }
        }
        double lc = mx[0] + mx[1] + mx[2];
        tc = fmin(tc, CFLScale * 2. / lc);
        if (c->parabolic) { double lv = mx[3] + mx[4] + mx[5]; tv = fmin(tv, DFLScale * 4. / lv); }
    }
    if (dt_conv) *dt_conv = tc;
    if (dt_visc) *dt_visc = tv;
    return fmin(tc, tv);
}

/* ------------------------------------------------------------------------------------------------ */
dgo *dgo_create(const dgo_config *cfg)
{
    dgo *s = (dgo *)calloc(1, sizeof(dgo));
    s->c = *cfg;
    s->n = cfg->N + 1; s->n2 = s->n * s->n; s->n3 = s->n2 * s->n;
    s->nDOF = (size_t)s->n3 * cfg->nElems;
    s->nFace = (size_t)s->n2 * cfg->nSides;
#define AL(x, cnt) s->x = (double *)calloc((size_t)(cnt) + 1, sizeof(double))
    AL(U, NV * s->nDOF); AL(Ut, NV * s->nDOF); AL(UPrim, NP * s->nDOF); AL(Ut_tmp, NV * s->nDOF);
    AL(U_master, NV * s->nFace); AL(U_slave, NV * s->nFace); AL(UPrim_master, NP * s->nFace); AL(UPrim_slave, NP * s->nFace);
    AL(Flux_master, NV * s->nFace); AL(Flux_slave, NV * s->nFace);
    AL(gradUx, NL * s->nDOF); AL(gradUy, NL * s->nDOF); AL(gradUz, NL * s->nDOF);
    AL(gradUx_master, NL * s->nFace); AL(gradUy_master, NL * s->nFace); AL(gradUz_master, NL * s->nFace);
    AL(gradUx_slave, NL * s->nFace); AL(gradUy_slave, NL * s->nFace); AL(gradUz_slave, NL * s->nFace);
    AL(f, NV * s->nDOF); AL(g, NV * s->nDOF); AL(h, NV * s->nDOF);
    AL(FluxX, NL * s->nFace); AL(FluxY, NL * s->nFace); AL(FluxZ, NL * s->nFace);
    AL(SpBaseFlow, NV * s->nDOF);
#undef AL
    /* the reference initialises U_slave/UPrim_slave of BC sides to 0 and never touches them; prim of a zero
     * state would divide by zero, so the slave arrays of sides without a slave element get a benign state. */
    for (size_t d = 0; d < s->nFace; d++) { s->U_slave[NV * d] = 1.; s->U_slave[NV * d + 4] = 1.; s->U_master[NV * d] = 1.; s->U_master[NV * d + 4] = 1.; }
    return s;
}

void dgo_destroy(dgo *s)
{
    if (!s) return;
    double *p[] = {s->U, s->Ut, s->UPrim, s->Ut_tmp, s->U_master, s->U_slave, s->UPrim_master, s->UPrim_slave, s->Flux_master,
                   s->Flux_slave, s->gradUx, s->gradUy, s->gradUz, s->gradUx_master, s->gradUy_master, s->gradUz_master,
                   s->gradUx_slave, s->gradUy_slave, s->gradUz_slave, s->f, s->g, s->h, s->FluxX, s->FluxY, s->FluxZ, s->SpBaseFlow};
    for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(p[i]);
    free(s);
}

/* array access for the tests: name -> pointer */
double *dgo_array(dgo *s, const char *name)
{
#define R(x) if (!strcmp(name, #x)) return s->x
    R(U); R(Ut); R(UPrim); R(Ut_tmp); R(U_master); R(U_slave); R(UPrim_master); R(UPrim_slave); R(Flux_master); R(Flux_slave);
    R(gradUx); R(gradUy); R(gradUz); R(gradUx_master); R(gradUy_master); R(gradUz_master);
    R(gradUx_slave); R(gradUy_slave); R(gradUz_slave); R(SpBaseFlow);
#undef R
    return NULL;
}

/* stand-alone entry points used by the golden tests (unitTests/ProlongToFace.f90, unitTests/SurfInt.f90) */
void dgo_prolong_to_face(dgo *s, int nVar, const double *Uvol, double *Um, double *Us) { prolong_to_face(s, nVar, Uvol, Um, Us); }
void dgo_surf_int(dgo *s, int nVar, const double *Fm, const double *Fs, double *Ut) { surf_int(s, nVar, Fm, Fs, Ut, 0, 0, 0); }
void dgo_lifting(dgo *s) { lifting_br1(s); }
void dgo_filter(dgo *s) { filter_u(s); }
/* sponge/pruettdamping.f90:69-92 TempFilterTimeDeriv: the base flow follows the solution with the time scale tempFilterWidth */
void dgo_temp_filter_time_deriv(dgo *s, double dt, double tempFilterWidth)
{
    const double fac = dt / tempFilterWidth;
    size_t nTot = NV * s->nDOF;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nTot; i++) s->SpBaseFlow[i] = s->SpBaseFlow[i] + (s->U[i] - s->SpBaseFlow[i]) * fac;
}
void dgo_set_forcing(dgo *s, int on, double dpdx, double BulkVel) { s->c.tcSource = on; s->c.dpdx = dpdx; s->c.BulkVel = BulkVel; }
/* testcase/channel/testcase.f90:241-271 CalcForcing: BulkVel = 1/Vol sum u wGPVol / sJ over the solution nodes */
double dgo_bulk_velocity(dgo *s, const double *wGP, double Vol)
{
    const dgo_config *c = &s->c;
    const int n = s->n;
    double b = 0.;
    for (int e = 0; e < c->nElems; e++)
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            size_t d = i + n * (j + n * (k + (size_t)n * e));
            b = b + s->U[NV * d + MOM1] / s->U[NV * d + DENS] * (wGP[i] * wGP[j] * wGP[k]) / c->sJ[d];
        }
    return b / Vol;
}
void dgo_u_mortar(dgo *s, int nVar, double *Um, double *Us) { u_mortar_all(s, nVar, Um, Us); }
void dgo_flux_mortar(dgo *s, int nVar, double *Fm, const double *Fs, int weak) { flux_mortar_all(s, nVar, Fm, Fs, weak); }
size_t dgo_sizeof_config(void) { return sizeof(dgo_config); }
