// Point-wise physics of the compressible Navier-Stokes DG operator (device functions).
//
// Hand-written for sm_100a FP64 CUDA cores. Formulas follow the reference (paths relative to
// /root/reference/src/equations/navierstokes): idealgas/eos.f90:212-239 (ConsToPrim), :467-489 (PrimToCons),
// :605-628 (PRESSURE_RIEMANN), idealgas/viscosity.f90 (muSuth), eos.h:110 (thermal conductivity),
// riemann.f90:216-289 (rotation), :707 LF, :738 HLLC, :809 Roe, :885 RoeEntropyFix, :994 RoeL2, :1075 HLL, :1126 HLLE,
// :1175 HLLEM, :1239 FluxAverage; splitflux.f90 volume/surface: :145/315 SD, :543/627 MO, :346/407 DU, :437/506 KG,
// :669/735 PI; flux.f90:619-697/700-745/779-822 viscous fluxes,
// idealgas/getboundaryflux.f90:262-487 boundary states (types 2,3,4,9,91,23,24,25,27).
#pragma once
#include <math.h>

namespace dgx {

enum { DENS = 0, MOM1, MOM2, MOM3, ENER };
enum { VEL1 = 1, VEL2, VEL3, PRES, TEMP };
// lifting variables (PP_OPTLIFT=1): u,v,w,T
enum { LV1 = 0, LV2, LV3, LT, NLIFT };

struct Eos {
    double kappa, R, Pr, mu0, Ts, Tref, ExpoSuth, cSuth;
    int viscLaw;
};

__device__ __forceinline__ void cons_to_prim(double* __restrict__ P, const double* __restrict__ U, const Eos& e) {
    const double sRho = 1.0 / U[DENS];
    P[DENS] = U[DENS];
    P[VEL1] = U[MOM1] * sRho;
    P[VEL2] = U[MOM2] * sRho;
    P[VEL3] = U[MOM3] * sRho;
    P[PRES] = (e.kappa - 1.0) * (U[ENER] - 0.5 * (U[MOM1] * P[VEL1] + U[MOM2] * P[VEL2] + U[MOM3] * P[VEL3]));
    P[TEMP] = P[PRES] * sRho / e.R;
}

__device__ __forceinline__ void prim_to_cons(const double* __restrict__ P, double* __restrict__ U, const Eos& e) {
    U[DENS] = P[DENS];
    U[MOM1] = P[VEL1] * P[DENS];
    U[MOM2] = P[VEL2] * P[DENS];
    U[MOM3] = P[VEL3] * P[DENS];
    U[ENER] = P[PRES] / (e.kappa - 1.0) + 0.5 * (U[MOM1] * P[VEL1] + U[MOM2] * P[VEL2] + U[MOM3] * P[VEL3]);
}

__device__ __forceinline__ double viscosity(const Eos& e, double T) {
    if (e.viscLaw == 0) return e.mu0;
    const double Tn = T * e.Tref;
    if (Tn >= e.Ts) return e.mu0 * pow(Tn, e.ExpoSuth) * (1.0 + e.Ts) / (Tn + e.Ts);
    return e.mu0 * Tn * e.cSuth;
}
__device__ __forceinline__ double conductivity(const Eos& e, double mu) { return mu * e.R * e.kappa / ((e.kappa - 1.0) * e.Pr); }

// ---------------------------------------------------------------------------------------------------------
// split fluxes
// node record used by the two-point volume fluxes: rho,u,v,w,p,X with X = H (PI), e=E/rho (KG), rhoE (SD, MO, DU)
__device__ __forceinline__ double split_sixth(int variant, const double* U, const double* P) {
    if (variant == 4) return (U[ENER] + P[PRES]) / U[DENS];
    if (variant == 3) return U[ENER] / U[DENS];
    return U[ENER];
}

// Flux = 1/2(MRef+M) . (f,g,h)#  -- metric sum passed as Ms = MRef+M (factor 1/2 applied here)
__device__ __forceinline__ void split_volume_flux(int variant, const double* a, const double* b, const double* Ms, double* F) {
    if (variant == 0) {  // SD: {rho u}, {rho u u}+{p}, {(rhoE+p)u}  (sums, factor 2 in D matrix)
        const double m1a = a[0] * a[1], m2a = a[0] * a[2], m3a = a[0] * a[3];
        const double m1b = b[0] * b[1], m2b = b[0] * b[2], m3b = b[0] * b[3];
        const double Epa = a[5] + a[4], Epb = b[5] + b[4];
        const double ps = a[4] + b[4];
        const double h1 = 0.5 * Ms[0], h2 = 0.5 * Ms[1], h3 = 0.5 * Ms[2];
        F[DENS] = h1 * (m1a + m1b) + h3 * (m3a + m3b) + h2 * (m2a + m2b);
        F[MOM1] = h1 * (m1a * a[1] + m1b * b[1] + ps) + h3 * (m1a * a[3] + m1b * b[3]) + h2 * (m1a * a[2] + m1b * b[2]);
        F[MOM2] = h1 * (m1a * a[2] + m1b * b[2]) + h3 * (m2a * a[3] + m2b * b[3]) + h2 * (m2a * a[2] + m2b * b[2] + ps);
        F[MOM3] = h1 * (m1a * a[3] + m1b * b[3]) + h3 * (m3a * a[3] + m3b * b[3] + ps) + h2 * (m2a * a[3] + m2b * b[3]);
        F[ENER] = h1 * (Epa * a[1] + Epb * b[1]) + h3 * (Epa * a[3] + Epb * b[3]) + h2 * (Epa * a[2] + Epb * b[2]);
        return;
    }
    if (variant == 1) {  // MO (splitflux.f90:543-622): {rho u}, {rho u}{u}+{p}, advective energy; contracted with 1/2 Ms first
        const double h1 = 0.5 * Ms[0], h2 = 0.5 * Ms[1], h3 = 0.5 * Ms[2];
        const double una = h1 * a[1] + h2 * a[2] + h3 * a[3], unb = h1 * b[1] + h2 * b[2] + h3 * b[3];  // contravariant velocities
        const double mna = a[0] * una, mnb = b[0] * unb;
        const double qa = a[1] * a[1] + a[2] * a[2] + a[3] * a[3], qb = b[1] * b[1] + b[2] * b[2] + b[3] * b[3];
        const double rhoepa = a[5] - 0.5 * a[0] * qa + a[4], rhoepb = b[5] - 0.5 * b[0] * qb + b[4];
        const double ms = mna + mnb, ps = a[4] + b[4];
        const double us = a[1] + b[1], vs = a[2] + b[2], ws = a[3] + b[3];
        F[DENS] = ms;
        F[MOM1] = 0.5 * ms * us + h1 * ps;
        F[MOM2] = 0.5 * ms * vs + h2 * ps;
        F[MOM3] = 0.5 * ms * ws + h3 * ps;
        F[ENER] = (rhoepa * una + rhoepb * unb) +
                  0.5 * ((mna * a[1] + mnb * b[1]) * us + (mna * a[2] + mnb * b[2]) * vs + (mna * a[3] + mnb * b[3]) * ws) -
                  0.5 * (mna * qa + mnb * qb);
        return;
    }
    if (variant == 2) {  // DU (splitflux.f90:346-402): {rho}{u}, {rho u}{u}+{p}, ({rho E}+{p}){u}
        const double h1 = 0.5 * Ms[0], h2 = 0.5 * Ms[1], h3 = 0.5 * Ms[2];
        const double vn = 0.5 * (h1 * (a[1] + b[1]) + h2 * (a[2] + b[2]) + h3 * (a[3] + b[3]));  // 1/2 of the contravariant velocity sum
        const double ps = a[4] + b[4];
        F[DENS] = (a[0] + b[0]) * vn;
        F[MOM1] = (a[0] * a[1] + b[0] * b[1]) * vn + h1 * ps;
        F[MOM2] = (a[0] * a[2] + b[0] * b[2]) * vn + h2 * ps;
        F[MOM3] = (a[0] * a[3] + b[0] * b[3]) * vn + h3 * ps;
        F[ENER] = (a[5] + b[5] + ps) * vn;
        return;
    }
    const double rs = a[0] + b[0];
    const double us = a[1] + b[1], vs = a[2] + b[2], ws = a[3] + b[3];
    const double ps = a[4] + b[4];
    const double xs = a[5] + b[5];
    // contravariant velocity sum: (1/2 Ms).(us,vs,ws); common factor 1/4 rho_s
    const double vn = 0.5 * (Ms[0] * us + Ms[1] * vs + Ms[2] * ws);
    const double q = 0.25 * rs * vn;   // 1/4 {rho}_s * contravariant sum
    F[DENS] = 2.0 * q;                 // 1/2 rs us M
    F[MOM1] = q * us + 0.5 * Ms[0] * ps;
    F[MOM2] = q * vs + 0.5 * Ms[1] * ps;
    F[MOM3] = q * ws + 0.5 * Ms[2] * ps;
    if (variant == 3) F[ENER] = q * xs + 0.5 * ps * vn;   // KG: {rho}{e}{u}+{p}{u}
    else F[ENER] = q * xs;                                 // PI: {rho}{H}{u}
}

// extended left/right state for the 1D (rotated) Riemann problem
struct Ext {
    double rho, sRho, ener, pres, v1, v2, v3, m1, m2, m3;
};

__device__ __forceinline__ void split_surface_flux(int variant, const Ext& L, const Ext& R, double* F) {
    if (variant == 1) {  // MO
        const double qL = L.v1 * L.v1 + L.v2 * L.v2 + L.v3 * L.v3, qR = R.v1 * R.v1 + R.v2 * R.v2 + R.v3 * R.v3;
        const double rhoepL = L.ener - 0.5 * L.rho * qL + L.pres, rhoepR = R.ener - 0.5 * R.rho * qR + R.pres;
        const double ms = L.m1 + R.m1;
        F[DENS] = 0.5 * ms;
        F[MOM1] = 0.25 * ms * (L.v1 + R.v1) + 0.5 * (L.pres + R.pres);
        F[MOM2] = 0.25 * ms * (L.v2 + R.v2);
        F[MOM3] = 0.25 * ms * (L.v3 + R.v3);
        F[ENER] = 0.5 * (rhoepL * L.v1 + rhoepR * R.v1) + 0.25 * (L.m1 * L.v1 + R.m1 * R.v1) * (L.v1 + R.v1) +
                  0.25 * (L.m1 * L.v2 + R.m1 * R.v2) * (L.v2 + R.v2) + 0.25 * (L.m1 * L.v3 + R.m1 * R.v3) * (L.v3 + R.v3) -
                  0.25 * (L.m1 * L.v1 * L.v1 + R.m1 * R.v1 * R.v1) - 0.25 * (L.m1 * L.v2 * L.v2 + R.m1 * R.v2 * R.v2) -
                  0.25 * (L.m1 * L.v3 * L.v3 + R.m1 * R.v3 * R.v3);
        return;
    }
    if (variant == 2) {  // DU
        const double us = L.v1 + R.v1;
        F[DENS] = 0.25 * (L.rho + R.rho) * us;
        F[MOM1] = 0.25 * (L.m1 + R.m1) * us + 0.5 * (L.pres + R.pres);
        F[MOM2] = 0.25 * (L.m2 + R.m2) * us;
        F[MOM3] = 0.25 * (L.m3 + R.m3) * us;
        F[ENER] = 0.25 * (L.ener + R.ener + L.pres + R.pres) * us;
        return;
    }
    if (variant == 0) {
        F[DENS] = 0.5 * (L.m1 + R.m1);
        F[MOM1] = 0.5 * (L.m1 * L.v1 + L.pres + R.m1 * R.v1 + R.pres);
        F[MOM2] = 0.5 * (L.m1 * L.v2 + R.m1 * R.v2);
        F[MOM3] = 0.5 * (L.m1 * L.v3 + R.m1 * R.v3);
        F[ENER] = 0.5 * ((L.ener + L.pres) * L.v1 + (R.ener + R.pres) * R.v1);
        return;
    }
    const double rs = L.rho + R.rho, us = L.v1 + R.v1;
    F[DENS] = 0.25 * rs * us;
    F[MOM1] = 0.125 * rs * (us * us) + 0.5 * (L.pres + R.pres);
    F[MOM2] = 0.125 * rs * us * (L.v2 + R.v2);
    F[MOM3] = 0.125 * rs * us * (L.v3 + R.v3);
    if (variant == 3) {
        const double eL = L.ener / L.rho, eR = R.ener / R.rho;
        F[ENER] = 0.125 * rs * (eL + eR) * us + 0.25 * (L.pres + R.pres) * us;
    } else {
        const double HL = (L.ener + L.pres) / L.rho, HR = (R.ener + R.pres) / R.rho;
        F[ENER] = 0.125 * rs * (HL + HR) * us;
    }
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void make_ext(Ext& X, const double* U, const double* P, const double* nv, const double* t1, const double* t2) {
    X.rho = U[DENS];
    X.sRho = 1.0 / X.rho;
    X.ener = U[ENER];
    X.pres = P[PRES];
    X.v1 = P[VEL1] * nv[0] + P[VEL2] * nv[1] + P[VEL3] * nv[2];
    X.v2 = P[VEL1] * t1[0] + P[VEL2] * t1[1] + P[VEL3] * t1[2];
    X.v3 = P[VEL1] * t2[0] + P[VEL2] * t2[1] + P[VEL3] * t2[2];
    X.m1 = X.rho * X.v1;
    X.m2 = X.rho * X.v2;
    X.m3 = X.rho * X.v3;
}

__device__ __forceinline__ void euler_flux_1d(const Ext& X, double* F) {
    F[DENS] = X.m1;
    F[MOM1] = X.m1 * X.v1 + X.pres;
    F[MOM2] = X.m1 * X.v2;
    F[MOM3] = X.m1 * X.v3;
    F[ENER] = (X.ener + X.pres) * X.v1;
}

// Riemann solver in the face frame. split>=0: central part is the split surface flux (riemann.f90:868-877).
__device__ __forceinline__ void riemann_solver(int riem, int split, double kappa, const Ext& L, const Ext& R, double* F) {
    const double UL[5] = {L.rho, L.m1, L.m2, L.m3, L.ener};
    const double UR[5] = {R.rho, R.m1, R.m2, R.m3, R.ener};
    double FL[5], FR[5];
    if (split < 0) { euler_flux_1d(L, FL); euler_flux_1d(R, FR); }
    if (riem == 0) {  // LF
        const double cL = sqrt(kappa * L.pres * L.sRho), cR = sqrt(kappa * R.pres * R.sRho);
        const double lam = fmax(fabs(R.v1), fabs(L.v1)) + fmax(cL, cR);
        if (split < 0) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = 0.5 * ((FL[v] + FR[v]) - lam * (UR[v] - UL[v]));
        } else {
            split_surface_flux(split, L, R, F);
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = F[v] - 0.5 * lam * (UR[v] - UL[v]);
        }
        return;
    }
    if (riem == 9) {  // Riemann_FluxAverage (SPLIT_DG only; dgx_create enforces it)
        split_surface_flux(split, L, R, F);
        return;
    }
    // Roe averages
    const double HL = (L.ener + L.pres) * L.sRho, HR = (R.ener + R.pres) * R.sRho;
    const double sl = sqrt(L.rho), sr = sqrt(R.rho);
    const double ss = 1.0 / (sl + sr);
    const double rv1 = (sr * R.v1 + sl * L.v1) * ss, rv2 = (sr * R.v2 + sl * L.v2) * ss, rv3 = (sr * R.v3 + sl * L.v3) * ss;
    const double RoeH = (sr * HR + sl * HL) * ss;
    const double absVel = rv1 * rv1 + rv2 * rv2 + rv3 * rv3;
    const double Roec = sqrt((kappa - 1.0) * (RoeH - 0.5 * absVel));
    if (riem == 4 || riem == 6 || riem == 7) {  // HLL, HLLE, HLLEM (non-split builds only)
        double Ssl, Ssr;
        if (riem == 4) { Ssl = rv1 - Roec; Ssr = rv1 + Roec; }
        else {
            const double beta = sqrt(0.5 * (kappa - 1.0) / kappa);
            const double cL = sqrt(kappa * L.pres * L.sRho), cR = sqrt(kappa * R.pres * R.sRho);
            Ssl = fmin(fmin(rv1 - Roec, L.v1 - beta * cL), 0.0);
            Ssr = fmax(fmax(rv1 + Roec, R.v1 + beta * cR), 0.0);
        }
        if (Ssl >= 0.0) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = FL[v];
        } else if (Ssr <= 0.0) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = FR[v];
        } else if (riem != 7) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = (Ssr * FL[v] - Ssl * FR[v] + Ssl * Ssr * (UR[v] - UL[v])) / (Ssr - Ssl);
        } else {
            const double RoeDens = sqrt(L.rho * R.rho);
            const double delta = Roec / (Roec + fabs(0.5 * (Ssl + Ssr)));
            const double A2 = (R.rho - L.rho) - (R.pres - L.pres) / (Roec * Roec);
            const double A3 = RoeDens * (R.v2 - L.v2), A4 = RoeDens * (R.v3 - L.v3);
            const double q2[5] = {1.0, rv1, rv2, rv3, 0.5 * absVel};
            const double q3[5] = {0.0, 0.0, 1.0, 0.0, rv2};
            const double q4[5] = {0.0, 0.0, 0.0, 1.0, rv3};
#pragma unroll
            for (int v = 0; v < 5; v++)
                F[v] = (Ssr * FL[v] - Ssl * FR[v] + Ssl * Ssr * (UR[v] - UL[v] - delta * (q2[v] * A2 + q3[v] * A3 + q4[v] * A4))) / (Ssr - Ssl);
        }
        return;
    }
    if (riem == 5) {  // HLLC
        const double Ssl = rv1 - Roec, Ssr = rv1 + Roec;
        if (Ssl >= 0.0) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = FL[v];
        } else if (Ssr <= 0.0) {
#pragma unroll
            for (int v = 0; v < 5; v++) F[v] = FR[v];
        } else {
            const double sMuL = Ssl - L.v1, sMuR = Ssr - R.v1;
            const double SStar = (R.pres - L.pres + L.m1 * sMuL - R.m1 * sMuR) / (L.rho * sMuL - R.rho * sMuR);
            if (Ssl <= 0.0 && SStar >= 0.0) {
                const double EStar = L.ener * L.sRho + (SStar - L.v1) * (SStar + L.pres * L.sRho / sMuL);
                const double fac = L.rho * sMuL / (Ssl - SStar);
                const double Us[5] = {fac, fac * SStar, fac * L.v2, fac * L.v3, fac * EStar};
#pragma unroll
                for (int v = 0; v < 5; v++) F[v] = FL[v] + Ssl * (Us[v] - UL[v]);
            } else {
                const double EStar = R.ener * R.sRho + (SStar - R.v1) * (SStar + R.pres * R.sRho / sMuR);
                const double fac = R.rho * sMuR / (Ssr - SStar);
                const double Us[5] = {fac, fac * SStar, fac * R.v2, fac * R.v3, fac * EStar};
#pragma unroll
                for (int v = 0; v < 5; v++) F[v] = FR[v] + Ssr * (Us[v] - UR[v]);
            }
        }
        return;
    }
    double a[5] = {rv1 - Roec, rv1, rv1, rv1, rv1 + Roec};
    const double r1[5] = {1.0, a[0], rv2, rv3, RoeH - rv1 * Roec};
    const double r2[5] = {1.0, rv1, rv2, rv3, 0.5 * absVel};
    const double r3[5] = {0.0, 0.0, 1.0, 0.0, rv2};
    const double r4[5] = {0.0, 0.0, 0.0, 1.0, rv3};
    const double r5[5] = {1.0, a[4], rv2, rv3, RoeH + rv1 * Roec};
    double Al[5];
    if (riem == 1 || riem == 2) {  // Roe, RoeL2
        double dU[6];
#pragma unroll
        for (int v = 0; v < 5; v++) dU[v] = UR[v] - UL[v];
        dU[5] = dU[4] - (dU[2] - rv2 * dU[0]) * rv2 - (dU[3] - rv3 * dU[0]) * rv3;
        if (riem == 2) {  // low Mach number fix (riemann.f90:1044-1046)
            const double Ma = sqrt(absVel) / (Roec * sqrt(kappa));
            dU[1] *= Ma; dU[2] *= Ma; dU[3] *= Ma;
        }
        Al[2] = dU[2] - rv2 * dU[0];
        Al[3] = dU[3] - rv3 * dU[0];
        Al[1] = (kappa - 1.0) / (Roec * Roec) * (dU[0] * (RoeH - rv1 * rv1) - dU[5] + rv1 * dU[1]);
        Al[0] = 0.5 / Roec * (dU[0] * (rv1 + Roec) - dU[1] - Roec * Al[1]);
        Al[4] = dU[0] - Al[0] - Al[1];
#pragma unroll
        for (int i = 0; i < 5; i++) a[i] = fabs(a[i]);
    } else {  // RoeEntropyFix
        const double cL = sqrt(kappa * L.pres * L.sRho), cR = sqrt(kappa * R.pres * R.sRho);
        const double RoeDens = sqrt(L.rho * R.rho);
        const double d1 = R.rho - L.rho, d2 = R.v1 - L.v1, d3 = R.v2 - L.v2, d4 = R.v3 - L.v3, d5 = R.pres - L.pres;
        const double tmp = 0.5 / (Roec * Roec);
        Al[0] = tmp * (d5 - RoeDens * Roec * d2);
        Al[1] = d1 - d5 * 2.0 * tmp;
        Al[2] = RoeDens * d3;
        Al[3] = RoeDens * d4;
        Al[4] = tmp * (d5 + RoeDens * Roec * d2);
        const double al[5] = {L.v1 - cL, L.v1, L.v1, L.v1, L.v1 + cL};
        const double ar[5] = {R.v1 - cR, R.v1, R.v1, R.v1, R.v1 + cR};
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const double da = fmax(0.0, fmax(a[i] - al[i], ar[i] - a[i]));
            if (fabs(a[i]) < da) a[i] = 0.5 * (a[i] * a[i] / da + da);
            else a[i] = fabs(a[i]);
        }
    }
    if (split < 0) {
#pragma unroll
        for (int v = 0; v < 5; v++)
            F[v] = 0.5 * ((FL[v] + FR[v]) - Al[0] * a[0] * r1[v] - Al[1] * a[1] * r2[v] - Al[2] * a[2] * r3[v] - Al[3] * a[3] * r4[v] -
                          Al[4] * a[4] * r5[v]);
    } else {
        split_surface_flux(split, L, R, F);
#pragma unroll
        for (int v = 0; v < 5; v++)
            F[v] = F[v] - 0.5 * (Al[0] * a[0] * r1[v] + Al[1] * a[1] * r2[v] + Al[2] * a[2] * r3[v] + Al[3] * a[3] * r4[v] +
                                 Al[4] * a[4] * r5[v]);
    }
}

// full advective numerical flux in global coordinates (riemann.f90:216-289)
__device__ __forceinline__ void riemann(int riem, int split, double kappa, double* Fout, const double* UL, const double* UR,
                                        const double* PL, const double* PR, const double* nv, const double* t1, const double* t2) {
    Ext L, R;
    make_ext(L, UL, PL, nv, t1, t2);
    make_ext(R, UR, PR, nv, t1, t2);
    double F[5];
    riemann_solver(riem, split, kappa, L, R, F);
    Fout[DENS] = F[DENS];
#pragma unroll
    for (int d = 0; d < 3; d++) Fout[MOM1 + d] = nv[d] * F[MOM1] + t1[d] * F[MOM2] + t2[d] * F[MOM3];
    Fout[ENER] = F[ENER];
}

// ---------------------------------------------------------------------------------------------------------
// viscous stress from lifted gradients g[d*4+v] (d = x,y,z; v = u,v,w,T)
struct Tau {
    double xx, yy, zz, xy, xz, yz, qx, qy, qz;  // q* = tau.v + lambda dT/dx*
};
__device__ __forceinline__ void stress(Tau& t, const double* P, const double* g, double mu, double lambda) {
    const double s23 = 2.0 / 3.0, s43 = 4.0 / 3.0;
    const double ux = g[0 * 4 + LV1], vx = g[0 * 4 + LV2], wx = g[0 * 4 + LV3], Tx = g[0 * 4 + LT];
    const double uy = g[1 * 4 + LV1], vy = g[1 * 4 + LV2], wy = g[1 * 4 + LV3], Ty = g[1 * 4 + LT];
    const double uz = g[2 * 4 + LV1], vz = g[2 * 4 + LV2], wz = g[2 * 4 + LV3], Tz = g[2 * 4 + LT];
    t.xx = mu * (s43 * ux - s23 * vy - s23 * wz);
    t.yy = mu * (-s23 * ux + s43 * vy - s23 * wz);
    t.zz = mu * (-s23 * ux - s23 * vy + s43 * wz);
    t.xy = mu * (uy + vx);
    t.xz = mu * (uz + wx);
    t.yz = mu * (vz + wy);
    t.qx = t.xx * P[VEL1] + t.xy * P[VEL2] + t.xz * P[VEL3] + lambda * Tx;
    t.qy = t.xy * P[VEL1] + t.yy * P[VEL2] + t.yz * P[VEL3] + lambda * Ty;
    t.qz = t.xz * P[VEL1] + t.yz * P[VEL2] + t.zz * P[VEL3] + lambda * Tz;
}
// viscous flux contracted with a (metric or normal) vector M: F[1..4] (DENS component is zero)
__device__ __forceinline__ void visc_flux_dir(const Tau& t, const double* M, double* F4) {
    F4[0] = -M[0] * t.xx - M[1] * t.xy - M[2] * t.xz;
    F4[1] = -M[0] * t.xy - M[1] * t.yy - M[2] * t.yz;
    F4[2] = -M[0] * t.xz - M[1] * t.yz - M[2] * t.zz;
    F4[3] = -M[0] * t.qx - M[1] * t.qy - M[2] * t.qz;
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double pressure_riemann(const double* P, double kappa) {
    if (P[VEL1] <= 0.0) {
        const double kf = 2.0 * kappa / (kappa - 1.0);
        return P[PRES] * pow(fmax(0.0001, (1.0 + 0.5 * (kappa - 1.0) * P[VEL1] / sqrt(kappa * P[PRES] / P[DENS]))), kf);
    }
    const double ar = 2.0 / ((kappa + 1.0) * P[DENS]);
    const double br = (kappa - 1.0) / (kappa + 1.0) * P[PRES];
    return P[PRES] + P[VEL1] / ar * 0.5 * (P[VEL1] + sqrt(P[VEL1] * P[VEL1] + 4.0 * ar * (P[PRES] + br)));
}

__device__ __forceinline__ bool is_riemann_bc(int t) { return t == 2 || t == 23 || t == 24 || t == 25 || t == 27; }
__device__ __forceinline__ bool is_wall_bc(int t) { return t == 3 || t == 4 || t == 9 || t == 91; }

// returns 0 if the BC type is supported
__device__ __forceinline__ int boundary_state(int bct, const Eos& e, double* out, const double* Pm, const double* Ref, const double* nv,
                                              const double* t1, const double* t2) {
    if (bct == 2) {
#pragma unroll
        for (int v = 0; v < 6; v++) out[v] = Ref[v];
        return 0;
    }
    if (!is_wall_bc(bct) && !is_riemann_bc(bct)) return 1;
    const double kappa = e.kappa, R = e.R;
    double b[6];
    b[DENS] = Pm[DENS];
    b[VEL1] = Pm[VEL1] * nv[0] + Pm[VEL2] * nv[1] + Pm[VEL3] * nv[2];
    b[VEL2] = Pm[VEL1] * t1[0] + Pm[VEL2] * t1[1] + Pm[VEL3] * t1[2];
    b[VEL3] = Pm[VEL1] * t2[0] + Pm[VEL2] * t2[1] + Pm[VEL3] * t2[2];
    b[PRES] = Pm[PRES];
    b[TEMP] = Pm[TEMP];
    if (is_wall_bc(bct)) {
        b[PRES] = pressure_riemann(b, kappa);
        if (bct == 3) {
            b[VEL1] = b[VEL2] = b[VEL3] = 0.0;
            b[TEMP] = Pm[TEMP];
            b[DENS] = b[PRES] / (b[TEMP] * R);
        } else if (bct == 4) {
            b[VEL1] = b[VEL2] = b[VEL3] = 0.0;
            b[TEMP] = Ref[TEMP];
            b[DENS] = b[PRES] / (b[TEMP] * R);
        } else {  // 9, 91
            b[VEL1] = 0.0;
            b[DENS] = Pm[DENS];
            b[TEMP] = b[PRES] / (b[DENS] * R);
        }
    } else if (bct == 27) {  // subsonic inflow; Ref = (Tt, a1, a2, a3, pt) with the unit direction a set up by the host
        const double Tt = Ref[0], pt = Ref[4];
        const double an = Ref[1] * nv[0] + Ref[2] * nv[1] + Ref[3] * nv[2];
        const double A = -1.0 * an;
        const double c = sqrt(kappa * b[PRES] / b[DENS]);
        const double Rplus = -b[VEL1] - 2.0 * c / (kappa - 1.0);
        const double tmp1 = A * A + 2.0 / (kappa - 1.0);
        const double tmp2 = 2.0 * Rplus;
        const double tmp3 = (kappa - 1.0) / 2.0 * (Rplus * Rplus) - kappa * R * Tt * (A * A);
        const double disc = sqrt(tmp2 * tmp2 - 4.0 * tmp1 * tmp3);
        const double cb = fmax((-tmp2 + disc) / (2.0 * tmp1), (-tmp2 - disc) / (2.0 * tmp1));
        const double Tb = cb * cb / (kappa * R);
        const double Ma = sqrt(2.0 / (kappa - 1.0) * (Tt / Tb - 1.0));
        const double pb = pt * pow(1.0 + 0.5 * (kappa - 1.0) * (Ma * Ma), -kappa / (kappa - 1.0));
        const double Um = Ma * sqrt(kappa * R * Tb);
        b[DENS] = pb / (R * Tb);
        b[VEL1] = Um * an;
        b[VEL2] = Um * (Ref[1] * t1[0] + Ref[2] * t1[1] + Ref[3] * t1[2]);
        b[VEL3] = Um * (Ref[1] * t2[0] + Ref[2] * t2[1] + Ref[3] * t2[2]);
        b[PRES] = pb;
        b[TEMP] = Tb;
    } else {  // 23, 24, 25: outflows (Carlson, NASA/TM-2011-217181)
        const double c = sqrt(kappa * b[PRES] / b[DENS]);
        const double Ma = b[VEL1] / c;
        const double ptot = b[PRES] + 0.5 * b[DENS] * (b[VEL1] * b[VEL1] + b[VEL2] * b[VEL2] + b[VEL3] * b[VEL3]);
        if (bct == 23) {
            const double MaOut = Ref[1];
            double pb;
            if (Ma < 1.0) {
                const double pt = b[PRES] * pow(1.0 + 0.5 * (kappa - 1.0) * Ma * Ma, kappa / (kappa - 1.0));
                pb = pt * pow(1.0 + 0.5 * (kappa - 1.0) * MaOut * MaOut, -kappa / (kappa - 1.0));
            } else pb = ptot;
            b[DENS] = kappa * pb / (c * c);
            b[PRES] = pb;
            b[TEMP] = b[PRES] / (R * b[DENS]);
        } else if (bct == 24) {
            if (Ma < 1.0) {
                const double pb = Ref[4];
                b[DENS] = kappa * pb / (c * c);
                b[PRES] = pb;
                b[TEMP] = b[PRES] / (R * b[DENS]);
            }
        } else {
            const double pb = (Ma < 1.0) ? Ref[4] : ptot;
            if (b[VEL1] < 0.0) { b[VEL1] = fabs(b[VEL1]); b[VEL2] = 0.0; b[VEL3] = 0.0; }
            b[DENS] = kappa * pb / (c * c);
            b[PRES] = Ref[4];
            b[TEMP] = b[PRES] / (R * b[DENS]);
        }
    }
    out[DENS] = b[DENS];
#pragma unroll
    for (int d = 0; d < 3; d++) out[VEL1 + d] = b[VEL1] * nv[d] + b[VEL2] * t1[d] + b[VEL3] * t2[d];
    out[PRES] = b[PRES];
    out[TEMP] = b[TEMP];
    return 0;
}

// slip wall, version 2 (BC 91, getboundaryflux.f90:716-781): wall-normal derivative of the tangential velocities and
// wall-tangential derivatives of the normal velocity set to zero; temperature gradient loses its normal component.
__device__ __forceinline__ void slip_wall_gradients_91(const double* gm, const double* nv, const double* t1, const double* t2, double* gf) {
    const double* tv[3] = {nv, t1, t2};
    const double B[3][3] = {{1.0 - nv[0] * nv[0], -nv[0] * nv[1], -nv[0] * nv[2]},
                            {-nv[0] * nv[1], 1.0 - nv[1] * nv[1], -nv[2] * nv[1]},
                            {-nv[0] * nv[2], -nv[2] * nv[1], 1.0 - nv[2] * nv[2]}};
#pragma unroll
    for (int d = 0; d < 3; d++) gf[d * 4 + LT] = B[d][0] * gm[0 * 4 + LT] + B[d][1] * gm[1 * 4 + LT] + B[d][2] * gm[2 * 4 + LT];
    double gw[3][3], ga[3][3];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int w = 0; w < 3; w++) gw[d][w] = tv[w][0] * gm[d * 4 + LV1] + tv[w][1] * gm[d * 4 + LV2] + tv[w][2] * gm[d * 4 + LV3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int w = 0; w < 3; w++) ga[a][w] = tv[a][0] * gw[0][w] + tv[a][1] * gw[1][w] + tv[a][2] * gw[2][w];
    ga[0][1] = 0.0; ga[0][2] = 0.0; ga[1][0] = 0.0; ga[2][0] = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int w = 0; w < 3; w++) gw[d][w] = nv[d] * ga[0][w] + t1[d] * ga[1][w] + t2[d] * ga[2][w];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int x = 0; x < 3; x++) gf[d * 4 + LV1 + x] = nv[x] * gw[d][0] + t1[x] * gw[d][1] + t2[x] * gw[d][2];
}

}  // namespace dgx
