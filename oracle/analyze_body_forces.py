"""CPU restatement (numpy) of CalcBodyForces, equations/navierstokes/calcbodyforces.f90:41-205: pressure and friction force on
every wall boundary condition (BC types 3, 4, 9: analyze_equation.f90:113-121) from the face data of the last
DGTimeDerivative_weakForm (UPrim_master, gradU*_master), integrated with wGPSurf * SurfElem.

TEST INFRASTRUCTURE ONLY (parity oracle of dgx_calc_body_forces). Pinned by the reference's
tutorials/naca0012/NACA0012_Re5000_AoA8_BodyForces_BC_wall_Reference.csv evaluated on the reference's own NACA state at t=10
(tests/test_oracle_goldens.py)."""
import numpy as np

WALL_TYPES = (3, 4, 9)


def _viscosity(eos, T):
    """VISCOSITY_PRIM: constant or Sutherland (idealgas/viscosity.f90:105-125 with the EOS_Vars of eos.f90:150-203)."""
    if eos.visc_law == 0:
        return np.full_like(T, eos.mu0)
    _, _, _, mu0, Ts, Tref, expo, cS = eos.eos_vars()
    Tn = T * Tref
    return np.where(Tn >= Ts, mu0 * Tn ** expo * (1.0 + Ts) / (Tn + Ts), mu0 * Tn * cS)


def calc_body_forces(case, UPrim_master, gradUx_master, gradUy_master, gradUz_master, lift_vel=(1, 2, 3)):
    """Face arrays [side,q,p,var] (reference memory order). Returns Fp, Fv of shape (nBCs,3); BodyForce = Fp + Fv."""
    m, geo, eos = case.mesh, case.geo, case.eos
    w = case.basis.wGP
    wS = w[:, None] * w[None, :]                                # wGPSurf(i,j)
    nBCs = m.BoundaryType.shape[0]
    Fp = np.zeros((nBCs, 3))
    Fv = np.zeros((nBCs, 3))
    lv = list(lift_vel)
    for s in range(m.nBCSides):
        iBC = int(m.BC[s]) - 1
        if int(m.BoundaryType[iBC, 0]) not in WALL_TYPES:
            continue
        dA = wS * geo["SurfElem"][s]
        nv = geo["NormVec"][s]
        P = UPrim_master[s]
        Fp[iBC] += np.einsum("qp,qpd->d", P[..., 4] * dA, nv)
        if case.parabolic:
            T = P[..., 5]
            mu = _viscosity(eos, T)
            G = np.stack([gradUx_master[s][..., lv], gradUy_master[s][..., lv], gradUz_master[s][..., lv]], axis=-1)  # G[q,p,i,j] = d v_i / d x_j
            div = G[..., 0, 0] + G[..., 1, 1] + G[..., 2, 2]
            tau = mu[..., None, None] * (G + np.swapaxes(G, -1, -2))
            for d in range(3):
                tau[..., d, d] -= 2.0 / 3.0 * mu * div
            Fv[iBC] -= np.einsum("qpij,qpj,qp->i", tau, nv, dA)     # calcbodyforces.f90:202 "force acting on the wall"
    return Fp, Fv


def calc_wall_velocity(case, UPrim_master, Surf):
    """CalcWallVelocity (analyze_equation.f90:435-499): maxV, minV, meanV per boundary condition."""
    m, geo = case.mesh, case.geo
    w = case.basis.wGP
    wS = w[:, None] * w[None, :]
    nBCs = m.BoundaryType.shape[0]
    maxV, minV, meanV = np.full(nBCs, -1.e14), np.full(nBCs, 1.e14), np.zeros(nBCs)
    for s in range(m.nBCSides):
        iBC = int(m.BC[s]) - 1
        if int(m.BoundaryType[iBC, 0]) not in WALL_TYPES:
            continue
        v = np.sqrt(np.sum(UPrim_master[s][..., 1:4] ** 2, axis=-1))
        maxV[iBC] = max(maxV[iBC], v.max())
        minV[iBC] = min(minV[iBC], v.min())
        meanV[iBC] += np.sum(v * wS * geo["SurfElem"][s])
    for b in range(nBCs):
        if int(m.BoundaryType[b, 0]) in WALL_TYPES:
            meanV[b] /= Surf[b]
    return maxV, minV, meanV
