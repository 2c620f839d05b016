"""CPU restatement (numpy) of the Taylor-Green vortex diagnostics, testcase/taylorgreenvortex/testcase.f90:283-515
(AnalyzeTestcase): the nTGVvars = 15 columns of the *_TGVAnalysis.csv file.

TEST INFRASTRUCTURE ONLY (parity oracle of dgx_analyze_tgv). Pinned by the reference's own
regressioncheck/checks/tgv/split/TGV_Re1600_Split_TGVAnalysis_Reference.csv (tests/test_oracle_goldens.py)."""
import numpy as np

COLUMNS = ("Dissipation Rate Incompressible", "Dissipation Rate Compressible", "Ekin incomp", "Ekin comp", "Enstrophy comp",
           "DR_u", "DR_S", "DR_Sd", "DR_p", "Maximum Vorticity", "Mean Temperature", "uprime", "Mean Entropy", "ED_S", "ED_D")


def _to_analyze(V, X):
    """ChangeBasis3D (changeBasis.t90): X[e,k,j,i,c] -> [e,K,J,I,c]."""
    Y = np.einsum("Ii,ekjic->ekjIc", V, X)
    Y = np.einsum("Jj,ekjIc->ekJIc", V, Y)
    return np.einsum("Kk,ekJIc->eKJIc", V, Y)


def analyze_tgv(case, U, gradUx, gradUy, gradUz, Vdm, wA, Vol, rho0=1.0, lift_vel=(1, 2, 3)):
    """U [e,k,j,i,5]; gradU* [e,k,j,i,nLift] with the velocity gradients at the indices lift_vel. Returns the 15 columns."""
    eos = case.eos
    kappa, R, mu0 = eos.kappa, eos.R, eos.mu0
    sJ = _to_analyze(Vdm, case.geo["sJ"][..., None])[..., 0]
    Ua = _to_analyze(Vdm, U)
    lv = list(lift_vel)
    G = np.stack([_to_analyze(Vdm, g[..., lv]) for g in (gradUx, gradUy, gradUz)], axis=-1)   # G[..., i, j] = d u_i / d x_j
    rho = Ua[..., 0]
    vel = Ua[..., 1:4] / rho[..., None]
    p = (kappa - 1.0) * (Ua[..., 4] - 0.5 * np.sum(Ua[..., 1:4] * vel, axis=-1))
    T = p / (rho * R)
    divU = G[..., 0, 0] + G[..., 1, 1] + G[..., 2, 2]
    S = 0.5 * (G + np.swapaxes(G, -1, -2))
    Sd = S - (1.0 / 3.0) * divU[..., None, None] * np.eye(3)
    vort = np.stack([G[..., 2, 1] - G[..., 1, 2], G[..., 0, 2] - G[..., 2, 0], G[..., 1, 0] - G[..., 0, 1]], axis=-1)
    w3 = wA[:, None, None] * wA[None, :, None] * wA[None, None, :]
    F = w3[None] / sJ
    if eos.visc_law == 0:
        mu = np.full_like(T, mu0)
    else:
        raise NotImplementedError("TGV analysis oracle: constant viscosity only")
    v2 = np.sum(vel * vel, axis=-1)
    w2 = np.sum(vort * vort, axis=-1)
    T_mean = np.sum(F * T) / Vol
    Entropy = np.sum(F * (-1.0 / (kappa - 1.0)) * rho * (np.log(p) - kappa * np.log(rho))) / Vol
    Ekin = np.sum(F * 0.5 * v2) / Vol
    Ekin_comp = np.sum(F * 0.5 * rho * v2) / Vol / rho0
    Enstr = np.sum(F * 0.5 * rho * w2) / (rho0 * Vol)
    DR_u = np.sum(F * np.sum(G * G, axis=(-1, -2))) * mu0 / Vol
    DR_S = np.sum(F * np.sum(S * S, axis=(-1, -2))) * 2.0 * mu0 / (rho0 * Vol)
    DR_Sd = np.sum(F * np.sum(Sd * Sd, axis=(-1, -2))) * 2.0 * mu0 / (rho0 * Vol)
    DR_p = -np.sum(F * p * divU) / (rho0 * Vol)
    ED_S = np.sum(F * mu * w2) / Vol
    ED_D = np.sum(F * mu * divU ** 2) * 4.0 / 3.0 / Vol
    vmax = float(np.sqrt(w2).max())
    uprime = np.sqrt(2.0 / 3.0 * Ekin)
    return np.array([DR_S, DR_Sd + DR_p, Ekin, Ekin_comp, Enstr, DR_u, DR_S, DR_Sd, DR_p, vmax, T_mean, uprime, Entropy, ED_S, ED_D])
