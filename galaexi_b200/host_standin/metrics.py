"""Volume and surface metrics of curved hexahedra (host, init time).

Mirror of /root/reference/src/mesh/metrics.f90:
  * BuildCoords        :40-105   Elem_xGP from NodeCoords (equidistant NGeo -> CL N -> solution nodes)
  * CalcMetrics        :171-552  dX/dxi on CL points, Jacobian via NGeoRef=3*NGeo projection, metrics
                                 in cross-product or curl form, interpolation CL -> solution nodes
  * CalcSurfMetrics    :611-741  face metrics of master sides, ordered with SideToVol2(flip=0)
  * SurfMetricsFromJa  :755-791  NormVec, TangVec1, TangVec2, SurfElem

Array convention: numpy C order with the reversed Fortran index list, i.e. Fortran
``Metrics_fTilde(3,0:N,0:N,0:N,nElems)`` is ``Metrics_fTilde[e,k,j,i,0:3]`` and ``NormVec(3,0:N,0:N,nSides)``
is ``NormVec[side,q,p,0:3]``: the raw buffers are what the Fortran host would pass through the C ABI.

The tensor-product basis changes are evaluated with BLAS-backed contractions, i.e. a different summation
order than the reference's scalar loops; the results agree to a few ulp. They are *inputs* of the hot path
(both the CUDA kernels and the CPU oracle receive the same arrays).
"""
from __future__ import annotations

import numpy as np

from . import basis as bs
from . import mappings as mp
from . import mortar as mo
from .mesh import Mesh

NormalDirs = (3, 2, 1, 2, 1, 3)       # mesh_vars.f90:112
TangDirs = (1, 3, 2, 3, 2, 1)         # mesh_vars.f90:114 (3D)
NormalSigns = (-1.0, -1.0, 1.0, 1.0, -1.0, 1.0)  # mesh_vars.f90:118


def change_basis_volume(V: np.ndarray, X: np.ndarray) -> np.ndarray:
    """X[e,k,j,i,c] -> Y[e,K,J,I,c] = sum V[I,i] V[J,j] V[K,k] X (ChangeBasis3D, changeBasis.t90)."""
    Y = np.einsum("Ii,ekjic->ekjIc", V, X, optimize=True)
    Y = np.einsum("Jj,ekjIc->ekJIc", V, Y, optimize=True)
    Y = np.einsum("Kk,ekJIc->eKJIc", V, Y, optimize=True)
    return Y


def change_basis_surf(V: np.ndarray, X: np.ndarray) -> np.ndarray:
    """X[e,b,a,c] -> Y[e,B,A,c] (ChangeBasis2D)."""
    Y = np.einsum("Aa,ebac->ebAc", V, X, optimize=True)
    Y = np.einsum("Bb,ebAc->eBAc", V, Y, optimize=True)
    return Y


def build_coords(NodeCoords: np.ndarray, NGeo: int, N: int, node_type: str) -> np.ndarray:
    """Elem_xGP[e,k,j,i,3] (metrics.f90:40-105, non-tree branch)."""
    V1 = bs.get_vandermonde(NGeo, bs.NODETYPE_VISU, N, bs.NODETYPE_CL)
    V2 = bs.get_vandermonde(N, bs.NODETYPE_CL, N, node_type)
    return change_basis_volume(V2 @ V1, NodeCoords)


class _Ops:
    def __init__(self, NGeo: int, N: int, node_type: str):
        self.NGeo, self.N, self.node_type = NGeo, N, node_type
        self.NGeoRef = 3 * NGeo
        G, CL, VISU = node_type, bs.NODETYPE_CL, bs.NODETYPE_VISU
        self.Vdm_EQNGeo_CLNGeo = bs.get_vandermonde(NGeo, VISU, NGeo, CL)
        self.Vdm_CLNGeo_CLN = bs.get_vandermonde(NGeo, CL, N, CL)
        self.DCL_NGeo = bs.polynomial_derivative_matrix(bs.get_nodes_and_weights(NGeo, CL)[0])
        self.Vdm_CLNGeo_NGeoRef = bs.get_vandermonde(NGeo, CL, self.NGeoRef, G)
        self.Vdm_NGeoRef_N = bs.get_vandermonde(self.NGeoRef, G, N, G, modal=True)
        self.DCL_N = bs.polynomial_derivative_matrix(bs.get_nodes_and_weights(N, CL)[0])
        self.Vdm_CLN_N = bs.get_vandermonde(N, CL, N, G)


def _deriv(D: np.ndarray, X: np.ndarray) -> np.ndarray:
    """dX[e,k,j,i,d1,c]: derivative of X[e,k,j,i,c] in direction d1 (0=xi,1=eta,2=zeta)."""
    d1 = np.einsum("Ii,ekjic->ekjIc", D, X, optimize=True)
    d2 = np.einsum("Jj,ekjic->ekJic", D, X, optimize=True)
    d3 = np.einsum("Kk,ekjic->eKjic", D, X, optimize=True)
    return np.stack([d1, d2, d3], axis=-2)


def _elem_geometry(ops: _Ops, NodeCoords: np.ndarray, crossProductMetrics: bool):
    """Per-element CL-point geometry: XCL_N[e,k,j,i,3], JaCL_N[e,k,j,i,d,c], detJac_N[e,k,j,i]."""
    N, NGeo = ops.N, ops.NGeo
    XCL_NGeo = change_basis_volume(ops.Vdm_EQNGeo_CLNGeo, NodeCoords)
    XCL_N = change_basis_volume(ops.Vdm_CLNGeo_CLN, XCL_NGeo)
    dXCL_NGeo = _deriv(ops.DCL_NGeo, XCL_NGeo)                       # [e,k,j,i,d1,c] == dXCL(d1,c)
    sh = dXCL_NGeo.shape
    dX_ref = change_basis_volume(ops.Vdm_CLNGeo_NGeoRef, dXCL_NGeo.reshape(sh[:4] + (9,)))
    dX_ref = dX_ref.reshape(dX_ref.shape[:4] + (3, 3))
    a = dX_ref
    det = (a[..., 0, 0] * (a[..., 1, 1] * a[..., 2, 2] - a[..., 2, 1] * a[..., 1, 2])
           + a[..., 1, 0] * (a[..., 2, 1] * a[..., 0, 2] - a[..., 0, 1] * a[..., 2, 2])
           + a[..., 2, 0] * (a[..., 0, 1] * a[..., 1, 2] - a[..., 1, 1] * a[..., 0, 2]))
    detJac_N = change_basis_volume(ops.Vdm_NGeoRef_N, det[..., None])[..., 0]
    if N >= NGeo:
        dXCL_N = change_basis_volume(ops.Vdm_CLNGeo_CLN, dXCL_NGeo.reshape(sh[:4] + (9,)))
        dXCL_N = dXCL_N.reshape(dXCL_N.shape[:4] + (3, 3))
    else:
        dXCL_N = _deriv(ops.DCL_N, XCL_N)
    d = dXCL_N  # d[..., a, b] == dXCL(a+1, b+1)
    Ja = np.zeros_like(d)  # Ja[..., a, b] == JaCL_N(a+1, b+1)
    if crossProductMetrics:
        Ja[..., 0, 0] = d[..., 1, 1] * d[..., 2, 2] - d[..., 1, 2] * d[..., 2, 1]
        Ja[..., 1, 0] = d[..., 2, 1] * d[..., 0, 2] - d[..., 2, 2] * d[..., 0, 1]
        Ja[..., 2, 0] = d[..., 0, 1] * d[..., 1, 2] - d[..., 0, 2] * d[..., 1, 1]
        Ja[..., 0, 1] = d[..., 1, 2] * d[..., 2, 0] - d[..., 1, 0] * d[..., 2, 2]
        Ja[..., 1, 1] = d[..., 2, 2] * d[..., 0, 0] - d[..., 2, 0] * d[..., 0, 2]
        Ja[..., 2, 1] = d[..., 0, 2] * d[..., 1, 0] - d[..., 0, 0] * d[..., 1, 2]
        Ja[..., 0, 2] = d[..., 1, 0] * d[..., 2, 1] - d[..., 1, 1] * d[..., 2, 0]
        Ja[..., 1, 2] = d[..., 2, 0] * d[..., 0, 1] - d[..., 2, 1] * d[..., 0, 0]
        Ja[..., 2, 2] = d[..., 0, 0] * d[..., 1, 1] - d[..., 0, 1] * d[..., 1, 0]
    else:
        X = XCL_N
        # R_CL_N(:,c,...) indexed R[..., a, c] with a = first (derivative-direction) index
        R = np.zeros_like(d)
        R[..., :, 0] = 0.5 * (X[..., 2, None] * d[..., :, 1] - X[..., 1, None] * d[..., :, 2])
        R[..., :, 1] = 0.5 * (X[..., 0, None] * d[..., :, 2] - X[..., 2, None] * d[..., :, 0])
        R[..., :, 2] = 0.5 * (X[..., 1, None] * d[..., :, 0] - X[..., 0, None] * d[..., :, 1])
        D = ops.DCL_N
        # JaCL(1,:) = -d/deta R(3,:) + d/dzeta R(2,:) ; JaCL(2,:) = -d/dzeta R(1,:) + d/dxi R(3,:)
        # JaCL(3,:) = -d/dxi R(2,:) + d/deta R(1,:)
        dxi = lambda A: np.einsum("Ii,ekjic->ekjIc", D, A, optimize=True)
        det_ = lambda A: np.einsum("Jj,ekjic->ekJic", D, A, optimize=True)
        dze = lambda A: np.einsum("Kk,ekjic->eKjic", D, A, optimize=True)
        Ja[..., 0, :] = -det_(R[..., 2, :]) + dze(R[..., 1, :])
        Ja[..., 1, :] = -dze(R[..., 0, :]) + dxi(R[..., 2, :])
        Ja[..., 2, :] = -dxi(R[..., 1, :]) + det_(R[..., 0, :])
    return XCL_N, Ja, detJac_N


def det_jac_ref(NodeCoords: np.ndarray, NGeo: int, node_type: str, chunk: int = 4096) -> np.ndarray:
    """detJac_Ref(1,0:3*NGeo,0:3*NGeo,0:3*NGeo,nElems) (metrics.f90:262-283): the Jacobian determinant on the 3*NGeo
    interpolation points of ``node_type``; [e,k,j,i]. Kept by the reference for the conservative restart projection
    (restart.f90:500-512); NodeCoords [e,k,j,i,3] on the equidistant NGeo points."""
    ops = _Ops(NGeo, NGeo, node_type)
    out = np.empty((NodeCoords.shape[0],) + (3 * NGeo + 1,) * 3)
    for s0 in range(0, NodeCoords.shape[0], chunk):
        XCL = change_basis_volume(ops.Vdm_EQNGeo_CLNGeo, NodeCoords[s0:s0 + chunk])
        d = _deriv(ops.DCL_NGeo, XCL)
        a = change_basis_volume(ops.Vdm_CLNGeo_NGeoRef, d.reshape(d.shape[:4] + (9,)))
        a = a.reshape(a.shape[:4] + (3, 3))
        out[s0:s0 + chunk] = (a[..., 0, 0] * (a[..., 1, 1] * a[..., 2, 2] - a[..., 2, 1] * a[..., 1, 2])
                              + a[..., 1, 0] * (a[..., 2, 1] * a[..., 0, 2] - a[..., 0, 1] * a[..., 2, 2])
                              + a[..., 2, 0] * (a[..., 0, 1] * a[..., 1, 2] - a[..., 1, 1] * a[..., 0, 2]))
    return out


def _face_slice(A: np.ndarray, loc: int, N: int) -> np.ndarray:
    """A[e,k,j,i,...] -> face array tmp[e,b,a,...] with (a,b) the two remaining volume indices in order."""
    if loc == mp.XI_MINUS:
        return A[:, :, :, 0]
    if loc == mp.XI_PLUS:
        return A[:, :, :, N]
    if loc == mp.ETA_MINUS:
        return A[:, :, 0, :]
    if loc == mp.ETA_PLUS:
        return A[:, :, N, :]
    if loc == mp.ZETA_MINUS:
        return A[:, 0, :, :]
    if loc == mp.ZETA_PLUS:
        return A[:, N, :, :]
    raise ValueError(loc)


def _surf_metrics_from_ja(Ja_face: np.ndarray, loc: int):
    """Ja_face[s,q,p,d,c] -> NormVec,TangVec1,TangVec2 [s,q,p,3], SurfElem[s,q,p] (metrics.f90:755-791)."""
    nd, td, sg = NormalDirs[loc - 1] - 1, TangDirs[loc - 1] - 1, NormalSigns[loc - 1]
    jn = Ja_face[..., nd, :]
    surf = np.sqrt(np.sum(jn ** 2, axis=-1))
    nv = sg * jn / surf[..., None]
    jt = Ja_face[..., td, :]
    t1 = jt - np.sum(jt * nv, axis=-1)[..., None] * nv
    t1 = t1 / np.sqrt(np.sum(t1 ** 2, axis=-1))[..., None]
    t2 = np.cross(nv, t1)
    return nv, t1, t2, surf


def calc_metrics(mesh: Mesh, N: int, node_type: str, crossProductMetrics: bool = False,
                 hopr: dict | None = None, chunk: int = 4096) -> dict:
    """Returns dict(Elem_xGP, Metrics_fTilde/gTilde/hTilde [e,k,j,i,3], sJ [e,k,j,i], NormVec, TangVec1,
    TangVec2 [side,q,p,3], SurfElem [side,q,p], Face_xGP).

    For MPI sides owned by the neighbour rank (YOUR sides) the reference receives the master's surface
    metrics over MPI (metrics.f90:553-575). Here the master element's geometry is evaluated directly from
    the global node coordinates in ``hopr`` -- the same arithmetic the owning rank performs.
    """
    n = N + 1
    NGeo = mesh.NGeo
    ops = _Ops(NGeo, N, node_type)
    nE, nS = mesh.nElems, mesh.nSides
    Mf = np.zeros((nE, n, n, n, 3))
    Mg = np.zeros((nE, n, n, n, 3))
    Mh = np.zeros((nE, n, n, n, 3))
    sJ = np.zeros((nE, n, n, n))
    Elem_xGP = np.zeros((nE, n, n, n, 3))
    NormVec = np.zeros((nS, n, n, 3))
    TangVec1 = np.zeros((nS, n, n, 3))
    TangVec2 = np.zeros((nS, n, n, 3))
    SurfElem = np.zeros((nS, n, n))
    Face_xGP = np.zeros((nS, n, n, 3))
    S2V2 = mp.build_mappings(N)["S2V2"]
    V_geo_N = ops.Vdm_CLN_N @ ops.Vdm_CLNGeo_CLN @ ops.Vdm_EQNGeo_CLNGeo

    has_mortar = mesh.MortarType is not None and bool(np.any(mesh.MortarType[:, 0] > 0))

    def surf(XCL_N, Ja, elem_sel, loc, side_ids, iMortar=None, mtype=None):
        """Surface metrics of the master sides `side_ids` of the elements `elem_sel` (local side loc). With iMortar /
        mtype (arrays) the result is the iMortar-th small side of that (remote) big side instead."""
        if len(elem_sel) == 0:
            return
        xf = change_basis_surf(ops.Vdm_CLN_N, _face_slice(XCL_N[elem_sel], loc, N))
        jf = _face_slice(Ja[elem_sel], loc, N)
        sh = jf.shape
        jf = change_basis_surf(ops.Vdm_CLN_N, jf.reshape(sh[:3] + (9,))).reshape(sh)
        a = S2V2[loc - 1, 0, :, :, 0]
        b = S2V2[loc - 1, 0, :, :, 1]
        xf = xf[:, b, a]
        jf = jf[:, b, a]
        if iMortar is not None:
            # YOUR side whose master is a virtual small side of a remote big mortar side
            for x in range(len(side_ids)):
                if iMortar[x] > 0:
                    mja, mx = mo.mortar_surf_metrics(int(mtype[x]), N, node_type, jf[x], xf[x])
                    jf[x], xf[x] = mja[iMortar[x] - 1], mx[iMortar[x] - 1]
        nv, t1, t2, se = _surf_metrics_from_ja(jf, loc)
        sid = side_ids - 1
        NormVec[sid], TangVec1[sid], TangVec2[sid], SurfElem[sid], Face_xGP[sid] = nv, t1, t2, se, xf
        if has_mortar and iMortar is None:
            # metrics.f90:727-741 + mortar_metrics.f90: small master sides get the interpolated big-side metrics
            for x in np.nonzero(mesh.MortarType[sid, 0] > 0)[0]:
                big = sid[x]
                mja, mx = mo.mortar_surf_metrics(int(mesh.MortarType[big, 0]), N, node_type, jf[x], xf[x])
                info = mesh.MortarInfo[mesh.MortarType[big, 1] - 1]
                for im in range(len(mja)):
                    if info[im, 1] > 0:
                        continue  # slave small sides (MPI YOUR) are built by their master rank
                    s2 = int(info[im, 0]) - 1
                    n2, u1, u2, s_e = _surf_metrics_from_ja(mja[im][None], loc)
                    NormVec[s2], TangVec1[s2], TangVec2[s2], SurfElem[s2], Face_xGP[s2] = n2[0], u1[0], u2[0], s_e[0], mx[im]

    for s0 in range(0, nE, chunk):
        s1 = min(nE, s0 + chunk)
        nc = mesh.NodeCoords[s0:s1]
        XCL_N, Ja, det = _elem_geometry(ops, nc, crossProductMetrics)
        if np.any(det <= 0.0):
            raise RuntimeError("Negative Jacobian found on Gauss point")
        scaled = det / det.reshape(det.shape[0], -1).max(axis=1)[:, None, None, None]
        if np.any(scaled < 0.01):
            raise RuntimeError("Scaled Jacobian lower then tolerance")
        sJ[s0:s1] = 1.0 / det
        Mf[s0:s1] = change_basis_volume(ops.Vdm_CLN_N, Ja[..., 0, :])
        Mg[s0:s1] = change_basis_volume(ops.Vdm_CLN_N, Ja[..., 1, :])
        Mh[s0:s1] = change_basis_volume(ops.Vdm_CLN_N, Ja[..., 2, :])
        Elem_xGP[s0:s1] = change_basis_volume(V_geo_N, nc)
        e2s = mesh.ElemToSide[s0:s1]
        for loc in range(1, 7):
            sel = np.nonzero(e2s[:, loc - 1, 1] == 0)[0]
            surf(XCL_N, Ja, sel, loc, e2s[sel, loc - 1, 0].astype(np.int64))

    # YOUR sides: geometry of the remote master element
    if mesh.nMPISides_YOUR > 0:
        if hopr is None:
            raise ValueError("hopr (global mesh) needed to build the surface metrics of YOUR MPI sides")
        NG0 = int(hopr["NGeo"])
        nn = (NG0 + 1) ** 3
        e2s = mesh.ElemToSide
        imort = mtyp = None
        if mesh.YourMaster is not None:
            ym = mesh.YourMaster
            side_ids, nb_elem, nb_loc, imort, mtyp = ym[:, 0], ym[:, 1] - 1, ym[:, 2], ym[:, 3], ym[:, 4]
        else:
            your = np.argwhere((e2s[:, :, 0] >= mesh.firstMPISide_YOUR) & (e2s[:, :, 0] <= mesh.lastMPISide_YOUR))
            ei = hopr["ElemInfo"]
            si = hopr["SideInfo"]
            rows = ei[mesh.offsetElem + your[:, 0], 2].astype(np.int64) + your[:, 1]
            nb_elem = si[rows, 2].astype(np.int64) - 1
            nb_loc = (si[rows, 3] // 10).astype(np.int64)
            side_ids = e2s[your[:, 0], your[:, 1], 0].astype(np.int64)
        for loc in range(1, 7):
            sel = np.nonzero(nb_loc == loc)[0]
            for c0 in range(0, len(sel), chunk):
                ss = sel[c0:c0 + chunk]
                idx = (nb_elem[ss][:, None] * nn + np.arange(nn)[None, :]).ravel()
                nc = hopr["NodeCoords"][idx].reshape(len(ss), NG0 + 1, NG0 + 1, NG0 + 1, 3)
                if NG0 != NGeo:
                    nc = nc[:, ::NG0, ::NG0, ::NG0, :]
                XCL_N, Ja, _ = _elem_geometry(ops, np.ascontiguousarray(nc), crossProductMetrics)
                surf(XCL_N, Ja, np.arange(len(ss)), loc, side_ids[ss], None if imort is None else imort[ss],
                     None if mtyp is None else mtyp[ss])

    return dict(Elem_xGP=Elem_xGP, Metrics_fTilde=Mf, Metrics_gTilde=Mg, Metrics_hTilde=Mh, sJ=sJ,
                NormVec=NormVec, TangVec1=TangVec1, TangVec2=TangVec2, SurfElem=SurfElem, Face_xGP=Face_xGP)
