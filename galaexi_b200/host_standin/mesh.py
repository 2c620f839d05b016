"""HOPR-layout meshes: synthetic generation, SFC partition, local side numbering and connectivity maps.

Host-side (init-time) mirror of the reference's mesh preparation. The arrays built here are exactly the
integer tables the hot path consumes: ``ElemToSide(3,6,nElems)``,
``SideToElem(5,nSides)``, ``BC(nBCSides)``, the side ranges and the MPI neighbour tables.

Follows (relative to /root/reference):
  * src/mesh/mesh_readin.f90:734-785  BuildPartition (contiguous SFC slices), :793-829 ELEMIPROC
  * src/mesh/mesh_readin.f90:202-563  ReadMesh (side objects, flips, connections, counts)
  * src/mesh/prepare_mesh.f90:60-500  setLocalSideIDs (first-touch numbering; MPI sides sorted by
                                      *negated* global side index, MINE/YOUR split by rank order)
  * src/mesh/prepare_mesh.f90:836-920 exchangeFlip
  * src/mesh/prepare_mesh.f90:688-775 fillMeshInfo
  * src/mesh/mesh.f90:259-283, 416-434 side ranges, E2S_IS_MASTER

The reference walks pointer lists element by element; here the same numbering is produced with array
operations ("a side is numbered when the element-major, locSide-minor walk first touches it").
``oracle/mesh_walk.py`` holds a literal loop restatement used by the tests to pin this bit-exactly.

Non-conforming (mortar) meshes (rejected by GALAEXI, src/mesh/mesh.f90:140-143, but supported by the inherited host
code) are handled by ``mesh_mortar.prepare_mesh_general``, a loop restatement that also serves as the cross-check of
the vectorised path on conforming meshes.

All 1-based reference indices (SideID, ElemID, locSide, BC index) are kept 1-based inside the integer
tables, so they can be handed to a Fortran host unchanged; 0 / -1 keep the reference meaning.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import mappings as mp
from . import mesh_mortar

# local side -> (axis, +/-): ZETA_MINUS=1, ETA_MINUS=2, XI_PLUS=3, ETA_PLUS=4, XI_MINUS=5, ZETA_PLUS=6
_SIDE_AXIS = {1: (2, -1), 2: (1, -1), 3: (0, +1), 4: (1, +1), 5: (0, -1), 6: (2, +1)}
_OPPOSITE = {1: 6, 6: 1, 2: 4, 4: 2, 3: 5, 5: 3}


# --------------------------------------------------------------------------------------------------
# space filling curve
# --------------------------------------------------------------------------------------------------
def hilbert_index_3d(ijk: np.ndarray, bits: int) -> np.ndarray:
    """Hilbert curve index of integer coordinates (Skilling's transpose algorithm), vectorised.

    HOPR sorts elements along a Hilbert curve before writing the mesh file, which is what makes the
    reference's contiguous-slice partition compact (mesh_readin.f90:766-778).
    """
    X = [ijk[:, d].astype(np.uint64).copy() for d in range(3)]
    M = np.uint64(1) << np.uint64(bits - 1)
    Q = M
    while Q > 1:
        P = Q - np.uint64(1)
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = (X[0] ^ X[i]) & P
            t = np.where(hit, np.uint64(0), t)
            X[0] ^= t
            X[i] ^= t
        Q >>= np.uint64(1)
    for i in range(1, 3):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - np.uint64(1)), t)
        Q >>= np.uint64(1)
    for i in range(3):
        X[i] ^= t
    h = np.zeros_like(X[0])
    for b in range(bits - 1, -1, -1):
        for i in range(3):
            h = (h << np.uint64(1)) | ((X[i] >> np.uint64(b)) & np.uint64(1))
    return h


# --------------------------------------------------------------------------------------------------
# synthetic HOPR-layout box mesh
# --------------------------------------------------------------------------------------------------
def _structured_flip(loc_master: int, loc_slave: int) -> int:
    """Flip of the slave side for two identically oriented hexes sharing a face (brute force on N=1)."""
    N = 1
    ax, sg = _SIDE_AXIS[loc_master]

    def pos(loc, i, j, k, shift):
        x = [float(i), float(j), float(k)]
        x[ax] += shift
        return tuple(x)

    for f in range(1, 5):
        ok = True
        for q in range(N + 1):
            for p in range(N + 1):
                a = mp.side_to_vol(N, 0, p, q, 0, loc_master)
                b = mp.side_to_vol(N, 0, p, q, f, loc_slave)
                # slave element is shifted by sg along ax
                if pos(loc_master, *a, 0.0) != pos(loc_slave, *b, float(sg)):
                    ok = False
            if not ok:
                break
        if ok:
            return f
    raise RuntimeError("no flip found")


def make_box_mesh(nelems=(4, 4, 4), x0=(-1.0, -1.0, -1.0), x1=(1.0, 1.0, 1.0), NGeo: int = 1,
                  bctype=None, sfc: bool = True, deform: float = 0.0, stretch=None) -> dict:
    """Cartesian (optionally sine-deformed / stretched) hex box in HOPR file layout.

    bctype: six entries in local-side order [z-, y-, x+, y+, x-, z+]; each is either the string
            ``"periodic"`` or a tuple ``(BCType, BCState)`` (e.g. (4,1) isothermal wall, (2,1) Dirichlet
            refstate, (3,0) adiabatic wall, (9,0) slip). Default: fully periodic.
    deform: amplitude a of  x += a * L_d/2 * sin(pi xh) sin(pi yh) sin(pi zh), with xh in [-1,1] the
            box-normalised coordinate (the reference's ``meshdeform``, mesh.f90:224-235, which is written
            for the [-1,1]^3 box). Needs NGeo>=2 to be represented as a curved mesh.
    stretch: optional callable (d, s in [0,1]) -> s' in [0,1] grading the node distribution.

    Returns dict(ElemInfo(nE,6), SideInfo(6 nE,5), NodeCoords(nE (NGeo+1)^3,3), BCNames, BCType(nBC,4),
    NGeo, Elem_IJK(nE,3), nElems_IJK(3)) -- the datasets of mesh_readin.f90:33-50.
    """
    ex, ey, ez = (int(v) for v in nelems)
    nE = ex * ey * ez
    if bctype is None:
        bctype = ["periodic"] * 6
    for a, b in ((1, 6), (2, 4), (3, 5)):
        if (bctype[a - 1] == "periodic") != (bctype[b - 1] == "periodic"):
            raise ValueError("periodic BCs must come in opposite pairs")
    ne_ax = (ex, ey, ez)
    for loc in range(1, 7):
        ax, _ = _SIDE_AXIS[loc]
        if bctype[loc - 1] == "periodic" and ne_ax[ax] < 2:
            raise ValueError("periodic directions need at least 2 elements")

    # element ordering
    I, J, K = np.meshgrid(np.arange(ex), np.arange(ey), np.arange(ez), indexing="ij")
    ijk = np.stack([I.ravel(), J.ravel(), K.ravel()], axis=1)
    if sfc:
        bits = max(1, int(np.ceil(np.log2(max(ex, ey, ez)))))
        order = np.argsort(hilbert_index_3d(ijk, bits), kind="stable")
    else:
        order = np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))  # x fastest
    ijk = ijk[order]
    gid = -np.ones((ex, ey, ez), dtype=np.int64)
    gid[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = np.arange(nE)  # 0-based global elem index

    # BC table (HOPR: BCType(nBCs,4) = [type, curveIndex, state, alpha])
    names = ["BC_z-", "BC_y-", "BC_x+", "BC_y+", "BC_x-", "BC_z+"]
    alpha_of = {5: 1, 3: -1, 2: 2, 4: -2, 1: 3, 6: -3}
    BCType = np.zeros((6, 4), dtype=np.int32)
    for loc in range(1, 7):
        bt = bctype[loc - 1]
        if bt == "periodic":
            BCType[loc - 1] = (1, 0, 0, alpha_of[loc])
        else:
            BCType[loc - 1] = (int(bt[0]), 0, int(bt[1]), 0)

    flips = {(a, _OPPOSITE[a]): _structured_flip(a, _OPPOSITE[a]) for a in (3, 4, 6)}

    ElemInfo = np.zeros((nE, 6), dtype=np.int32)
    ElemInfo[:, 0] = 108
    ElemInfo[:, 1] = 1
    ElemInfo[:, 2] = 6 * np.arange(nE)
    ElemInfo[:, 3] = 6 * (np.arange(nE) + 1)
    nn = (NGeo + 1) ** 3
    ElemInfo[:, 4] = nn * np.arange(nE)
    ElemInfo[:, 5] = nn * (np.arange(nE) + 1)

    SideInfo = np.zeros((nE, 6, 5), dtype=np.int32)
    SideInfo[:, :, 0] = 4 if NGeo == 1 else 7  # bilinear / curved quad (informational only)
    # neighbour element and BC per local side
    nb = -np.ones((nE, 6), dtype=np.int64)
    bcid = np.zeros((nE, 6), dtype=np.int32)
    for loc in range(1, 7):
        ax, sg = _SIDE_AXIS[loc]
        c = ijk.copy()
        c[:, ax] += sg
        out = (c[:, ax] < 0) | (c[:, ax] >= ne_ax[ax])
        per = bctype[loc - 1] == "periodic"
        c[:, ax] %= ne_ax[ax]
        nbe = gid[c[:, 0], c[:, 1], c[:, 2]]
        if per:
            nb[:, loc - 1] = nbe
            bcid[:, loc - 1] = np.where(out, loc, 0)
        else:
            nb[:, loc - 1] = np.where(out, -1, nbe)
            bcid[:, loc - 1] = np.where(out, loc, 0)
    # unique side ids by first touch (element-major, locSide-minor); master = the "+" side owner
    e_idx = np.repeat(np.arange(nE), 6)
    s_idx = np.tile(np.arange(6), nE)
    nbf = nb.ravel()
    partner = np.where(nbf >= 0, nbf * 6 + (np.array([_OPPOSITE[s + 1] - 1 for s in range(6)])[s_idx]), -1)
    t = np.arange(nE * 6)
    key = np.where(partner >= 0, np.minimum(t, partner), t)
    uniq, inv = np.unique(key, return_inverse=True)
    side_gid = (inv + 1).astype(np.int32)  # 1-based global unique side id
    is_plus = np.isin(s_idx + 1, (3, 4, 6))
    has_nb = nbf >= 0
    oriented = np.where(has_nb, is_plus, True)
    SI = SideInfo.reshape(nE * 6, 5)
    SI[:, 1] = np.where(oriented, side_gid, -side_gid)
    SI[:, 2] = np.where(has_nb, nbf + 1, 0)
    opp = np.array([_OPPOSITE[s + 1] for s in range(6)])[s_idx]
    fl = np.zeros(nE * 6, dtype=np.int32)
    for (a, b), f in flips.items():
        # slave side b (minus) sees flip f; master side a stores the flip too (HOPR stores the relative
        # flip on both; only the non-oriented one is used, mesh_readin.f90:363-370)
        fl[(s_idx + 1 == b) & has_nb] = f
        fl[(s_idx + 1 == a) & has_nb] = f
    SI[:, 3] = np.where(has_nb, 10 * opp + fl, 0)
    SI[:, 4] = bcid.ravel()

    # node coordinates: (NGeo+1)^3 equidistant nodes per element, i fastest
    xi = np.arange(NGeo + 1) / float(NGeo)
    lo = np.asarray(x0, float)
    hi = np.asarray(x1, float)

    def grade(d, s):
        return s if stretch is None else stretch(d, s)

    axes = []
    for d in range(3):
        s = (ijk[:, d][:, None] + xi[None, :]) / float(ne_ax[d])  # (nE, NGeo+1) in [0,1]
        s = grade(d, s)
        axes.append(s)
    # node (i,j,k) of elem e: x=axes[0][e,i], y=axes[1][e,j], z=axes[2][e,k]; memory order k,j,i slowest->fastest
    sx = np.broadcast_to(axes[0][:, None, None, :], (nE, NGeo + 1, NGeo + 1, NGeo + 1))
    sy = np.broadcast_to(axes[1][:, None, :, None], (nE, NGeo + 1, NGeo + 1, NGeo + 1))
    sz = np.broadcast_to(axes[2][:, :, None, None], (nE, NGeo + 1, NGeo + 1, NGeo + 1))
    S = np.stack([sx, sy, sz], axis=-1)  # (nE,k,j,i,3) in [0,1]
    if deform != 0.0:
        h = 2.0 * S - 1.0
        bump = deform * np.sin(np.pi * h[..., 0]) * np.sin(np.pi * h[..., 1]) * np.sin(np.pi * h[..., 2])
        S = S + 0.5 * bump[..., None]
    X = lo + S * (hi - lo)
    NodeCoords = np.ascontiguousarray(X.reshape(nE * nn, 3))

    return dict(NGeo=NGeo, ElemInfo=ElemInfo, SideInfo=SI.copy(), NodeCoords=NodeCoords,
                BCNames=names, BCType=BCType, Elem_IJK=(ijk + 1).astype(np.int32),
                nElems_IJK=np.array(ne_ax, dtype=np.int32), isMortarMesh=0)


# --------------------------------------------------------------------------------------------------
# partition
# --------------------------------------------------------------------------------------------------
def build_partition(nGlobalElems: int, nProcs: int) -> np.ndarray:
    """offsetElemMPI(0:nProcs) (mesh_readin.f90:766-778)."""
    if nGlobalElems < nProcs:
        raise ValueError("Number of elements is smaller than number of processors")
    n = nGlobalElems // nProcs
    r = nGlobalElems - n * nProcs
    off = np.array([n * p + min(p, r) for p in range(nProcs)] + [nGlobalElems], dtype=np.int64)
    return off


def elem_to_proc(elem_gid_1based: np.ndarray, offsetElemMPI: np.ndarray) -> np.ndarray:
    """ELEMIPROC (mesh_readin.f90:793-829): owner rank of a 1-based global element id."""
    return (np.searchsorted(offsetElemMPI, elem_gid_1based, side="left") - 1).astype(np.int64)


# --------------------------------------------------------------------------------------------------
# prepared (per-rank) mesh
# --------------------------------------------------------------------------------------------------
@dataclass
class Mesh:
    nGlobalElems: int
    nElems: int
    offsetElem: int
    nSides: int
    nBCSides: int
    nInnerSides: int
    nMPISides: int
    nMPISides_MINE: int
    nMPISides_YOUR: int
    NGeo: int
    # side ranges (1-based inclusive, reference names)
    firstBCSide: int = 0
    lastBCSide: int = 0
    firstMortarInnerSide: int = 0
    lastMortarInnerSide: int = 0
    firstInnerSide: int = 0
    lastInnerSide: int = 0
    firstMPISide_MINE: int = 0
    lastMPISide_MINE: int = 0
    firstMPISide_YOUR: int = 0
    lastMPISide_YOUR: int = 0
    firstMortarMPISide: int = 0
    lastMortarMPISide: int = 0
    # tables: numpy C-order arrays whose memory equals the Fortran arrays
    ElemToSide: np.ndarray = None   # (nElems,6,3)  == ElemToSide(3,6,nElems)
    SideToElem: np.ndarray = None   # (nSides,5)    == SideToElem(5,nSides)
    BC: np.ndarray = None           # (nBCSides,)   index into BoundaryType
    AnalyzeSide: np.ndarray = None
    SideToGlobalSide: np.ndarray = None
    BoundaryType: np.ndarray = None  # (nBCs,3) = [BC_TYPE, BC_STATE, BC_ALPHA]
    BoundaryName: list = field(default_factory=list)
    NodeCoords: np.ndarray = None   # (nElems,NGeo+1,NGeo+1,NGeo+1,3) == NodeCoords(3,0:NGeo,..,nElems)
    # MPI neighbour tables
    nNbProcs: int = 0
    NbProc: np.ndarray = None
    nMPISides_Proc: np.ndarray = None
    nMPISides_MINE_Proc: np.ndarray = None
    nMPISides_YOUR_Proc: np.ndarray = None
    offsetMPISides_MINE: np.ndarray = None  # (0:nNbProcs)
    offsetMPISides_YOUR: np.ndarray = None
    offsetElemMPI: np.ndarray = None
    myRank: int = 0
    nProcs: int = 1
    # non-conforming interfaces (mesh.f90:318-322): MortarType(2,nSides) = [type (1..3 big, -1 small, 0), index into
    # MortarInfo]; MortarInfo(2,4,nMortarSides) = [SideID, flip] of the small sides
    nMortarSides: int = 0
    nMortarInnerSides: int = 0
    nMortarMPISides: int = 0
    MortarType: np.ndarray = None   # (nSides,2)
    MortarInfo: np.ndarray = None   # (nMortarSides,4,2)
    YourMaster: np.ndarray = None   # general path only: (nYOUR,5) [SideID, master elem, locSide, iMortar, mortar type]


def _apply_user_bcs(BCNames, BCType, user_bcs):
    """readBCs (mesh_readin.f90:60-169): ini BoundaryName/BoundaryType override type+state by name."""
    BCType = BCType.copy()
    if user_bcs:
        for name, (t, s) in user_bcs.items():
            hit = [i for i, n in enumerate(BCNames) if n.strip().lower() == name.strip().lower()]
            if not hit:
                raise ValueError(f"Boundary condition specified in parameter file has not been found: {name}")
            for i in hit:
                if t == 1 and BCType[i, 0] != 1:
                    raise ValueError("Remapping non-periodic to periodic BCs is not possible!")
                BCType[i, 0] = t
                BCType[i, 2] = s
    BoundaryType = np.stack([BCType[:, 0], BCType[:, 2], BCType[:, 3]], axis=1).astype(np.int32)
    return BoundaryType


def prepare_mesh(hopr: dict, nProcs: int = 1, myRank: int = 0, useCurveds: bool = True,
                 user_bcs: dict | None = None, general: bool = False) -> Mesh:
    """ReadMesh + setLocalSideIDs + exchangeFlip + fillMeshInfo for one rank (see module docstring)."""
    ElemInfo = hopr["ElemInfo"]
    SideInfo = hopr["SideInfo"]
    nGlobal = ElemInfo.shape[0]
    offMPI = build_partition(nGlobal, nProcs)
    offsetElem = int(offMPI[myRank])
    nElems = int(offMPI[myRank + 1] - offMPI[myRank])
    BoundaryType = _apply_user_bcs(hopr["BCNames"], hopr["BCType"], user_bcs)
    nBCs = BoundaryType.shape[0]
    if general or mesh_mortar.has_mortars(hopr):
        # non-conforming interfaces: host-FLEXI bookkeeping (GALAEXI aborts here, src/mesh/mesh.f90:140-143)
        return mesh_mortar.prepare_mesh_general(hopr, nProcs, myRank, useCurveds, BoundaryType, offMPI, elem_to_proc, Mesh)

    ei = ElemInfo[offsetElem:offsetElem + nElems]
    if np.any(ei[:, 3] - ei[:, 2] != 6):
        raise NotImplementedError("only conforming hexahedra with 6 sides (no mortars)")
    first = ei[:, 2].astype(np.int64)
    rows = (first[:, None] + np.arange(6)[None, :]).ravel()
    si = SideInfo[rows]                       # (nElems*6,5)
    if np.any(si[:, 2] < 0) or np.any(si[:, 0] < 0):
        raise NotImplementedError("mortar sides found")
    nS6 = nElems * 6
    ind = np.abs(si[:, 1]).astype(np.int64)
    oriented = si[:, 1] > 0
    flip = np.where(oriented, 0, si[:, 3] % 10).astype(np.int64)
    if np.any((flip < 0) | (flip > 4)):
        raise ValueError("NodeID doesnt belong to side")
    nbElem = si[:, 2].astype(np.int64)        # 1-based global, 0 = none
    BCindex = si[:, 4].astype(np.int64)

    # ---- connections (mesh_readin.f90:380-440)
    is_real_bc = np.zeros(nS6, dtype=bool)
    hasbc = BCindex != 0
    btype = np.where(hasbc, BoundaryType[np.maximum(BCindex, 1) - 1, 0], 0)
    is_real_bc = hasbc & (btype != 1) & (btype != 100)
    flip = np.where(is_real_bc, 0, flip)
    conn = -np.ones(nS6, dtype=np.int64)      # flat index of connected local side
    NbProc = -np.ones(nS6, dtype=np.int64)
    cand = (~is_real_bc) & (nbElem != 0)
    local_nb = cand & (nbElem > offsetElem) & (nbElem <= offsetElem + nElems)
    remote_nb = cand & ~local_nb
    if np.any(remote_nb):
        NbProc[remote_nb] = elem_to_proc(nbElem[remote_nb], offMPI)
    # local: find side of neighbour element with same global side index ("last match wins", see
    # the EXIT placement in mesh_readin.f90:407-419)
    idx = np.nonzero(local_nb)[0]
    if idx.size:
        nbl = nbElem[idx] - 1 - offsetElem
        nb_inds = ind.reshape(nElems, 6)[nbl]             # (n,6)
        match = nb_inds == ind[idx][:, None]
        # exclude the side itself unless it is the only match
        self_pos = np.where(nbl == idx // 6, idx % 6, -1)
        match_ns = match.copy()
        rows_self = np.nonzero(self_pos >= 0)[0]
        match_ns[rows_self, self_pos[rows_self]] = False
        use = np.where(match_ns.any(axis=1)[:, None], match_ns, match)
        if not use.any(axis=1).all():
            raise RuntimeError("neighbour side not found")
        last = 5 - np.argmax(use[:, ::-1], axis=1)
        conn[idx] = nbl * 6 + last
    # the reference only connects if not already connected; with consistent meshes this is symmetric
    # ---- periodic BC remap (prepare_mesh.f90:83-107)
    PeriodicBCMap = np.full(nBCs, -2, dtype=np.int64)
    for i in range(nBCs):
        if BoundaryType[i, 0] != 1:
            PeriodicBCMap[i] = -1
        elif BoundaryType[i, 2] > 0:
            PeriodicBCMap[i] = -1
        elif BoundaryType[i, 2] < 0:
            for j in range(nBCs):
                if BoundaryType[j, 0] != 1:
                    continue
                if BoundaryType[j, 2] == -BoundaryType[i, 2]:
                    PeriodicBCMap[i] = j + 1
    if np.any(PeriodicBCMap == -2):
        raise RuntimeError("Periodic connection not found.")
    m = BCindex >= 1
    pm = np.where(m, PeriodicBCMap[np.maximum(BCindex, 1) - 1], -1)
    BCindex = np.where(m & (pm != -1), pm, BCindex)

    # ---- counts (mesh_readin.f90:585-640)
    t = np.arange(nS6)
    has_conn = conn >= 0
    is_mpi = NbProc >= 0
    is_bc = (~has_conn) & (~is_mpi)
    if np.any(is_bc & ~is_real_bc):
        raise RuntimeError("side without connection that is not a BC side")
    first_touch = has_conn & (t <= conn)
    nInnerSides = int(first_touch.sum())
    nBCSides = int(is_bc.sum())
    nMPISides = int(is_mpi.sum())
    nSides = nInnerSides + nBCSides + nMPISides

    # ---- local side ids (prepare_mesh.f90:141-190)
    SideID = -np.ones(nS6, dtype=np.int64)
    bc_t = t[is_bc]
    SideID[bc_t] = np.arange(1, nBCSides + 1)
    in_t = t[first_touch]
    ids = nBCSides + np.arange(1, nInnerSides + 1)
    SideID[in_t] = ids
    SideID[conn[in_t]] = ids

    # ---- MPI sides (prepare_mesh.f90:196-320)
    nbprocs = np.unique(NbProc[is_mpi]) if nMPISides else np.zeros(0, dtype=np.int64)
    nNb = len(nbprocs)
    nMPISides_Proc = np.array([int((NbProc == p).sum()) for p in nbprocs], dtype=np.int64)
    mine_proc = np.array([(c // 2) if myRank < p else (c - c // 2) for p, c in zip(nbprocs, nMPISides_Proc)],
                         dtype=np.int64)
    your_proc = nMPISides_Proc - mine_proc
    off_mine = np.zeros(nNb + 1, dtype=np.int64)
    off_your = np.zeros(nNb + 1, dtype=np.int64)
    off_mine[0] = nInnerSides + nBCSides
    off_mine[1:] = off_mine[0] + np.cumsum(mine_proc)
    off_your[0] = off_mine[nNb]
    off_your[1:] = off_your[0] + np.cumsum(your_proc)
    for ib, p in enumerate(nbprocs):
        sel = np.nonzero(NbProc == p)[0]
        # non-mortar sides enter the sort with negated global index -> descending global index
        order = np.argsort(-ind[sel], kind="stable")
        pos = np.empty(len(sel), dtype=np.int64)
        pos[order] = np.arange(1, len(sel) + 1)
        if myRank < p:
            mine = pos <= mine_proc[ib]
            sid = np.where(mine, pos + off_mine[ib], pos - mine_proc[ib] + off_your[ib])
        else:
            your = pos <= your_proc[ib]
            sid = np.where(your, pos + off_your[ib], pos - your_proc[ib] + off_mine[ib])
        SideID[sel] = sid
    nMINE = int(mine_proc.sum())
    nYOUR = int(your_proc.sum())
    if np.any(SideID < 1):
        raise RuntimeError("not all SideIDs are set!")

    # ---- exchangeFlip (prepare_mesh.f90:836-920): MINE sides get flip 0; YOUR sides whose file flip is 0
    # take the master's file flip. The master's file flip is the relative flip stored in its SIDE_Flip entry.
    if nMPISides:
        is_your = is_mpi & (SideID > off_your[0])
        is_mine = is_mpi & ~is_your
        rel = (si[:, 3] % 10).astype(np.int64)   # relative flip as stored on this side's SideInfo row
        need = is_your & (flip == 0)
        if np.any(need & (rel == 0)):
            raise RuntimeError("problem in exchangeflip")
        flip = np.where(need, rel, flip)
        flip = np.where(is_mine, 0, flip)

    # ---- fillMeshInfo (prepare_mesh.f90:688-745)
    ElemToSide = np.zeros((nElems, 6, 3), dtype=np.int32)
    ElemToSide[:, :, 0] = SideID.reshape(nElems, 6)
    ElemToSide[:, :, 1] = flip.reshape(nElems, 6)
    SideToElem = -np.ones((nSides, 5), dtype=np.int32)
    AnalyzeSide = np.zeros(nSides, dtype=np.int32)
    BC = np.zeros(nBCSides, dtype=np.int32)
    e_of = (t // 6 + 1).astype(np.int32)
    l_of = (t % 6 + 1).astype(np.int32)
    root = flip == 0
    SideToElem[SideID[root] - 1, 0] = e_of[root]
    SideToElem[SideID[root] - 1, 2] = l_of[root]
    AnalyzeSide[SideID[root] - 1] = BCindex[root]
    sl = ~root
    SideToElem[SideID[sl] - 1, 1] = e_of[sl]
    SideToElem[SideID[sl] - 1, 3] = l_of[sl]
    SideToElem[SideID[sl] - 1, 4] = flip[sl]
    isb = SideID <= nBCSides
    BC[SideID[isb] - 1] = BCindex[isb]
    # E2S_IS_MASTER + SideToGlobalSide (mesh.f90:416-434)
    ism = SideToElem[SideID - 1, 0] == e_of
    iss = SideToElem[SideID - 1, 1] == e_of
    if not np.all(ism | iss):
        raise RuntimeError("Seems like an error in side connectivity!")
    ElemToSide[:, :, 2] = np.where(ism, 1, 0).reshape(nElems, 6)
    SideToGlobalSide = np.zeros(nSides, dtype=np.int32)
    SideToGlobalSide[SideID - 1] = ind

    # ---- node coordinates
    NGeo = int(hopr["NGeo"])
    nn = (NGeo + 1) ** 3
    nc = hopr["NodeCoords"][offsetElem * nn:(offsetElem + nElems) * nn].reshape(nElems, NGeo + 1, NGeo + 1, NGeo + 1, 3)
    if not useCurveds and NGeo > 1:
        nc = nc[:, ::NGeo, ::NGeo, ::NGeo, :]
        NGeo = 1
    nc = np.ascontiguousarray(nc, dtype=np.float64)

    m = Mesh(nGlobalElems=nGlobal, nElems=nElems, offsetElem=offsetElem, nSides=nSides, nBCSides=nBCSides,
             nInnerSides=nInnerSides, nMPISides=nMPISides, nMPISides_MINE=nMINE, nMPISides_YOUR=nYOUR, NGeo=NGeo)
    m.firstBCSide = 1
    m.firstMortarInnerSide = m.firstBCSide + nBCSides
    m.firstInnerSide = m.firstMortarInnerSide
    m.firstMPISide_MINE = m.firstInnerSide + nInnerSides
    m.firstMPISide_YOUR = m.firstMPISide_MINE + nMINE
    m.firstMortarMPISide = m.firstMPISide_YOUR + nYOUR
    m.lastBCSide = m.firstMortarInnerSide - 1
    m.lastMortarInnerSide = m.firstInnerSide - 1
    m.lastInnerSide = m.firstMPISide_MINE - 1
    m.lastMPISide_MINE = m.firstMPISide_YOUR - 1
    m.lastMPISide_YOUR = m.firstMortarMPISide - 1
    m.lastMortarMPISide = nSides
    m.ElemToSide = ElemToSide
    m.SideToElem = SideToElem
    m.BC = BC
    m.AnalyzeSide = AnalyzeSide
    m.SideToGlobalSide = SideToGlobalSide
    m.BoundaryType = BoundaryType
    m.BoundaryName = list(hopr["BCNames"])
    m.NodeCoords = nc
    m.nNbProcs = nNb
    m.NbProc = nbprocs.astype(np.int32)
    m.nMPISides_Proc = nMPISides_Proc
    m.nMPISides_MINE_Proc = mine_proc
    m.nMPISides_YOUR_Proc = your_proc
    m.offsetMPISides_MINE = off_mine
    m.offsetMPISides_YOUR = off_your
    m.offsetElemMPI = offMPI
    m.myRank = myRank
    m.nProcs = nProcs
    m.MortarType = np.zeros((nSides, 2), dtype=np.int32)
    m.MortarInfo = -np.ones((1, 4, 2), dtype=np.int32)
    return m
