"""Per-rank mesh preparation for HOPR meshes WITH non-conforming (mortar) interfaces.

GALAEXI itself aborts on mortar meshes (src/mesh/mesh.f90:140-143); the mortar bookkeeping below follows
the inherited host FLEXI code that is still in the reference tree, paths relative to
/root/reference/src:
  * mesh/mesh_readin.f90:300-470   side objects incl. the virtual small sides of a big mortar side, connections
  * mesh/mesh_readin.f90:568-650   side counts
  * mesh/prepare_mesh.f90:60-520   setLocalSideIDs: [BC | inner mortar | inner | MPI MINE | MPI YOUR | MPI mortar],
                                   MPI sides sorted by signed global index (conforming sides first), big mortar
                                   sides without a YOUR small side moved to the inner mortars
  * mesh/prepare_mesh.f90:836-920  exchangeFlip
  * mesh/prepare_mesh.f90:688-775  fillMeshInfo: ElemToSide, SideToElem, MortarType(2,nSides), MortarInfo(2,4,nMortarSides)

The walk is element-major / local-side-minor / small-side-minor ("slot" order == SideInfo row order), written
with explicit loops: mortar meshes used for parity work are small. ``mesh.prepare_mesh`` dispatches here when
the mesh file contains a big mortar side; on conforming meshes both paths give identical tables
(tests/test_host_goldens.py).
"""
from __future__ import annotations

import numpy as np


class _Slot:
    __slots__ = ("elem", "loc", "iMortar", "row", "MortarType", "nMortars", "ind", "flip", "BCindex", "conn", "virt",
                 "NbProc", "SideID", "tmp", "small", "nbElem")

    def __init__(self, elem, loc, iMortar, row):
        self.elem, self.loc, self.iMortar, self.row = elem, loc, iMortar, row
        self.MortarType = 0
        self.nMortars = 0
        self.ind = 0
        self.flip = 0
        self.BCindex = 0
        self.conn = None     # connected local slot
        self.virt = False    # connection is a virtual side on another rank
        self.NbProc = -1
        self.SideID = -1
        self.tmp = 0
        self.small = []      # virtual small sides of a big mortar side
        self.nbElem = 0

    @property
    def connected(self):
        return self.conn is not None or self.virt


def has_mortars(hopr: dict) -> bool:
    return bool(np.any(hopr["SideInfo"][:, 2] < 0))


def prepare_mesh_general(hopr: dict, nProcs: int, myRank: int, useCurveds: bool, BoundaryType: np.ndarray,
                         offMPI: np.ndarray, elem_to_proc, Mesh):
    ElemInfo, SideInfo = hopr["ElemInfo"], hopr["SideInfo"]
    nGlobal = ElemInfo.shape[0]
    offsetElem = int(offMPI[myRank])
    nElems = int(offMPI[myRank + 1] - offMPI[myRank])
    nBCs = BoundaryType.shape[0]
    first, last = offsetElem + 1, offsetElem + nElems

    # ---- side objects (mesh_readin.f90:320-375)
    prim = [[None] * 6 for _ in range(nElems)]      # primary sides
    for ie in range(nElems):
        r = int(ElemInfo[offsetElem + ie, 2])        # first row - 1 (0-based row of the first side)
        for loc in range(1, 7):
            a = _Slot(ie, loc, 0, r)
            prim[ie][loc - 1] = a
            nb = int(SideInfo[r, 2])
            if nb < 0:
                a.MortarType = -nb
                a.nMortars = 4 if a.MortarType == 1 else 2
            if SideInfo[r, 0] < 0:
                a.MortarType = -1
            if a.MortarType <= 0:
                a.ind = abs(int(SideInfo[r, 1]))
                if SideInfo[r, 1] > 0:
                    a.flip = 0
                else:
                    a.flip = int(SideInfo[r, 3]) % 10
                    if a.flip < 0 or a.flip > 4:
                        raise ValueError("NodeID doesnt belong to side")
                r += 1
            else:
                r += 1
                for im in range(1, a.nMortars + 1):
                    s = _Slot(ie, loc, im, r)
                    if SideInfo[r, 1] < 0:
                        raise ValueError("Problem in Mortar readin,should be flip=0")
                    s.flip = 0
                    s.ind = abs(int(SideInfo[r, 1]))
                    a.small.append(s)
                    r += 1
        if r != int(ElemInfo[offsetElem + ie, 3]):
            raise ValueError("SideInfo rows of an element do not match its ElemInfo range")

    def slots_of(a):
        """DO iMortar=0,nMortars with the loop bound taken before the body runs."""
        return [a] + list(a.small)

    def walk():
        for ie in range(nElems):
            for loc in range(6):
                a = prim[ie][loc]
                for s in [a] + list(a.small):
                    yield a, s

    # ---- connections (mesh_readin.f90:388-455)
    for ie in range(nElems):
        for loc in range(6):
            a0 = prim[ie][loc]
            for s in slots_of(a0):
                s.nbElem = int(SideInfo[s.row, 2])
                s.BCindex = int(SideInfo[s.row, 4])
                if s.BCindex != 0:
                    bt = int(BoundaryType[s.BCindex - 1, 0])
                    if bt != 1 and bt != 100:
                        s.flip = 0
                        if s.iMortar == 0:
                            s.MortarType = 0
                            s.nMortars = 0
                            s.small = []
                        continue
                if s.MortarType > 0:
                    continue
                if s.connected:
                    continue
                if s.nbElem != 0:
                    if first <= s.nbElem <= last:
                        for b0 in prim[s.nbElem - first]:
                            for b in slots_of(b0):
                                if b.ind == s.ind:          # EXIT leaves the small-side loop only: the last match wins
                                    s.conn = b
                                    b.conn = s
                                    break
                    else:
                        s.virt = True
                        s.NbProc = int(elem_to_proc(np.array([s.nbElem]), offMPI)[0])

    # ---- counts (mesh_readin.f90:568-650)
    nBCSides = nMortarSides = nSides = nMPISides = 0
    MPISideCount = {}
    for _, s in walk():
        s.tmp = 0
    for _, s in walk():
        if s.tmp == 0:
            nSides += 1
            s.tmp = -1
            if s.conn is not None:
                s.conn.tmp = -1
            if s.BCindex != 0 and not s.connected and s.MortarType == 0:
                nBCSides += 1
            if s.MortarType > 0:
                nMortarSides += 1
            if s.NbProc != -1:
                nMPISides += 1
                MPISideCount[s.NbProc] = MPISideCount.get(s.NbProc, 0) + 1
    nInnerSides = nSides - nBCSides - nMPISides - nMortarSides
    nbprocs = sorted(MPISideCount)
    nNb = len(nbprocs)
    nMPISides_Proc = np.array([MPISideCount[p] for p in nbprocs], dtype=np.int64)

    # ---- periodic BC remap (prepare_mesh.f90:83-123)
    PeriodicBCMap = np.full(nBCs, -2, dtype=np.int64)
    for i in range(nBCs):
        if BoundaryType[i, 0] != 1:
            PeriodicBCMap[i] = -1
        elif BoundaryType[i, 2] > 0:
            PeriodicBCMap[i] = -1
        elif BoundaryType[i, 2] < 0:
            for j in range(nBCs):
                if BoundaryType[j, 0] == 1 and BoundaryType[j, 2] == -BoundaryType[i, 2]:
                    PeriodicBCMap[i] = j + 1
    if np.any(PeriodicBCMap == -2):
        raise RuntimeError("Periodic connection not found.")
    for _, s in walk():
        s.SideID = -1
        if s.BCindex >= 1 and PeriodicBCMap[s.BCindex - 1] != -1:
            s.BCindex = int(PeriodicBCMap[s.BCindex - 1])

    # ---- big mortar sides: inner or MPI (prepare_mesh.f90:125-150)
    nMortarInnerSides = nMortarMPISides = 0
    for ie in range(nElems):
        for a in prim[ie]:
            a.tmp = 0
            if a.nMortars > 0:
                if any(m.NbProc != -1 for m in a.small):
                    a.tmp = -1
                    nMortarMPISides += 1
                else:
                    nMortarInnerSides += 1
    if nMortarInnerSides + nMortarMPISides != nMortarSides:
        raise RuntimeError("nInner+nMPI mortars <> nMortars.")

    # ---- non-MPI side ids (prepare_mesh.f90:152-195)
    iSide = 0
    iBCSide = 0
    iMortarInnerSide = nBCSides
    iInnerSide = nBCSides + nMortarInnerSides
    iMortarMPISide = nSides - nMortarMPISides
    for a, s in walk():
        if s.SideID == -1 and s.NbProc == -1:
            if s.conn is not None:
                iInnerSide += 1
                iSide += 1
                s.SideID = iInnerSide
                s.conn.SideID = iInnerSide
            elif s.MortarType > 0:
                if s.tmp == -1:
                    iMortarMPISide += 1
                    s.SideID = iMortarMPISide
                else:
                    iMortarInnerSide += 1
                    iSide += 1
                    s.SideID = iMortarInnerSide
            else:
                iBCSide += 1
                iSide += 1
                s.SideID = iBCSide
    if iSide != nInnerSides + nBCSides + nMortarInnerSides:
        raise RuntimeError("not all SideIDs are set!")

    # ---- MPI side ids (prepare_mesh.f90:200-320)
    mine_proc = np.array([(c // 2) if myRank < p else (c - c // 2) for p, c in zip(nbprocs, nMPISides_Proc)], dtype=np.int64)
    your_proc = nMPISides_Proc - mine_proc
    off_mine = np.zeros(nNb + 1, dtype=np.int64)
    off_your = np.zeros(nNb + 1, dtype=np.int64)
    off_mine[0] = nInnerSides + nBCSides + nMortarInnerSides
    off_mine[1:] = off_mine[0] + np.cumsum(mine_proc)
    off_your[0] = off_mine[nNb]
    off_your[1:] = off_your[0] + np.cumsum(your_proc)
    for ib, p in enumerate(nbprocs):
        sel = [s for _, s in walk() if s.NbProc == p]
        # conforming sides enter the sort with negated global index: they come first, in descending index order
        key = [(-s.ind if (s.iMortar == 0 and s.MortarType == 0) else s.ind) for s in sel]
        order = sorted(range(len(sel)), key=lambda x: key[x])
        for pos0, x in enumerate(order):
            pos = pos0 + 1
            s = sel[x]
            if myRank < p:
                s.SideID = pos + off_mine[ib] if pos <= mine_proc[ib] else pos - mine_proc[ib] + off_your[ib]
            else:
                s.SideID = pos + off_your[ib] if pos <= your_proc[ib] else pos - your_proc[ib] + off_mine[ib]
            s.SideID = int(s.SideID)

    # ---- big mortar sides whose small sides are all local or MINE become inner mortars (prepare_mesh.f90:340-420)
    if nMortarSides > 0 and nNb > 0:
        for _, s in walk():
            s.tmp = 0
        add = 0
        for ie in range(nElems):
            for a in prim[ie]:
                if a.nMortars > 0:
                    a.tmp = -1
                    if any(m.SideID > off_your[0] for m in a.small):
                        a.tmp = -2
                    if a.tmp == -1:
                        add += 1
        add -= nMortarInnerSides
        if add > 0:
            lastMortarInner = nBCSides + nMortarInnerSides
            for _, s in walk():
                if s.tmp == 0 and s.SideID > lastMortarInner:
                    s.SideID += add
                    s.tmp = 1
            off_mine += add
            off_your += add
            nMortarInnerSides += add
            nMortarMPISides -= add
            iMortarMPISide = nSides - nMortarMPISides
            iMortarInnerSide = nBCSides
            for ie in range(nElems):
                for a in prim[ie]:
                    if a.tmp == -2:
                        iMortarMPISide += 1
                        a.SideID = iMortarMPISide
                    elif a.tmp == -1:
                        iMortarInnerSide += 1
                        a.SideID = iMortarInnerSide
    if any(s.SideID < 1 for _, s in walk()):
        raise RuntimeError("not all SideIDs are set!")
    nMINE, nYOUR = int(mine_proc.sum()), int(your_proc.sum())

    # ---- exchangeFlip (prepare_mesh.f90:836-920). The flip a MINE side sends is its file flip; it is read here from
    # the neighbour element's SideInfo row with the same global side index.
    if nNb > 0:
        for _, s in walk():
            if s.NbProc == -1:
                continue
            if s.SideID > off_your[0]:
                if s.flip == 0:
                    r0, r1 = int(ElemInfo[s.nbElem - 1, 2]), int(ElemInfo[s.nbElem - 1, 3])
                    rows = [r for r in range(r0, r1) if abs(int(SideInfo[r, 1])) == s.ind and SideInfo[r, 2] >= 0]
                    fm = 0
                    for r in rows:
                        if SideInfo[r, 1] < 0:
                            fm = int(SideInfo[r, 3]) % 10
                    if fm == 0:
                        raise RuntimeError("problem in exchangeflip")
                    s.flip = fm
            else:
                s.flip = 0

    # ---- master of every YOUR side (for the surface metrics the reference receives over MPI, metrics.f90:553-575):
    # [SideID, global master element, its local side, iMortar (0: the side itself), mortar type of that local side]
    your_master = []
    for _, s in walk():
        if s.NbProc != -1 and s.SideID > off_your[0]:
            r = int(ElemInfo[s.nbElem - 1, 2])
            hit = None
            for loc in range(1, 7):
                nb = int(SideInfo[r, 2])
                nm = (4 if nb == -1 else 2) if nb < 0 else 0
                if nm == 0 and abs(int(SideInfo[r, 1])) == s.ind:
                    hit = (s.SideID, s.nbElem, loc, 0, 0)
                for im in range(1, nm + 1):
                    if abs(int(SideInfo[r + im, 1])) == s.ind:
                        hit = (s.SideID, s.nbElem, loc, im, -nb)
                r += 1 + nm
            if hit is None:
                raise RuntimeError("master side of a YOUR side not found")
            your_master.append(hit)

    # ---- fillMeshInfo (prepare_mesh.f90:688-775)
    firstMortarInnerSide = nBCSides + 1
    lastMortarInnerSide = nBCSides + nMortarInnerSides
    firstMortarMPISide = nSides - nMortarMPISides + 1
    ElemToSide = np.zeros((nElems, 6, 3), dtype=np.int32)
    SideToElem = -np.ones((nSides, 5), dtype=np.int32)
    AnalyzeSide = np.zeros(nSides, dtype=np.int32)
    BC = np.zeros(nBCSides, dtype=np.int32)
    MortarType = np.zeros((nSides, 2), dtype=np.int32)
    MortarInfo = -np.ones((max(nMortarSides, 1), 4, 2), dtype=np.int32)
    SideToGlobalSide = np.zeros(nSides, dtype=np.int32)
    for ie in range(nElems):
        for a in prim[ie]:
            ElemToSide[ie, a.loc - 1, 0] = a.SideID
            ElemToSide[ie, a.loc - 1, 1] = a.flip
            if a.flip == 0:
                SideToElem[a.SideID - 1, 0] = ie + 1
                SideToElem[a.SideID - 1, 2] = a.loc
                AnalyzeSide[a.SideID - 1] = a.BCindex
            else:
                SideToElem[a.SideID - 1, 1] = ie + 1
                SideToElem[a.SideID - 1, 3] = a.loc
                SideToElem[a.SideID - 1, 4] = a.flip
            if a.SideID <= nBCSides:
                BC[a.SideID - 1] = a.BCindex
    for ie in range(nElems):
        for a in prim[ie]:
            MortarType[a.SideID - 1, 0] = a.MortarType
            if a.nMortars > 0:
                idx = a.SideID + 1 - (firstMortarInnerSide if a.SideID <= lastMortarInnerSide
                                      else firstMortarMPISide - nMortarInnerSides)
                MortarType[a.SideID - 1, 1] = idx
                for m in a.small:
                    MortarInfo[idx - 1, m.iMortar - 1, 0] = m.SideID
                    MortarInfo[idx - 1, m.iMortar - 1, 1] = m.flip
            # mesh.f90:405-425
            SideToGlobalSide[a.SideID - 1] = abs(int(SideInfo[a.row, 1]))
            if SideToElem[a.SideID - 1, 0] == ie + 1:
                ElemToSide[ie, a.loc - 1, 2] = 1
            elif SideToElem[a.SideID - 1, 1] == ie + 1:
                ElemToSide[ie, a.loc - 1, 2] = 0
            else:
                raise RuntimeError("Seems like an error in side connectivity!")
    for _, s in walk():
        if s.iMortar > 0:
            SideToGlobalSide[s.SideID - 1] = s.ind

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
    m.firstInnerSide = m.firstMortarInnerSide + nMortarInnerSides
    m.firstMPISide_MINE = m.firstInnerSide + nInnerSides
    m.firstMPISide_YOUR = m.firstMPISide_MINE + nMINE
    m.firstMortarMPISide = m.firstMPISide_YOUR + nYOUR
    m.lastBCSide = m.firstMortarInnerSide - 1
    m.lastMortarInnerSide = m.firstInnerSide - 1
    m.lastInnerSide = m.firstMPISide_MINE - 1
    m.lastMPISide_MINE = m.firstMPISide_YOUR - 1
    m.lastMPISide_YOUR = m.firstMortarMPISide - 1
    m.lastMortarMPISide = nSides
    if m.firstMortarMPISide != firstMortarMPISide:
        raise RuntimeError("side ranges inconsistent")
    m.ElemToSide, m.SideToElem, m.BC, m.AnalyzeSide = ElemToSide, SideToElem, BC, AnalyzeSide
    m.SideToGlobalSide = SideToGlobalSide
    m.BoundaryType = BoundaryType
    m.BoundaryName = list(hopr["BCNames"])
    m.NodeCoords = nc
    m.nNbProcs = nNb
    m.NbProc = np.array(nbprocs, dtype=np.int32)
    m.nMPISides_Proc = nMPISides_Proc
    m.nMPISides_MINE_Proc = mine_proc
    m.nMPISides_YOUR_Proc = your_proc
    m.offsetMPISides_MINE = off_mine
    m.offsetMPISides_YOUR = off_your
    m.offsetElemMPI = offMPI
    m.myRank, m.nProcs = myRank, nProcs
    m.nMortarSides, m.nMortarInnerSides, m.nMortarMPISides = nMortarSides, nMortarInnerSides, nMortarMPISides
    m.MortarType, m.MortarInfo = MortarType, MortarInfo
    m.YourMaster = np.array(your_master, dtype=np.int64).reshape(-1, 5)
    return m
