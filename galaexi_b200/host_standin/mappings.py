"""Integer volume<->side index maps (bit-exact mirror of the reference tables).

Follows /root/reference/src/mesh/mappings.f90:
  * Flip_S2M / Flip_M2S   :237-279
  * CGNS_VolToSide        :288-328
  * CGNS_SideToVol        :337-377
  * CGNS_SideToVol2       :388-425
  * VolToSide/SideToVol/SideToVol2 :436-500
  * buildMappings         :75-230 (S2V2, S2V2_inv, V2S, S2V, FS2M + the two self checks)

Local side numbering (src/flexi.h:97-102): ZETA_MINUS=1, ETA_MINUS=2, XI_PLUS=3, ETA_PLUS=4,
XI_MINUS=5, ZETA_PLUS=6. Flips 0..4.

Array layout: numpy arrays use C order with the *reversed* Fortran index list, so the raw memory is
identical to the reference's column-major arrays: Fortran ``S2V2(1:2,0:N,0:N,0:4,1:6)`` is
``S2V2[locSide-1, flip, q, p, 0:2]`` here.
"""
from __future__ import annotations

import numpy as np

ZETA_MINUS, ETA_MINUS, XI_PLUS, ETA_PLUS, XI_MINUS, ZETA_PLUS = 1, 2, 3, 4, 5, 6


def flip_s2m(N: int, p: int, q: int, flip: int) -> tuple[int, int]:
    if flip == 0:
        return p, q
    if flip == 1:
        return q, p
    if flip == 2:
        return N - p, q
    if flip == 3:
        return N - q, N - p
    if flip == 4:
        return p, N - q
    raise ValueError(flip)


flip_m2s = flip_s2m  # mappings.f90:274-279: the flip permutations are involutions


def cgns_vol_to_side(N: int, i: int, j: int, k: int, s: int) -> tuple[int, int, int]:
    if s == XI_MINUS:
        return k, j, i
    if s == XI_PLUS:
        return j, k, N - i
    if s == ETA_MINUS:
        return i, k, j
    if s == ETA_PLUS:
        return N - i, k, N - j
    if s == ZETA_MINUS:
        return j, i, k
    if s == ZETA_PLUS:
        return i, j, N - k
    raise ValueError(s)


def cgns_side_to_vol(N: int, l: int, p: int, q: int, s: int) -> tuple[int, int, int]:
    if s == XI_MINUS:
        return l, q, p
    if s == XI_PLUS:
        return N - l, p, q
    if s == ETA_MINUS:
        return p, l, q
    if s == ETA_PLUS:
        return N - p, N - l, q
    if s == ZETA_MINUS:
        return q, p, l
    if s == ZETA_PLUS:
        return p, q, N - l
    raise ValueError(s)


def cgns_side_to_vol2(N: int, p: int, q: int, s: int) -> tuple[int, int]:
    if s == XI_MINUS:
        return q, p
    if s == XI_PLUS:
        return p, q
    if s == ETA_MINUS:
        return p, q
    if s == ETA_PLUS:
        return N - p, q
    if s == ZETA_MINUS:
        return q, p
    if s == ZETA_PLUS:
        return p, q
    raise ValueError(s)


def vol_to_side(N, i, j, k, flip, s):
    p, q, l = cgns_vol_to_side(N, i, j, k, s)
    p2, q2 = flip_s2m(N, p, q, flip)
    return p2, q2, l


def side_to_vol(N, l, p, q, flip, s):
    p2, q2 = flip_m2s(N, p, q, flip)
    return cgns_side_to_vol(N, l, p2, q2, s)


def side_to_vol2(N, p, q, flip, s):
    p2, q2 = flip_m2s(N, p, q, flip)
    return cgns_side_to_vol2(N, p2, q2, s)


def build_mappings(N: int):
    """Returns dict with S2V2, S2V2_inv, V2S, S2V, FS2M (int32; see module docstring for layout).

    Runs the reference's two consistency checks (mappings.f90:191-225).
    """
    n = N + 1
    S2V2 = np.zeros((6, 5, n, n, 2), dtype=np.int32)
    S2V2_inv = np.full((6, 5, n, n, 2), -1, dtype=np.int32)
    V2S = np.zeros((6, 5, n, n, n, 3), dtype=np.int32)
    S2V = np.zeros((6, 5, n, n, n, 3), dtype=np.int32)
    FS2M = np.zeros((5, n, n, 2), dtype=np.int32)
    for s in range(1, 7):
        for f in range(5):
            for q in range(n):
                for p in range(n):
                    S2V2[s - 1, f, q, p, :] = side_to_vol2(N, p, q, f, s)
            for q in range(n):
                for p in range(n):
                    a, b = S2V2[s - 1, f, q, p, :]
                    S2V2_inv[s - 1, f, b, a, 0] = p
                    S2V2_inv[s - 1, f, b, a, 1] = q
            for k in range(n):
                for j in range(n):
                    for i in range(n):
                        V2S[s - 1, f, k, j, i, :] = vol_to_side(N, i, j, k, f, s)
                        # S2V(:,l,p,q,f,s): first index l, then p, q
                        S2V[s - 1, f, k, j, i, :] = side_to_vol(N, i, j, k, f, s)
    for f in range(5):
        for q in range(n):
            for p in range(n):
                FS2M[f, q, p, :] = flip_s2m(N, p, q, f)
    # self checks
    for f in range(5):
        for s in range(1, 7):
            for q in range(n):
                for p in range(n):
                    i, j, k = S2V[s - 1, f, q, p, 0, :]
                    pq = V2S[s - 1, f, k, j, i, :]
                    if pq[0] != p or pq[1] != q:
                        raise RuntimeError("SideToVol does not fit to VolToSide")
            for k in range(n):
                for j in range(n):
                    for i in range(n):
                        pq = V2S[s - 1, f, k, j, i, :]
                        a, b = S2V2[s - 1, f, pq[1], pq[0], :]
                        if s in (XI_MINUS, XI_PLUS):
                            ok = (a == j and b == k)
                        elif s in (ETA_MINUS, ETA_PLUS):
                            ok = (a == i and b == k)
                        else:
                            ok = (a == i and b == j)
                        if not ok:
                            raise RuntimeError("SideToVol2 does not fit to VolToSide")
    return dict(S2V2=S2V2, S2V2_inv=S2V2_inv, V2S=V2S, S2V=S2V, FS2M=FS2M)
