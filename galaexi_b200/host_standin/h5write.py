"""Minimal pure-Python writer for the HDF5 subset FLEXI/GALAEXI state files use (the counterpart of h5lite).

The reference writes its files through libhdf5 (src/io_hdf5/hdf5_output.f90, io_hdf5.f90:212-270); no HDF5 library
exists in this environment, so the on-disk structures are emitted directly, and they are the same ones libhdf5 1.8/1.10
emits for these files (checked structure by structure against the reference's shipped state files, see
tests/test_state_io.py): superblock version 0 behind a userblock of 512*2^k bytes, version-1 object headers, an
old-style root group (symbol-table B-tree + local heap + symbol nodes), contiguous little-endian datasets with a
version-1 dataspace (with maximum dimensions), version-1 datatypes (32-bit signed integers, IEEE doubles, space-padded
fixed-length Fortran strings), version-1 attribute messages stored as one-dimensional arrays.

Layout of a written file (addresses relative to the end of the userblock, the file's base address):
  superblock | root object header (all root attributes in its first block) | B-tree node | local heap | symbol nodes |
  dataset object headers | raw data (8-byte aligned)
Datasets are streamed to disk with ndarray.tofile, so a 64^3-element state (2.3 GB) needs no second copy in memory.
"""
from __future__ import annotations

import struct
import time

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 4        # symbols per symbol node: 2*K
_INTERNAL_K = 16   # children per B-tree node: 2*K


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dtype_msg(a: np.ndarray) -> bytes:
    """Version-1 datatype message of the element type of ``a``."""
    k = a.dtype.kind
    if k == "i" and a.dtype.itemsize in (4, 8):
        return struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, a.dtype.itemsize, 0, 8 * a.dtype.itemsize)
    if k == "f" and a.dtype.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if k == "S":
        return struct.pack("<BBBBI", 0x13, 0x02, 0, 0, a.dtype.itemsize)      # space padded, ASCII (H5T_FORTRAN_S1)
    raise TypeError(f"unsupported dtype {a.dtype}")


def _dspace_msg(shape: tuple[int, ...]) -> bytes:
    """Version-1 simple dataspace with maximum dimensions equal to the dimensions (what H5Screate_simple writes)."""
    r = len(shape)
    return struct.pack("<BBBB4x", 1, r, 1, 0) + struct.pack(f"<{r}Q", *shape) + struct.pack(f"<{r}Q", *shape)


def _msg(mtype: int, payload: bytes, flags: int = 0) -> bytes:
    p = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(p), flags) + p


def _as_array(v) -> np.ndarray:
    """Attribute values: str / list of str -> fixed-length space-padded strings; ints -> int32; floats -> float64."""
    if isinstance(v, (bytes, str)):
        v = [v]
    if isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], (bytes, str)):
        bs = [s.encode() if isinstance(s, str) else s for s in v]
        width = max(1, max(len(s) for s in bs))
        return np.array([s.ljust(width) for s in bs], dtype=f"S{width}")
    a = np.atleast_1d(np.asarray(v))
    if a.dtype.kind == "S":
        return a
    if a.dtype.kind in "iub":
        return a.astype("<i4")
    return a.astype("<f8")


def _attr_msg(name: str, value) -> bytes:
    a = _as_array(value)
    nm = name.encode() + b"\x00"
    dt = _dtype_msg(a)
    ds = _dspace_msg((a.size,))
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()
    if len(body) > 65000:
        raise ValueError(f"attribute {name} too large for an object-header message")
    return _msg(0x000C, body)


def _object_header(msgs: list[bytes]) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def fortran_str(s: str, width: int = 255) -> bytes:
    """CHARACTER(LEN=255) as the reference stores most string attributes (blank padded)."""
    return s.encode().ljust(width)


class H5Writer:
    """Collect root attributes and datasets, then ``write(path)``.

    ``create_dataset(name, array)`` takes the array in C order with the dimensions in file order, i.e. reversed with
    respect to the Fortran dimensions the reference passes (DG_Solution: (nElems, N+1, N+1, N+1, nVar))."""

    def __init__(self, userblock: bytes = b""):
        self.userblock = userblock
        self.attrs: list[tuple[str, object]] = []
        self.datasets: dict[str, np.ndarray] = {}
        self.dataset_attrs: dict[str, list[tuple[str, object]]] = {}
        self.mtime = int(time.time())

    def set_attr(self, name: str, value, dataset: str | None = None):
        lst = self.attrs if dataset is None else self.dataset_attrs.setdefault(dataset, [])
        for i, (k, _) in enumerate(lst):
            if k == name:
                lst[i] = (name, value)
                return
        lst.append((name, value))

    def create_dataset(self, name: str, a: np.ndarray):
        a = np.asarray(a)
        if a.dtype.kind == "f":
            a = a.astype("<f8", copy=False)
        elif a.dtype.kind in "iu":
            a = a.astype("<i4", copy=False)
        self.datasets[name] = np.ascontiguousarray(a)

    @staticmethod
    def userblock_size(nbytes: int) -> int:
        """io_hdf5.f90:259-263: the next 512*2^k that holds the userblock text."""
        if nbytes <= 0:
            return 0
        size = 512
        while size < nbytes:
            size *= 2
        return size

    def reserve_dataset(self, name: str, shape: tuple[int, ...], dtype="<f8"):
        """A dataset whose data is written later (by several ranks, each its own byte range): see ``layout``."""
        self.datasets[name] = _Reserved(tuple(int(x) for x in shape), np.dtype(dtype))

    def _plan(self, data_start: int | None, reserve: int):
        """Metadata block, {name: data address}, end-of-file address (all relative to the base address)."""
        names = sorted(self.datasets, key=lambda s: s.encode())
        if len(names) > 2 * _LEAF_K * 2 * _INTERNAL_K:
            raise ValueError("too many datasets for a single-level group B-tree")
        # ---- local heap data: the empty string at offset 0, then the names (8-byte aligned, NUL terminated)
        heap = bytearray(8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)                      # one free block: next = H5HL_FREE_NULL, its own size
        # ---- sizes and addresses
        root_msgs = [_attr_msg(k, v) for k, v in self.attrs]
        root_hdr_len = 16 + sum(len(m) for m in root_msgs) + 8 + 16        # + symbol table message
        chunks = [names[i:i + 2 * _LEAF_K] for i in range(0, len(names), 2 * _LEAF_K)] or [[]]
        a_root = 96
        a_btree = a_root + root_hdr_len
        btree_len = 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8
        a_heap = a_btree + btree_len
        a_heap_data = a_heap + 32
        a_snod = a_heap_data + len(heap)
        snod_len = 8 + 2 * _LEAF_K * 40
        a_obj = a_snod + snod_len * len(chunks)
        ds_hdr = {}
        for nm in names:
            a = self.datasets[nm]
            msgs = [_msg(0x0001, _dspace_msg(a.shape)), _msg(0x0003, _dtype_msg(a), 1),
                    _msg(0x0005, struct.pack("<BBBBI", 2, 1, 2, 1, 0), 1),   # fill value v2: early alloc, write if set, size 0
                    None,                                                   # layout, filled in below
                    _msg(0x0012, struct.pack("<B3xI", 1, self.mtime))]
            msgs += [_attr_msg(k, v) for k, v in self.dataset_attrs.get(nm, [])]
            ds_hdr[nm] = msgs
        hdr_len = {nm: 16 + sum(len(m) for m in ms if m is not None) + 8 + 24 for nm, ms in ds_hdr.items()}
        a_hdr = {}
        pos = a_obj
        for nm in names:
            a_hdr[nm] = pos
            pos += hdr_len[nm]
        meta_len = pos
        if data_start is None:
            data_start = meta_len + reserve
        if data_start < meta_len or data_start % 8:
            raise ValueError("metadata does not fit in front of the data")
        pos = data_start
        a_data = {}
        for nm in names:
            pos += -pos % 8
            a_data[nm] = pos
            pos += self.datasets[nm].nbytes
        eof = pos
        ub = self.userblock_size(len(self.userblock))

        out = bytearray()
        out += _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        out += struct.pack("<QQQQ", ub, _UNDEF, ub + eof, _UNDEF)
        out += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)
        assert len(out) == a_root
        out += _object_header(root_msgs + [_msg(0x0011, struct.pack("<QQ", a_btree, a_heap))])
        assert len(out) == a_btree
        # B-tree node (group node, level 0): key[0] = "", key[i+1] = largest name of child i
        bt = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(chunks) if names else 0, _UNDEF, _UNDEF)
        bt += struct.pack("<Q", 0)
        for i, ch in enumerate(chunks):
            if not ch:
                break
            bt += struct.pack("<QQ", a_snod + i * snod_len, name_off[ch[-1]])
        out += bt.ljust(btree_len, b"\x00")
        out += struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), free_off, a_heap_data)
        out += heap
        for ch in chunks:
            sn = struct.pack("<4sBBH", b"SNOD", 1, 0, len(ch))
            for nm in ch:
                sn += struct.pack("<QQII16x", name_off[nm], a_hdr[nm], 0, 0)
            out += sn.ljust(snod_len, b"\x00")
        assert len(out) == a_obj
        for nm in names:
            msgs = list(ds_hdr[nm])
            msgs[3] = _msg(0x0008, struct.pack("<BBQQ", 3, 1, a_data[nm], self.datasets[nm].nbytes), 1)
            h = _object_header(msgs)
            assert len(h) == hdr_len[nm], (len(h), hdr_len[nm])
            out += h
        return bytes(out), a_data, eof, data_start

    def layout(self, reserve: int = 0) -> tuple[dict[str, int], int]:
        """({dataset: absolute file offset of its data}, data_start) for a file written with the same ``reserve`` bytes
        of room behind the metadata (for attributes added by a later ``write(..., metadata_only=True)``)."""
        _, a_data, _, data_start = self._plan(None, reserve)
        ub = self.userblock_size(len(self.userblock))
        return {k: ub + v for k, v in a_data.items()}, data_start

    def write(self, path: str, reserve: int = 0, data_start: int | None = None, metadata_only: bool = False):
        """Write the file; with ``metadata_only`` the metadata block of an existing file is rewritten in place (the data
        has to start at the same ``data_start``)."""
        out, a_data, eof, _ = self._plan(data_start, reserve)
        ub = self.userblock_size(len(self.userblock))
        with open(path, "r+b" if metadata_only else "wb") as f:
            if ub:
                f.write(self.userblock.ljust(ub, b"\x00"))
            f.seek(ub)
            f.write(out)
            if metadata_only:
                return
            for nm, a in self.datasets.items():
                if isinstance(a, np.ndarray):
                    f.seek(ub + a_data[nm])
                    a.tofile(f)
            f.truncate(ub + eof)


class _Reserved:
    """Shape and type of a dataset whose data is written separately."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = shape, dtype
        self.nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize


def write_hopr_mesh(path: str, hopr: dict):
    """A mesh in the HOPR file layout the reference reads (mesh/mesh_readin.f90:33-50, 100-160, 280-320, 470-560, 845-855):
    attributes Version, Ngeo, nElems, nSides, nNodes, nUniqueSides, nUniqueNodes, nBCs; data sets BCNames, BCType, ElemInfo,
    SideInfo, NodeCoords, GlobalNodeIDs, ElemBarycenters, ElemWeight, ElemCounter and, for structured meshes, Elem_IJK /
    nElems_IJK. ``hopr`` is the dict of read_hopr_mesh / make_box_mesh (arrays in file order). The reference itself reads only
    Ngeo, BCNames, BCType, ElemInfo, SideInfo, NodeCoords (and the optional tree / IJK data); the rest is written for the posti
    tools and HOPR's own consistency."""
    ei = np.asarray(hopr["ElemInfo"], dtype=np.int32)
    si = np.asarray(hopr["SideInfo"], dtype=np.int32)
    nc = np.asarray(hopr["NodeCoords"], dtype=np.float64)
    NGeo = int(hopr["NGeo"])
    nE, nn = ei.shape[0], (NGeo + 1) ** 3
    if nc.shape != (nE * nn, 3):
        raise ValueError("NodeCoords does not hold (NGeo+1)^3 nodes per element")
    w = H5Writer()
    gid = hopr.get("GlobalNodeIDs")
    if gid is None:
        # unique nodes: coordinates equal up to a tolerance relative to the mesh extent, numbered in order of first appearance
        ext = float(np.max(nc.max(axis=0) - nc.min(axis=0))) or 1.0
        key = np.round((nc - nc.min(axis=0)) / (ext * 1e-9)).astype(np.int64)
        _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
        order = np.argsort(np.argsort(first))
        gid = (order[inv.ravel()] + 1).astype(np.int32)
    gid = np.asarray(gid, dtype=np.int32)
    ind = si[:, 1]
    for k, v in (("Version", 1.0), ("Ngeo", NGeo), ("nElems", nE), ("nSides", si.shape[0]), ("nNodes", nc.shape[0]),
                 ("nUniqueSides", int(np.count_nonzero(ind > 0)) if np.any(ind < 0) else len(np.unique(np.abs(ind)))),
                 ("nUniqueNodes", int(gid.max()) if gid.size else 0), ("nBCs", len(hopr["BCNames"]))):
        w.set_attr(k, v)
    if int(hopr.get("isMortarMesh", 0)):
        if hopr.get("TreeCoords") is None:
            raise ValueError("a non-conforming (mortar) mesh needs its octree data: NgeoTree, nTrees, xiMinMax, ElemToTree, TreeCoords")
        w.set_attr("isMortarMesh", 1)
        w.set_attr("NgeoTree", int(hopr["NgeoTree"]))
        w.set_attr("nTrees", int(hopr["nTrees"]))
        w.create_dataset("xiMinMax", np.asarray(hopr["xiMinMax"], dtype=np.float64))
        w.create_dataset("ElemToTree", np.asarray(hopr["ElemToTree"], dtype=np.int32))
        w.create_dataset("TreeCoords", np.asarray(hopr["TreeCoords"], dtype=np.float64))
    w.create_dataset("BCNames", np.array([fortran_str(str(s)) for s in hopr["BCNames"]], dtype="S255"))
    w.create_dataset("BCType", np.asarray(hopr["BCType"], dtype=np.int32))
    w.create_dataset("ElemInfo", ei)
    w.create_dataset("SideInfo", si)
    w.create_dataset("NodeCoords", nc)
    w.create_dataset("GlobalNodeIDs", gid)
    corners = nc.reshape(nE, NGeo + 1, NGeo + 1, NGeo + 1, 3)[:, ::NGeo, ::NGeo, ::NGeo].reshape(nE, 8, 3)
    w.create_dataset("ElemBarycenters", corners.mean(axis=1))     # HOPR: mean of the eight corner nodes
    w.create_dataset("ElemWeight", np.ones(nE))
    types = np.array([104, 204, 105, 115, 205, 106, 116, 206, 108, 118, 208], dtype=np.int32)
    cnt = np.array([int(np.count_nonzero(ei[:, 0] == t)) for t in types], dtype=np.int32)
    w.create_dataset("ElemCounter", np.stack([types, cnt], axis=1))
    if hopr.get("Elem_IJK") is not None and hopr.get("nElems_IJK") is not None:
        w.create_dataset("Elem_IJK", np.asarray(hopr["Elem_IJK"], dtype=np.int32))
        w.create_dataset("nElems_IJK", np.asarray(hopr["nElems_IJK"], dtype=np.int32))
    w.write(path)
