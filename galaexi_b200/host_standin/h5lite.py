"""Minimal pure-Python reader for the HDF5 subset HOPR meshes and FLEXI/GALAEXI state files use.

The reference reads/writes these through libhdf5 (src/io_hdf5/hdf5_input.f90, hdf5_output.f90); no
HDF5 library exists in this environment, so the layout is parsed directly. Supported: superblock v0
(optionally behind a userblock of 512*2^k bytes: state files carry a 4096-byte text userblock,
hdf5_output.f90:84-211), v1 object headers with continuation blocks, old-style groups (symbol table
B-tree + local heap), contiguous or compact dataset layout, fixed-point / IEEE float / fixed-length
string datatypes, v1 attribute messages. No chunking, no filters, little-endian only.

File layouts consumed (mesh_readin.f90:33-50, hdf5_output.f90:84-211):
  mesh : attrs Ngeo,nElems,...; ElemInfo(nElems,6) SideInfo(nSides,5) NodeCoords(nNodes,3)
         BCNames(nBCs) BCType(nBCs,4)
  state: attrs N, Time, MeshFile, NodeType, VarNames...; DG_Solution(nElems,N+1,N+1,N+1,nVar)
"""
from __future__ import annotations

import mmap
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            try:
                self.buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)   # multi-GB states: pages on demand
            except (ValueError, OSError):
                self.buf = f.read()
        self.base = self._find_base()
        b = self.base
        ver = self.buf[b + 8]
        if ver != 0:
            raise NotImplementedError(f"superblock version {ver}")
        if self.buf[b + 13] != 8 or self.buf[b + 14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths")
        # root symbol table entry at base+56: link name offset(8), object header address(8)
        root_hdr = struct.unpack_from("<Q", self.buf, b + 56 + 8)[0]
        self.userblock = self.buf[:b]
        self.root_msgs = self._read_header(root_hdr)
        self.objects = self._read_group(self.root_msgs)
        self._cache: dict[str, list] = {}

    def close(self):
        """Release the mapping (idempotent)."""
        buf, self.buf = getattr(self, "buf", None), b""
        if isinstance(buf, mmap.mmap):
            buf.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ low level
    def _find_base(self) -> int:
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == _SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise ValueError("not an HDF5 file")

    def _read_header(self, addr: int) -> list[tuple[int, bytes]]:
        """All messages (type, payload) of a v1 object header, following continuations."""
        a = self.base + addr
        ver, _, nmsgs, _refc, size = struct.unpack_from("<BBHII", self.buf, a)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        msgs: list[tuple[int, bytes]] = []
        blocks = [(a + 16, size)]
        while blocks and len(msgs) < nmsgs:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, pos)
                payload = self.buf[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((self.base + coff, clen))
                msgs.append((mtype, payload))
        return msgs

    def _read_group(self, msgs) -> dict[str, int]:
        out: dict[str, int] = {}
        for mtype, p in msgs:
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", p, 0)
                heap_a = self.base + heap
                assert self.buf[heap_a:heap_a + 4] == b"HEAP"
                heap_data = self.base + struct.unpack_from("<Q", self.buf, heap_a + 24)[0]
                self._walk_btree(self.base + btree, heap_data, out)
        return out

    def _walk_btree(self, a: int, heap_data: int, out: dict[str, int]):
        assert self.buf[a:a + 4] == b"TREE", "bad B-tree node"
        ntype, level, nent = struct.unpack_from("<BBH", self.buf, a + 4)
        # keys and children interleaved from a+24: key0, child0, key1, child1, ...
        pos = a + 24
        for i in range(nent):
            child = struct.unpack_from("<Q", self.buf, pos + 8)[0]
            pos += 16
            if level > 0:
                self._walk_btree(self.base + child, heap_data, out)
            else:
                self._read_snod(self.base + child, heap_data, out)

    def _read_snod(self, a: int, heap_data: int, out: dict[str, int]):
        assert self.buf[a:a + 4] == b"SNOD"
        nsym = struct.unpack_from("<H", self.buf, a + 6)[0]
        for i in range(nsym):
            e = a + 8 + 40 * i
            name_off, hdr = struct.unpack_from("<QQ", self.buf, e)
            s = heap_data + name_off
            t = self.buf.find(b"\x00", s)
            out[self.buf[s:t].decode()] = hdr

    # ------------------------------------------------------------------ datatypes
    @staticmethod
    def _dtype(p: bytes):
        cls = p[0] & 0x0F
        size = struct.unpack_from("<I", p, 4)[0]
        if cls == 0:
            signed = (p[1] >> 3) & 1
            return np.dtype(f"<{'i' if signed else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"<f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise NotImplementedError(f"datatype class {cls}")

    @staticmethod
    def _dspace(p: bytes) -> tuple[int, ...]:
        ver, rank, flags = p[0], p[1], p[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", p, off + 8 * i)[0] for i in range(rank))

    # ------------------------------------------------------------------ public
    def keys(self):
        return list(self.objects.keys())

    def _msgs(self, name: str):
        if name not in self._cache:
            self._cache[name] = self._read_header(self.objects[name])
        return self._cache[name]

    def dataset_shape(self, name: str) -> tuple[int, ...]:
        for mtype, p in self._msgs(name):
            if mtype == 0x01:
                return self._dspace(p)
        raise KeyError(name)

    def dataset_rows(self, name: str, start: int, count: int) -> np.ndarray:
        """Rows [start, start+count) along the slowest dimension (an element range: what one rank reads)."""
        return self.dataset(name, start, count)

    def dataset(self, name: str, start: int = 0, count: int | None = None) -> np.ndarray:
        dt = shape = None
        data = None
        for mtype, p in self._msgs(name):
            if mtype == 0x01:
                shape = self._dspace(p)
            elif mtype == 0x03:
                dt = self._dtype(p)
            elif mtype == 0x08:
                ver, cls = p[0], p[1]
                if ver != 3:
                    raise NotImplementedError(f"layout version {ver}")
                if cls == 1:
                    addr, size = struct.unpack_from("<QQ", p, 2)
                    data = (self.base + addr, size) if addr != 0xFFFFFFFFFFFFFFFF else (0, 0)
                elif cls == 0:
                    size = struct.unpack_from("<H", p, 2)[0]
                    data = bytes(p[4:4 + size])
                else:
                    raise NotImplementedError("chunked layout")
        skip = 0
        if shape and (start or count is not None):
            count = shape[0] - start if count is None else count
            if start < 0 or count < 0 or start + count > shape[0]:
                raise IndexError(f"rows {start}:{start + count} outside {name}{shape}")
            row = int(np.prod(shape[1:], dtype=np.int64))
            skip = start * row * dt.itemsize
            shape = (count,) + tuple(shape[1:])
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if isinstance(data, tuple):
            arr = np.frombuffer(self.buf, dtype=dt, count=n, offset=data[0] + skip) if data[1] else np.zeros(n, dt)
        else:
            arr = np.frombuffer(data, dtype=dt, count=n, offset=skip)
        return arr.reshape(shape).copy()

    def attrs(self, name: str | None = None) -> dict[str, np.ndarray]:
        msgs = self.root_msgs if name is None else self._msgs(name)
        out = {}
        for mtype, p in msgs:
            if mtype != 0x0C:
                continue
            ver = p[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", p, 2)
            if ver == 1:
                pad = lambda x: (x + 7) & ~7
                o = 8
                nm = p[o:o + nsz].split(b"\x00")[0].decode()
                o += pad(nsz)
                dt = self._dtype(p[o:o + tsz])
                o += pad(tsz)
                shape = self._dspace(p[o:o + ssz])
                o += pad(ssz)
            else:
                raise NotImplementedError(f"attribute version {ver}")
            n = int(np.prod(shape)) if shape else 1
            out[nm] = np.frombuffer(p, dtype=dt, count=n, offset=o).reshape(shape).copy()
        return out


def read_hopr_mesh(path: str) -> dict:
    """HOPR mesh file -> dict of the arrays mesh_readin.f90 consumes (C-order shapes)."""
    f = H5File(path)
    a = f.attrs()
    out = dict(
        NGeo=int(a["Ngeo"].ravel()[0]),
        ElemInfo=f.dataset("ElemInfo").astype(np.int32),
        SideInfo=f.dataset("SideInfo").astype(np.int32),
        NodeCoords=f.dataset("NodeCoords").astype(np.float64),
        BCNames=[s.decode().strip() for s in f.dataset("BCNames").ravel()],
        BCType=f.dataset("BCType").astype(np.int32),
        isMortarMesh=int(a["isMortarMesh"].ravel()[0]) if "isMortarMesh" in a else 0,
    )
    if out["isMortarMesh"] and "TreeCoords" in f.objects:
        # octree data of non-conforming meshes (mesh_readin.f90:515-560): kept so that the mesh can be written again
        out.update(NgeoTree=int(a["NgeoTree"].ravel()[0]), nTrees=int(a["nTrees"].ravel()[0]), xiMinMax=f.dataset("xiMinMax"),
                   ElemToTree=f.dataset("ElemToTree"), TreeCoords=f.dataset("TreeCoords"))
    return out


def read_state(path: str) -> dict:
    """State file -> dict(DG_Solution (nElems,Nz+1,N+1,N+1,nVar), attrs)."""
    f = H5File(path)
    a = f.attrs()
    out = dict(DG_Solution=f.dataset("DG_Solution"), attrs=a, userblock=f.userblock)
    if "ElemData" in f.objects:
        out["ElemData"] = f.dataset("ElemData")
    return out
