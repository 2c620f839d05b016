"""State files in the reference layout + restart: galaexi_b200/host_standin/h5write.py, state_io.py.

No HDF5 library exists here, so the writer is pinned against what libhdf5 itself wrote: tests/golden/state_h5_structs.json
holds the raw superblock / attribute messages / dataset header messages / heap / B-tree bytes of the reference's shipped
cavity state file (tools/make_golden.py:state_h5_structs), and the writer has to emit the same bytes for the same content
(addresses and time stamps excepted). Restart follows restart/restart.f90:304-560.
"""
import datetime
import json
import os
import struct

import numpy as np
import pytest

from galaexi_b200.host_standin import basis as bs
from galaexi_b200.host_standin import h5lite, h5write, metrics, state_io

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ref_structs():
    return json.load(open(os.path.join(GOLD, "state_h5_structs.json")))


@pytest.fixture(scope="module")
def cavity_state():
    z = np.load(os.path.join(GOLD, "cavity3d_state.npz"))
    return z["DG_Solution"]


def _write_cavity(tmp_path, U, **kw):
    ed = {"myRank": np.repeat([0.0, 1.0], 32), "IndValue": 0.0, "dt": np.linspace(2.89e-3, 2.9e-3, 64)}
    return state_io.write_state(U, 2, "GAUSS", "cavity_Re100", "cavity4x4x4_mesh.h5", 1.0, 1.0, out_dir=str(tmp_path),
                                elem_data=ed, ini_text="ProjectName=cavity_Re100\nN=2\n", **kw), ed


def _raw_msgs(path):
    """(root attribute messages by name, dataset messages by dataset) of a written file, raw bytes."""
    f = h5lite.H5File(path)
    b, buf = f.base, f.buf
    attrs = {}
    for t, pl in f.root_msgs:
        if t == 0x0C:
            nsz = struct.unpack_from("<H", pl, 2)[0]
            attrs[bytes(pl[8:8 + nsz - 1]).decode()] = bytes(pl)
    ds = {}
    for nm, addr in f.objects.items():
        a = b + addr
        size = struct.unpack_from("<I", buf, a + 8)[0]
        pos, msgs = a + 16, []
        while pos < a + 16 + size:
            t, sz, fl = struct.unpack_from("<HHB", buf, pos)
            msgs.append((t, fl, bytes(buf[pos + 8:pos + 8 + sz])))
            pos += 8 + sz
        ds[nm] = msgs
    return f, attrs, ds


def test_timestamp_and_names():
    assert state_io.timestamp("cavity_Re100_State", 1.0) == "cavity_Re100_State_0000001.000000000"
    assert state_io.state_file_name("NACA0012_Re5000_AoA8_3D", 10.0) == "NACA0012_Re5000_AoA8_3D_State_0000010.000000000.h5"
    assert state_io.state_file_name("p", 0.123456789) == "p_State_0000000.123456789.h5"
    assert h5write.H5Writer.userblock_size(2439) == 4096       # the cavity file: 2 439 bytes of text behind a 4 096-byte block
    assert h5write.H5Writer.userblock_size(512) == 512 and h5write.H5Writer.userblock_size(513) == 1024


def test_attribute_messages_match_libhdf5(tmp_path, cavity_state, ref_structs):
    """Every root attribute message is byte-identical to the one libhdf5 wrote for the same value."""
    path, _ = _write_cavity(tmp_path, cavity_state)
    _, attrs, _ = _raw_msgs(path)
    ref = {k: bytes.fromhex(v) for k, v in ref_structs["root_attr_msgs"].items()}
    assert set(ref) == set(attrs)
    for k in ref:
        if k == "TIME":   # wall-clock of the write: same type, shape and size, different values
            assert len(attrs[k]) == len(ref[k]) and attrs[k][:56] == ref[k][:56]
        else:
            assert attrs[k] == ref[k], k


def test_dataset_headers_match_libhdf5(tmp_path, cavity_state, ref_structs):
    """Dataspace and datatype messages of DG_Solution / ElemData are byte-identical; the layout message has the same
    version, class and size; the data is contiguous at the recorded address."""
    path, ed = _write_cavity(tmp_path, cavity_state)
    f, _, ds = _raw_msgs(path)
    for nm in ("DG_Solution", "ElemData"):
        ref = {t: (fl, bytes.fromhex(h)) for t, fl, h in ref_structs["datasets"][nm]["msgs"]}
        mine = {t: (fl, p) for t, fl, p in ds[nm]}
        for t in (0x0001, 0x0003):
            assert mine[t] == ref[t], (nm, hex(t))
        assert mine[0x0005][0] == ref[0x0005][0] and mine[0x0005][1][0] == 2 and len(mine[0x0005][1]) == 8
        fl, lay = mine[0x0008]
        rfl, rlay = ref[0x0008]
        assert fl == rfl and lay[:2] == rlay[:2] and lay[10:18] == rlay[10:18]      # version 3, contiguous, same byte size
        addr = struct.unpack_from("<Q", lay, 2)[0]
        assert addr % 8 == 0 and f.base + addr + struct.unpack_from("<Q", lay, 10)[0] <= len(f.buf)
    assert np.array_equal(f.dataset("DG_Solution"), cavity_state)
    got = f.dataset("ElemData")
    assert np.array_equal(got[:, 0], ed["myRank"]) and np.array_equal(got[:, 2], ed["dt"]) and not got[:, 1].any()


def test_superblock_group_structures_match_libhdf5(tmp_path, cavity_state, ref_structs):
    path, _ = _write_cavity(tmp_path, cavity_state)
    buf = open(path, "rb").read()
    ub = 512
    assert buf[:23] == b"{[( START USERBLOCK )]}" and buf[ub:ub + 8] == b"\x89HDF\r\n\x1a\n"
    ref_sb = bytes.fromhex(ref_structs["superblock"])
    sb = buf[ub:ub + 96]
    assert sb[:24] == ref_sb[:24]                                      # versions, offset/length sizes, group K values, flags
    base, free, eof, drv = struct.unpack_from("<QQQQ", sb, 24)
    rbase, rfree, reof, rdrv = struct.unpack_from("<QQQQ", ref_sb, 24)
    assert (base, free, drv) == (ub, rfree, rdrv) and rbase == ref_structs["userblock_size"]
    assert eof == len(buf) and reof == ref_structs["file_size"]        # end-of-file address is absolute in both
    name_off, hdr, cache, _ = struct.unpack_from("<QQII", sb, 56)
    assert (name_off, hdr, cache) == struct.unpack_from("<QQII", ref_sb, 56)[:3] == (0, 96, 1)
    bt, heap = struct.unpack_from("<QQ", sb, 80)
    # local heap: same header fields and the same name segment as libhdf5 (names 8-byte aligned behind the empty name)
    rh = bytes.fromhex(ref_structs["heap_header"])
    h = buf[ub + heap:ub + heap + 32]
    assert h[:8] == rh[:8]
    dsz, fr, daddr = struct.unpack_from("<QQQ", h, 8)
    rd = bytes.fromhex(ref_structs["heap_data"])
    assert fr == struct.unpack_from("<Q", rh, 16)[0] == 40
    assert buf[ub + daddr:ub + daddr + 40] == rd[:40]
    nxt, fsz = struct.unpack_from("<QQ", buf, ub + daddr + fr)
    assert nxt == 1 and fr + fsz == dsz                                 # one free block closing the segment
    # B-tree node and symbol node heads
    rb = bytes.fromhex(ref_structs["btree_head"])
    b = buf[ub + bt:ub + bt + 48]
    assert b[:32] == rb[:32] and b[40:48] == rb[40:48]                  # TREE, type 0, level 0, 1 entry, no siblings, keys 0 / 24
    sn = struct.unpack_from("<Q", b, 32)[0]
    rs = bytes.fromhex(ref_structs["snod"])
    s = buf[ub + sn:ub + sn + 88]
    assert s[:8] == rs[:8]                                              # SNOD, version 1, 2 symbols
    for i in range(2):
        e, r = s[8 + 40 * i:48 + 40 * i], rs[8 + 40 * i:48 + 40 * i]
        assert e[:8] == r[:8] and e[16:] == r[16:]                       # name offset, cache type 0, empty scratch pad
    # every object header starts 8-byte aligned inside the file
    f = h5lite.H5File(path)
    assert all(a % 8 == 0 for a in f.objects.values())


def test_time_attribute_is_written_last(tmp_path, cavity_state):
    """MarkWriteSuccessfull: the skeleton has no TIME; adding it rewrites only the metadata block."""
    w = state_io._skeleton("p", "State", "m.h5", 2, 2, "GAUSS", 0.0, 0.1, 64, 5, state_io.STR_VAR_NAMES, [], b"")
    off, data_start = w.layout(reserve=256)
    path = str(tmp_path / "p.h5")
    w.write(path, data_start=data_start)
    assert not state_io.read_state_attrs(path)["complete"]
    with open(path, "r+b") as f:
        f.seek(off["DG_Solution"])
        cavity_state.tofile(f)
    size = os.path.getsize(path)
    state_io.mark_write_successful(w, path, data_start, datetime.datetime(2019, 1, 9, 16, 35, 11, 765000))
    assert os.path.getsize(path) == size
    info = state_io.read_state_attrs(path)
    assert info["complete"] and np.array_equal(h5lite.read_state(path)["DG_Solution"], cavity_state)
    assert list(h5lite.H5File(path).attrs()["TIME"]) == [2019, 1, 9, 0, 16, 35, 11, 765]


def test_two_rank_write_equals_single(tmp_path, cavity_state):
    """Each rank writes its own element range (GatheredWriteArray with offsetElem)."""
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(); d2.mkdir()
    now = datetime.datetime(2020, 1, 1)
    kw = dict(N=2, node_type="GAUSS", project="p", mesh_file="m.h5", t=0.5, t_next=1.0, now=now)
    ed = np.arange(64.0)
    p1 = state_io.write_state(cavity_state, out_dir=str(d1), elem_data={"x": ed}, **kw)
    # rank 1 before rank 0's final mark: emulate the collective order by hand
    w = state_io.write_state(cavity_state[:40], out_dir=str(d2), elem_data={"x": ed[:40]}, nGlobalElems=64, rank=0, **kw)
    state_io.write_state(cavity_state[40:], out_dir=str(d2), elem_data={"x": ed[40:]}, nGlobalElems=64, offsetElem=40, rank=1, **kw)
    a, b = open(p1, "rb").read(), open(w, "rb").read()
    f = h5lite.H5File(p1)
    assert len(a) == len(b)
    g = h5lite.H5File(w)
    assert np.array_equal(g.dataset("DG_Solution"), f.dataset("DG_Solution")) and np.array_equal(g.dataset("ElemData"), f.dataset("ElemData"))
    assert np.array_equal(g.dataset_rows("DG_Solution", 40, 24), cavity_state[40:])


def test_restart_same_degree_and_reset_time(tmp_path, cavity_state):
    path, _ = _write_cavity(tmp_path, cavity_state)
    U, t = state_io.restart(path, 2, "GAUSS", nGlobalElems=64)
    assert t == 1.0 and np.array_equal(U, cavity_state)
    U1, t1 = state_io.restart(path, 2, "GAUSS", offsetElem=32, nElems=32, ResetTime=True)
    assert t1 == 0.0 and np.array_equal(U1, cavity_state[32:])
    with pytest.raises(RuntimeError, match="Dimensions of restart file do not match"):
        state_io.restart(path, 2, "GAUSS", nGlobalElems=65)


def _poly_state(X, deg):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    U = np.empty(X.shape[:-1] + (5,))
    U[..., 0] = 1.0 + 0.1 * x ** deg + 0.05 * y * z
    U[..., 1] = 0.3 * x * y ** (deg - 1)
    U[..., 2] = -0.2 * z ** deg
    U[..., 3] = 0.1 * (x + y + z) ** deg
    U[..., 4] = 2.5 + 0.2 * x * y * z
    return U


def _cart_coords(nE, N, node_type):
    """nE unit-cube elements in a row: NodeCoords (NGeo=1) and the interpolation points of (N, node_type)."""
    xi = bs.get_nodes_and_weights(N, node_type)[0]
    NC = np.zeros((nE, 2, 2, 2, 3))
    for e in range(nE):
        for k in range(2):
            for j in range(2):
                for i in range(2):
                    NC[e, k, j, i] = (e + i, j, k)
    X = np.zeros((nE, N + 1, N + 1, N + 1, 3))
    h = 0.5 * (xi + 1.0)
    for e in range(nE):
        X[e, ..., 0] = e + h[None, None, :]
        X[e, ..., 1] = h[None, :, None]
        X[e, ..., 2] = h[:, None, None]
    return NC, X


def test_restart_interpolates_degree_and_node_type(tmp_path):
    """A polynomial of degree 3 written on N=3 Gauss-Lobatto is reproduced exactly on N=5 Gauss (restart.f90:513-523) and
    on N=3 Gauss (node type change only); restart from N=5 down to N=3 goes through the conservative branch (JU projected
    with J from detJac_Ref, :500-512) and is exact too since the data is of degree 3 and J is constant."""
    NC, X3 = _cart_coords(3, 3, bs.NODETYPE_GL)
    U3 = _poly_state(X3, 3)
    p = state_io.write_state(U3, 3, bs.NODETYPE_GL, "rs", "m.h5", 0.25, 0.5, out_dir=str(tmp_path))
    _, X5 = _cart_coords(3, 5, bs.NODETYPE_G)
    U5, t = state_io.restart(p, 5, bs.NODETYPE_G)
    assert t == 0.25 and np.abs(U5 - _poly_state(X5, 3)).max() < 1e-13
    _, X3g = _cart_coords(3, 3, bs.NODETYPE_G)
    assert np.abs(state_io.restart(p, 3, bs.NODETYPE_G)[0] - _poly_state(X3g, 3)).max() < 1e-13
    # down: 5 (Gauss) -> 3 (GL), conservative projection
    p5 = state_io.write_state(_poly_state(X5, 3), 5, bs.NODETYPE_G, "rs5", "m.h5", 0.0, 0.5, out_dir=str(tmp_path))
    dj = metrics.det_jac_ref(NC, 1, bs.NODETYPE_GL)
    assert dj.shape == (3, 4, 4, 4) and np.allclose(dj, 0.125)
    sJ = np.full((3, 4, 4, 4), 8.0)
    Ud, _ = state_io.restart(p5, 3, bs.NODETYPE_GL, sJ=sJ, detJac_Ref=dj, NGeo=1)
    assert np.abs(Ud - U3).max() < 1e-13
    with pytest.raises(ValueError, match="conservative projection"):
        state_io.restart(p5, 3, bs.NODETYPE_GL)


def test_restart_projection_conserves_on_curved_element(tmp_path):
    """N_Restart > N on a deformed element: the integral of every conserved variable, sum w J U, is kept by the projection
    (the point of projecting J*U, restart.f90:497-512) while the plain interpolation of U is not conservative."""
    NGeo, NR, N = 2, 6, 3
    nt = bs.NODETYPE_G
    xe = np.linspace(-1, 1, NGeo + 1)
    NC = np.zeros((1, 3, 3, 3, 3))
    for k in range(3):
        for j in range(3):
            for i in range(3):
                x = np.array([xe[i], xe[j], xe[k]])
                NC[0, k, j, i] = x + np.array([0.15 * (1 - 0.5 * x[0] ** 2) * (1 - x[1] ** 2) * (1 - 0.3 * x[2]),
                                               -0.1 * (1 - x[0] ** 2) * (1 + 0.4 * x[1]) * (1 - x[2] ** 2),
                                               0.08 * (1 + 0.3 * x[0]) * (1 - x[1] ** 2) * (1 - 0.6 * x[2] ** 2)])
    dj = metrics.det_jac_ref(NC, NGeo, nt)
    VJ = bs.get_vandermonde(3 * NGeo, nt, NR, nt, modal=True)
    JR = metrics.change_basis_volume(VJ, dj[..., None])[..., 0]
    VN = bs.get_vandermonde(3 * NGeo, nt, N, nt, modal=True)
    JN = metrics.change_basis_volume(VN, dj[..., None])[..., 0]
    xR, wR = bs.get_nodes_and_weights(NR, nt)[:2]
    xN, wN = bs.get_nodes_and_weights(N, nt)[:2]
    X = np.stack(np.meshgrid(xR, xR, xR, indexing="ij")[::-1], axis=-1)[None]
    UR = 1.0 + _poly_state(X, 5) ** 2
    p = state_io.write_state(UR, NR, nt, "cv", "m.h5", 0.0, 1.0, out_dir=str(tmp_path))
    U, _ = state_io.restart(p, N, nt, sJ=1.0 / JN, detJac_Ref=dj, NGeo=NGeo)
    w3R = wR[:, None, None] * wR[None, :, None] * wR[None, None, :]
    w3N = wN[:, None, None] * wN[None, :, None] * wN[None, None, :]
    IR = np.einsum("kji,kjiv->v", w3R * JR[0], UR[0])
    IN = np.einsum("kji,kjiv->v", w3N * JN[0], U[0])
    assert np.abs(IN / IR - 1.0).max() < 1e-13
    Uplain = metrics.change_basis_volume(bs.get_vandermonde(NR, nt, N, nt, modal=True), UR)
    assert np.abs(np.einsum("kji,kjiv->v", w3N * JN[0], Uplain[0]) / IR - 1.0).max() > 1e-6


def test_write_state_nout_projection(tmp_path):
    """NOut != N (hdf5_output.f90:131-149): U_out = P(J U) / P(J); on a Cartesian element this is the plain projection."""
    NC, X = _cart_coords(2, 4, bs.NODETYPE_G)
    U = _poly_state(X, 2)
    sJ = np.full(U.shape[:-1], 8.0)
    p = state_io.write_state(U, 4, bs.NODETYPE_G, "no", "m.h5", 0.0, 1.0, out_dir=str(tmp_path), sJ=sJ, NOut=2)
    info = state_io.read_state_attrs(p)
    # attribute N is the computation degree, the data is on NOut: N_Restart comes from the data extent (hdf5_input.f90:386)
    assert info["shape"] == (2, 3, 3, 3, 5) and info["NComputation"] == 4 and info["N"] == 2
    _, X2 = _cart_coords(2, 2, bs.NODETYPE_G)
    assert np.abs(h5lite.read_state(p)["DG_Solution"] - _poly_state(X2, 2)).max() < 1e-13
    # ... so a restart from this file works on NOut directly and is interpolated to the computation degree (degree-2 data: exact)
    U2, t2 = state_io.restart(p, 2, bs.NODETYPE_G)
    assert t2 == 0.0 and np.abs(U2 - _poly_state(X2, 2)).max() < 1e-13
    U4, _ = state_io.restart(p, 4, bs.NODETYPE_G)
    assert U4.shape == U.shape and np.abs(U4 - U).max() < 1e-12
    with pytest.raises(ValueError):
        state_io.write_state(U, 4, bs.NODETYPE_G, "no", "m.h5", 0.0, 1.0, out_dir=str(tmp_path), NOut=2)


def test_flush_files_follows_next_file_chain(tmp_path, cavity_state):
    times = [0.0, 0.5, 1.0, 1.5]
    for a, b in zip(times, times[1:] + [2.0]):
        state_io.write_state(cavity_state[:2], 2, "GAUSS", "fl", "m.h5", a, b, out_dir=str(tmp_path))
    other = state_io.write_state(cavity_state[:2], 2, "GAUSS", "other", "m.h5", 1.0, 2.0, out_dir=str(tmp_path))
    gone = state_io.flush_files("fl", 0.5, out_dir=str(tmp_path))
    assert [os.path.basename(g) for g in gone] == [state_io.state_file_name("fl", t) for t in (0.5, 1.0, 1.5)]
    assert os.path.exists(tmp_path / state_io.state_file_name("fl", 0.0)) and os.path.exists(other)
    assert state_io.flush_files("fl", 7.0, out_dir=str(tmp_path)) == []


def test_writer_many_datasets_and_types(tmp_path):
    """More than one symbol node (> 8 objects), integer and string datasets, dataset attributes."""
    w = h5write.H5Writer()
    arrs = {f"d{i:02d}": np.arange(i + 1, dtype=np.float64) * 1.5 for i in range(19)}
    for k, v in arrs.items():
        w.create_dataset(k, v)
    w.create_dataset("ints", np.arange(12, dtype=np.int64).reshape(3, 4))
    w.create_dataset("names", np.array([b"BC_wall ", b"BC_inlet"], dtype="S8"))
    w.set_attr("nElems", 7)
    w.set_attr("unit", "m", dataset="d03")
    p = str(tmp_path / "many.h5")
    w.write(p)
    f = h5lite.H5File(p)
    assert sorted(f.keys()) == sorted(list(arrs) + ["ints", "names"])
    for k, v in arrs.items():
        assert np.array_equal(f.dataset(k), v)
    assert np.array_equal(f.dataset("ints"), np.arange(12).reshape(3, 4)) and f.dataset("ints").dtype == np.int32
    assert list(f.dataset("names")) == [b"BC_wall ", b"BC_inlet"]
    assert int(f.attrs()["nElems"][0]) == 7 and f.attrs("d03")["unit"][0] == b"m"


REF = "/root/reference/regressioncheck/checks"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("rel,N,nt,nE,t", [("parabolic/cavity_3D/reggie_cavity_Re100_State_0000001.000000000.h5", 2, "GAUSS", 64, 1.0),
                                            ("naca/3D/NACA0012_Re5000_AoA8_3D_Referenz_0000010.000000000.h5", 3, "GAUSS", 652, 10.0),
                                            ("h5diff/cavity/cavity_reference_State_0000000.000200000.h5", 3, "GAUSS", 36, 2e-4)])
def test_restart_reads_files_written_by_libhdf5(rel, N, nt, nE, t):
    """InitRestart + Restart on the reference's own state files (written by libhdf5 through FLEXI): attributes, element ranges of
    two 'ranks', and interpolation to another degree all work on the library's container layout (object-header continuation
    blocks, 4 096-byte userblock)."""
    path = os.path.join(REF, rel)
    info = state_io.read_state_attrs(path)
    assert (info["N"], info["NodeType"], info["nGlobalElems"], info["Time"], info["complete"]) == (N, nt, nE, t, True)
    U, tr = state_io.restart(path, N, nt, nGlobalElems=nE)
    assert tr == t and U.shape == (nE, N + 1, N + 1, N + 1, 5) and np.all(U[..., 0] > 0)
    h = nE // 2
    Ua, _ = state_io.restart(path, N, nt, offsetElem=0, nElems=h)
    Ub, _ = state_io.restart(path, N, nt, offsetElem=h, nElems=nE - h)
    assert np.array_equal(np.concatenate([Ua, Ub]), U)
    Uup, _ = state_io.restart(path, N + 2, "GAUSS-LOBATTO")
    back = metrics.change_basis_volume(bs.get_vandermonde(N + 2, "GAUSS-LOBATTO", N, nt, modal=True), Uup)
    assert np.abs(back - U).max() <= 1e-11 * np.abs(U).max()


def test_hopr_mesh_writer_round_trip_of_generated_meshes(tmp_path):
    """write_hopr_mesh: a generated (curved, periodic / wall) box in the HOPR layout reads back to the same arrays and builds the
    same case tables; unique node / side counts as HOPR counts them."""
    from galaexi_b200.host_standin import mesh as ms
    h = ms.make_box_mesh((3, 2, 2), NGeo=2, deform=0.05, bctype=["periodic", (4, 1), "periodic", (4, 1), "periodic", "periodic"])
    p = str(tmp_path / "box_mesh.h5")
    h5write.write_hopr_mesh(p, h)
    g = h5lite.read_hopr_mesh(p)
    for k in ("ElemInfo", "SideInfo", "NodeCoords", "BCType"):
        assert np.array_equal(g[k], h[k]), k
    assert g["NGeo"] == 2 and g["BCNames"] == [str(s) for s in h["BCNames"]]
    f = h5lite.H5File(p)
    a = f.attrs()
    assert int(a["nElems"][0]) == 12 and int(a["nSides"][0]) == 72 and int(a["nNodes"][0]) == 12 * 27
    assert int(a["nUniqueNodes"][0]) == 7 * 5 * 5                      # physical nodes (periodic copies are distinct in HOPR)
    assert int(a["nUniqueSides"][0]) == 36 + 3 * 2                     # 72 element sides, 60 of them paired (periodic in x and z), 12 walls
    assert np.array_equal(f.dataset("Elem_IJK"), h["Elem_IJK"]) and list(f.dataset("nElems_IJK")) == [3, 2, 2]
    m1 = ms.prepare_mesh(h)
    m2 = ms.prepare_mesh(g)
    assert np.array_equal(m1.ElemToSide, m2.ElemToSide) and np.array_equal(m1.SideToElem, m2.SideToElem) and np.array_equal(m1.BC, m2.BC)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
@pytest.mark.parametrize("rel", ["regressioncheck/checks/tgv/split/CART_HEX_PERIODIC_008_mesh.h5",
                                 "regressioncheck/checks/naca/3D/NACA0012_652_Ng2_mesh.h5",
                                 "tutorials/convtest/CART_HEX_PERIODIC_MORTAR_002_mesh.h5",
                                 "regressioncheck/checks/parabolic/cavity_3D/cavity4x4x4_mesh.h5"])
def test_hopr_mesh_writer_reproduces_reference_mesh_files(tmp_path, rel):
    """Reading a HOPR-written mesh and writing it again gives the same data sets (all of them: also the ones only HOPR / posti
    use -- GlobalNodeIDs handed through, ElemBarycenters, ElemWeight, ElemCounter recomputed) and the same attributes
    (nUniqueSides, nUniqueNodes ... recomputed)."""
    src = os.path.join("/root/reference", rel)
    f = h5lite.H5File(src)
    m = h5lite.read_hopr_mesh(src)
    for k in ("Elem_IJK", "nElems_IJK", "GlobalNodeIDs"):
        if k in f.keys():
            m[k] = f.dataset(k)
    p = str(tmp_path / "m.h5")
    h5write.write_hopr_mesh(p, m)
    g = h5lite.H5File(p)
    assert set(f.keys()) == set(g.keys())                              # incl. the octree data of the mortar mesh
    for k in g.keys():
        a, b = f.dataset(k), g.dataset(k)
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind == "f":
            assert np.allclose(a, b, rtol=0, atol=1e-13), k
        else:
            assert np.array_equal(a, b), k
    fa, ga = f.attrs(), g.attrs()
    assert set(fa) == set(ga)
    for k in fa:
        assert np.array_equal(fa[k], ga[k]) and fa[k].dtype == ga[k].dtype, k


def test_baseflow_file_round_trip_and_interpolation(tmp_path):
    """WriteBaseflow / ReadBaseFlow (hdf5_output.f90:527-603, sponge.f90:468-518)."""
    NC, X = _cart_coords(3, 3, bs.NODETYPE_G)
    B = _poly_state(X, 3)
    p = state_io.write_baseflow(B, 3, bs.NODETYPE_G, "naca", "m.h5", 2.0, 2.5, out_dir=str(tmp_path))
    assert os.path.basename(p) == "naca_BaseFlow_0000002.000000000.h5"
    f = h5lite.H5File(p)
    a = f.attrs()
    assert a["File_Type"][0] == b"BaseFlow" and a["NextFile"][0].decode().strip() == "naca_BaseFlow_0000002.500000000.h5"
    assert f.base == 0 and f.keys() == ["DG_Solution"] and "TIME" in a
    assert np.array_equal(state_io.read_baseflow(p, 3, bs.NODETYPE_G, nGlobalElems=3), B)
    assert np.array_equal(state_io.read_baseflow(p, 3, bs.NODETYPE_G, offsetElem=1, nElems=2), B[1:])
    _, X5 = _cart_coords(3, 5, bs.NODETYPE_GL)
    assert np.abs(state_io.read_baseflow(p, 5, bs.NODETYPE_GL) - _poly_state(X5, 3)).max() < 1e-13
    with pytest.raises(RuntimeError, match="Baseflow file does not match solution"):
        state_io.read_baseflow(p, 3, bs.NODETYPE_G, nGlobalElems=4)
