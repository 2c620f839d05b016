#!/usr/bin/env python
"""Extracts the reference's own golden vectors into small fixtures under tests/golden/.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tools/make_golden.py
Sources (all from flexi-framework/galaexi, read-only):
  unitTests/*.bin                                   Fortran sequential unformatted records (4-byte markers)
  regressioncheck/checks/tgv/split/*.csv + mesh     TGV N=7 GL split-form regression (first rows)
  regressioncheck/checks/parabolic/cavity_3D/*      cavity N=2 Gauss NS+BR1 reference state + mesh
  regressioncheck/checks/naca/3D mesh               curved NGeo=2 mesh (for metric checks)
  tutorials/convtest/CART_HEX_PERIODIC_MORTAR_*     non-conforming meshes (mortar types 1, 2, 3)
Only data files are converted (to .npz); no reference source code is copied.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from galaexi_b200.host_standin import h5lite  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def fort_records(path):
    b = open(path, "rb").read()
    o, recs = 0, []
    while o < len(b):
        n = int(np.frombuffer(b, "<i4", 1, o)[0])
        recs.append(b[o + 4:o + 4 + n])
        assert int(np.frombuffer(b, "<i4", 1, o + 4 + n)[0]) == n
        o += 8 + n
    return recs


def main():
    os.makedirs(OUT, exist_ok=True)
    U = os.path.join(REF, "unitTests")
    g = {}
    r = fort_records(os.path.join(U, "NodesAndWeights.bin"))[0]
    g["nodes_xi_w_wbary"] = np.frombuffer(r, "<f8").reshape(3, 4, 10, 11)     # [xi|w|wBary][G,GL,CL,VISU][N-1][0:10]
    r = fort_records(os.path.join(U, "DerivativeMatrix.bin"))[0]
    g["D"] = np.frombuffer(r, "<f8").reshape(40, 11, 11)                      # [10*(type-1)+N-1][col][row]
    r = fort_records(os.path.join(U, "UnittestElementData3D.bin"))[0]
    o = 0
    g["ued_nElems"] = np.frombuffer(r, "<i4", 1, o); o += 4
    g["ued_SideToElem"] = np.frombuffer(r, "<i4", 30, o).reshape(6, 5); o += 120
    g["ued_ranges"] = np.frombuffer(r, "<i4", 3, o); o += 12
    g["ued_S2V2"] = np.frombuffer(r, "<i4", 6000, o).reshape(6, 5, 10, 10, 2); o += 24000
    for nm in ("L_Minus", "L_Plus", "L_HatPlus", "L_HatMinus"):
        g["ued_" + nm] = np.frombuffer(r, "<f8", 10, o); o += 80
    g["ued_sJ"] = np.frombuffer(r, "<f8", 2000, o).reshape(2, 10, 10, 10); o += 16000
    assert o == len(r)
    g["p2f_Uvol"] = np.frombuffer(fort_records(os.path.join(U, "ProlongToFaceUvol.bin"))[0], "<f8").reshape(10, 10, 10)
    for nt in ("G", "GL"):
        rr = fort_records(os.path.join(U, f"ProlongToFace_{nt}3D.bin"))
        g[f"p2f_{nt}"] = np.frombuffer(rr[0], "<f8").reshape(6, 10, 10)       # [side][q][p]
    g["si_Flux"] = np.frombuffer(fort_records(os.path.join(U, "SurfIntFlux.bin"))[0], "<f8").reshape(6, 10, 10)
    for nt in ("G", "GL"):
        rr = fort_records(os.path.join(U, f"SurfInt_{nt}3D.bin"))
        g[f"si_{nt}"] = np.frombuffer(rr[0], "<f8").reshape(10, 10, 10)       # [k][j][i]
    r = fort_records(os.path.join(U, "ChangeBasis.bin"))[0]
    a = np.frombuffer(r, "<f8")
    n3d = 3 * 6 * 6 * 6 * 6 * 4
    g["cb_UOut"] = a[:n3d].reshape(4, 6, 6, 6, 6, 3)          # [slot][elem][k][j][i][var]  (ChangeBasis.f90:27-32)
    g["cb_UOut2D"] = a[n3d:].reshape(3, 6, 6, 6, 3)           # [slot][elem][j][i][var]
    assert a.size == n3d + 3 * 6 * 6 * 6 * 3
    rr = fort_records(os.path.join(U, "Vandermonde.bin"))
    g["vdm_raw"] = np.frombuffer(b"".join(rr), "<f8")
    np.savez_compressed(os.path.join(OUT, "unit_goldens.npz"), **g)

    # regression meshes + references
    def mesh_npz(src, dst):
        m = h5lite.read_hopr_mesh(os.path.join(REF, src))
        np.savez_compressed(os.path.join(OUT, dst), NGeo=m["NGeo"], ElemInfo=m["ElemInfo"], SideInfo=m["SideInfo"],
                            NodeCoords=m["NodeCoords"], BCNames=np.array(m["BCNames"]), BCType=m["BCType"])
    mesh_npz("regressioncheck/checks/tgv/split/CART_HEX_PERIODIC_008_mesh.h5", "tgv_split_mesh.npz")
    mesh_npz("regressioncheck/checks/parabolic/cavity_3D/cavity4x4x4_mesh.h5", "cavity3d_mesh.npz")
    mesh_npz("regressioncheck/checks/naca/3D/NACA0012_652_Ng2_mesh.h5", "naca_mesh.npz")
    mesh_npz("tutorials/convtest/CART_HEX_PERIODIC_002_mesh.h5", "cart_periodic_002_mesh.npz")
    mesh_npz("tutorials/convtest/CART_HEX_PERIODIC_004_mesh.h5", "cart_periodic_004_mesh.npz")
    mesh_npz("tutorials/convtest/CART_HEX_PERIODIC_008_mesh.h5", "cart_periodic_008_mesh.npz")
    # non-conforming (mortar) meshes of the convergence-test tutorial: types 1 (1->4), 2 and 3 (1->2), periodic
    for nm in ("001", "002", "004", "008"):
        mesh_npz(f"tutorials/convtest/CART_HEX_PERIODIC_MORTAR_{nm}_mesh.h5", f"cart_mortar_{nm}_mesh.npz")
    csv = np.loadtxt(os.path.join(REF, "regressioncheck/checks/tgv/split/TGV_Re1600_Split_TGVAnalysis_Reference.csv"),
                     delimiter=",", skiprows=1, converters=lambda s: float(s.replace("E+0", "E+").replace("E-0", "E-")))
    np.savez_compressed(os.path.join(OUT, "tgv_split_csv.npz"), rows=csv)   # all 555 analyze rows (t = 0 ... 13)
    # tgv/oInt: N=11, OverintegrationType=1 (cut-off filter on JU_t), NUnder=7, 4^3 elements: pins the overintegration step
    mesh_npz("regressioncheck/checks/tgv/oInt/CART_HEX_PERIODIC_004_mesh.h5", "tgv_oint_mesh.npz")
    csv = np.loadtxt(os.path.join(REF, "regressioncheck/checks/tgv/oInt/TGV_Re1600_OInt_TGVAnalysis_Reference.csv"),
                     delimiter=",", skiprows=1, converters=lambda s: float(s.replace("E+0", "E+").replace("E-0", "E-")))
    np.savez_compressed(os.path.join(OUT, "tgv_oint_csv.npz"), rows=csv)    # all 437 analyze rows (t = 0 ... 13)
    st = h5lite.read_state(os.path.join(REF, "regressioncheck/checks/parabolic/cavity_3D/reggie_cavity_Re100_State_0000001.000000000.h5"))
    np.savez_compressed(os.path.join(OUT, "cavity3d_state.npz"), DG_Solution=st["DG_Solution"], Time=st["attrs"]["Time"])
    st = h5lite.read_state(os.path.join(REF, "regressioncheck/checks/naca/3D/NACA0012_Re5000_AoA8_3D_Referenz_0000010.000000000.h5"))
    np.savez_compressed(os.path.join(OUT, "naca3d_state.npz"), DG_Solution=st["DG_Solution"], Time=st["attrs"]["Time"])
    state_h5_structs()
    unit_goldens_emm()
    print("wrote", sorted(os.listdir(OUT)))


def unit_goldens_emm():
    """unitTests/SurfInt_GL3D_EMM.bin: the SurfInt unit test of a Gauss-Lobatto build with FLEXI_EXACT_MASSMATRIX (first record:
    Ut of the DG element; the second record is the FV variant)."""
    rr = fort_records(os.path.join(REF, "unitTests", "SurfInt_GL3D_EMM.bin"))
    np.savez_compressed(os.path.join(OUT, "unit_goldens_emm.npz"), si_GL_EMM=np.frombuffer(rr[0], "<f8").reshape(10, 10, 10))


def state_h5_structs():
    """The HDF5 structures libhdf5 wrote into the reference's cavity state file, as hex strings: superblock, the root
    attribute messages, the object-header messages of DG_Solution / ElemData, the local heap and the B-tree / symbol node
    heads. tests/test_state_io.py compares what galaexi_b200/host_standin/h5write.py emits with them."""
    import json
    import struct
    f = h5lite.H5File(os.path.join(REF, "regressioncheck/checks/parabolic/cavity_3D/reggie_cavity_Re100_State_0000001.000000000.h5"))
    b, buf = f.base, f.buf
    out = dict(userblock_size=b, file_size=len(buf), superblock=bytes(buf[b:b + 96]).hex())
    out["root_attr_msgs"] = {}
    for t, pl in f.root_msgs:
        if t == 0x0C:
            nsz = struct.unpack_from("<H", pl, 2)[0]
            out["root_attr_msgs"][bytes(pl[8:8 + nsz - 1]).decode()] = bytes(pl).hex()
    out["datasets"] = {}
    for nm, addr in f.objects.items():
        a = b + addr
        nmsgs, size = struct.unpack_from("<H", buf, a + 2)[0], struct.unpack_from("<I", buf, a + 8)[0]
        pos, msgs = a + 16, []
        while pos < a + 16 + size:
            t, sz, fl = struct.unpack_from("<HHB", buf, pos)
            msgs.append([t, fl, bytes(buf[pos + 8:pos + 8 + sz]).hex()])
            pos += 8 + sz
        out["datasets"][nm] = dict(header=bytes(buf[a:a + 16]).hex(), msgs=msgs)
    heap_a = b + struct.unpack_from("<Q", buf, b + 56 + 32)[0]
    dsz, free, daddr = struct.unpack_from("<QQQ", buf, heap_a + 8)
    out["heap_header"] = bytes(buf[heap_a:heap_a + 32]).hex()
    out["heap_data"] = bytes(buf[b + daddr:b + daddr + dsz]).hex()
    bt = b + struct.unpack_from("<Q", buf, b + 56 + 24)[0]
    out["btree_head"] = bytes(buf[bt:bt + 48]).hex()
    sn = b + struct.unpack_from("<Q", buf, bt + 32)[0]
    out["snod"] = bytes(buf[sn:sn + 8 + 2 * 40]).hex()
    with open(os.path.join(OUT, "state_h5_structs.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
