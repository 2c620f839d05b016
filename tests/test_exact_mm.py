"""FLEXI_EXACT_MASSMATRIX (-DEXACT_MM, src/CMakeLists.txt:150-156): Gauss-Lobatto nodes with the exact mass matrix.

Reference: interpolation/basis.f90:762-807 (PolynomialMassMatrix), dg/dg.f90:181-242 (D_Hat, L_Hat from Minv), dg/surfint.t90:74-104
and dg/lifting/lifting_br1.t90:238-270 (the full-L_Hat form for `PP_NodeType==1 || (PP_NodeType==2 && defined(EXACT_MM))`),
timedisc_vars.f90:105-109 (Gauss CFL / DFL tables). The host hands such a case to the library and the oracle as
dgx_config.nodeType = 1 (Case.op_node_type) with the Gauss-Lobatto operator tables; pinned by unitTests/SurfInt_GL3D_EMM.bin
(tests/test_oracle_goldens.py)."""
import numpy as np
import pytest

import cases
from galaexi_b200.host_standin import basis as bs
from galaexi_b200.host_standin import equation as eq
from galaexi_b200.host_standin import timedisc as td
from galaexi_b200.host_standin import timeloop


@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 9])
def test_exact_mass_matrix_is_the_integral_of_lagrange_products(N):
    x, w, wb = bs.get_nodes_and_weights(N, bs.NODETYPE_GL)
    M, Minv = bs.polynomial_mass_matrix(N, x, w, True)
    xg, wg, _ = bs.get_nodes_and_weights(N + 2, bs.NODETYPE_G)           # exact for degree 2N
    L = np.array([bs.lagrange_interpolation_polys(float(t), x, wb) for t in xg])
    assert np.abs(M - np.einsum("q,qi,qj->ij", wg, L, L)).max() < 1e-14
    assert np.abs(M @ Minv - np.eye(N + 1)).max() < 1e-13
    Md, _ = bs.polynomial_mass_matrix(N, x, w, False)
    assert np.array_equal(Md, np.diag(w)) and np.abs(M - Md).max() > 1e-3      # the lumped matrix is a different one
    b = bs.init_dg_basis(N, bs.NODETYPE_GL, True)
    assert np.array_equal(b.L_Minus, np.eye(N + 1)[0]) and np.array_equal(b.L_Plus, np.eye(N + 1)[N])
    assert np.allclose(b.L_HatMinus, Minv[:, 0], rtol=0, atol=1e-13) and np.count_nonzero(b.L_HatMinus) == N + 1
    # SBP property with the exact mass matrix: M D + (M D)^T = B = diag(-1, 0, ..., 0, 1)
    Q = M @ b.D
    B = np.zeros((N + 1, N + 1)); B[0, 0] = -1.0; B[N, N] = 1.0
    assert np.abs(Q + Q.T - B).max() < 1e-12


def test_options_and_tables():
    with pytest.raises(ValueError, match="only works on FLEXI_NODETYPE==GAUSS-LOBATTO"):
        bs.init_dg_basis(3, bs.NODETYPE_G, True)
    with pytest.raises(ValueError, match="EXACT_MM with SplitDG is not built"):
        cases.tgv_box_case(E=2, N=3, exact_mm=True)
    c, _ = cases.tgv_box_case(E=2, N=3, split=None, riemann="Roe", exact_mm=True)
    assert c.node_type == bs.NODETYPE_GL and c.op_node_type == 1 and c.exact_mm
    assert cases.tgv_box_case(E=2, N=3, split=None, riemann="Roe")[0].op_node_type == 2
    assert cases.tgv_box_case(E=2, N=3, split=None, riemann="Roe", node_type="GAUSS")[0].op_node_type == 1
    g = td.set_timedisc("carpenterrk4-5", 3, bs.NODETYPE_G, 0.9, 0.9)
    e = td.set_timedisc("carpenterrk4-5", 3, bs.NODETYPE_GL, 0.9, 0.9, exact_mm=True)
    assert (e.CFLScale, e.DFLScale) == (g.CFLScale, g.DFLScale)


def _oracle(c):
    from oracle.oracle import Oracle
    return Oracle(c)


def test_oracle_free_stream_and_conservation_on_curved_mesh():
    c, U0 = cases.tgv_box_case(E=3, N=4, NGeo=2, deform=0.05, perturb=1e-3, split=None, riemann="Roe", exact_mm=True)
    o = _oracle(c)
    Ufs = eq.ini_refstate(c.geo["Elem_xGP"], c.RefStatePrim[0], c.eos)
    o.set_state(Ufs)
    Ut = o.time_derivative(0.0)
    assert np.abs(Ut).max() <= 1e-9 * np.abs(Ufs).max()
    # conservation on the periodic box: sum_e int J Ut = 0 with the quadrature the scheme itself is exact for (M exact: use
    # the mass matrix, i.e. 1^T M x 1^T M x 1^T M applied to J Ut)
    o.set_state(U0)
    Ut = o.time_derivative(0.0).copy()
    x, w, _ = bs.get_nodes_and_weights(c.N, bs.NODETYPE_GL)
    M, _ = bs.polynomial_mass_matrix(c.N, x, w, True)
    m1 = M.sum(axis=0)
    W = m1[:, None, None] * m1[None, :, None] * m1[None, None, :]
    tot = np.einsum("kji,ekjiv->v", W, Ut / c.geo["sJ"][..., None])
    scale = np.einsum("kji,ekjiv->v", W, np.abs(Ut) / c.geo["sJ"][..., None])
    assert np.all(np.abs(tot) <= 1e-11 * scale)
    o.close()


def test_prolongation_on_gl_nodes_is_extraction():
    """op_node_type 1 uses the interpolating ProlongToFace; with L_Minus / L_Plus unit vectors it must give exactly the face data
    of the Gauss-Lobatto extraction (prolongtoface.t90:168-344)."""
    c1, U0 = cases.tgv_box_case(E=2, N=3, NGeo=2, deform=0.05, perturb=1e-2, split=None, riemann="Roe", exact_mm=True)
    c2, _ = cases.tgv_box_case(E=2, N=3, NGeo=2, deform=0.05, perturb=1e-2, split=None, riemann="Roe")
    faces = []
    for c in (c1, c2):
        o = _oracle(c)
        o.set_state(U0)
        o.time_derivative(0.0)
        faces.append((o.array("U_master").copy(), o.array("U_slave").copy()))
        o.close()
    assert np.array_equal(faces[0][0], faces[1][0]) and np.array_equal(faces[0][1], faces[1][1])


def test_density_wave_exact_mm_converges_and_beats_collocation():
    """convtest set-up (exact function 2): order N+1 with the exact mass matrix, and a smaller error than the lumped
    Gauss-Lobatto scheme on the same mesh (which under-integrates the mass matrix)."""
    errs = {}
    for mesh in ("cart_periodic_002", "cart_periodic_004"):
        for mm in (True, False):
            c, U0 = cases.convtest_case(mesh, N=3, node_type=bs.NODETYPE_GL, exact_mm=mm)
            o = _oracle(c)
            o.set_state(U0)

            class _Op:
                def calc_timestep(self):
                    return o.calc_timestep()

                def rk_step(self, t, dt):
                    o.rk_step(t, dt)
            t, _ = timeloop.advance(_Op(), 0.0, 0.1)
            errs[(mesh, mm)] = cases.l2_error(c, o.array("U").copy(), t)[0]
            o.close()
    order = np.log2(errs[("cart_periodic_002", True)] / errs[("cart_periodic_004", True)])
    assert order > 3.5, (order, errs)
    assert errs[("cart_periodic_004", True)] < errs[("cart_periodic_004", False)]
