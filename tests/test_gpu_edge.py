"""Edge cases of the /gpu/b200 path and size-independent properties at BASELINE sizes."""
import ctypes as C

import numpy as np
import pytest

from helpers import OracleProblem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def G():
    import gpu_helpers
    return gpu_helpers


@pytest.mark.parametrize("n", [(1, 1, 1), (1, 1, 3), (5, 1, 1)])
def test_single_element_and_tail_groups(G, n):
    g = G.GpuProblem("hyperFS", n, 4)
    o = OracleProblem("hyperFS", n, 4)
    assert rel_err(g.residual(), o.residual_fine(o.u_fine)) < TOL
    x = np.random.default_rng(0).standard_normal(o.lsize)
    assert rel_err(g.jacobian(len(g.degrees) - 1, x), o.jacobian(x)) < TOL


@pytest.mark.parametrize("problem,nu,E", [("linElas", 0.45, 2.5e3), ("hyperSS", 0.1, 7.0), ("hyperFS", 0.49, 1e-2)])
def test_other_material_parameters(G, problem, nu, E):
    import helpers
    old = helpers.PHYS
    helpers.PHYS = (nu, E)
    try:
        g = G.GpuProblem(problem, 3, 2, nu=nu, E=E)
        o = OracleProblem(problem, 3, 2)
        assert rel_err(g.residual(), o.residual_fine(o.u_fine)) < TOL
        x = np.random.default_rng(1).standard_normal(o.lsize)
        assert rel_err(g.jacobian(len(g.degrees) - 1, x), o.jacobian(x)) < TOL
        assert rel_err(g.diagonal(len(g.degrees) - 1), o.diagonal()) < TOL
    finally:
        helpers.PHYS = old


def test_smoother_context_swap_in_getdiag(G):
    """GetDiag_Ceed swaps in physSmoother around the assembly and restores phys (matops.c:215-232)."""
    import torch
    import helpers
    from ceedpetscsolid_b200 import matops
    from ceedpetscsolid_b200.ceed import Physics
    g = G.GpuProblem("hyperFS", 3, 2)
    o = OracleProblem("hyperFS", 3, 2)
    g.residual()
    lvl = len(g.degrees) - 1
    dm = matops.LevelDM(g.mesh, g.degrees[lvl], bc_faces=None)
    smooth = Physics(0.2, 1.0)
    user = matops.setup_jacobian_ctx(dm, g.ceed, g.data[lvl], g.phys, physSmoother=smooth)
    D = dm.create_global_vector()
    matops.GetDiag_Ceed(user, D)
    old = helpers.PHYS
    helpers.PHYS = (0.2, 1.0)
    try:
        dref = o.diagonal()
    finally:
        helpers.PHYS = old
    assert rel_err(D.cpu().numpy(), dref) < TOL
    x = np.random.default_rng(2).standard_normal(o.lsize)
    X, Y = dm.create_global_vector(), dm.create_global_vector()
    X.copy_(torch.from_numpy(x))
    matops.ApplyJacobian_Ceed(user, X, Y)
    assert rel_err(Y.cpu().numpy(), o.jacobian(x)) < TOL  # back on the original context


def test_vector_semantics(G):
    import torch
    from ceedpetscsolid_b200 import ceed as libceed
    c = libceed.Ceed("/gpu/b200")
    v = c.Vector(1000)
    h = np.arange(1000, dtype=np.float64)
    v.set_array(h, libceed.MEM_HOST, libceed.COPY_VALUES)
    h[:] = -1  # COPY_VALUES: the caller's array is no longer referenced
    np.testing.assert_array_equal(v.to_numpy(), np.arange(1000))
    assert abs(v.norm(libceed.NORM_2) - np.linalg.norm(np.arange(1000))) < 1e-9
    assert v.norm(libceed.NORM_MAX) == 999 and v.norm(libceed.NORM_1) == 999 * 500
    v.set_value(4.0)
    v.reciprocal()
    np.testing.assert_array_equal(v.to_numpy(), np.full(1000, 0.25))
    d = torch.full((1000,), 2.0, dtype=torch.float64, device="cuda")
    v.set_array(d)            # borrow a device array ...
    v.reciprocal()            # ... results are written in place
    v.take_array()
    torch.cuda.synchronize()
    assert float(d[7]) == 0.5
    with pytest.raises(libceed.CeedError, match="no valid data"):
        v.to_numpy()          # after TakeArray the vector holds nothing


def test_standalone_restriction_and_basis_apply(G):
    """CeedElemRestrictionApply / CeedBasisApply (generic kernels) against the oracle."""
    from ceedpetscsolid_b200 import ceed as libceed
    from ceedpetscsolid_b200.mesh import BoxMesh
    from oracle import oracle
    c = libceed.Ceed("/gpu/b200")
    mesh = BoxMesh(n=(3, 2, 2))
    P, Q, nel = 4, 5, 12
    off = mesh.offsets(3)
    r = c.ElemRestriction(nel, P ** 3, 3, 1, mesh.lsize(3), off)
    rng = np.random.default_rng(3)
    L = rng.standard_normal(mesh.lsize(3))
    lv, ev = c.Vector(L.size), c.Vector(nel * 3 * P ** 3)
    lv.set_array(L, libceed.MEM_HOST, libceed.COPY_VALUES)
    r.apply(lv, ev)
    E = ev.to_numpy().reshape(nel, 3, P ** 3)
    np.testing.assert_array_equal(E, L[off[:, None, :] + np.arange(3)[None, :, None]])
    b = c.BasisTensorH1Lagrange(3, 3, P, Q)
    B, D, qr, qw = oracle.basis_1d(P, Q, 0)
    np.testing.assert_allclose(b.interp1d, B, atol=1e-14)
    np.testing.assert_allclose(b.grad1d, D, atol=1e-13)
    np.testing.assert_allclose(b.qweight1d, qw, atol=1e-15)
    qv = c.Vector(nel * 9 * Q ** 3)
    b.apply(nel, libceed.NOTRANSPOSE, libceed.EVAL_GRAD, ev, qv)
    ref = oracle.basis_apply(nel, 3, P, Q, B, D, qw, 0, 2, E)
    assert rel_err(qv.to_numpy(), ref.reshape(-1)) < TOL
    back = c.Vector(nel * 3 * P ** 3)
    b.apply(nel, libceed.TRANSPOSE, libceed.EVAL_GRAD, qv, back)
    assert rel_err(back.to_numpy(), oracle.basis_apply(nel, 3, P, Q, B, D, qw, 1, 2, ref).reshape(-1)) < TOL
    lv2 = c.Vector(L.size)
    lv2.set_value(0.0)
    r.apply(back, lv2, libceed.TRANSPOSE)
    ref_l = np.zeros(L.size)
    oracle.lib().oracle_restrict_scatter_add(nel, P ** 3, 3, 1, off.ctypes.data_as(C.c_void_p),
                                            np.ascontiguousarray(back.to_numpy()).ctypes.data_as(C.c_void_p),
                                            ref_l.ctypes.data_as(C.c_void_p))
    assert rel_err(lv2.to_numpy(), ref_l) < TOL


def test_empty_restriction_and_operator(G):
    """nelem = 0 (a rank that owns no elements): every call is a no-op, nothing is launched out of bounds."""
    from ceedpetscsolid_b200 import ceed as libceed
    c = libceed.Ceed("/gpu/b200")
    r = c.ElemRestriction(0, 27, 3, 1, 30, np.zeros(0, dtype=np.int32))
    rq = c.StridedElemRestriction(0, 27, 10, 0)
    b = c.BasisTensorH1Lagrange(3, 3, 3, 3)
    qd = c.Vector(0)
    qf = c.QFunction(1, "qfunctions/linElas.h:LinElasdF")
    qf.add_input("deltadu", 9, libceed.EVAL_GRAD)
    qf.add_input("qdata", 10, libceed.EVAL_NONE)
    qf.add_output("deltadv", 9, libceed.EVAL_GRAD)
    qf.set_context(libceed.Physics(0.3, 1.0))
    op = c.Operator(qf)
    op.set_field("deltadu", r, b, libceed.VECTOR_ACTIVE)
    op.set_field("qdata", rq, libceed.BASIS_COLLOCATED, qd)
    op.set_field("deltadv", r, b, libceed.VECTOR_ACTIVE)
    x, y = c.Vector(30), c.Vector(30)
    x.set_value(1.0)
    y.set_value(7.0)
    op.apply(x, y)
    np.testing.assert_array_equal(y.to_numpy(), np.zeros(30))  # CeedOperatorApply zeroes its output
    op.linear_assemble_diagonal(y)
    np.testing.assert_array_equal(y.to_numpy(), np.zeros(30))


@pytest.mark.parametrize("problem,p,n", [("hyperSS", 3, 32), ("hyperFS", 4, 64), ("hyperFS", 4, 80)])
def test_properties_at_baseline_sizes(G, problem, p, n):
    """BASELINE configs[1], [2] and the ~100 M-DoF box of configs[3] on one GPU: linearity, symmetry, rigid-body null space, positive diagonal --
    size-independent properties where the oracle is too slow to run."""
    import torch
    g = G.GpuProblem(problem, n, p)
    g.residual()
    lvl = len(g.degrees) - 1
    nl = 3 * g.mesh.num_nodes(p)
    rng = np.random.default_rng(4)
    v, w = rng.standard_normal(nl), rng.standard_normal(nl)
    Jv, Jw = g.jacobian(lvl, v), g.jacobian(lvl, w)
    Jc = g.jacobian(lvl, 0.3 * v - 1.7 * w)
    assert rel_err(Jc, 0.3 * Jv - 1.7 * Jw) < 1e-12
    a, b = float(np.dot(w, Jv)), float(np.dot(v, Jw))
    assert abs(a - b) < 1e-11 * abs(a)
    t = np.tile([0.3, -0.2, 0.7], nl // 3)
    assert np.linalg.norm(g.jacobian(lvl, t)) < 1e-12 * np.linalg.norm(Jv)
    d = g.diagonal(lvl)
    assert d.min() > 0
    # the diagonal against unit-vector probing of the same operator at a few dofs
    for i in rng.choice(nl, 3, replace=False):
        e = np.zeros(nl)
        e[i] = 1.0
        assert abs(g.jacobian(lvl, e)[i] - d[i]) < 1e-12 * d.max()
    del g
    torch.cuda.empty_cache()


@pytest.mark.parametrize("problem,p,n,perm", [("hyperFS", 2, (16, 32, 33), None), ("hyperSS", 4, (16, 16, 70), None),
                                             ("hyperFS", 2, (16, 32, 32), 5), ("linElas", 1, (40, 40, 21), None)])
def test_host_resident_vectors_take_the_pipelined_path(G, problem, p, n, perm):
    """-memtype host (matops.c:40-50 with host arrays): page-locked x and y -> chunked H2D | kernel | D2H pipeline.
    Same numbers as the device-resident apply (atomics reorder: 1e-13), for ordered and for random numberings."""
    import torch
    from ceedpetscsolid_b200 import ceed as libceed
    g = G.GpuProblem(problem, n, p, node_perm_seed=perm)
    g.residual()
    lsize = 3 * g.mesh.num_nodes(p)
    rng = np.random.default_rng(4)
    x = rng.standard_normal(lsize)
    for op in (g.fine.opJacob, g.fine.opApply):
        xin = x if op is g.fine.opJacob else g.u_fine
        ref = g.apply(op, xin, lsize)
        for pinned in (True, False):
            xh = torch.from_numpy(xin.copy())
            yh = torch.full((lsize,), np.nan, dtype=torch.float64)
            if pinned:
                xh, yh = xh.pin_memory(), yh.pin_memory()
            xc, yc = g.ceed.Vector(lsize), g.ceed.Vector(lsize)
            xc.set_array(xh, libceed.MEM_HOST)
            yc.set_array(yh, libceed.MEM_HOST)
            n0 = libceed.launch_count()
            op.apply(xc, yc)
            launches = libceed.launch_count() - n0
            xc.take_array(libceed.MEM_HOST)
            yc.take_array(libceed.MEM_HOST)
            assert (launches >= 2) == pinned, launches    # one kernel per element chunk vs a single kernel
            assert rel_err(yh.numpy(), ref) < 1e-13
            # x is current on the device afterwards: a second, device-side use must not need the host copy
            xc.destroy(); yc.destroy()


def test_partitioned_apply_without_a_halo_equals_apply_plus_mask():
    """CeedOperatorApplyPartitionedB200 on one rank (halo = NULL): zero, fused operator, Dirichlet rows zeroed --
    bit-for-bit what CeedOperatorApply followed by the mask gives when the scatter is deterministic, and to round-off
    with atomics; element counts that are not a multiple of the group size are refused for the interface split."""
    import torch
    import gpu_helpers as G
    from ceedpetscsolid_b200 import ceed as libceed
    for resource in ("/gpu/b200", "/gpu/b200:deterministic"):
        g = G.GpuProblem("hyperFS", (3, 2, 2), 4, resource=resource)
        g.residual()
        fine = len(g.degrees) - 1
        op = g.data[fine].opJacob
        n = g.mesh.lsize(4)
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        y1, y2 = torch.full_like(x, 7.0), torch.full_like(x, -3.0)
        mask = torch.arange(0, n, 5, dtype=torch.int32, device="cuda")
        xc, yc = g.ceed.Vector(n), g.ceed.Vector(n)
        xc.set_array(x); yc.set_array(y1)
        op.apply(xc, yc)
        yc.take_array()
        y1[mask.long()] = 0.0
        yc.set_array(y2)
        op.apply_partitioned(xc, yc, 0, None, mask)
        yc.take_array()
        torch.cuda.synchronize()
        if "deterministic" in resource:
            assert torch.equal(y1, y2)
        else:
            assert float((y1 - y2).norm() / y1.norm()) < 1e-14
        yc.set_array(y2)
        with pytest.raises(libceed.CeedError, match="n_interface"):
            op.apply_partitioned(xc, yc, 3, None, mask)          # not a multiple of the element-group size
        yc.take_array(); xc.take_array()


def test_borrowed_host_output_is_current_when_apply_returns():
    """/cpu/self writes a CEED_USE_POINTER host array in place and the reference's ViewDiagnosticQuantities reads it
    between CeedOperatorApply and CeedVectorTakeArray (misc.c:258-268): the borrowed host output must already hold the
    result when CeedOperatorApply / CeedOperatorApplyAdd / CeedElemRestrictionGetMultiplicity return."""
    import gpu_helpers as G
    from ceedpetscsolid_b200 import ceed as libceed
    g = G.GpuProblem("hyperSS", (3, 2, 2), 2)
    g.residual()
    fine = len(g.degrees) - 1
    op = g.data[fine].opJacob
    n = g.mesh.lsize(2)
    x = np.random.default_rng(5).standard_normal(n)
    y, y_ref = np.full(n, 7.0), np.empty(n)
    xc, yc, rc = g.ceed.Vector(n), g.ceed.Vector(n), g.ceed.Vector(n)
    xc.set_array(x, libceed.MEM_HOST, libceed.USE_POINTER)
    rc.set_array(y_ref, libceed.MEM_HOST, libceed.USE_POINTER)
    op.apply(xc, rc)
    rc.take_array(libceed.MEM_HOST)
    assert np.linalg.norm(y_ref) > 0
    yc.set_array(y, libceed.MEM_HOST, libceed.USE_POINTER)
    op.apply(xc, yc)
    assert rel_err(y, y_ref) < 1e-14                 # before TakeArray
    op.apply_add(xc, yc)
    assert rel_err(y, 2 * y_ref) < 1e-14
    g.data[fine].Erestrictu.get_multiplicity(yc)
    assert y.min() >= 1.0 and y.max() <= 8.0 and np.all(y == np.round(y))
    yc.take_array(libceed.MEM_HOST)
    xc.take_array(libceed.MEM_HOST)
