"""Newton-Krylov-p-MG harness: host logic on the CPU oracle (no GPU), and GPU-vs-oracle parity of
iteration counts and of the solution under the same harness."""
import numpy as np
import pytest
import torch

from ceedpetscsolid_b200 import solver
from ceedpetscsolid_b200.elasticity import AppCtx, bc_clamp


def test_pcg_and_chebyshev_on_a_small_spd_matrix():
    rng = np.random.default_rng(0)
    n = 60
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.linspace(1.0, 50.0, n)
    Am = torch.from_numpy(Q @ np.diag(lam) @ Q.T)
    A = lambda x, y: y.copy_(Am @ x)
    V = solver.Vec()
    b = torch.from_numpy(rng.standard_normal(n))
    x = torch.zeros(n, dtype=torch.float64)
    its, reason, _ = solver.pcg(V, A, b, x, rtol=1e-12)
    assert reason == "rtol" and its <= n
    assert torch.linalg.norm(Am @ x - b) < 1e-9 * torch.linalg.norm(b)
    dinv = 1.0 / torch.diagonal(Am)
    lmax = solver.estimate_lambda_max(V, A, dinv, n, torch.device("cpu"), its=30)
    true = np.linalg.eigvalsh((torch.diag(dinv) @ Am).numpy().astype(float)).real.max() if False else \
        np.max(np.real(np.linalg.eigvals(torch.diag(dinv).numpy() @ Am.numpy())))
    assert abs(lmax - true) < 0.05 * true
    sm = solver.ChebyshevJacobi(V, A, n, torch.device("cpu"), its=3)
    sm.setup(torch.diagonal(Am).clone())
    x0 = torch.zeros(n, dtype=torch.float64)
    sm.apply(b, x0, zero_guess=True)
    e0 = torch.linalg.norm(torch.linalg.solve(Am, b))
    e1 = torch.linalg.norm(x0 - torch.linalg.solve(Am, b))
    assert e1 < e0  # a smoother reduces the error


def test_bc_clamp_matches_reference_expression():
    """src/boundary.c:53-74 incl. its precedence quirk in the first component"""
    xyz = np.array([[0.3, 0.7, 1.0]])
    cl = [0.1, -0.2, 0.05, 0.0, 0.0, 1.0, 0.25]
    u = bc_clamp(xyz, 0.5, cl)[0]
    x, y, z = xyz[0]
    th = 0.25 * np.pi * 0.5
    c, s = np.cos(th), np.sin(th)
    kx, ky, kz = 0.0, 0.0, 1.0
    assert np.isclose(u[0], 0.05 + s * (-kz * y + ky * z) + (1 - c) * (-ky * ky + kz * kz * x + kx * ky * y + kx * kz * z))
    assert np.isclose(u[1], -0.1 + s * (kz * x - kx * z) + (1 - c) * (kx * ky * x - (kx * kx + kz * kz) * y + ky * kz * z))
    assert np.isclose(u[2], 0.025)


@pytest.mark.parametrize("problem,degree,steps,masked", [("linElas", 2, 1, False), ("hyperFS", 2, 2, True)])
def test_newton_krylov_pmg_converges_on_the_oracle(problem, degree, steps, masked):
    from oracle_levels import oracle_solve
    app = AppCtx(problem=problem, degree=degree, n=(3, 3, 3), num_steps=steps)
    out, U = oracle_solve(app, masked=masked)
    assert out["converged"], out
    assert out["snes_its"] >= steps and out["ksp_its"] > 0
    if problem == "linElas":
        assert out["snes_its"] <= 2  # linear problem: one Newton step (+ at most one check step)
    assert out["ksp_its"] / out["snes_its"] < 40  # p-MG preconditioned CG, mesh-independent-ish counts


@pytest.mark.gpu
@pytest.mark.parametrize("deterministic", [False, True], ids=["atomics", "deterministic"])
@pytest.mark.parametrize("problem,degree,n,steps,masked", [("hyperFS", 2, (4, 4, 4), 2, True), ("hyperSS", 3, (3, 3, 3), 1, True),
                                                           ("hyperFS", 4, (2, 2, 3), 2, True), ("hyperFS", 2, (4, 4, 4), 2, False)])
def test_gpu_and_oracle_agree_on_iteration_counts_and_solution(problem, degree, n, steps, masked, deterministic):
    """deterministic: "/gpu/b200:deterministic" -- every transposed restriction sums in the serial /cpu/self order
    (matops.c:46 behind CeedOperatorApply), so the iteration counts do not rest on FP64 atomics being benign."""
    from ceedpetscsolid_b200.elasticity import Elasticity
    from oracle_levels import oracle_solve
    app = AppCtx(problem=problem, degree=degree, n=n, num_steps=steps, perturb=0.05,
                 clamp={(2, 0): [0, 0, 0, 0, 0, 1, 0], (2, 1): [0.01, 0, -0.04, 0, 0, 1, 0.02]})
    ref, Uref = oracle_solve(app, masked=masked)
    el = Elasticity(app, masked=masked, deterministic=deterministic)
    assert el.ceed.is_deterministic == deterministic
    out = el.solve()
    assert out["converged"] and ref["converged"]
    assert out["snes_its"] == ref["snes_its"], (out, ref)
    assert out["ksp_its"] == ref["ksp_its"], (out, ref)
    u = el.U.cpu().numpy()
    assert np.linalg.norm(u - Uref.numpy()) < 1e-9 * np.linalg.norm(Uref.numpy())


@pytest.mark.gpu
def test_reference_acceptance_run_mms(tmp_path):
    """The reference's only test (elasticity.c:36): `-test -degree 3 -nu 0.3 -E 1 -dm_plex_box_faces 3,3,3`
    = linElas with the manufactured forcing and BCMMS on the whole boundary; passes iff the relative
    L2 error against the manufactured solution is at most 5 % (elasticity.c:800-811)."""
    from ceedpetscsolid_b200.elasticity import Elasticity
    app = AppCtx(problem="linElas", degree=3, n=(3, 3, 3), nu=0.3, E=1.0, num_steps=1, test_mode=True)
    el = Elasticity(app)
    out = el.solve()
    assert out["converged"]
    err = el.mms_l2_error()
    assert err < 0.05, err
    assert err < 5e-3  # what this discretisation actually delivers
    # finer mesh: the error must drop (degree-3 elements: ~h^4)
    el2 = Elasticity(AppCtx(problem="linElas", degree=3, n=(6, 6, 6), nu=0.3, E=1.0, num_steps=1, test_mode=True))
    assert el2.solve()["converged"]
    err2 = el2.mms_l2_error()
    assert err2 < err / 8, (err, err2)


@pytest.mark.gpu
def test_forcing_operators_match_the_oracle():
    """SetupMMSForce / SetupConstantForce through the generic operator path (INTERP in, INTERP^T out)."""
    import ctypes as C
    from ceedpetscsolid_b200 import ceed as libceed, setuplibceed
    from ceedpetscsolid_b200.mesh import BoxMesh
    from oracle import oracle
    mesh = BoxMesh(n=(3, 2, 2), perturb=0.1, seed=0)
    c = libceed.Ceed("/gpu/b200")
    deg, P, Q = 2, 3, 3
    phys = libceed.Physics(0.3, 1.0)
    data = setuplibceed.CeedData()
    setuplibceed.setup_fine_level(c, mesh, "linElas", deg, phys, data)
    nel = mesh.nelem
    Bx, Dx, _, qw = oracle.basis_1d(2, Q, 0)
    Bu, Du, _, _ = oracle.basis_1d(P, Q, 0)
    qdata = oracle.setup_geo(nel, Q, mesh.offsets(1), mesh.coord_lvector())
    xe = mesh.coord_lvector()[mesh.offsets(1)[:, None, :] + np.arange(3)[None, :, None]]
    xq = oracle.basis_apply(nel, 3, 2, Q, Bx, Dx, qw, 0, 1, xe).reshape(nel, 3, Q ** 3)
    for forcing, ctx in (("mms", oracle.Physics(0.3, 1.0)), ("constant", (C.c_double * 3)(0.3, -1.0, 2.5))):
        fq = np.zeros((nel, 3, Q ** 3))
        for e in range(nel):
            name = "SetupMMSForce" if forcing == "mms" else "SetupConstantForce"
            f = C.cast(oracle.qf(name, oracle.default_which()), oracle.QFN)
            ins = [np.ascontiguousarray(xq[e]), np.ascontiguousarray(qdata[e])]
            inp = (C.POINTER(C.c_double) * 2)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in ins])
            outp = (C.POINTER(C.c_double) * 1)(fq[e].ctypes.data_as(C.POINTER(C.c_double)))
            assert f(C.cast(C.pointer(ctx), C.c_void_p), Q ** 3, inp, outp) == 0
        fe = oracle.basis_apply(nel, 3, P, Q, Bu, Du, qw, 1, 1, fq.reshape(nel, -1))
        ref = np.zeros(mesh.lsize(deg))
        np.add.at(ref, (mesh.offsets(deg)[:, None, :] + np.arange(3)[None, :, None]).reshape(-1), fe.reshape(-1))
        fc = c.Vector(mesh.lsize(deg))
        setuplibceed.setup_forcing(c, mesh, data, forcing, phys, (0.3, -1.0, 2.5), fc)
        got = fc.to_numpy()
        assert np.linalg.norm(got - ref) < 1e-12 * np.linalg.norm(ref), forcing


def test_coo_to_stencil_map_and_stencil_matvec_on_cpu():
    """MatSetValuesCOO stand-in (ColoredCoarseMatrix._assemble_coo) against a dense assembly of the same element
    matrices, and against the coloured assembly of the same operator; CPU tensors, random symmetric element blocks."""
    from ceedpetscsolid_b200 import matops, solver
    from ceedpetscsolid_b200.mesh import BoxMesh
    mesh = BoxMesh(n=(3, 2, 4))
    dm = matops.LevelDM(mesh, 1, bc_faces=[(2, 0)], device="cpu")
    off = mesh.offsets(1)                                  # (E, 8) component-0 dof of each element node
    E = off.shape[0]
    rng = np.random.default_rng(2)
    Ke = rng.standard_normal((E, 24, 24))
    Ke = Ke + Ke.transpose(0, 2, 1)
    eld = (off[:, :, None] + np.arange(3)[None, None, :]).reshape(E, 24)   # element dof = node*3 + comp
    n = dm.lsize
    A = np.zeros((n, n))
    for e in range(E):
        A[np.ix_(eld[e], eld[e])] += Ke[e]

    class Coo:
        elem_nodes = torch.from_numpy(off // 3)

        @staticmethod
        def values():  # CeedOperatorLinearAssemble layout: [e][col][row]
            return torch.from_numpy(np.ascontiguousarray(Ke.transpose(0, 2, 1)).reshape(-1))

    def local_apply(x, y):
        y.copy_(torch.from_numpy(A @ x.numpy()))

    m_coo = solver.ColoredCoarseMatrix(dm, local_apply, coo=Coo())
    m_col = solver.ColoredCoarseMatrix(dm, local_apply)
    m_coo.assemble()
    m_col.assemble()
    assert torch.allclose(m_coo.svals, m_col.svals, rtol=0, atol=1e-12)
    x = torch.from_numpy(rng.standard_normal(n))
    y = torch.zeros(n, dtype=torch.float64)
    m_coo.local_mult(x, y)
    assert np.allclose(y.numpy(), A @ x.numpy(), rtol=0, atol=1e-11)
    # global action and diagonal with the Dirichlet face eliminated
    X = torch.from_numpy(rng.standard_normal(dm.nglobal))
    Y = torch.zeros(dm.nglobal, dtype=torch.float64)
    m_coo.mult(X, Y)
    free = dm._fo_host
    assert np.allclose(Y.numpy(), A[np.ix_(free, free)] @ X.numpy(), rtol=0, atol=1e-11)
    D = torch.zeros(dm.nglobal, dtype=torch.float64)
    m_coo.diagonal(D)
    assert np.allclose(D.numpy(), np.diag(A)[free], rtol=0, atol=1e-12)


def test_h_multigrid_coarse_solve_is_mesh_independent_on_the_oracle_operator():
    """GAMG stand-in: one V(2,2) cycle of the geometric h-multigrid on the assembled p = 1 linear-elasticity matrix
    (oracle operator, clamped face) as CG preconditioner: the iteration count must not grow with the mesh, unlike Jacobi."""
    from ceedpetscsolid_b200 import matops
    from ceedpetscsolid_b200.elasticity import build_h_dms
    from ceedpetscsolid_b200.mesh import BoxMesh
    from helpers import PHYS
    from oracle import oracle
    its = {}
    for n in (4, 8):
        mesh = BoxMesh(n=(n, n, n), perturb=0.05, seed=0)
        dm = matops.LevelDM(mesh, 1, bc_faces=[(2, 0)], device="cpu", masked=True)
        B, D, _, _ = oracle.basis_1d(2, 2, 0)
        qdata = oracle.setup_geo(mesh.nelem, 2, mesh.offsets(1), mesh.coord_lvector())
        off = mesh.offsets(1)

        def local_apply(x, y):
            y.copy_(torch.from_numpy(oracle.operator_apply("linElas", True, PHYS, mesh.nelem, 2, 2, B, D, off, qdata, None,
                                                           x.numpy())))

        V = solver.Vec()
        V.consistent = {}
        A = solver.ColoredCoarseMatrix(dm, local_apply)
        A.assemble()
        h_dms = build_h_dms(mesh, (1, 1, 1), 0, 1, [(2, 0)], "cpu", masked=True)
        assert len(h_dms) == (1 if n == 4 else 2)
        for d in [dm] + h_dms:
            V.consistent[d.nglobal] = d.make_consistent
        hmg = solver.HMultigrid(V, A, [dm] + h_dms)
        hmg.setup()
        b = torch.from_numpy(np.random.default_rng(n).standard_normal(dm.nglobal))
        dm.zero_constrained(b)
        x = torch.zeros_like(b)
        M = lambda r, z: hmg.solve(r, z)
        k_mg, reason, _ = solver.pcg(V, A.mult, b, x, M=M, rtol=1e-8, maxit=200)
        assert reason == "rtol"
        r = torch.zeros_like(b)
        A.mult(x, r)
        assert torch.linalg.norm(r - b) < 1e-6 * torch.linalg.norm(b)
        D_ = torch.zeros_like(b)
        A.diagonal(D_)
        dinv = 1.0 / D_
        x.zero_()
        k_jac, _, _ = solver.pcg(V, A.mult, b, x, M=lambda r, z: torch.mul(dinv, r, out=z), rtol=1e-8, maxit=500)
        its[n] = (k_mg, k_jac)
    assert its[8][0] <= its[4][0] + 4, its          # multigrid: flat
    assert its[8][1] >= 1.5 * its[4][1], its        # Jacobi: grows like 1/h
    assert its[8][0] < its[8][1] / 3, its


def test_sync_free_pcg_survives_exact_convergence_inside_a_check_block():
    """A diagonal operator makes Jacobi-PCG exact after one iteration; the loop only looks at the residual every
    check_every iterations, so the remaining iterations of the block run with r = 0 (rz = pAp = 0).  The step
    lengths are guarded: the solution stays exact instead of turning into NaN (coarsest h-multigrid levels with a
    single interior node hit exactly this)."""
    V = solver.Vec()
    d = torch.tensor([2.0, 2.0, 2.0], dtype=torch.float64)
    A = lambda x, y: y.copy_(d * x)
    b = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    x = torch.zeros(3, dtype=torch.float64)
    work = [torch.zeros(3, dtype=torch.float64) for _ in range(4)]
    its = solver.jacobi_pcg_nosync(V, A, 1.0 / d, b, x, work, rtol=1e-12, maxit=50)
    assert its == 10
    assert torch.equal(x, b / d)


@pytest.mark.gpu
def test_sync_free_pcg_survives_exact_convergence_on_the_device():
    V = solver.Vec()
    d = torch.tensor([2.0, 2.0, 2.0], dtype=torch.float64, device="cuda")
    A = lambda x, y: y.copy_(d * x)
    b = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device="cuda")
    x = torch.zeros_like(b)
    work = [torch.zeros_like(b) for _ in range(4)]
    its = solver.jacobi_pcg_nosync(V, A, 1.0 / d, b, x, work, rtol=1e-12, maxit=50)
    assert its == 10 and torch.equal(x, b / d)


def test_dot_product_weights_of_equal_length_must_agree():
    """Vec looks inner-product weights up by vector length: registering a DIFFERENT weight for a length that already
    has one is refused (it would silently change every dot product of the first DM)."""
    V = solver.Vec()
    w1 = torch.tensor([1.0, 0.5, 1.0])
    V.set_weight(3, w1)
    V.set_weight(3, w1.clone())                      # same values: fine
    with pytest.raises(ValueError, match="different dot-product weights"):
        V.set_weight(3, torch.tensor([1.0, 1.0, 0.5]))
    assert V.dot(torch.ones(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)) == 2.5
