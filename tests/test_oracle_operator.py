"""Validate the libCEED-layer restatement in oracle/ceed_oracle.c through properties
(the reference holds no golden vectors for this layer: SURVEY.md 8(c), "parity unpinned")."""
import numpy as np
import pytest

from helpers import PHYS, OracleProblem, rel_err
from oracle import oracle


@pytest.mark.parametrize("Q", [2, 3, 4, 5, 6, 8])
def test_quadrature_exactness(Q):
    x, w = oracle.gauss(Q)
    assert np.all(np.diff(x) > 0)
    for k in range(2 * Q):
        exact = 0.0 if k % 2 else 2.0 / (k + 1)
        assert abs(np.dot(w, x ** k) - exact) < 1e-14
    xl, wl = oracle.lobatto(Q)
    assert xl[0] == -1 and xl[-1] == 1
    for k in range(2 * Q - 2):
        exact = 0.0 if k % 2 else 2.0 / (k + 1)
        assert abs(np.dot(wl, xl ** k) - exact) < 1e-14


@pytest.mark.parametrize("P,Q", [(2, 3), (3, 3), (2, 4), (3, 4), (4, 4), (2, 5), (3, 5), (5, 5), (5, 6)])
def test_basis_reproduces_monomials(P, Q):
    B, D, qr, _ = oracle.basis_1d(P, Q, 0)
    nodes, _ = oracle.lobatto(P)
    for k in range(P):
        np.testing.assert_allclose(B @ nodes ** k, qr ** k, atol=1e-13)
        dk = k * qr ** (k - 1) if k else np.zeros(Q)
        np.testing.assert_allclose(D @ nodes ** k, dk, atol=1e-12)


@pytest.mark.parametrize("P,Q", [(2, 3), (3, 4), (5, 5), (3, 5)])
def test_tensor_basis_vs_dense_kronecker_and_adjoint(P, Q):
    rng = np.random.default_rng(0)
    B, D, _, qw = oracle.basis_1d(P, Q, 0)
    nelem, ncomp = 3, 3
    u = rng.standard_normal((nelem, ncomp, P ** 3))
    v = oracle.basis_apply(nelem, ncomp, P, Q, B, D, qw, 0, 2, u).reshape(nelem, 3, ncomp, Q ** 3)
    G = [np.kron(B, np.kron(B, D)), np.kron(B, np.kron(D, B)), np.kron(D, np.kron(B, B))]  # z,y,x order
    for d in range(3):
        ref = np.einsum("qn,ecn->ecq", G[d], u)
        np.testing.assert_allclose(v[:, d], ref, atol=1e-12)
    vi = oracle.basis_apply(nelem, ncomp, P, Q, B, D, qw, 0, 1, u).reshape(nelem, ncomp, Q ** 3)
    np.testing.assert_allclose(vi, np.einsum("qn,ecn->ecq", np.kron(B, np.kron(B, B)), u), atol=1e-12)
    w = rng.standard_normal(v.shape)
    ut = oracle.basis_apply(nelem, ncomp, P, Q, B, D, qw, 1, 2, w.reshape(nelem, -1)).reshape(u.shape)
    assert abs(np.sum(ut * u) - np.sum(v * w)) < 1e-11 * abs(np.sum(v * w))
    wt = oracle.basis_apply(nelem, 1, P, Q, B, D, qw, 0, 4, None)
    np.testing.assert_allclose(wt[0], np.kron(qw, np.kron(qw, qw)), atol=1e-15)


def test_geometric_factors_sum_to_volume():
    pr = OracleProblem("linElas", 3, 2, perturb=0.2)
    assert abs(pr.qdata[:, 0, :].sum() - 1.0) < 1e-13


@pytest.mark.parametrize("problem,p", [("linElas", 2), ("hyperSS", 2), ("hyperFS", 2), ("hyperFS", 3)])
def test_jacobian_is_derivative_of_residual(problem, p):
    pr = OracleProblem(problem, 2, p)
    rng = np.random.default_rng(1)
    d = rng.standard_normal(pr.lsize)
    u0 = pr.u_fine.copy()
    h = 1e-6
    fp = pr.residual_fine(u0 + h * d).copy()
    fm = pr.residual_fine(u0 - h * d).copy()
    pr.residual_fine(u0)  # restore gradu at the linearisation point (SURVEY hard part 8)
    jd = pr.jacobian(d)
    assert rel_err(jd, (fp - fm) / (2 * h)) < 5e-9


@pytest.mark.parametrize("problem", ["linElas", "hyperSS", "hyperFS"])
def test_jacobian_symmetric_and_rigid_translation(problem):
    pr = OracleProblem(problem, 2, 3, pl=2)
    rng = np.random.default_rng(2)
    v, w = rng.standard_normal(pr.lsize), rng.standard_normal(pr.lsize)
    a, b = np.dot(v, pr.jacobian(w)), np.dot(w, pr.jacobian(v))
    assert abs(a - b) < 1e-13 * abs(a)
    t = np.tile([0.3, -0.2, 0.7], pr.lsize // 3)
    assert np.linalg.norm(pr.jacobian(t)) < 1e-13 * np.linalg.norm(pr.jacobian(v))


@pytest.mark.parametrize("problem,p,pl", [("linElas", 2, 2), ("hyperFS", 2, 1), ("hyperSS", 3, 2)])
def test_diagonal_equals_unit_vector_probing(problem, p, pl):
    pr = OracleProblem(problem, 2, p, pl=pl)
    diag = pr.diagonal()
    rng = np.random.default_rng(3)
    for i in rng.choice(pr.lsize, 12, replace=False):
        e = np.zeros(pr.lsize)
        e[i] = 1.0
        assert abs(pr.jacobian(e)[i] - diag[i]) < 1e-13 * np.max(np.abs(diag))


def test_arbitrary_offsets_permutation_is_consistent():
    """PETSc numbers dofs by mesh point, not lexicographically (SURVEY hard part 4)."""
    a = OracleProblem("hyperFS", 2, 2)
    b = OracleProblem("hyperFS", 2, 2, node_perm_seed=5)
    x = np.random.default_rng(4).standard_normal(a.lsize)
    np.testing.assert_allclose(b.jacobian(b.to_perm(x)), b.to_perm(a.jacobian(x)), rtol=0, atol=1e-14)


def test_port_and_reference_qfunctions_agree_through_the_operator():
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    x = np.random.default_rng(6).standard_normal(OracleProblem("hyperFS", 2, 2).lsize)
    ya = OracleProblem("hyperFS", 2, 2, which="ref").jacobian(x)
    yb = OracleProblem("hyperFS", 2, 2, which="port").jacobian(x)
    assert rel_err(yb, ya) < 1e-14


def test_transfer_adjoint_and_constants():
    pr = OracleProblem("linElas", 2, 2)
    offc, offf = pr.mesh.offsets(1), pr.mesh.offsets(2)
    lc, lf = pr.mesh.lsize(1), pr.mesh.lsize(2)
    mult = oracle.multiplicity(pr.nelem, 27, 3, lf, offf)
    ones_c = np.ones(lc)
    pf = oracle.transfer(False, pr.nelem, 2, 3, offc, offf, ones_c, lf) / mult
    np.testing.assert_allclose(pf, 1.0, atol=1e-14)
    rng = np.random.default_rng(8)
    c, f = rng.standard_normal(lc), rng.standard_normal(lf)
    lhs = np.dot(f, oracle.transfer(False, pr.nelem, 2, 3, offc, offf, c, lf))
    rhs = np.dot(c, oracle.transfer(True, pr.nelem, 2, 3, offc, offf, f, lc))
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
