"""Strain-energy and nodal-diagnostic operators (setuplibceed.c:645-737, matops.c:247-300, misc.c:217-300) through
the generic device path, against the oracle: basis evaluation, the reference's Energy / Diagnostic QFunctions,
transposed interpolation / collocated output, restriction transpose.  FP64, 1e-12."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _oracle_post(problem, mesh, p, u, which=None):
    from oracle import oracle
    which = which or oracle.default_which()
    name = {"linElas": "LinElas", "hyperSS": "HyperSS", "hyperFS": "HyperFS"}[problem]
    P = Q = p + 1
    nel = mesh.nelem
    phys = oracle.Physics(0.3, 1.0)
    off = mesh.offsets(p)
    ue = u[(off[:, None, :] + np.arange(3)[None, :, None])]                       # (nel, 3, P^3)
    # ---- energy: Gauss points
    B, D, _, qw = oracle.basis_1d(P, Q, 0)
    qdata = oracle.setup_geo(nel, Q, mesh.offsets(1), mesh.coord_lvector(), which)
    dq = oracle.basis_apply(nel, 3, P, Q, B, D, qw, 0, 2, ue.reshape(nel, -1)).reshape(nel, 9, Q ** 3)
    eq = np.zeros((nel, Q ** 3))
    for e in range(nel):
        (o,) = oracle.call_qf(name + "Energy", which, phys, Q ** 3, [np.ascontiguousarray(dq[e]), np.ascontiguousarray(qdata[e])], [1])
        eq[e] = o.reshape(-1)
    ee = oracle.basis_apply(nel, 1, P, Q, B, D, qw, 1, 1, eq)                      # (nel, P^3)
    energy = float(ee.sum())
    # ---- diagnostics: collocated at the GLL nodes (basis P -> P, Gauss-Lobatto), own geometric factors
    Bd, Dd, _, qwd = oracle.basis_1d(P, P, 1)
    Bx, Dx, _, _ = oracle.basis_1d(2, P, 1)
    xe = mesh.coord_lvector()[mesh.offsets(1)[:, None, :] + np.arange(3)[None, :, None]]
    dxq = oracle.basis_apply(nel, 3, 2, P, Bx, Dx, qwd, 0, 2, xe.reshape(nel, -1)).reshape(nel, 9, P ** 3)
    wq = oracle.basis_apply(nel, 3, 2, P, Bx, Dx, qwd, 0, 4, None).reshape(nel, 1, P ** 3)
    uq = oracle.basis_apply(nel, 3, P, P, Bd, Dd, qwd, 0, 1, ue.reshape(nel, -1)).reshape(nel, 3, P ** 3)
    duq = oracle.basis_apply(nel, 3, P, P, Bd, Dd, qwd, 0, 2, ue.reshape(nel, -1)).reshape(nel, 9, P ** 3)
    nn = mesh.num_nodes(p)
    acc, mult = np.zeros((nn, 8)), np.zeros(nn)
    for e in range(nel):
        (qd,) = oracle.call_qf("SetupGeo", which, None, P ** 3, [np.ascontiguousarray(dxq[e]), np.ascontiguousarray(wq[e])], [10])
        (dg,) = oracle.call_qf(name + "Diagnostic", which, phys, P ** 3,
                               [np.ascontiguousarray(uq[e]), np.ascontiguousarray(duq[e]), qd], [8])
        nodes = off[e] // 3
        np.add.at(acc, nodes, dg.T)
        np.add.at(mult, nodes, 1.0)
    return energy, acc / mult[:, None]


@pytest.mark.parametrize("problem,p,n", [("linElas", 2, (3, 2, 2)), ("hyperSS", 3, (2, 2, 3)), ("hyperFS", 2, (3, 3, 2)),
                                         ("hyperFS", 4, (2, 2, 2))])
def test_energy_and_diagnostics_match_the_oracle(problem, p, n):
    import torch
    from ceedpetscsolid_b200 import matops, setuplibceed
    from gpu_helpers import GpuProblem
    g = GpuProblem(problem, n, p, scale=0.05)
    data = g.fine
    setuplibceed.setup_energy(g.ceed, g.mesh, problem, data, g.phys)
    setuplibceed.setup_diagnostic(g.ceed, g.mesh, problem, data, g.phys)
    assert not data.opEnergy.is_fused and not data.opDiagnostic.is_fused       # generic device path
    dm = matops.LevelDM(g.mesh, p, bc_faces=None)                              # no constraints: X is the whole state
    user = matops.setup_jacobian_ctx(dm, g.ceed, data, g.phys)
    X = torch.from_numpy(g.u_fine.copy()).cuda()
    energy = matops.ComputeStrainEnergy(user, data.opEnergy, X)
    diag = matops.ComputeDiagnosticQuantities(user, data.opDiagnostic, data.ErestrictDiagnostic, X).cpu().numpy()
    e_ref, d_ref = _oracle_post(problem, g.mesh, p, g.u_fine)
    assert abs(energy - e_ref) < TOL * abs(e_ref)
    assert rel_err(diag[:, :3], g.u_fine.reshape(-1, 3)) < TOL                  # displacement passes through
    for k in range(3, 8):
        assert rel_err(diag[:, k], d_ref[:, k]) < TOL, k
    if problem == "hyperFS":
        assert np.all(diag[:, 6] > 0.5)                                        # volume ratio J around 1
