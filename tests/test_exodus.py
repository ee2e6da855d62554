"""-mesh file.exo (setupdm.c:40-68): Exodus II reader/writer, topological high-order numbering on unstructured hex8
meshes, side-set boundary conditions; the operators and the solver on a tube mesh against the oracle."""
import glob
import os

import numpy as np
import pytest

from ceedpetscsolid_b200.elasticity import AppCtx
from ceedpetscsolid_b200.exodus import HexMesh, read_exodus, tube_mesh, write_exodus
from ceedpetscsolid_b200.mesh import BoxMesh
from helpers import OracleProblem, rel_err

REF_MESHES = sorted(glob.glob("/root/reference/meshes/Tube8_*.exo") + glob.glob("/root/reference/meshes/cyl-hole_672e_*.exo")
                    + glob.glob("/root/reference/meshes/cube8_8e_6ss_s.exo"))


def _entity_counts(conn):
    E_loc = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
    F_loc = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 2, 4, 6), (1, 3, 5, 7)]
    edges = {tuple(sorted((e[a], e[b]))) for e in conn for a, b in E_loc}
    faces = {tuple(sorted(e[list(f)])) for e in conn for f in F_loc}
    return len(edges), len(faces)


@pytest.mark.parametrize("angle", [2 * np.pi, np.pi / 2])
def test_topological_numbering_counts_entities_and_agrees_on_coordinates(angle):
    m = tube_mesh(2, 8, 3, angle=angle)
    ne, nf = _entity_counts(m.connect)
    for p in (1, 2, 3, 4):
        ids, nn = m._level(p)
        assert nn == m.vertices.shape[0] + (p - 1) * ne + (p - 1) ** 2 * nf + (p - 1) ** 3 * m.nelem
        assert sorted(np.unique(ids)) == list(range(nn))
        # every element that holds a node computes the same physical point for it
        from ceedpetscsolid_b200.mesh import gll_nodes
        P = p + 1
        r = (gll_nodes(P) + 1) / 2
        w1 = np.stack([1 - r, r], axis=1)
        W = np.einsum("ai,bj,ck->cbakji", w1, w1, w1).reshape(P ** 3, 8)
        xe = np.einsum("nv,evd->end", W, m.vertices[m.connect])
        assert np.abs(xe - m.node_coords(p)[ids]).max() < 1e-14
        on1 = m.node_coords(p)[m.boundary_mask(p, [1])]
        assert on1.shape[0] > 0 and np.all(np.abs(on1[:, 2]) < 1e-14)          # side set 1 is the z = 0 face


def test_exodus_round_trip(tmp_path):
    m = tube_mesh(1, 6, 2)
    f = str(tmp_path / "tube.exo")
    write_exodus(f, m.vertices, m.connect, m.sidesets)
    d = read_exodus(f)
    assert np.array_equal(d["connect"], m.connect) and np.allclose(d["coords"], m.vertices)
    for k in m.sidesets:
        assert np.array_equal(d["sidesets"][k][0], m.sidesets[k][0]) and np.array_equal(d["sidesets"][k][1], m.sidesets[k][1])


@pytest.mark.skipif(not REF_MESHES, reason="/root/reference not mounted")
def test_reference_meshes_are_readable():
    for f in REF_MESHES:
        hm = HexMesh.from_file(f)
        X = hm.vertices[hm.connect]
        det = np.einsum("ei,ei->e", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 4] - X[:, 0])
        assert np.all(det > 0), f                                               # tensor ordering is right-handed
        for sid in hm.sidesets:
            el, sd = hm.sidesets[sid]
            bel, bsd = hm.boundary_faces()
            ext = set(zip(bel.tolist(), bsd.tolist()))
            assert all((a, b) in ext for a, b in zip(el.tolist(), sd.tolist())), (f, sid)   # side sets lie on the boundary
    with pytest.raises(ValueError, match="HEX8"):
        read_exodus("/root/reference/meshes/cylinder27_1440e_1ns_us.exo")


def test_box_as_unstructured_mesh_gives_the_same_operator():
    """a box handed over as (coords, connectivity): different node numbering, same Jacobian action"""
    b = BoxMesh(n=(3, 2, 2), perturb=0.08, seed=0)
    conn = b.offsets(1) // 3
    h = HexMesh(b.vertices.reshape(-1, 3), conn)
    p = 3
    ob, oh = OracleProblem("hyperFS", (3, 2, 2), p), OracleProblem("hyperFS", None, p, mesh=h)
    # map box nodes -> unstructured nodes by coordinates
    cb, ch = b.node_coords(p), h.node_coords(p)
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(ch).query(cb)
    assert dist.max() < 1e-12 and len(set(idx)) == cb.shape[0]
    x = np.random.default_rng(0).standard_normal(ob.lsize)
    xh = np.zeros_like(x).reshape(-1, 3)
    xh[idx] = x.reshape(-1, 3)
    yb, yh = ob.jacobian(x), oh.jacobian(xh.reshape(-1))
    assert rel_err(yh.reshape(-1, 3)[idx], yb.reshape(-1, 3)) < 1e-12


def _tube_app(problem="hyperFS", degree=2, steps=2):
    m = tube_mesh(1, 6, 2, r0=0.6, r1=1.0, length=1.0)
    return AppCtx(problem=problem, degree=degree, num_steps=steps, mesh=m,
                  clamp={1: [0, 0, 0, 0, 0, 1, 0], 2: [0.02, 0, -0.05, 0, 0, 1, 0.05]})


def test_solver_on_a_tube_with_side_set_clamps_converges_on_the_oracle():
    from oracle_levels import oracle_solve
    out, U = oracle_solve(_tube_app())
    assert out["converged"] and out["snes_its"] >= 2 and out["ksp_its"] > 0


@pytest.mark.gpu
def test_tube_gpu_matches_oracle():
    from ceedpetscsolid_b200.elasticity import Elasticity
    from gpu_helpers import GpuProblem
    from oracle_levels import oracle_solve
    # operators on the unstructured numbering
    m = tube_mesh(2, 8, 3)
    for problem, p in (("hyperFS", 4), ("hyperSS", 3)):
        g, o = GpuProblem(problem, None, p, mesh=m), OracleProblem(problem, None, p, mesh=m)
        assert rel_err(g.residual(), o.residual_fine(o.u_fine)) < 1e-12
        x = np.random.default_rng(1).standard_normal(o.lsize)
        fine = len(g.degrees) - 1
        assert g.data[fine].opJacob.is_fused
        assert rel_err(g.jacobian(fine, x), o.jacobian(x)) < 1e-12
        assert rel_err(g.diagonal(fine), o.diagonal()) < 1e-12
    # Newton-Krylov-p-MG with the ELL coarse matrix from CeedOperatorLinearAssemble
    app = _tube_app()
    ref, Uref = oracle_solve(app)
    el = Elasticity(app)
    from ceedpetscsolid_b200.solver import SparseCoarseMatrix
    assert isinstance(el.pc.coarse, SparseCoarseMatrix) and el.pc.coarse.coo is not None and el.pc.hmg is None
    out = el.solve()
    assert out["converged"] and ref["converged"]
    assert (out["snes_its"], out["ksp_its"]) == (ref["snes_its"], ref["ksp_its"]), (out, ref)
    u = el.U.cpu().numpy()
    assert np.linalg.norm(u - Uref.numpy()) < 1e-9 * np.linalg.norm(Uref.numpy())
    assert el.strain_energy() > 0


def test_ell_coarse_matrix_from_element_matrices_on_cpu():
    """SparseCoarseMatrix: COO element matrices -> ELL (MatSetValuesCOO stand-in on an unstructured mesh)"""
    import torch
    from ceedpetscsolid_b200 import matops, solver
    m = tube_mesh(2, 7, 2)
    dm = matops.LevelDM(m, 1, bc_faces=[1], device="cpu")
    off = m.offsets(1)
    E = off.shape[0]
    rng = np.random.default_rng(5)
    Ke = rng.standard_normal((E, 24, 24))
    Ke = Ke + Ke.transpose(0, 2, 1)
    eld = (off[:, :, None] + np.arange(3)[None, None, :]).reshape(E, 24)
    n = dm.lsize
    A = np.zeros((n, n))
    for e in range(E):
        A[np.ix_(eld[e], eld[e])] += Ke[e]

    class Coo:
        elem_nodes = torch.from_numpy(off // 3)

        @staticmethod
        def values():
            return torch.from_numpy(np.ascontiguousarray(Ke.transpose(0, 2, 1)).reshape(-1))

    sm = solver.SparseCoarseMatrix(dm, None, coo=Coo())
    sm.assemble()
    x = torch.from_numpy(rng.standard_normal(n))
    y = torch.zeros(n, dtype=torch.float64)
    sm.local_mult(x, y)
    assert np.allclose(y.numpy(), A @ x.numpy(), rtol=0, atol=1e-11)
    D = torch.zeros(dm.nglobal, dtype=torch.float64)
    sm.diagonal(D)
    assert np.allclose(D.numpy(), np.diag(A)[dm._fo_host], rtol=0, atol=1e-12)
    assert sm.nslots <= 81 and sm.nslots >= 27 * 3 // 3


@pytest.mark.skipif(not os.path.exists("/root/reference/meshes/Tube8_32e_2ss_us.exo"), reason="/root/reference not mounted")
def test_reference_tube_mesh_solves_with_its_side_sets_clamped():
    """README.rst:63 style run (`-mesh Tube8_... -bc_clamp 998,999 -bc_clamp_999_translate ...`) on the oracle operators"""
    from oracle_levels import oracle_solve
    m = HexMesh.from_file("/root/reference/meshes/Tube8_32e_2ss_us.exo")
    ext = float(np.ptp(m.vertices, axis=0).max())
    app = AppCtx(problem="hyperFS", degree=2, num_steps=2, mesh=m, nu=0.3, E=1e6,
                 clamp={998: [0, 0, 0, 0, 0, 1, 0], 999: [0, -0.02 * ext, 0.03 * ext, 0, 0, 1, 0]})
    out, U = oracle_solve(app)
    assert out["converged"] and out["snes_its"] <= 8
    # the clamped end moved by the prescribed translation, the fixed end did not: check through the residual's state
    assert float(U.abs().max()) > 0.02 * ext
