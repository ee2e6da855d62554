"""Partitioned ApplyJacobian_Ceed (one rank per GPU, halo exchange) vs. the serial CPU oracle on the same global mesh.

Run under torchrun (rank 0 prints PASS/FAIL and the exit code is non-zero on failure):
    torchrun --standalone --nproc-per-node 2 tests/mgpu_check.py
Used by tests/test_gpu_multi.py, and -- through partitioned_parity() -- by bench.py, which puts the result into the
"parity" key of its JSON line so that every driver-run scaling record carries a multi-GPU parity proof.
Environment: MGPU_SHARED = 0 | 1 | masked (DM layout), MGPU_HALO = nccl | p2p.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def partitioned_parity(rank, world, local, layout="masked", halo_mode="p2p", problem="hyperFS", p=2, n=None, repeats=5,
                       overlap=True):
    """Returns (rel_err, ok, description) on every rank (computed on rank 0 and broadcast).  dist must be initialised
    with the nccl backend when world > 1; world == 1 checks the same code path without a halo."""
    from ceedpetscsolid_b200 import ceed as libceed
    from ceedpetscsolid_b200 import matops, setuplibceed
    from ceedpetscsolid_b200.halo import Halo
    from ceedpetscsolid_b200.mesh import BoxMesh, grid_for, smooth_displacement

    masked = layout == "masked"
    shared = layout in ("1", "masked", "shared")
    if n is None:
        n = (12, 10, 10) if masked else (6, 4, 4)
    grid = grid_for(world)
    gmesh = BoxMesh(n=n, perturb=0.08, seed=0)
    mesh = gmesh.brick(grid, rank, interface_first=masked and overlap) if world > 1 else gmesh
    ceed = libceed.Ceed(f"/gpu/b200:device_id={local}")
    degrees, data, phys = setuplibceed.setup_all(ceed, mesh, problem, p)
    fine = len(degrees) - 1
    halo = None
    if world > 1:
        halo = Halo(gmesh, grid, rank, p, dist)
        if halo_mode == "p2p" and not halo.enable_p2p():
            return float("nan"), False, False, f"peer-memory halo could not be set up: {halo._p2p_error}"
    dm = matops.LevelDM(mesh, p, bc_faces="all", halo=halo, shared=shared, masked=masked)
    user = matops.setup_jacobian_ctx(dm, ceed, data[fine], phys)
    user.overlap = overlap
    old_min, matops.OVERLAP_MIN_INTERIOR = matops.OVERLAP_MIN_INTERIOR, 0   # exercise the overlapped exchange on this small mesh
    if world > 1 and masked and overlap:
        assert 0 < mesh.n_interface < mesh.nelem
    u = torch.from_numpy(smooth_displacement(mesh.node_coords(p)).reshape(-1)).cuda()
    uc, rc = ceed.Vector(u.numel()), ceed.Vector(u.numel())
    r = torch.zeros_like(u)
    uc.set_array(u); rc.set_array(r)
    data[fine].opApply.apply(uc, rc)
    uc.take_array(); rc.take_array()

    # global numbering of this rank's free owned dofs
    N, GN = mesh.nodes_per_dim(p), gmesh.nodes_per_dim(p)
    oz, oy, ox = (mesh.origin[2] * p, mesh.origin[1] * p, mesh.origin[0] * p)
    z, y, x = np.meshgrid(np.arange(N[2]) + oz, np.arange(N[1]) + oy, np.arange(N[0]) + ox, indexing="ij")
    gid = (x + GN[0] * (y + GN[1] * z)).reshape(-1)
    gdof = (gid[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    mydofs = gdof[dm.free_owned_idx.cpu().numpy()]

    xglob = np.random.default_rng(5).standard_normal(gmesh.lsize(p))
    gbc = np.repeat(gmesh.boundary_mask(p, "all"), 3)
    xglob[gbc] = 0.0
    X, Y = dm.create_global_vector(), dm.create_global_vector()
    X.copy_(torch.from_numpy(xglob[mydofs]))
    for _ in range(repeats):   # several exchanges: generation counter and window parity
        matops.ApplyJacobian_Ceed(user, X, Y)
    torch.cuda.synchronize()
    matops.OVERLAP_MIN_INTERIOR = old_min
    if halo is not None and halo.p2p_failed_anywhere():
        halo.close()
        return float("nan"), False, False, "a peer-memory halo exchange timed out waiting for a neighbour"
    # gather (dof id, value) pairs on rank 0
    parts = [(mydofs, Y.cpu().numpy())]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (mydofs, Y.cpu().numpy()))
    err, ok, bitwise = float("nan"), True, True
    if rank == 0:
        from helpers import OracleProblem, rel_err
        o = OracleProblem(problem, n, p)
        yref = o.jacobian(xglob)
        ypar = np.full_like(yref, np.nan)
        seen = np.zeros(yref.size, dtype=int)
        for d, v in parts:
            again = seen[d] > 0
            bitwise = bitwise and bool(np.array_equal(ypar[d][again], v[again]))   # interface copies: bit-identical
            ypar[d] = v
            seen[d] += 1
        free = ~gbc
        err = float(rel_err(ypar[free], yref[free]))
        ok = bool(np.all(seen[free] >= 1) and (shared or np.all(seen[free] == 1)) and err < 1e-12
                  and (np.all(ypar[gbc] == 0.0) if masked else np.all(seen[gbc] == 0)) and (bitwise or not shared))
    if world > 1:
        t = torch.tensor([err, 1.0 if ok else 0.0, 1.0 if bitwise else 0.0], dtype=torch.float64, device="cuda")
        dist.broadcast(t, 0)
        err, ok, bitwise = float(t[0].item()), bool(t[1].item() > 0.5), bool(t[2].item() > 0.5)
    if halo is not None:
        halo.close()
    desc = (f"{problem} degree {p} Jacobian MatMult on a {n[0]}x{n[1]}x{n[2]} box, bricks {'x'.join(map(str, grid))}, "
            f"{'masked' if masked else ('shared' if shared else 'owner/ghost')} layout, "
            f"halo {'none' if world == 1 else halo_mode}, {repeats} applies, vs the serial CPU oracle")
    return err, ok, bitwise, desc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    layout = os.environ.get("MGPU_SHARED", "0")
    halo_mode = os.environ.get("MGPU_HALO", "nccl")
    err, ok, bitwise, desc = partitioned_parity(rank, world, local, layout=layout, halo_mode=halo_mode,
                                                repeats=5 if halo_mode == "p2p" else 1)
    if rank == 0:
        print(f"mgpu_check world={world}: {desc}: rel err {err:.2e}, interface copies bit-identical: {bitwise} "
              f"-> {'PASS' if ok else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
