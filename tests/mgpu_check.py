"""Run under torchrun (one rank per GPU): partitioned ApplyJacobian_Ceed (NCCL halo exchange)
vs. the serial CPU oracle on the same global mesh.  Rank 0 prints PASS/FAIL and exits non-zero
on failure.  Used by tests/test_gpu_multi.py and by hand:
    torchrun --standalone --nproc-per-node 2 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ceedpetscsolid_b200 import ceed as libceed
    from ceedpetscsolid_b200 import matops, setuplibceed
    from ceedpetscsolid_b200.halo import Halo
    from ceedpetscsolid_b200.mesh import BoxMesh, grid_for, smooth_displacement

    problem, p, n = "hyperFS", 2, ((12, 10, 10) if os.environ.get("MGPU_SHARED", "0") == "masked" else (6, 4, 4))
    grid = grid_for(world)
    gmesh = BoxMesh(n=n, perturb=0.08, seed=0)
    mesh = gmesh.brick(grid, rank, interface_first=os.environ.get("MGPU_SHARED", "0") == "masked")
    ceed = libceed.Ceed(f"/gpu/b200:device_id={local}")
    degrees, data, phys = setuplibceed.setup_all(ceed, mesh, problem, p)
    fine = len(degrees) - 1
    halo = Halo(gmesh, grid, rank, p, dist)
    p2p = os.environ.get("MGPU_HALO", "nccl") == "p2p"
    if p2p:
        halo.enable_p2p()
    shared = os.environ.get("MGPU_SHARED", "0") in ("1", "masked")
    masked = os.environ.get("MGPU_SHARED", "0") == "masked"
    dm = matops.LevelDM(mesh, p, bc_faces="all", halo=halo, shared=shared, masked=masked)
    user = matops.setup_jacobian_ctx(dm, ceed, data[fine], phys)
    matops.OVERLAP_MIN_INTERIOR = 0   # exercise the overlapped exchange even on this small mesh
    assert not masked or 0 < mesh.n_interface < mesh.nelem
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
    for _ in range(5 if p2p else 1):   # several exchanges: generation counters and window parity
        matops.ApplyJacobian_Ceed(user, X, Y)
    torch.cuda.synchronize()
    halo.check_p2p()
    # gather (dof id, value) pairs on rank 0
    parts = [None] * world
    dist.all_gather_object(parts, (mydofs, Y.cpu().numpy()))
    ok = True
    if rank == 0:
        from helpers import OracleProblem, rel_err
        o = OracleProblem(problem, n, p)
        yref = o.jacobian(xglob)
        ypar = np.zeros_like(yref)
        seen = np.zeros(yref.size, dtype=int)
        for d, v in parts:
            ypar[d] = v
            seen[d] += 1
        free = ~gbc
        err = rel_err(ypar[free], yref[free])
        ok = bool(np.all(seen[free] >= 1) and (shared or np.all(seen[free] == 1)) and err < 1e-12
                  and (np.all(ypar[gbc] == 0.0) if masked else np.all(seen[gbc] == 0)))
        print(f"mgpu_check world={world} bricks={grid} shared={shared} masked={masked} halo={'p2p' if p2p else 'nccl'}: rel err {err:.2e} -> {'PASS' if ok else 'FAIL'}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
