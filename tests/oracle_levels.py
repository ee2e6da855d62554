"""Oracle-backed level operations for the solver harness (tests only): the same interface as
ceedpetscsolid_b200.elasticity.GpuLevel / GpuTransfer, with every operator evaluated by the CPU oracle."""
import numpy as np
import torch

from ceedpetscsolid_b200 import matops, setuplibceed, solver
from ceedpetscsolid_b200.elasticity import Elasticity
from ceedpetscsolid_b200.mesh import BoxMesh, grid_for
from oracle import oracle


class OracleShared:
    def __init__(self, app, masked=False, dist=None, rank=0, world=1):
        self.app, self.masked = app, masked
        self.dist, self.rank, self.world = dist, rank, world
        self.gmesh = app.mesh if getattr(app, "mesh", None) is not None else BoxMesh(n=app.n, perturb=app.perturb, seed=0)
        self.grid = grid_for(world)
        self.mesh = self.gmesh.brick(self.grid, rank) if world > 1 else self.gmesh
        self.degrees = setuplibceed.level_degrees(app.degree, app.multigrid)
        p = app.degree
        self.Q = p + 1 + app.qextra
        self.qdata = oracle.setup_geo(self.mesh.nelem, self.Q, self.mesh.offsets(1), self.mesh.coord_lvector())
        self.has_gradu = oracle.PROBLEMS[app.problem][2]
        self.gradu = np.zeros((self.mesh.nelem, 9, self.Q ** 3)) if self.has_gradu else None
        self.phys = (app.nu, app.E)


class OracleLevel:
    def __init__(self, sh, level, is_fine):
        self.sh = sh
        app, mesh = sh.app, sh.mesh
        self.deg = sh.degrees[level]
        self.P = self.deg + 1
        self.B, self.D, _, _ = oracle.basis_1d(self.P, sh.Q, 0)
        self.off = mesh.offsets(self.deg)
        halo = None
        if sh.world > 1:
            from ceedpetscsolid_b200.halo import Halo
            halo = Halo(sh.gmesh, sh.grid, sh.rank, self.deg, sh.dist, device="cpu")
        self.dm = matops.LevelDM(mesh, self.deg, bc_faces=list(app.clamp.keys()), halo=halo, device="cpu", masked=sh.masked,
                                 shared=True)
        self.n, self.device = self.dm.nglobal, self.dm.device
        self.Xloc, self.Yloc = self.dm.create_local_vector(matops.MEM_HOST), self.dm.create_local_vector(matops.MEM_HOST)
        self.bc_values = Elasticity._bc_values_fn(self.dm, app) if is_fine else None

    def local_apply(self, xloc, yloc):
        sh = self.sh
        y = oracle.operator_apply(sh.app.problem, True, sh.phys, sh.mesh.nelem, self.P, sh.Q, self.B, self.D, self.off,
                                  sh.qdata, sh.gradu, xloc.numpy())
        yloc.copy_(torch.from_numpy(y))

    def jacobian(self, X, Y):
        self.dm.zero_and_global_to_local(X, self.Xloc)
        self.local_apply(self.Xloc, self.Yloc)
        self.dm.local_to_global(self.Yloc, Y)

    def diagonal(self, D):
        sh = self.sh
        d = oracle.operator_diagonal(sh.app.problem, sh.phys, sh.mesh.nelem, self.P, sh.Q, self.B, self.D, self.off,
                                     sh.qdata, sh.gradu, self.dm.lsize)
        self.dm.local_to_global(torch.from_numpy(d), D)
        self.dm.fix_diagonal(D)

    def bc_increment_rhs(self, F, load_prev, load):
        self.Xloc.zero_()
        self.dm.insert_boundary_values(self.Xloc, self.bc_values(load) - self.bc_values(load_prev))
        self.local_apply(self.Xloc, self.Yloc)
        self.dm.local_to_global(self.Yloc, F)

    def residual(self, U, F, load):
        sh = self.sh
        self.Xloc.zero_()
        self.dm.insert_boundary_values(self.Xloc, self.bc_values(load))
        self.dm.global_to_local(U, self.Xloc)
        y = oracle.operator_apply(sh.app.problem, False, sh.phys, sh.mesh.nelem, self.P, sh.Q, self.B, self.D, self.off,
                                  sh.qdata, sh.gradu, self.Xloc.numpy())
        self.dm.local_to_global(torch.from_numpy(y), F)


class OracleTransfer:
    def __init__(self, sh, lc, lf):
        self.sh, self.c, self.f = sh, lc, lf
        mesh = sh.mesh
        mult = torch.from_numpy(oracle.multiplicity(mesh.nelem, lf.P ** 3, 3, lf.dm.lsize, lf.off))
        if lf.dm.halo is not None:          # multiplicity summed over ranks (misc.c:115-143)
            lf.dm.halo.sum_and_share(mult)
        self.minv = 1.0 / mult
        self.fl, self.cl = lf.dm.create_local_vector(matops.MEM_HOST), lc.dm.create_local_vector(matops.MEM_HOST)

    def prolong(self, Xc, Yf):
        self.cl.zero_()
        self.c.dm.global_to_local(Xc, self.cl)
        y = oracle.transfer(False, self.sh.mesh.nelem, self.c.P, self.f.P, self.c.off, self.f.off, self.cl.numpy(),
                            self.f.dm.lsize)
        self.f.dm.local_to_global(torch.from_numpy(y) * self.minv, Yf)

    def restrict(self, Xf, Yc):
        self.fl.zero_()
        self.f.dm.global_to_local(Xf, self.fl)
        y = oracle.transfer(True, self.sh.mesh.nelem, self.c.P, self.f.P, self.c.off, self.f.off,
                            (self.fl * self.minv).numpy(), self.c.dm.lsize)
        self.c.dm.local_to_global(torch.from_numpy(y), Yc)


def oracle_solve(app, log=None, coarse="hmg", masked=True, dist=None, rank=0, world=1, **kw):
    """dist/rank/world: brick-partitioned run over torch.distributed (gloo in the CPU tests)"""
    from ceedpetscsolid_b200.elasticity import build_h_dms
    sh = OracleShared(app, masked, dist, rank, world)
    L = len(sh.degrees)
    levels = [OracleLevel(sh, l, l == L - 1) for l in range(L)]
    transfers = [None] + [OracleTransfer(sh, levels[l - 1], levels[l]) for l in range(1, L)]
    V = solver.Vec(dist if world > 1 else None)
    V.consistent = {}
    faces = "all" if app.test_mode else list(app.clamp.keys())
    h_dms = build_h_dms(sh.gmesh, sh.grid, rank, world, faces, "cpu", dist, masked=masked) \
        if coarse == "hmg" and getattr(sh.gmesh, "structured", True) else None
    for dm in [lev.dm for lev in levels] + list(h_dms or []):
        if dm.dot_weight is not None:
            V.set_weight(dm.nglobal, dm.dot_weight)
        if dm.dot_weight is not None or dm.masked:
            V.consistent[dm.nglobal] = dm.make_consistent
    pc = solver.PMultigrid(V, levels, transfers, h_dms=h_dms)
    U = levels[-1].dm.create_global_vector(matops.MEM_HOST)
    out = solver.newton_solve(V, levels[-1], pc, U, num_increments=app.num_steps, log=log, **kw)
    return out, U
