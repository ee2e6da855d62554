"""GPU-side test set-up: the same synthetic problems as tests/helpers.py, built through the
libCEED API of libceed_b200.so exactly as the reference's setuplibceed.c would."""
import numpy as np
import torch

from ceedpetscsolid_b200 import ceed as libceed
from ceedpetscsolid_b200 import setuplibceed
from ceedpetscsolid_b200.mesh import BoxMesh, smooth_displacement


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class GpuProblem:
    def __init__(self, problem, n, p, perturb=0.08, qextra=0, scale=0.02, multigrid="logarithmic",
                 node_perm_seed=None, nu=0.3, E=1.0, mesh=None, resource="/gpu/b200"):
        n = (n, n, n) if np.isscalar(n) else n
        self.problem, self.p, self.qextra = problem, p, qextra
        self.mesh = mesh if mesh is not None else BoxMesh(n=n, perturb=perturb, seed=0)
        self.ceed = libceed.Ceed(resource)
        self.perm = None
        if node_perm_seed is not None:
            self.perm = np.random.default_rng(node_perm_seed).permutation(self.mesh.num_nodes(p))
        self.degrees, self.data, self.phys = setuplibceed.setup_all(
            self.ceed, self.mesh, problem, p, nu, E, qextra, multigrid, node_perm=self.perm)
        self.fine = self.data[-1]
        self.u_fine = smooth_displacement(self.mesh.node_coords(p), scale).reshape(-1)
        if self.perm is not None:
            tmp = np.empty_like(self.u_fine).reshape(-1, 3)
            tmp[self.perm] = self.u_fine.reshape(-1, 3)
            self.u_fine = tmp.reshape(-1)

    def apply(self, op, x, nout):
        """CeedOperatorApply with borrowed device arrays (matops.c:40-50)."""
        xin = dev(x)
        yout = torch.zeros(nout, dtype=torch.float64, device="cuda")
        xc, yc = self.ceed.Vector(xin.numel()), self.ceed.Vector(nout)
        xc.set_array(xin)
        yc.set_array(yout)
        op.apply(xc, yc)
        xc.take_array()
        yc.take_array()
        xc.destroy(); yc.destroy()
        torch.cuda.synchronize()
        return yout.cpu().numpy()

    def residual(self, u=None):
        u = self.u_fine if u is None else u
        return self.apply(self.fine.opApply, u, u.size)

    def jacobian(self, level, x):
        return self.apply(self.data[level].opJacob, x, x.size)

    def diagonal(self, level):
        n = 3 * self.mesh.num_nodes(self.degrees[level])
        d = torch.zeros(n, dtype=torch.float64, device="cuda")
        dc = self.ceed.Vector(n)
        dc.set_array(d)
        self.data[level].opJacob.linear_assemble_diagonal(dc)
        dc.take_array()
        dc.destroy()
        torch.cuda.synchronize()
        return d.cpu().numpy()

    def strided_to_plain(self, rstr, vec):
        """Backend-strided L-vector -> [elem][comp][q] (CeedElemRestrictionApply NOTRANSPOSE)."""
        ev = self.ceed.Vector(rstr.nelem * rstr.ncomp * rstr.elemsize)
        rstr.apply(vec, ev)
        out = ev.to_numpy().reshape(rstr.nelem, rstr.ncomp, rstr.elemsize)
        ev.destroy()
        return out
