"""Shared test set-up: synthetic box problems evaluated with the oracle."""
import numpy as np

from ceedpetscsolid_b200.mesh import BoxMesh, smooth_displacement
from oracle import oracle

PHYS = (0.3, 1.0)  # nu, E  (TESTARGS, /root/reference/elasticity.c:36)


class OracleProblem:
    """Box problem at fine degree `p` with level degree `pl` (P = pl+1, Q = p+1+qextra)."""

    def __init__(self, problem, n, p, pl=None, perturb=0.08, qextra=0, scale=0.02, which=None,
                 node_perm_seed=None, mesh=None):
        self.problem, self.p, self.pl = problem, p, (p if pl is None else pl)
        self.which = which
        n = (n, n, n) if np.isscalar(n) else n
        self.mesh = mesh if mesh is not None else BoxMesh(n=n, perturb=perturb, seed=0)
        self.nelem = self.mesh.nelem
        self.P, self.Pf, self.Q = self.pl + 1, p + 1, p + 1 + qextra
        self.perm = None
        if node_perm_seed is not None:
            self.perm = np.random.default_rng(node_perm_seed).permutation(self.mesh.num_nodes(self.pl))
        self.offsets = self.mesh.offsets(self.pl, node_perm=self.perm)
        self.offsets_fine = self.mesh.offsets(p)
        self.lsize = self.mesh.lsize(self.pl)
        self.lsize_fine = self.mesh.lsize(p)
        self.B, self.D, _, self.qw = oracle.basis_1d(self.P, self.Q, 0)
        self.Bf, self.Df, _, _ = oracle.basis_1d(self.Pf, self.Q, 0)
        self.xoffsets = self.mesh.offsets(1)
        self.xcoord = self.mesh.coord_lvector()
        self.qdata = oracle.setup_geo(self.nelem, self.Q, self.xoffsets, self.xcoord, which)
        self.has_gradu = oracle.PROBLEMS[problem][2]
        self.gradu = np.zeros((self.nelem, 9, self.Q ** 3)) if self.has_gradu else None
        # state on the FINE level (gradu is always produced by the fine residual operator)
        self.u_fine = smooth_displacement(self.mesh.node_coords(p), scale).reshape(-1)
        self.residual_fine(self.u_fine)

    def to_perm(self, v_lex):
        """Re-order a lexicographic level L-vector into the (optionally permuted) numbering."""
        if self.perm is None:
            return v_lex
        out = np.empty_like(v_lex).reshape(-1, 3)
        out[self.perm] = v_lex.reshape(-1, 3)
        return out.reshape(-1)

    def residual_fine(self, u):
        """Residual operator on the fine level (writes gradu as a side effect)."""
        return oracle.operator_apply(self.problem, False, PHYS, self.nelem, self.Pf, self.Q, self.Bf,
                                     self.Df, self.offsets_fine, self.qdata, self.gradu, u, self.which)

    def jacobian(self, x):
        return oracle.operator_apply(self.problem, True, PHYS, self.nelem, self.P, self.Q, self.B, self.D,
                                     self.offsets, self.qdata, self.gradu, x, self.which)

    def diagonal(self):
        return oracle.operator_diagonal(self.problem, PHYS, self.nelem, self.P, self.Q, self.B, self.D,
                                        self.offsets, self.qdata, self.gradu, self.lsize, self.which)


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)
