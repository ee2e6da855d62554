#!/usr/bin/env python
"""Kernel-only timings of the p-multigrid transfer kernels (CUDA events, best of 3 x 10 calls, L-vectors > L2 on the
fine side): plain prolongation / restriction (sum semantics) and the forms the solver uses (prolongation that stores
the interpolant, restriction with the inverse multiplicity applied inside the kernel).
CEED_B200_LIB selects a tuning build.  One line per (config, level pair)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kernel_table import timeit  # noqa: E402
from ceedpetscsolid_b200 import ceed as libceed  # noqa: E402
from ceedpetscsolid_b200 import setuplibceed  # noqa: E402
from ceedpetscsolid_b200.mesh import BoxMesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="hyperSS:3:64,hyperFS:4:64")
    args = ap.parse_args()
    name = os.environ.get("CEED_B200_LIB", "default").split("libceed_b200")[-1]
    for cfg in args.configs.split(","):
        problem, p, n = cfg.split(":")
        p, n = int(p), int(n)
        mesh = BoxMesh(n=(n, n, n), perturb=0.08, seed=0)
        ceed = libceed.Ceed("/gpu/b200")
        degrees, data, phys = setuplibceed.setup_all(ceed, mesh, problem, p)
        for level in range(1, len(degrees)):
            pc, pf = degrees[level - 1], degrees[level]
            nc, nf = 3 * mesh.num_nodes(pc), 3 * mesh.num_nodes(pf)
            c = torch.randn(nc, dtype=torch.float64, device="cuda")
            f = torch.randn(nf, dtype=torch.float64, device="cuda")
            cc, fc, mc = ceed.Vector(nc), ceed.Vector(nf), ceed.Vector(nf)
            cc.set_array(c); fc.set_array(f)
            d = data[level]
            d.Erestrictu.get_multiplicity(mc)
            mc.reciprocal()
            t = {}
            t["prolong"] = timeit(lambda: d.opProlong.apply_add(cc, fc))
            t["restrict"] = timeit(lambda: d.opRestrict.apply_add(fc, cc))
            d.opProlong.set_transfer_scaling(mc, inject=True)
            d.opRestrict.set_transfer_scaling(mc)
            t["prolong(inject)"] = timeit(lambda: d.opProlong.apply_add(cc, fc))
            t["restrict(scaled)"] = timeit(lambda: d.opRestrict.apply_add(fc, cc))
            d.opProlong.set_transfer_scaling(None)
            d.opRestrict.set_transfer_scaling(None)
            cc.take_array(); fc.take_array()
            print(f"{name:16s} {problem} p={p} {n}^3 {pc}->{pf}  " + "  ".join(f"{k} {v:.4f} ms" for k, v in t.items()), flush=True)
        del data, ceed
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
