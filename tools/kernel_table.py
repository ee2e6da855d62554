#!/usr/bin/env python
"""Kernel-only timings of every fused kernel on the BASELINE configs (CUDA events, inputs > L2):
Jacobian apply on every p-multigrid level, residual, diagonal, prolongation/restriction.
Prints a markdown table with algorithmic GB/s against the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceedpetscsolid_b200 import ceed as libceed  # noqa: E402
from ceedpetscsolid_b200 import setuplibceed  # noqa: E402
from ceedpetscsolid_b200.mesh import BoxMesh, smooth_displacement  # noqa: E402


def timeit(fn, reps=10, warm=3, rounds=3):
    """best of `rounds` batches of `reps` calls: the power cap lowers the SM clock for a while after an FP64-heavy
    kernel (diagonal, residual), which would otherwise leak into the next row of the table"""
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(rounds):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="linElas:2:96,hyperSS:3:64,hyperFS:4:64")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(f"| config | kernel | (P,Q) | elements | DoFs (level) | ms | GDoF/s | alg. B/elem | alg. GB/s | frac of {peak:.0f} GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for cfg in args.configs.split(","):
        problem, p, n = cfg.split(":")
        p, n = int(p), int(n)
        mesh = BoxMesh(n=(n, n, n), perturb=0.08, seed=0)
        ceed = libceed.Ceed("/gpu/b200")
        degrees, data, phys = setuplibceed.setup_all(ceed, mesh, problem, p)
        fine = len(degrees) - 1
        Q = p + 1
        g = 0 if problem == "linElas" else 9
        u = torch.from_numpy(smooth_displacement(mesh.node_coords(p)).reshape(-1)).cuda()
        r = torch.zeros_like(u)
        uc, rc = ceed.Vector(u.numel()), ceed.Vector(u.numel())
        uc.set_array(u); rc.set_array(r)
        ms = timeit(lambda: data[fine].opApply.apply_add(uc, rc))
        nel = mesh.nelem
        def row(kern, P, ms, dofs, bpe):
            gb = bpe * nel / ms / 1e6
            print(f"| {problem} p={p} {n}^3 | {kern} | ({P},{Q}) | {nel} | {dofs} | {ms:.3f} | {dofs / ms / 1e6:.2f} | {bpe} | {gb:.0f} | {gb / peak:.2f} |")
        # residual: reads qdata (10), writes gradu (g), offsets, x, y
        row("residual (fused)", p + 1, ms, u.numel(), 8 * Q ** 3 * (10 + g) + 4 * (p + 1) ** 3 + 48 * p ** 3)
        uc.take_array(); rc.take_array()
        for level, deg in enumerate(degrees):
            nl = 3 * mesh.num_nodes(deg)
            x = torch.randn(nl, dtype=torch.float64, device="cuda")
            y = torch.zeros_like(x)
            xc, yc = ceed.Vector(nl), ceed.Vector(nl)
            xc.set_array(x); yc.set_array(y)
            ms = timeit(lambda: data[level].opJacob.apply_add(xc, yc))
            row("Jacobian (fused)", deg + 1, ms, nl, 8 * Q ** 3 * (10 + g) + 4 * (deg + 1) ** 3 + 48 * deg ** 3)
            ms = timeit(lambda: data[level].opJacob.linear_assemble_diagonal(yc), reps=3, warm=1)
            row("diagonal (fused)", deg + 1, ms, nl, 8 * Q ** 3 * (10 + g) + 4 * (deg + 1) ** 3 + 24 * deg ** 3)
            if level > 0:
                nc = 3 * mesh.num_nodes(degrees[level - 1])
                c = torch.randn(nc, dtype=torch.float64, device="cuda")
                cc = ceed.Vector(nc)
                cc.set_array(c)
                ms = timeit(lambda: data[level].opProlong.apply_add(cc, yc))
                bt = 4 * (deg + 1) ** 3 + 4 * (degrees[level - 1] + 1) ** 3 + 24 * deg ** 3 + 24 * degrees[level - 1] ** 3
                row(f"prolong {degrees[level-1]}->{deg}", deg + 1, ms, nl, bt)
                ms = timeit(lambda: data[level].opRestrict.apply_add(yc, cc))
                row(f"restrict {deg}->{degrees[level-1]}", deg + 1, ms, nl, bt)
                cc.take_array()
            xc.take_array(); yc.take_array()
        del data, ceed
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
