#!/usr/bin/env python
"""Times the per-Newton-step set-up pieces of the p-MG preconditioner in isolation (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceedpetscsolid_b200 import solver  # noqa: E402
from ceedpetscsolid_b200.elasticity import AppCtx, Elasticity  # noqa: E402


def timeit(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"  {name:50s} {a.elapsed_time(b) / reps:9.3f} ms")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    app = AppCtx(problem="hyperFS", degree=4, n=(n, n, n), num_steps=1, perturb=0.05)
    el = Elasticity(app)
    pc = el.pc
    F = torch.zeros_like(el.U)
    el.levels[-1].residual(el.U, F, 1.0)
    pc.setup()
    coo = el.levels[0].coo
    timeit("CeedOperatorLinearAssemble (k_assemble_p1)", coo.values)
    timeit("COO -> stencil (index_add_)", lambda: (pc.coarse.svals.zero_(), pc.coarse.svals.view(-1).index_add_(0, pc.coarse.coo_dest, coo.vals)))
    timeit("coarse.assemble() total", pc.coarse.assemble)
    timeit("h-MG setup total", pc.hmg.setup)
    for l in range(1, len(el.levels)):
        timeit(f"smoother setup L{l} (eig estimate)", lambda l=l: pc.smoothers[l].setup(pc.diag[l]))
        timeit(f"jacobian L{l}", lambda l=l: el.levels[l].jacobian(pc.x[l], pc.t[l]))
        timeit(f"diagonal L{l}", lambda l=l: el.levels[l].diagonal(pc.diag[l]))
    timeit("diagonal L0", lambda: el.levels[0].diagonal(pc.diag[0]))
    timeit("pc.setup() total", pc.setup)
    timeit("V-cycle", lambda: pc.apply(F, pc.t[-1]))


if __name__ == "__main__":
    main()
