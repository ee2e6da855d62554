#!/usr/bin/env python
"""Where does the Newton-Krylov-p-MG solve spend its time?  Wraps the solver's building blocks with
synchronising timers (perturbs the total a little; use for shares)."""
import os
import sys
import time
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceedpetscsolid_b200 import solver  # noqa: E402
from ceedpetscsolid_b200.elasticity import AppCtx, Elasticity  # noqa: E402

T = defaultdict(float)
C = defaultdict(int)


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] += time.perf_counter() - t
        C[name] += 1
        return r
    return w


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    app = AppCtx(problem="hyperFS", degree=4, n=(n, n, n), num_steps=steps, perturb=0.05,
                 clamp={(2, 0): [0, 0, 0, 0, 0, 1, 0], (2, 1): [0, 0, -0.1 * steps / 10, 0, 0, 1, 0]})
    el = Elasticity(app)
    pc = el.pc
    for l, lev in enumerate(el.levels):
        lev.jacobian = timed(f"jacobian L{l}", lev.jacobian)
        lev.diagonal = timed(f"diagonal L{l}", lev.diagonal)
    el.levels[-1].residual = timed("residual", el.levels[-1].residual)
    for l in range(1, len(el.levels)):
        pc.smoothers[l].A = el.levels[l].jacobian
        pc.smoothers[l].apply = timed(f"smoother L{l} (incl. its jacobians)", pc.smoothers[l].apply)
        pc.smoothers[l].setup = timed(f"smoother setup L{l} (eig estimate)", pc.smoothers[l].setup)
        el.transfers[l].prolong = timed(f"prolong L{l}", el.transfers[l].prolong)
        el.transfers[l].restrict = timed(f"restrict L{l}", el.transfers[l].restrict)
    pc.coarse.assemble = timed("coarse assemble (81 applies)", pc.coarse.assemble)
    pc._coarse_solve = timed("coarse solve", pc._coarse_solve)
    if pc.hmg is not None:
        pc.hmg.setup = timed("h-MG setup (Galerkin colouring + eig estimates)", pc.hmg.setup)
    pc.apply = timed("V-cycle total", pc.apply)
    out = el.solve()
    tot = out["time_s"]
    print(f"box {n}^3, {steps} load steps: {tot:.2f} s, snes {out['snes_its']}, ksp {out['ksp_its']}, coarse its {out['coarse_its']}")
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
        print(f"  {k:42s} {v:8.3f} s  {100 * v / tot:5.1f}%  calls {C[k]:6d}  {1e3 * v / C[k]:8.3f} ms/call")


if __name__ == "__main__":
    main()
