#!/usr/bin/env python
"""Kernel-only time of the fused hyperFS p=4 Jacobian / residual / diagonal at 64^3 for one build of the library
(CEED_B200_LIB selects a tuning variant).  Prints one line: name jac_ms res_ms diag_ms."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceedpetscsolid_b200 import ceed as libceed  # noqa: E402
from ceedpetscsolid_b200 import setuplibceed  # noqa: E402
from ceedpetscsolid_b200.mesh import BoxMesh, smooth_displacement  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = BoxMesh(n=(n, n, n), perturb=0.08, seed=0)
ceed = libceed.Ceed("/gpu/b200")
degrees, data, phys = setuplibceed.setup_all(ceed, mesh, "hyperFS", 4)
fine = len(degrees) - 1
u = torch.from_numpy(smooth_displacement(mesh.node_coords(4)).reshape(-1)).cuda()
r = torch.zeros_like(u)
uc, rc = ceed.Vector(u.numel()), ceed.Vector(u.numel())
uc.set_array(u); rc.set_array(r)
res = timeit(lambda: data[fine].opApply.apply_add(uc, rc))
x = torch.randn_like(u)
y = torch.zeros_like(u)
xc, yc = ceed.Vector(u.numel()), ceed.Vector(u.numel())
xc.set_array(x); yc.set_array(y)
jac = min(timeit(lambda: data[fine].opJacob.apply_add(xc, yc)) for _ in range(3))
diag = timeit(lambda: data[fine].opJacob.linear_assemble_diagonal(yc), reps=5, warm=1)
print(f"{os.environ.get('CEED_B200_LIB', 'default').split('libceed_b200')[-1]:24s} ahead={os.environ.get('B200_OFFSETS_AHEAD', 'dflt'):6s} "
      f"x_ahead={os.environ.get('B200_X_AHEAD', 'dflt'):5s} slab_ahead={os.environ.get('B200_SLAB_AHEAD', 'dflt'):5s} "
      f"jacobian {jac:.4f} ms  residual {res:.4f} ms  diagonal {diag:.4f} ms")
