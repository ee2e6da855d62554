"""Generate tests/golden/ref_host_code_golden.npz from the REFERENCE's own host code and QFunctions.

Run in the build container (where /root/reference is mounted), after `make -C oracle refdriver_cpu`:

    python tests/golden/make_ref_host_golden.py

oracle/_ref/ref_driver_cpu links the reference's src/setuplibceed.c, src/matops.c and src/misc.c -- compiled unchanged
from where they lie -- with the generic CPU restatement of the libCEED calls they make (oracle/ceed_cpu.c) and runs the
reference's SetupLibceedFineLevel / SetupLibceedLevel, FormResidual_Ceed, ApplyJacobian_Ceed, GetDiag_Ceed,
Prolong_Ceed, Restrict_Ceed, ComputeStrainEnergy and ViewDiagnosticQuantities with the reference's QFunctions on the
seeded cases of tests/ref_host_code.py.  Everything the driver writes (residual; Jacobian product and diagonal on every
p-multigrid level; prolongation and restriction between the levels; strain energy; nodal diagnostics) is stored per case.
The .npz travels with the repository: on the GPU box, where /root/reference does not exist, the CUDA path
(tests/test_reference_host_code_on_gpu.py) and the oracle (tests/test_reference_host_code_on_cpu.py) are compared
with these vectors.
"""
import os
import pathlib
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_host_code as R  # noqa: E402


def main():
    assert os.path.isdir("/root/reference/src"), "run where the reference tree is mounted"
    assert os.path.exists(R.DRIVER_CPU), "make -C oracle refdriver_cpu first"
    out = {}
    for problem, n, degrees in R.GOLDEN_CASES:
        with tempfile.TemporaryDirectory() as d:
            _, _, _, res = R.run_driver(R.DRIVER_CPU, pathlib.Path(d), problem, n, degrees, 0, "/cpu/self")
        out[R.case_key(problem, n, degrees)] = res
        print(R.case_key(problem, n, degrees), res.size, "doubles")
    np.savez_compressed(R.GOLDEN, **out)
    print("wrote", R.GOLDEN, os.path.getsize(R.GOLDEN), "bytes")


if __name__ == "__main__":
    main()
