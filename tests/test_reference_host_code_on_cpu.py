"""The reference's OWN host code and OWN QFunctions on the CPU -- no GPU involved.

oracle/_ref/ref_driver_cpu (`make -C oracle refdriver_cpu`, built where /root/reference is mounted) links the
reference's src/setuplibceed.c, src/matops.c and src/misc.c -- compiled unchanged, nothing copied -- with the same
driver and miniature PETSc as the GPU variant, but against oracle/ceed_cpu.c: a GENERIC restatement of the libCEED
calls those files make (objects wired field by field, element-by-element apply through the caller's QFunction pointer,
LinearAssembleDiagonal by unit inputs).  libCEED is not vendored in the reference, so this
is how the fixed-shape oracle of oracle/ceed_oracle.c is anchored on the reference's own call sites: every libCEED
object here is created and wired by the reference's SetupLibceedFineLevel / SetupLibceedLevel, applied by the
reference's FormResidual_Ceed / ApplyJacobian_Ceed / GetDiag_Ceed / Prolong_Ceed / Restrict_Ceed /
ComputeStrainEnergy / ViewDiagnosticQuantities, and the point functions are the reference's -- and the results must
equal what the oracle computes from the mesh and the vectors alone.
"""
import os

import pytest

from ref_host_code import DRIVER_CPU, GOLDEN_CASES, ROOT, run_and_check


@pytest.mark.parametrize("problem,n,degrees", [("hyperFS", (3, 2, 2), [1, 2, 4]), ("hyperSS", (2, 2, 3), [1, 2, 3]),
                                               ("linElas", (3, 3, 2), [1, 2]), ("hyperFS", (2, 2, 2), [1, 3])])
def test_reference_host_code_on_the_generic_cpu_restatement_equals_the_oracle(tmp_path, problem, n, degrees):
    if not os.path.exists(DRIVER_CPU):
        pytest.skip("oracle/_ref/ref_driver_cpu not built (needs /root/reference at build time)")
    run_and_check(DRIVER_CPU, tmp_path, problem, n, degrees, 0, "/cpu/self")


@pytest.mark.parametrize("problem,n,degrees,forcing", [("linElas", (3, 2, 2), [1, 2], "mms"), ("linElas", (2, 2, 2), [1, 3], "constant"),
                                                       ("hyperFS", (2, 2, 2), [1, 2], "constant")])
def test_reference_forcing_and_true_solution_setup_equal_the_oracle(tmp_path, problem, n, degrees, forcing):
    """-forcing constant / mms (setuplibceed.c:550-640): the reference's forcing operator (SetupConstantForce /
    SetupMMSForce: x INTERP, qdata, force INTERP^T) and, for MMS, its nodal true solution (MMSTrueSoln on a P=2 -> GLL
    basis, divided by the multiplicity through CeedVectorGetArray) on the generic CPU restatement, besides everything
    the other cases check"""
    if not os.path.exists(DRIVER_CPU):
        pytest.skip("oracle/_ref/ref_driver_cpu not built (needs /root/reference at build time)")
    run_and_check(DRIVER_CPU, tmp_path, problem, n, degrees, 0, "/cpu/self", forcing=forcing)


def test_cpu_driver_is_test_infrastructure_only():
    """the generic CPU restatement lives under oracle/ and is linked into nothing the product ships"""
    mk = open(os.path.join(ROOT, "ceedpetscsolid_b200", "csrc", "Makefile")).read()
    assert "ceed_cpu" not in mk and "oracle" not in mk
    assert open(os.path.join(ROOT, "oracle", "ceed_cpu.c")).read().startswith("/* TEST INFRASTRUCTURE ONLY")


@pytest.mark.parametrize("problem,n,degrees", GOLDEN_CASES)
def test_oracle_reproduces_the_committed_output_of_the_reference_host_code(problem, n, degrees):
    """no driver, no reference tree: the oracle alone against tests/golden/ref_host_code_golden.npz, i.e. against what the
    reference's set-up / MatShell functions and QFunctions produced on the CPU when the fixture was generated"""
    import numpy as np
    from helpers import rel_err
    from ref_host_code import GOLDEN, case_key, make_case, oracle_outputs, split_like
    o, frees, xs = make_case(problem, n, degrees)
    parts = oracle_outputs(problem, degrees, o, frees, xs)
    with np.load(GOLDEN) as g:
        res = g[case_key(problem, n, degrees)]
    for (name, want), got in zip(parts, split_like(res, parts, o.mesh.num_nodes(degrees[-1]))):
        assert rel_err(got, want) < 1e-12, name
