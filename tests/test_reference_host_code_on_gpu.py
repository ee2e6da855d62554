"""The reference's OWN host code on /gpu/b200.

oracle/_ref/ref_driver (built by `make -C oracle refdriver` where /root/reference is mounted; the binary travels to
the GPU box) links the reference's src/setuplibceed.c, src/matops.c and src/misc.c -- compiled unchanged, nothing
copied -- against this repository's <ceed.h> / libceed_b200.so and a functional miniature of the PETSc calls they make
(tests/c/petsc_mini).  The reference's SetupLibceedFineLevel / SetupLibceedLevel build every libCEED object (with the
reference's own QFunction pointers: the backend's QFunction guard runs on them), SetupJacobianCtx /
SetupProlongRestrictCtx wire the MatShell contexts, and FormResidual_Ceed, ApplyJacobian_Ceed, GetDiag_Ceed,
Prolong_Ceed, Restrict_Ceed, ComputeStrainEnergy and ViewDiagnosticQuantities are the reference's functions.  This test writes the mesh
(closure indices with essential-BC dofs encoded as -(loc+1), as DMPlex hands them to CreateRestrictionPlex) and
the vectors, runs the driver with host and with device memtype, and compares every output with the CPU oracle.
"""
import os

import pytest

from ref_host_code import DRIVER, ROOT, run_and_check


@pytest.mark.gpu
@pytest.mark.parametrize("resource", ["/gpu/b200", "/gpu/b200:deterministic"])
@pytest.mark.parametrize("memtype", [0, 1], ids=["memtype_host", "memtype_device"])
@pytest.mark.parametrize("problem,n,degrees", [("hyperFS", (3, 2, 2), [1, 2, 4]), ("hyperSS", (2, 2, 3), [1, 2, 3]),
                                               ("linElas", (3, 3, 2), [1, 2])])
def test_reference_setup_and_matshell_callbacks_drive_the_backend(tmp_path, problem, n, degrees, memtype, resource):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver not built (needs /root/reference at build time)")
    run_and_check(DRIVER, tmp_path, problem, n, degrees, memtype, resource)


def test_driver_sources_exist_and_build_recipe_names_the_reference_files():
    """CPU-side sanity: the recipe compiles the reference files where they lie and nothing of them is in this repo"""
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    for f in ("setuplibceed.c", "matops.c", "misc.c"):
        assert f"$(REF)/src/{f}" in mk
        assert not os.path.exists(os.path.join(ROOT, "tests", "c", f))
    assert os.path.exists(os.path.join(ROOT, "tests", "c", "ref_driver.c"))
    assert os.path.exists(os.path.join(ROOT, "tests", "c", "petsc_mini", "petsc_mini.c"))
