"""Drop-in boundary, caller side: the reference's own libCEED-facing sources compile UNCHANGED against this
repository's <ceed.h>.

`gcc -fsyntax-only` of /root/reference/elasticity.c, src/setuplibceed.c, src/matops.c, src/misc.c and src/boundary.c
with -Iinclude (this repo's ceed.h) and a test-only stand-in for the PETSc headers (tests/c/petsc_stub: types and
macros only; PETSc is not in the image).  Passes iff there is no error, every Ceed* function the reference calls is
declared (no implicit declaration), and no call of a Ceed* function draws a type diagnostic.  Skipped where the
reference tree is not mounted (the GPU box)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SOURCES = ["elasticity.c", "src/setuplibceed.c", "src/matops.c", "src/misc.c", "src/boundary.c"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("src", SOURCES)
def test_reference_source_compiles_against_this_ceed_h(src):
    cmd = ["/usr/bin/gcc", "-std=gnu99", "-fsyntax-only", "-Wall", "-Wno-unused", "-I" + os.path.join(ROOT, "tests", "c", "petsc_stub"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(REF, src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    diag = r.stderr
    assert r.returncode == 0, diag[-3000:]
    assert " error: " not in diag
    implicit = set(re.findall(r"implicit declaration of function '(\w+)'", diag))
    assert not [n for n in implicit if n.startswith("Ceed")], f"Ceed functions missing from include/ceed/ceed.h: {implicit}"
    # diagnostics about argument / assignment types are only acceptable for the undeclared PETSc functions
    bad = []
    lines = diag.splitlines()
    for i, line in enumerate(lines):
        if re.search(r"warning: .*(incompatible pointer|makes (pointer|integer) from|too (few|many) arguments|discards)", line):
            ctx = " ".join(lines[i:i + 3])
            if re.search(r"\bCeed\w+\(", ctx):
                bad.append(ctx)
    assert not bad, bad[:3]
    if src != "src/boundary.c":
        assert re.search(r"\bCeed\w+\(", open(os.path.join(REF, src)).read())   # the file does use the API


def test_ceed_h_covers_every_ceed_identifier_the_reference_uses():
    """independent of the compiler: every Ceed* / CEED_* token in the reference's C sources appears in include/ceed/ceed.h"""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not mounted")
    hdr = open(os.path.join(ROOT, "include", "ceed", "ceed.h")).read()
    used = set()
    for src in SOURCES + ["elasticity.h"]:
        text = re.sub(r"/\*.*?\*/|//[^\n]*", "", open(os.path.join(REF, src)).read(), flags=re.S)
        used |= set(re.findall(r"\b(Ceed[A-Z]\w*|CEED_[A-Z_]+)\b", text))
    used -= {"CeedData", "CeedData_private", "CeedDataDestroy"}   # the reference's own struct / function (elasticity.h:218-240)
    used -= {"CEED_MEM_"}                                        # option-name prefix in a PetscOptionsEnum call (cloptions-style)
    missing = sorted(n for n in used if not re.search(r"\b" + n + r"\b", hdr))
    assert not missing, missing
