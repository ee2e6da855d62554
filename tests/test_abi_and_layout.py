"""CPU-side checks of the drop-in boundary: libceed_b200.so loads without a GPU and exports
every symbol declared in include/*.h; the q-blocked layout is a bijection; the shared-memory
lattice of the fused kernels is bank-conflict free for every line orientation."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ceedpetscsolid_b200", "libceed_b200.so")


def _declared(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(Ceed[A-Z]\w+|b200_\w+)\s*\(", src))
    names -= {"CeedQFunctionUser", "CeedErrorHandler", "b200_physics"}
    data = set(re.findall(r"CEED_EXTERN\s+[^;()]*?\b(CEED_[A-Z_]+|CeedMemTypes|CeedEvalModes)\s*(?:\[\d*\])?\s*;", src))
    return names, data


def test_library_loads_and_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    funcs, data = set(), set()
    for h in ("include/ceed/ceed.h", "include/b200_kernels.h"):
        f, d = _declared(os.path.join(ROOT, h))
        funcs |= f
        data |= d
    assert len(funcs) > 90 and "CeedOperatorApply" in funcs and "b200_apply_jacobian" in funcs
    missing = [n for n in sorted(funcs | data) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_no_cpu_fallback_and_loud_failure_without_gpu():
    from ceedpetscsolid_b200 import ceed as libceed
    with pytest.raises(libceed.CeedError, match="no CPU fallback"):
        libceed.Ceed("/cpu/self")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(libceed.CeedError, match="no CUDA device"):
            libceed.Ceed("/gpu/b200")


def _eb(Q):
    return 16 if Q <= 3 else 4


def _qblocked(nelem, ncomp, Q, e, c, q):
    EB, T, Q3 = _eb(Q), Q * Q, Q ** 3
    g, ei = divmod(e, EB)
    ebn = min(EB, nelem - g * EB)
    return g * EB * ncomp * Q3 + (c * Q + q % Q) * (ebn * T) + (q // Q) * ebn + ei


@pytest.mark.parametrize("Q,nelem,ncomp", [(2, 37, 10), (3, 16, 9), (4, 13, 10), (5, 8, 17), (5, 21, 10)])
def test_qblocked_layout_is_a_bijection(Q, nelem, ncomp):
    """mirror of qblocked_index (csrc/b200_common.cuh): exactly nelem*ncomp*Q^3 slots, none reused"""
    from ceedpetscsolid_b200 import ceed as libceed
    assert libceed.lib.b200_elems_per_block(Q) == _eb(Q)
    idx = [_qblocked(nelem, ncomp, Q, e, c, q) for e in range(nelem) for c in range(ncomp) for q in range(Q ** 3)]
    assert sorted(idx) == list(range(nelem * ncomp * Q ** 3))


@pytest.mark.parametrize("Q", [2, 3, 4, 5])
def test_shared_lattice_is_bank_conflict_free(Q):
    """tid = t*EB + e, odd lattice strides (1, QP, QP^2), element stride SE = 16/EB mod 16 (EB=8, 4) or
    odd (EB=16): every half-warp of every line orientation touches 16 distinct 8-byte banks."""
    EB = _eb(Q)
    QP = Q if Q % 2 else Q + 1
    SE = 9 * QP ** 3
    SE = SE + ((16 // EB - SE % 16) + 16) % 16 if EB < 16 else SE | 1
    T = Q * Q
    for orient in "xyz":
        for w in range(0, T * EB, 16):
            banks = set()
            for tid in range(w, min(w + 16, T * EB)):
                t, e = divmod(tid, EB)
                a, b = t % Q, t // Q
                off = {"x": b * QP * QP + a * QP, "y": b * QP * QP + a, "z": b * QP + a}[orient]
                banks.add((e * SE + off) % 16)
            assert len(banks) == min(16, T * EB - w), (orient, w)


def test_fused_apply_lattice_of_the_headline_kernel_is_conflict_free_including_the_scatter_sweep():
    """P = Q = 5 (hyperFS degree 4): the strides the kernel really uses (b200_apply_smem_layout) give conflict-free
    line accesses in all three orientations; the last stage writes the nodal output interlaced [node][component] at
    the start of the element's region, which makes both that write and the final scatter sweep (lanes over
    (element, node, component), component fastest) conflict-free except where a half-warp straddles two elements."""
    from ceedpetscsolid_b200 import ceed as libceed
    out = (ctypes.c_int * 6)()
    assert libceed.lib.b200_apply_smem_layout(5, 5, out) == 0
    SY, SZ, SC, SE, EB, NT = list(out)
    Q = P = 5
    T = Q * Q
    assert (SY, SZ, SC, EB, NT) == (5, 25, 125, 4, 128) and SE >= 9 * SC and SE % 16 == 4
    for orient in "xyz":
        for idx in range(Q):
            for w in range(0, T * EB, 16):
                banks = set()
                for tid in range(w, min(w + 16, T * EB)):
                    t, e = divmod(tid, EB)
                    a, b = t % Q, t // Q
                    off = {"x": idx + a * SY + b * SZ, "y": a + idx * SY + b * SZ, "z": a + b * SY + idx * SZ}[orient]
                    banks.add((e * SE + off) % 16)
                assert len(banks) == min(16, T * EB - w), (orient, idx, w)
    # last stage: thread (j, k) writes node (i, j, k), component c to e*SE + ((k*P + j)*P + i)*3 + c
    for i in range(P):
        for c in range(3):
            for w in range(0, T * EB, 16):
                banks = set()
                for tid in range(w, min(w + 16, T * EB)):
                    t, e = divmod(tid, EB)
                    j, k = t % Q, t // Q
                    banks.add((e * SE + ((k * P + j) * P + i) * 3 + c) % 16)
                assert len(banks) == min(16, T * EB - w), (i, c, w)
    # scatter sweep: consecutive lanes read consecutive words inside an element
    total, wavefronts, ideal = EB * 3 * P ** 3, 0, 0
    for w in range(0, total, 16):
        cnt = {}
        for f in range(w, min(w + 16, total)):
            el, r = divmod(f, 3 * P ** 3)
            bank = (el * SE + r) % 16
            cnt[bank] = cnt.get(bank, 0) + 1
        wavefronts += max(cnt.values())
        ideal += 1
    assert ideal == 94 and wavefronts <= ideal + (EB - 1), (wavefronts, ideal)


def _build_capi(tmp):
    import subprocess
    exe = os.path.join(tmp, "capi_smoke")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O1", "-Wall", "-Werror=implicit-function-declaration",
                    "-Wno-unused-parameter", "-Wno-unused-variable", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "capi_smoke.c"), "-o", exe,
                    "-L" + os.path.join(ROOT, "ceedpetscsolid_b200"), "-lceed_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "ceedpetscsolid_b200"),
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                    "-lm"], check=True)
    return exe


def test_c_program_written_against_ceed_h_compiles_and_links(tmp_path):
    """the reference's call sequence (setuplibceed.c / matops.c) in plain C99 against <ceed.h>"""
    assert os.path.exists(_build_capi(str(tmp_path)))


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(tmp_path):
    import subprocess
    r = subprocess.run([_build_capi(str(tmp_path))], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "capi_smoke OK" in r.stdout, r.stdout + r.stderr
