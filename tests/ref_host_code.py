"""Shared by tests/test_reference_host_code_on_gpu.py and tests/test_reference_host_code_on_cpu.py: the input file of
tests/c/ref_driver.c (mesh as DMPlex would hand it to the reference, vectors) and the comparison of every output of the
reference's MatShell callbacks with the CPU oracle."""
import os
import subprocess

import numpy as np

from helpers import PHYS, OracleProblem, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")            # against libceed_b200.so
DRIVER_CPU = os.path.join(ROOT, "oracle", "_ref", "ref_driver_cpu")    # against oracle/ceed_cpu.c
TOL = 1e-12
PROBLEM_ID = {"linElas": 0, "hyperSS": 1, "hyperFS": 2}
FORCING_ID = {0: 0, None: 0, "none": 0, "constant": 1, "mms": 2}      # elasticity.h:64-66
FORCING_VECTOR = (0.3, -1.0, 2.5)                                      # tests/c/ref_driver.c


def closure(offsets, ncomp_stride, ncomp, bc_nodes=None):
    """closure indices of every cell: interlaced dofs of each node in tensor order, essential-BC dofs as -(loc+1)"""
    node = offsets // ncomp_stride                                               # (nelem, P^3)
    loc = node[:, :, None] * ncomp + np.arange(ncomp)[None, None, :]
    if bc_nodes is not None:
        loc = np.where(bc_nodes[node][:, :, None], -(loc + 1), loc)
    return np.ascontiguousarray(loc.reshape(offsets.shape[0], -1), dtype=np.int32)


def write_input(path, mesh, problem, degrees, memtype, u_fine, xs, forcing=0):
    p = degrees[-1]
    with open(path, "wb") as f:
        def wi(*v):
            np.asarray(v, dtype=np.int32).tofile(f)
        wi(0x42323030, len(degrees), mesh.nelem, PROBLEM_ID[problem] | (FORCING_ID[forcing] << 8), memtype)
        wi(*degrees)
        xoff = mesh.offsets(1)
        wi(mesh.lsize(1), mesh.lsize(1))
        closure(xoff, 3, 3).tofile(f)
        np.ascontiguousarray(mesh.coord_lvector(), dtype=np.float64).tofile(f)
        maps = []
        for deg in degrees:
            bc = mesh.boundary_mask(deg, "all")
            free = ~np.repeat(bc, 3)
            l2g = np.full(mesh.lsize(deg), -1, dtype=np.int32)
            l2g[free] = np.arange(int(free.sum()), dtype=np.int32)
            wi(mesh.lsize(deg), int(free.sum()))
            closure(mesh.offsets(deg), 3, 3, bc).tofile(f)
            l2g.tofile(f)
            maps.append(free)
        nn = mesh.num_nodes(p)
        wi(nn, nn)
        closure(mesh.offsets(p), 3, 1).tofile(f)
        wi(8 * nn, 8 * nn)
        closure(mesh.offsets(p), 3, 8).tofile(f)
        bc_idx = np.flatnonzero(~maps[-1]).astype(np.int32)
        wi(bc_idx.size)
        bc_idx.tofile(f)
        np.ascontiguousarray(u_fine[bc_idx], dtype=np.float64).tofile(f)
        np.ascontiguousarray(u_fine[maps[-1]], dtype=np.float64).tofile(f)
        for x in xs:
            np.ascontiguousarray(x, dtype=np.float64).tofile(f)
    return maps


GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_host_code_golden.npz")
GOLDEN_CASES = [("hyperFS", (3, 2, 2), [1, 2, 4]), ("hyperSS", (2, 2, 3), [1, 2, 3]), ("linElas", (3, 3, 2), [1, 2])]


def case_key(problem, n, degrees):
    return f"{problem}_n{'x'.join(map(str, n))}_p{'-'.join(map(str, degrees))}"


def make_case(problem, n, degrees):
    """the seeded case: (OracleProblem, free-dof mask of every level, global x vector of every level)"""
    o = OracleProblem(problem, n, degrees[-1])
    rng = np.random.default_rng(17)
    frees = [~np.repeat(o.mesh.boundary_mask(deg, "all"), 3) for deg in degrees]
    xs = [rng.standard_normal(int(fr.sum())) for fr in frees]
    return o, frees, xs


def run_driver(driver, tmp_path, problem, n, degrees, memtype, resource, forcing=0):
    """write the seeded case, run `driver` on it; returns (OracleProblem, free-dof masks, x vectors, raw output)"""
    o, frees, xs = make_case(problem, n, degrees)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_input(inp, o.mesh, problem, degrees, memtype, o.u_fine, xs, forcing)
    r = subprocess.run([driver, inp, out, resource], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ref_driver OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return o, frees, xs, np.fromfile(out, dtype=np.float64)


def oracle_outputs(problem, degrees, o, frees, xs, forcing=0):
    """what the driver writes, computed by the oracle from the mesh and the vectors alone: [(name, vector), ...] in the
    driver's order"""
    from oracle import oracle
    from test_gpu_postprocess import _oracle_post
    mesh, p = o.mesh, degrees[-1]
    # residual at the smooth state (boundary values inserted by the reference's FormResidual_Ceed at load 1)
    parts = [("residual", o.residual_fine(o.u_fine)[frees[-1]])]
    locals_x = []
    for l, deg in enumerate(degrees):
        P = deg + 1
        B, D, _, _ = oracle.basis_1d(P, o.Q, 0)
        off, lsize = mesh.offsets(deg), mesh.lsize(deg)
        xl = np.zeros(lsize)
        xl[frees[l]] = xs[l]
        locals_x.append(xl)
        yo = oracle.operator_apply(problem, True, PHYS, o.nelem, P, o.Q, B, D, off, o.qdata, o.gradu, xl)
        parts.append((f"jacobian p={deg}", yo[frees[l]]))
        do = oracle.operator_diagonal(problem, PHYS, o.nelem, P, o.Q, B, D, off, o.qdata, o.gradu, lsize)
        parts.append((f"diagonal p={deg}", do[frees[l]]))
    for l in range(1, len(degrees)):
        pc, pf = degrees[l - 1], degrees[l]
        offc, offf, lc, lf = mesh.offsets(pc), mesh.offsets(pf), mesh.lsize(pc), mesh.lsize(pf)
        mult = oracle.multiplicity(o.nelem, (pf + 1) ** 3, 3, lf, offf)
        minv = np.where(frees[l], 1.0 / mult, 0.0)            # misc.c:115-143: L2G, G2L, reciprocal (0 stays 0)
        yp = oracle.transfer(False, o.nelem, pc + 1, pf + 1, offc, offf, locals_x[l - 1], lf) * minv
        parts.append((f"prolong {pc}->{pf}", yp[frees[l]]))
        yr = oracle.transfer(True, o.nelem, pc + 1, pf + 1, offc, offf, locals_x[l] * minv, lc)
        parts.append((f"restrict {pf}->{pc}", yr[frees[l - 1]]))
    e_ref, d_ref = _oracle_post(problem, mesh, p, o.u_fine)
    parts.append(("strain energy", np.array([e_ref])))
    # ViewDiagnosticQuantities (misc.c:217-300): (u, pressure, two strain invariants, volume ratio, energy density) per node
    diag = np.concatenate([o.u_fine.reshape(-1, 3), d_ref[:, 3:8]], axis=1)
    for k in range(8):
        parts.append((f"diagnostic column {k}", diag[:, k]))
    if FORCING_ID[forcing]:
        parts += forcing_outputs(o, p, forcing)
    return parts


def forcing_outputs(o, p, forcing):
    """the forcing L-vector (setuplibceed.c:550-583: x INTERP, qdata, force INTERP^T) and, for MMS, the nodal true
    solution averaged over the elements sharing a node (setuplibceed.c:585-640), from oracle pieces"""
    import ctypes as C
    from oracle import oracle
    mesh, nel, P, Q = o.mesh, o.nelem, p + 1, o.Q
    which = oracle.default_which()
    Bx, Dx, _, qw = oracle.basis_1d(2, Q, 0)
    Bu, Du, _, _ = oracle.basis_1d(P, Q, 0)
    xe = mesh.coord_lvector()[mesh.offsets(1)[:, None, :] + np.arange(3)[None, :, None]]          # (nel, 3, 8)
    xq = oracle.basis_apply(nel, 3, 2, Q, Bx, Dx, qw, 0, 1, xe).reshape(nel, 3, Q ** 3)
    ctx = oracle.Physics(*PHYS) if forcing == "mms" else (C.c_double * 3)(*FORCING_VECTOR)
    name = "SetupMMSForce" if forcing == "mms" else "SetupConstantForce"
    fq = np.zeros((nel, 3, Q ** 3))
    for e in range(nel):
        (fq[e],) = oracle.call_qf(name, which, ctx, Q ** 3, [xq[e], o.qdata[e]], [3])
    fe = oracle.basis_apply(nel, 3, P, Q, Bu, Du, qw, 1, 1, fq.reshape(nel, -1))
    idx = (mesh.offsets(p)[:, None, :] + np.arange(3)[None, :, None]).reshape(-1)
    force = np.zeros(mesh.lsize(p))
    np.add.at(force, idx, fe.reshape(-1))
    parts = [("forcing L-vector", force)]
    if forcing == "mms":
        Bt, Dt, _, qwt = oracle.basis_1d(2, P, 1)                                                 # vertices -> GLL nodes
        xn = oracle.basis_apply(nel, 3, 2, P, Bt, Dt, qwt, 0, 1, xe).reshape(nel, 3, P ** 3)
        te = np.zeros((nel, 3, P ** 3))
        for e in range(nel):
            (te[e],) = oracle.call_qf("MMSTrueSoln", which, None, P ** 3, [xn[e]], [3])
        true = np.zeros(mesh.lsize(p))
        np.add.at(true, idx, te.reshape(-1))
        true /= oracle.multiplicity(nel, P ** 3, 3, mesh.lsize(p), mesh.offsets(p))
        parts.append(("nodal true solution", true))
    return parts


def split_like(res, parts, nnodes):
    """cut the driver's raw output into the pieces of oracle_outputs (the diagnostic block is [node][8] in the file)"""
    out, pos, block = [], 0, None
    for name, v in parts:
        if name.startswith("diagnostic column"):
            if block is None:
                block = res[pos:pos + 8 * nnodes].reshape(nnodes, 8)
                pos += 8 * nnodes
            out.append(block[:, int(name.split()[-1])])
            continue
        out.append(res[pos:pos + v.size])
        pos += v.size
    assert pos == res.size, "driver output has an unexpected length"
    return out


def check_against_golden(problem, n, degrees, res):
    if not os.path.exists(GOLDEN):
        return False
    key = case_key(problem, n, degrees)
    with np.load(GOLDEN) as g:
        if key not in g.files:
            return False
        assert res.shape == g[key].shape and rel_err(res, g[key]) < TOL, ("golden", key)
    return True


def run_and_check(driver, tmp_path, problem, n, degrees, memtype, resource, forcing=0):
    """run `driver` on a seeded case and compare everything it writes (1e-12) with the oracle and -- for the cases of
    GOLDEN_CASES -- with the committed output of the reference's own host code and QFunctions on the CPU
    (tests/golden/ref_host_code_golden.npz, generator tests/golden/make_ref_host_golden.py)"""
    o, frees, xs, res = run_driver(driver, tmp_path, problem, n, degrees, memtype, resource, forcing)
    if not FORCING_ID[forcing]:
        check_against_golden(problem, n, degrees, res)
    parts = oracle_outputs(problem, degrees, o, frees, xs, forcing)
    for (name, want), got in zip(parts, split_like(res, parts, o.mesh.num_nodes(degrees[-1]))):
        assert rel_err(got, want) < TOL, name
