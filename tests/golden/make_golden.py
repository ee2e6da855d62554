"""Generate tests/golden/qf_golden.npz from the REFERENCE's own QFunctions.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It calls oracle/_ref/libref_qf.so -- the reference's qfunctions/*.h compiled from the
sources where they lie by oracle/Makefile -- on seeded inputs and stores inputs and
outputs.  The .npz travels with the repo; nothing at test time reads /root/reference.

Two input families:
  * "katF": the SURVEY.md Appendix F known-answer inputs (Q=2, nu=0.3, E=1);
  * "rnd":  64 seeded random admissible points (|grad u| ~ 0.05, perturbed Jacobians).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def kat_inputs():
    Q = 2
    q = np.arange(Q)
    J = np.zeros((3, 3, Q)); ug = np.zeros((3, 3, Q)); dug = np.zeros((3, 3, Q))
    for a in range(3):
        for b in range(3):
            J[a, b] = (0.5 if a == b else 0.0) + 0.01 * (1 + a + 2 * b + q)
            ug[a, b] = 0.02 * np.sin(1 + a + 3 * b + q)
            dug[a, b] = np.cos(0.5 + 2 * a + b + q)
    w = 0.125 * (q + 1.0)
    return Q, J, w, ug, dug


def rnd_inputs(Q=64, seed=7):
    rng = np.random.default_rng(seed)
    J = 0.3 * np.eye(3)[:, :, None] + 0.05 * rng.standard_normal((3, 3, Q))
    w = 0.05 + 0.1 * rng.random(Q)
    ug = 0.02 * rng.standard_normal((3, 3, Q))
    dug = rng.standard_normal((3, 3, Q))
    return Q, J, w, ug, dug


def run(which, Q, J, w, ug, dug, nu=0.3, E=1.0):
    phys = oracle.Physics(nu, E)
    out = {}
    (qd,) = oracle.call_qf("SetupGeo", which, None, Q, [J.reshape(9, Q), w.reshape(1, Q)], [10])
    out["qdata"] = qd
    (out["LinElasF"],) = oracle.call_qf("LinElasF", which, phys, Q, [ug.reshape(9, Q), qd], [9])
    (out["LinElasdF"],) = oracle.call_qf("LinElasdF", which, phys, Q, [dug.reshape(9, Q), qd], [9])
    for name in ("HyperSS", "HyperFS"):
        f, gradu = oracle.call_qf(name + "F", which, phys, Q, [ug.reshape(9, Q), qd], [9, 9])
        (df,) = oracle.call_qf(name + "dF", which, phys, Q, [dug.reshape(9, Q), qd, gradu], [9])
        out[name + "F"], out[name + "F_gradu"], out[name + "dF"] = f, gradu, df
    # one-shot post-processing QFunctions (setuplibceed.c:645-737): strain energy, nodal diagnostics
    upt = np.ascontiguousarray(J.reshape(9, Q)[3:6] - 0.1)   # any numbers: the displacement passes through
    for name in ("LinElas", "HyperSS", "HyperFS"):
        (out[name + "Energy"],) = oracle.call_qf(name + "Energy", which, phys, Q, [ug.reshape(9, Q), qd], [1])
        (out[name + "Diagnostic"],) = oracle.call_qf(name + "Diagnostic", which, phys, Q, [upt, ug.reshape(9, Q), qd], [8])
    # forcing / manufactured solution: coordinates = the first three rows of J (any numbers will do)
    xyz = J.reshape(9, Q)[:3] + 0.25
    (out["SetupMMSForce"],) = oracle.call_qf("SetupMMSForce", which, phys, Q, [xyz, qd], [3])
    (out["MMSTrueSoln"],) = oracle.call_qf("MMSTrueSoln", which, None, Q, [xyz], [3])
    fv = (oracle.C.c_double * 3)(0.3, -1.0, 2.5)
    f = oracle.C.cast(oracle.qf("SetupConstantForce", which), oracle.QFN)
    inp = (oracle.C.POINTER(oracle.C.c_double) * 2)(xyz.ctypes.data_as(oracle.C.POINTER(oracle.C.c_double)),
                                                    qd.ctypes.data_as(oracle.C.POINTER(oracle.C.c_double)))
    o = np.zeros((3, Q))
    outp = (oracle.C.POINTER(oracle.C.c_double) * 1)(o.ctypes.data_as(oracle.C.POINTER(oracle.C.c_double)))
    assert f(oracle.C.cast(fv, oracle.C.c_void_p), Q, inp, outp) == 0
    out["SetupConstantForce"] = o
    return out


def main():
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle)"
    data = {}
    for tag, inputs in (("katF", kat_inputs()), ("rnd", rnd_inputs())):
        Q, J, w, ug, dug = inputs
        data[f"{tag}_J"], data[f"{tag}_w"], data[f"{tag}_ug"], data[f"{tag}_dug"] = J, w, ug, dug
        for k, v in run("ref", Q, J, w, ug, dug).items():
            data[f"{tag}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "qf_golden.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
