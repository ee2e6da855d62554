"""Pin the oracle's QFunction layer: restated port vs the committed golden vectors
(generated from the reference's own qfunctions/*.h by tests/golden/make_golden.py),
vs SURVEY.md Appendix F known answers, and vs the live oracle/_ref build when present."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402
from oracle import oracle  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "qf_golden.npz"))
NAMES = ["qdata", "LinElasF", "LinElasdF", "HyperSSF", "HyperSSF_gradu", "HyperSSdF",
         "HyperFSF", "HyperFSF_gradu", "HyperFSdF", "SetupMMSForce", "MMSTrueSoln", "SetupConstantForce",
         "LinElasEnergy", "LinElasDiagnostic", "HyperSSEnergy", "HyperSSDiagnostic", "HyperFSEnergy", "HyperFSDiagnostic"]

# SURVEY.md Appendix F (reference headers, gcc -O2): values at q = 1 unless noted
APPF = {
    "qdata@0": [0.019300000000000001, 1.974093264248705, -0.062176165803108821, -0.098445595854922269,
                -0.094559585492227968, 1.8730569948186528, -0.1593264248704663, -0.16321243523316065,
                -0.19170984455958548, 1.7797927461139897],
    "LinElasF": [0.0024545549266132695, -0.00050421114103224661, 0.00020321213877966986,
                 -0.00063036192242240567, -0.00054278581777295574, 0.00071638955580486361,
                 -4.2364554239563128e-05, 0.00068536585869476553, -0.001214636981782328],
    "LinElasdF": [0.018877087385668199, -0.056535256265106859, -0.013611828459057259,
                  -0.052009335042073758, -0.023203934403375984, 0.042013279006704783,
                  -0.0032218596138637989, 0.051277223385279153, 0.036026141066563891],
    "HyperSSF": [0.0024671415037638097, -0.0010334424991581821, 0.00032632036358469177,
                 -0.0011086826134297241, -0.00057282855647524582, 0.0013153991188547976,
                 0.00014953850319105924, 0.0013114372303044874, -0.0012998880873280358],
    "HyperSSdF": [0.021923572987849402, -0.11414655469162784, -0.024349736333968024,
                  -0.1030330290582379, -0.024356374649553123, 0.088237847582257889,
                  -0.0049250274429076899, 0.10000437425104267, 0.03153284736258713],
    "HyperFSF": [0.0024599763786419867, -0.0010259250610393023, 0.00033823400487648753,
                 -0.0011027199110915065, -0.00058173793626198095, 0.0013035006834497701,
                 0.00014528957128861697, 0.0013169971398477058, -0.0012924152480576203],
    "HyperFSdF": [0.0057954610606112322, -0.11128411560583007, -0.022790268092146453,
                  -0.10218930020665726, -0.039033383625925643, 0.091534031801229751,
                  -0.00066411056675458845, 0.1049627832299927, 0.017206687883524967],
}


def _inputs(tag):
    J, w, ug, dug = (GOLD[f"{tag}_{k}"] for k in ("J", "w", "ug", "dug"))
    return J.shape[-1], J, w, ug, dug


@pytest.mark.parametrize("tag", ["katF", "rnd"])
def test_port_matches_golden(tag):
    out = make_golden.run("port", *_inputs(tag))
    for name in NAMES:
        ref = GOLD[f"{tag}_{name}"]
        err = np.max(np.abs(out[name] - ref)) / np.max(np.abs(ref))
        assert err < 5e-14, (name, err)


def test_golden_matches_survey_appendix_f():
    for key, vals in APPF.items():
        if key == "qdata@0":
            got = GOLD["katF_qdata"][:, 0]
        else:
            got = GOLD["katF_" + key][:, 1]
        np.testing.assert_allclose(got, vals, rtol=2e-13, atol=1e-18, err_msg=key)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("tag", ["katF", "rnd"])
def test_live_reference_matches_golden(tag):
    out = make_golden.run("ref", *_inputs(tag))
    for name in NAMES:
        np.testing.assert_allclose(out[name], GOLD[f"{tag}_{name}"], rtol=1e-13, atol=1e-18)


def test_log1p_series_is_the_models_arithmetic():
    """hyperSS.h:37-42: the truncated series (not libm log1p) is part of the model."""
    Q, J, w, ug, dug = _inputs("katF")
    out = make_golden.run("port", Q, J, w, ug, dug)
    g = out["HyperSSF_gradu"].reshape(3, 3, Q)
    tr = g[0, 0] + g[1, 1] + g[2, 2]
    y = tr / (2 + tr)
    series = 2 * (y + y ** 3 / 3 + y ** 5 / 5 + y ** 7 / 7)
    assert np.max(np.abs(series - np.log1p(tr))) < 1e-7
