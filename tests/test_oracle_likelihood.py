"""Oracle (CPU) likelihood — the dense models.py restatement — against (1) the committed golden surfaces,
(2) what the reference pins (README.md:77-86: t001/HD = 15|41, PP 1) and (3) an independent closed-form
evaluation of SURVEY.md Appendix B written here from scratch, point by point.  No GPU needed."""
import glob
import json
import math
import os

import numpy as np
import pytest

from oracle import likelihood_oracle as lko
from tredparse_b200.meta import TREDsRepo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(GOLDEN, "likelihood_*.json")))
EPS, EPS2 = math.exp(-10), math.exp(-100)


def _models():
    md = json.load(open(os.path.join(ROOT, "tredparse_b200", "data", "models.json")))
    step = {int(k): np.array(v) for k, v in md["step_size_by_period"].items()}
    for i in range(6, 18):
        step[i] = step[6]
    return step, md["stutter_weights"]


class _PE:
    def __init__(self, g, t, ref, minpe):
        self.global_lens, self.target_lens, self.ref, self.MINPE = g, t, ref, minpe


def _oracle(doc):
    inp = doc["inputs"]
    step, w = _models()
    tred = TREDsRepo()[inp["tred"]]
    counts = {"FULL": {int(k): v for k, v in inp["FULL"].items()},
              "PREF": {int(k): v for k, v in inp["PREF"].items()}}
    pe = _PE(inp["global_lens"], inp["target_lens"], inp["pe_ref"], inp["MINPE"]) if inp["global_lens"] else None
    lk = lko.LikelihoodOracle(tred, inp["period"], inp["READLEN"], counts, inp["rept"], inp["ploidy"],
                              inp["depth"], pe, step, w, maxinsert=inp["maxinsert"], fullsearch=inp["fullsearch"])
    lk.call()
    return lk


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p)[11:-5])
def test_oracle_reproduces_golden(path):
    doc = json.load(open(path))
    out = doc["outputs"]
    lk = _oracle(doc)
    assert [int(x) for x in lk.alleles] == out["alleles"]
    assert lk.CI == out["CI"] and lk.label == out["label"]
    assert lk.PP == pytest.approx(out["PP"], rel=1e-12, abs=1e-15)
    assert lk.lik == pytest.approx(out["lik"], rel=1e-12)
    assert len(lk.surface) == out["n_points"]
    if "surface" in out:
        got = np.array([row[4] for row in lk.surface])
        gold = np.array([row[4] for row in out["surface"]])
        assert np.allclose(got, gold, rtol=1e-12, atol=0)
    for k in ("P_h1", "P_h2", "P_h1h2"):
        a, b = getattr(lk, k), out[k]
        assert set(a) == set(b)
        assert all(a[x] == pytest.approx(b[x], rel=1e-9) for x in a)


def test_reference_readme_call_t001_HD():
    doc = json.load(open(os.path.join(GOLDEN, "likelihood_t001_HD.json")))
    lk = _oracle(doc)
    assert [int(x) for x in lk.alleles] == [15, 41]            # README.md:79
    assert round(lk.PP, 6) == 1.0 and lk.label == "risk"
    assert doc["outputs"]["n_points"] == 521                    # 2 x 261 candidates


from oracle.closed_form import closed_form_surface as _closed_form_surface  # noqa: E402


@pytest.mark.parametrize("name", ["t001_HD", "t002_DM1", "t002_DM1_max1200", "t001_HD_haploid", "t001_HD_nope"])
def test_dense_oracle_equals_closed_form(name):
    doc = json.load(open(os.path.join(GOLDEN, "likelihood_{}.json".format(name))))
    inp, out = doc["inputs"], doc["outputs"]
    step, w = _models()
    ml = _closed_form_surface(inp, step, w)
    rows = out["surface"]
    pick = rows if len(rows) <= 600 else rows[::max(1, len(rows) // 600)]
    for (m1, m2, m3, m4, m, h1, h2) in pick:
        got = ml(h1, h2, out["run_pe"])
        assert got[0] == pytest.approx(m1, rel=1e-10, abs=1e-12)
        assert got[1] == pytest.approx(m2, rel=1e-10, abs=1e-12)
        assert got[2] == pytest.approx(m3, rel=1e-10, abs=1e-12)
        assert got[3] == pytest.approx(m4, rel=1e-10, abs=1e-12)
        assert sum(got) == pytest.approx(m, rel=1e-10)


def test_candidate_list_keeps_duplicates_quirk_Q9():
    """extended_range = base + range(...) is a list: spanning keys above max_partial appear twice."""
    doc = json.load(open(os.path.join(GOLDEN, "likelihood_t002_DM1.json")))
    h2 = doc["outputs"]["h2range"]
    assert len(h2) >= len(set(h2))
    assert doc["outputs"]["n_points"] == sum(1 for a in doc["outputs"]["h1range"] for b in h2 if a <= b)
