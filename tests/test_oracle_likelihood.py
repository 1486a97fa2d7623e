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
FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_likelihood_*.json")))
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


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p)[15:-5])
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
    doc = json.load(open(os.path.join(GOLDEN, "ref_likelihood_t001_HD.json")))
    lk = _oracle(doc)
    assert [int(x) for x in lk.alleles] == [15, 41]            # README.md:79
    assert round(lk.PP, 6) == 1.0 and lk.label == "risk"
    assert doc["outputs"]["n_points"] == 521                    # 2 x 261 candidates


from oracle.closed_form import closed_form_surface as _closed_form_surface  # noqa: E402


@pytest.mark.parametrize("name", ["t001_HD", "t002_DM1", "t002_DM1_max1200", "t001_HD_haploid", "t001_HD_nope"])
def test_dense_oracle_equals_closed_form(name):
    doc = json.load(open(os.path.join(GOLDEN, "ref_likelihood_{}.json".format(name))))
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
    doc = json.load(open(os.path.join(GOLDEN, "ref_likelihood_t002_DM1.json")))
    h2 = doc["outputs"]["h2range"]
    assert len(h2) >= len(set(h2))
    assert doc["outputs"]["n_points"] == sum(1 for a in doc["outputs"]["h1range"] for b in h2 if a <= b)


@pytest.mark.parametrize("name", ["t001_HD", "t002_DM1", "t002_DM1_max1200"])
def test_separability_the_far_region_tables_rely_on(name):
    """grid.cu replaces most points of a long-expansion surface by table look-ups.  The identities behind it,
    checked here on the independent closed form (exact equality — the same expression is evaluated):
      * h2 >= H1 = max(max spanning key + 19, max_partial, READLEN - 9, READLEN + 1): the spanning and partial
        terms do not depend on h2;
      * h2 >= H2 = max(H1, pe_ref + 1000 - min{target length >= MINPE}): nor does the paired-end term;
      * everywhere: the repeat-only term depends on max(h1-L,1) + max(h2-L,1) only."""
    doc = json.load(open(os.path.join(GOLDEN, "ref_likelihood_{}.json".format(name))))
    inp, out = doc["inputs"], doc["outputs"]
    step, w = _models()
    ml = _closed_form_surface(inp, step, w)
    K, L = inp["period"], inp["READLEN"]
    ks = max([int(k) * K for k in inp["FULL"]] or [-10 ** 6])
    mp = max([L - 18] + [int(k) * K for k in inp["PREF"]])
    H1 = max(ks + 19, mp, L - 9, L + 1)
    tl = [x + 1000 if x < 0 else x for x in inp["target_lens"]]
    tmin = min([x for x in tl if x >= inp["MINPE"]] or [None]) if out["run_pe"] else None
    H2 = max(H1, inp["pe_ref"] + 1000 - tmin) if tmin is not None else H1
    rng = np.random.default_rng(11)
    far = [H2 + int(x) for x in rng.integers(0, 2000, size=6)]
    mid = [H1 + int(x) for x in rng.integers(0, max(1, H2 - H1), size=6)]
    for h1 in [K, 5 * K, ks, max(K, ks - K), mp, H1 - 1, H1, H1 + 7 * K, H2, H2 + 40 * K]:
        if h1 <= 0:
            continue
        a = [ml(h1, h2, out["run_pe"]) for h2 in mid + far if h2 >= h1]
        assert len({(x[0], x[1]) for x in a}) <= 1, h1                      # span, partial: h1 only
        b = [ml(h1, h2, out["run_pe"]) for h2 in far if h2 >= h1]
        assert len({x[3] for x in b}) <= 1, h1                              # paired-end: h1 only
    for _ in range(50):
        d = int(rng.integers(2, 3000))
        pts = [(h1, d - max(h1 - L, 1) + L) for h1 in (L - 30, L + 1, L + d // 2) if 1 <= d - max(h1 - L, 1)]
        pts = [(h1, h2) for h1, h2 in pts if h1 > 0 and h2 > L and h1 <= h2]
        assert len({ml(h1, h2, False)[2] for h1, h2 in pts}) <= 1
