"""Parity of the CUDA likelihood grid / KDE (through the C ABI) against the dense CPU oracle and the
committed golden surfaces.  Needs a GPU: run with -m gpu."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-9          # north_star: log-likelihood surface within 1e-9 relative in FP64


class _PE:
    def __init__(self, g, t, ref, minpe):
        self.global_lens, self.target_lens, self.ref, self.MINPE = g, t, ref, minpe


def _run_case(doc):
    from tredparse_b200.meta import TREDsRepo
    from tredparse_b200 import models
    inp = doc["inputs"]
    tred = TREDsRepo()[inp["tred"]]
    period = inp["period"]
    obs_s = {int(k) * period: v for k, v in inp["FULL"].items()}
    obs_p = {int(k) * period: v for k, v in inp["PREF"].items()}
    has_pe = len(inp["global_lens"]) >= 100 and len(inp["target_lens"]) >= 5
    pdf = models.pe_kde([inp["global_lens"]])[0] if has_pe else None
    batch = models.GridBatch()
    i = batch.add(tred, period, inp["READLEN"], obs_s, obs_p, inp["rept"], inp["ploidy"], inp["depth"],
                  pdf, inp["target_lens"], inp["pe_ref"], inp["MINPE"], maxinsert=inp["maxinsert"],
                  fullsearch=inp["fullsearch"])
    assert i == 0
    batch.run()
    return batch, batch.summarize(0), tred, period


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_likelihood_*.json"))),
                         ids=lambda p: os.path.basename(p)[15:-5])
def test_surface_and_call_match_golden(path):
    doc = json.load(open(path))
    out = doc["outputs"]
    batch, s, tred, period = _run_case(doc)
    meta = batch.meta[0]
    assert list(meta[3]) == out["h1range"] and list(meta[4]) == out["h2range"]
    assert bool(meta[5]) == out["run_pe"]
    assert s["n_points"] == out["n_points"]
    # surface, point by point in the reference's evaluation order
    S = batch.surface_of(0)
    ploidy = doc["inputs"]["ploidy"]
    got = []
    for i1, h1 in enumerate(meta[3]):
        for i2, h2 in enumerate([h1] if ploidy == 1 else meta[4]):
            if h1 > h2:
                assert S[i1, i2] == -np.inf
                continue
            got.append(S[i1, i2])
    gold = np.array([row[4] for row in out["surface"]])
    got = np.array(got)
    assert got.shape == gold.shape
    assert np.max(np.abs(got - gold) / np.abs(gold)) < RTOL
    # call, CI, PP, label, marginals
    assert sorted(x // period for x in s["alleles"]) == out["alleles"]
    assert abs(s["lik"] - out["lik"]) <= RTOL * abs(out["lik"])
    assert "{}-{}|{}-{}".format(*s["CIs"]) == out["CI"]
    assert abs(s["PP"] - out["PP"]) <= 1e-9
    from tredparse_b200.models import calc_label
    assert calc_label(tred, out["alleles"]) == out["label"]
    for name in ("P_h1", "P_h2", "P_h1h2"):
        assert set(s[name].keys()) == set(out[name].keys()), name
        for k, v in out[name].items():
            assert abs(s[name][k] - v) <= 1e-9 * max(v, 1e-300) + 1e-15, (name, k)


def test_kde_matches_scipy():
    from scipy.stats import gaussian_kde
    from tredparse_b200 import models
    rng = np.random.default_rng(3)
    sets = [np.clip(rng.normal(350, 75, 2800).astype(int), 0, 999),
            np.clip(rng.normal(420, 90, 150).astype(int), -50, 999),
            json.load(open(os.path.join(GOLDEN, "ref_likelihood_t001_HD.json")))["inputs"]["global_lens"]]
    got = models.pe_kde(sets)
    for x, g in zip(sets, got):
        pdf = gaussian_kde(np.asarray(x, dtype=float)).evaluate(np.arange(1000))
        pdf = pdf / pdf.sum()
        assert np.max(np.abs(g - pdf) / np.maximum(pdf, 1e-300)) < 1e-9
        assert abs(g.sum() - 1) < 1e-12


def test_batched_problems_equal_single_problem_runs():
    """Many problems in one launch give the same numbers as one launch each (pool offsets)."""
    from tredparse_b200 import models
    from tredparse_b200.meta import TREDsRepo
    docs = [json.load(open(p)) for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_likelihood_*.json")))]
    singles = [_run_case(d)[1] for d in docs]
    batch = models.GridBatch()
    repo = TREDsRepo()
    for d in docs:
        inp = d["inputs"]
        period = inp["period"]
        has_pe = len(inp["global_lens"]) >= 100 and len(inp["target_lens"]) >= 5
        pdf = models.pe_kde([inp["global_lens"]])[0] if has_pe else None
        batch.add(repo[inp["tred"]], period, inp["READLEN"], {int(k) * period: v for k, v in inp["FULL"].items()},
                  {int(k) * period: v for k, v in inp["PREF"].items()}, inp["rept"], inp["ploidy"], inp["depth"],
                  pdf, inp["target_lens"], inp["pe_ref"], inp["MINPE"], maxinsert=inp["maxinsert"],
                  fullsearch=inp["fullsearch"])
    batch.run()
    for i, s in enumerate(singles):
        b = batch.summarize(i)
        assert b["alleles"] == s["alleles"] and b["lik"] == s["lik"] and b["PP"] == s["PP"] and b["CIs"] == s["CIs"]


def test_no_evidence_problem():
    from tredparse_b200 import models
    from tredparse_b200.meta import TREDsRepo
    batch = models.GridBatch()
    assert batch.add(TREDsRepo()["HD"], 3, 150, {}, {}, 0, 2, 30.0, None, [], 57, 77) == -1
