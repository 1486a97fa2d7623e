"""BASELINE config "long-expansion stress": DM1 / FXS candidate grids to 1000+ repeats with 150 bp and 250 bp
reads, at FULL size (maxinsert 1000 -> 500,500 grid points per problem with --fullsearch).  The dense oracle
cannot cover that in seconds, so parity is checked through size-independent properties:
  * random points of the surface vs the independent closed form (oracle/closed_form.py), 1e-9 relative;
  * the full-search surface restricted to the default-search candidates == the default-search surface;
  * max / arg-max / point count of the reductions recomputed on the host from the returned surface;
  * the all-device pipeline's call == the per-stage host path's call on the same problems.
Needs a GPU: run with -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-9


def _models():
    md = json.load(open(os.path.join(ROOT, "tredparse_b200", "data", "models.json")))
    step = {int(k): np.array(v) for k, v in md["step_size_by_period"].items()}
    for i in range(6, 18):
        step[i] = step[6]
    return step, md["stutter_weights"]


CASES = [("DM1", (13, 1000), 250, 1000), ("DM1", (12, 500), 150, 1000), ("FXS", (30, 800), 150, 1000),
         ("FXS", (250,), 250, 1200),
         ("DM1", (13, 1200), 250, 1500),      # config 5's larger grid: 1,125,750 points
         ("DM1", (5, 700), 150, 2100)]        # > 2048 candidate columns: the reduction exchanges column sums in two rounds


@pytest.fixture(scope="module")
def evidence():
    """reads -> (FULL, PREF, REPT) tallies through the family kernel, once per case"""
    from tredparse_b200 import simulate, cohort
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    out = []
    for i, (name, alleles, readlen, maxinsert) in enumerate(CASES):
        pr = simulate.simulate_problem(repo[name], alleles, readlen=readlen, cov_per_hap=15, seed=4000 + i)
        batch = cohort.CohortBatch([pr], maxinsert=maxinsert, fullsearch=True)
        res = batch.run_host(want_hist=True)
        out.append((pr, maxinsert, res["hist"][0], res["calls"][0]))
    return out


@pytest.mark.parametrize("k", range(len(CASES)))
def test_fullsearch_surface_properties(evidence, k):
    from tredparse_b200 import models, cohort
    from oracle.closed_form import closed_form_surface
    pr, maxinsert, hist, call = evidence[k]
    t, P, L = pr.tred, len(pr.tred.repeat), pr.readlen
    full = {int(u): int(c) for u, c in enumerate(hist[0]) if c}
    pref = {int(u): int(c) for u, c in enumerate(hist[1]) if c}
    rept = int(hist[2].sum())
    obs_s = {u * P: c for u, c in full.items()}
    obs_p = {u * P: c for u, c in pref.items()}
    ref = t.repeat_end - t.repeat_start + 1
    minpe = ref - 1 + 20
    has_pe = len(pr.global_lens) >= 100 and len(pr.target_lens) >= 5
    pdf = models.pe_kde([list(pr.global_lens)])[0] if has_pe else None
    gb = models.GridBatch()
    i_full = gb.add(t, P, L, obs_s, obs_p, rept, pr.ploidy, pr.depth, pdf, list(pr.target_lens), ref, minpe,
                    maxinsert=maxinsert, fullsearch=True)
    i_def = gb.add(t, P, L, obs_s, obs_p, rept, pr.ploidy, pr.depth, pdf, list(pr.target_lens), ref, minpe,
                   maxinsert=maxinsert, fullsearch=False)
    assert i_full == 0 and i_def == 1
    gb.run()
    S = gb.surface_of(0)
    h1r, h2r, run_pe = gb.meta[0][3], gb.meta[0][4], gb.meta[0][5]
    n = maxinsert
    assert list(h1r) == [P * i for i in range(1, n + 1)]
    if pr.ploidy == 2:
        assert S.shape == (n, n) and int(np.isfinite(S).sum()) == n * (n + 1) // 2 == gb.results[0]["n_points"]
        assert np.all(np.isneginf(S[np.tril_indices(n, -1)]))               # h1 > h2 is never evaluated
    else:
        assert S.shape == (n, 1) and gb.results[0]["n_points"] == n
    # (1) random points vs the closed form
    step, w = _models()
    inp = {"period": P, "READLEN": L, "FULL": full, "PREF": pref, "rept": rept, "depth": pr.depth,
           "global_lens": [int(x) for x in pr.global_lens], "target_lens": [int(x) for x in pr.target_lens],
           "pe_ref": ref, "MINPE": minpe}
    ml = closed_form_surface(inp, step, w)
    rng = np.random.default_rng(k)
    for _ in range(300):
        i1 = int(rng.integers(0, n))
        i2 = int(rng.integers(i1, n)) if pr.ploidy == 2 else 0
        h1 = h1r[i1]
        h2 = h2r[i2] if pr.ploidy == 2 else h1
        want = sum(ml(h1, h2, run_pe))
        assert abs(S[i1, i2] - want) <= RTOL * abs(want), (h1, h2, S[i1, i2], want)
    # (2) default-search surface == the same points of the full-search surface
    D = gb.surface_of(1)
    d1, d2 = gb.meta[1][3], gb.meta[1][4]
    for a, h1 in enumerate(d1):
        for b, h2 in enumerate([h1] if pr.ploidy == 1 else d2):
            if h1 > h2 or h2 > P * n:
                continue
            f = S[h1 // P - 1, (h2 // P - 1) if pr.ploidy == 2 else 0]
            assert D[a, b] == f or abs(D[a, b] - f) <= 1e-12 * abs(f)
    # (3) reductions recomputed on the host
    R = gb.results[0]
    finite = np.where(np.isfinite(S), S, -np.inf)
    assert R["max_ml"] == finite.max()
    i1, i2 = np.unravel_index(int(np.argmax(finite)), S.shape)            # first maximum == smallest h1 (Q10)
    assert (R["arg_i1"], R["arg_i2"]) == (i1, i2)
    W = np.exp(finite - finite.max())
    assert abs(R["sum_all"] - W.sum()) <= 1e-9 * R["sum_all"]
    Pd = gb.problems[0]
    ph1 = gb.marg[Pd["off_ph1"]:Pd["off_ph1"] + Pd["n_h1"]]
    ph2 = gb.marg[Pd["off_ph2"]:Pd["off_ph2"] + Pd["n_h2"]]
    np.testing.assert_allclose(ph1, W.sum(axis=1), rtol=1e-9, atol=1e-300)     # marginals (models.py:277-284)
    np.testing.assert_allclose(ph2, W.sum(axis=0), rtol=1e-9, atol=1e-300)
    H1 = np.array(h1r)[:, None]
    H2 = np.array(h2r)[None, :] if pr.ploidy == 2 else H1
    lo, hi = np.minimum(H1, H2) // P, np.maximum(H1, H2) // P
    if t.is_expansion:
        patho = (lo >= t.cutoff_risk) if t.is_recessive else (hi >= t.cutoff_risk)
    else:
        patho = (hi <= t.cutoff_risk) if t.is_recessive else (lo <= t.cutoff_risk)
    want_path = (W * np.broadcast_to(patho, W.shape)).sum()                     # PP numerator (models.py:342-368)
    assert abs(R["sum_path"] - want_path) <= 1e-9 * max(want_path, 1e-300)
    # (4) the all-device pipeline made the same call
    s = gb.summarize(0, want_joint=False)
    c = cohort.decode_call(call)
    assert c["alleles"] == sorted(x // P for x in s["alleles"])
    assert abs(c["lik"] - s["lik"]) <= RTOL * abs(s["lik"]) and abs(c["PP"] - s["PP"]) < 1e-9
    assert c["CI"] == "{}-{}|{}-{}".format(*s["CIs"])
