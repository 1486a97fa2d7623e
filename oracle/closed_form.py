"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Closed-form evaluation of one point of the reference's log-likelihood surface, written from the normative
statement in SURVEY.md Appendix B (tredparse/models.py:149-221, 426-473) independently of both the dense
oracle (likelihood_oracle.py) and the CUDA kernel.  It is cheap enough (O(keys) per point) to spot-check
surfaces with 10^5 - 10^6 points (BASELINE config "long-expansion stress"), where the dense oracle would
take minutes per problem.  tests/test_oracle_likelihood.py pins it against the dense oracle's golden surfaces.
"""
import math

import numpy as np

EPS, EPS2 = math.exp(-10), math.exp(-100)


def closed_form_surface(inp, step, w):
    """inp: the "inputs" dict of a tests/golden/ref_likelihood_*.json document (units-keyed counts);
    returns ml(h1, h2, run_pe) -> [ml1, ml2, ml3, ml4] with h in bp."""
    K, L = inp["period"], inp["READLEN"]
    t1, t2 = L - 9, L - 18
    S = {int(k) * K: v for k, v in inp["FULL"].items()}
    T = {int(k) * K: v for k, v in inp["PREF"].items()}
    U, D = inp["rept"], inp["depth"] / 2.0
    mp = max([t2] + list(T))
    stp = step[K]

    def sigma(h):
        z = w[0] + w[1] * K + w[2] * (h // K) + w[3] * 0.68 + w[4] * 1.0
        return 1.0 / (1.0 + math.exp(-z))

    def PS(h, k):
        if not (0 <= k < 1000):
            return 0.0
        if h + 19 <= 1000:
            d = k - h + 18
        else:
            if k < h - 18:
                return 0.0
            d = k - 963
        if d < 0 or d > 36:
            return 0.0
        sg = sigma(h)
        return (1 - sg) if d == 18 else stp[d] * sg

    def PT(h, k):
        hc = min(h, mp)
        c = 1.0 / (hc + 1)
        return (c if k < hc else 0.0) + c * PS(hc, k)

    have_pe = len(inp["global_lens"]) >= 100 and len(inp["target_lens"]) >= 5
    if have_pe:
        from scipy.stats import gaussian_kde
        g = gaussian_kde(inp["global_lens"]).evaluate(np.arange(1000))
        g = g / g.sum()

    def R(h, x):
        if x < 0:
            x += 1000
        y = x + h - inp["pe_ref"]
        if x < inp["MINPE"] or y < 0 or y >= 1000:
            return EPS
        return g[y]

    def lg(v):
        return math.log(max(v, EPS))

    def ml(h1, h2, run_pe):
        out = [0.0, 0.0, 0.0, 0.0]
        if S:
            s1, s2 = max(0, t2 - h1), max(0, t2 - h2)
            a = s1 / (s1 + s2) if s1 + s2 else 0.5
            out[0] = sum(c * lg(a * PS(h1, k) + (1 - a) * PS(h2, k)) for k, c in S.items())
        if T:
            s1, s2 = min(h1, t1), min(h2, t1)
            a = s1 / (s1 + s2) if s1 + s2 else 0.5
            out[1] = sum(c * lg(a * PT(h1, k) + (1 - a) * PT(h2, k)) for k, c in T.items())
        mu = (max(h1 - L, 1) + max(h2 - L, 1)) * D / L
        pk = (U * math.log(mu) if U else 0.0) - math.lgamma(U + 1) - mu
        out[2] = math.log(max(math.exp(pk), EPS2))
        if run_pe:
            out[3] = sum(lg(0.5 * R(h1, v) + 0.5 * R(h2, v)) for v in inp["target_lens"])
        return out
    return ml
