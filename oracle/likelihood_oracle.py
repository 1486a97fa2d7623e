"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Dense restatement of the statistical half of the reference's hot path, ``tredparse/models.py``
(citations are into /root/reference/).  It deliberately keeps the reference's *dense* formulation —
every candidate allele owns length-1000 probability vectors, every grid point takes the log of whole
vectors — so that it is independent of the closed forms the CUDA kernels evaluate
(SURVEY.md Appendix B).  Python-2 integer ``/`` is written ``//``.

  step / stutter model files ............... models.py:42-84
  spanning / partial pdfs, mixing weights ... models.py:149-190      (quirks Q6, Q7, Q8)
  spanning / partial / repeat / PE terms .... models.py:192-221, 426-473
  candidate ranges + grid ................... models.py:223-273        (quirk Q9: duplicated candidates)
  marginals, CI, sparsify, PP, label ........ models.py:277-392        (quirks Q10, Q11)

Parity status: PINNED only at the level the reference pins anything — the README.md:77-86 calls
(t001/HD = 15|41 with PP 1; tests/test_oracle_likelihood.py) — the finer surface has no reference
vector, and scipy.stats.gaussian_kde / poisson.pmf are unpinned third-party numerics:
scipy 1.18 in FP64 is what this oracle evaluates.
"""
from collections import defaultdict
from math import exp

import numpy as np
from scipy.stats import gaussian_kde, poisson

SPAN = 1000
FLANKMATCH = 9
MAX_PERIOD = 6
SMALL_VALUE = exp(-10)
REALLY_SMALL_VALUE = exp(-100)
MIN_SPANNING_PAIRS = 5


def load_step_model(path):
    """-> {period: np.array(37)} (models.py:46-61): 6 scalars, ProbIncrease, then 6 'PeriodNModel' rows;
    periods 6..17 reuse the period-6 row."""
    with open(path) as fp:
        lines = fp.read().split("\n")
    steps = {}
    for i in range(MAX_PERIOD):
        steps[i + 1] = np.array([float(x) for x in lines[MAX_PERIOD + 1 + i].split()[1:]])
    for i in range(MAX_PERIOD, 3 * MAX_PERIOD):
        steps[i] = steps[MAX_PERIOD]
    return steps


def load_noise_model(path):
    """-> list of 5 logistic weights (models.py:68-77): skip 6 header lines, keep non-empty rows."""
    with open(path) as fp:
        rows = fp.read().split("\n")[MAX_PERIOD:]
    return [float(r) for r in (x.strip() for x in rows) if r]


def noise_predict(weights, x):
    z = weights[0]
    for b, xx in zip(weights[1:], x):
        z += b * xx
    return 1.0 / (1 + exp(-1 * z))


def safe_log(v):
    v[v < SMALL_VALUE] = SMALL_VALUE
    return np.log(v)


def mean_std(a):
    if not a:
        return ""
    a = np.array(a)
    return "{:.0f}+/-{:.0f}bp".format(a.mean(), a.std())


def histogram(a, bins=40):
    if not a:
        return ""
    ar, br = np.histogram(a, bins=bins, range=(0, SPAN))
    return ",".join("{}:{}".format(int(b), n) for (n, b) in zip(ar, br))


class PEModelOracle:
    def __init__(self, global_lens, target_lens, ref, MINPE):
        self.MINPE = MINPE
        kde = gaussian_kde(global_lens)
        pdf = kde.evaluate(np.arange(SPAN))
        self.pdf = pdf / pdf.sum()
        self.target_lens = list(target_lens)
        self.ref = ref
        self.memo = {}

    def roll(self, h):
        if h in self.memo:
            return self.memo[h]
        shift = self.ref - h
        p = np.roll(self.pdf, shift)
        if shift > 0:
            p[:shift] = SMALL_VALUE
        elif shift < 0:
            p[shift:] = SMALL_VALUE
        p[:self.MINPE] = SMALL_VALUE
        self.memo[h] = p
        return p

    def evaluate(self, h1, h2):
        mix = 0.5 * self.roll(h1) + (1 - 0.5) * self.roll(h2)
        mm = safe_log(mix)
        return sum(mm[tl] for tl in self.target_lens)


class LikelihoodOracle:
    """IntegratedCaller restated.  Inputs are plain values instead of a BamParser:

    counts      {"FULL": {units: n}, "PREF": {units: n}}   (PREF already includes POST, Q12)
    rept        number of REPT reads
    depth       mean depth (half_depth = depth / 2)
    pe          object with global_lens, target_lens, ref, MINPE  (or None)
    tred        object with cutoff_prerisk, cutoff_risk, is_expansion, is_recessive
    """

    def __init__(self, tred, period, readlen, counts, rept, ploidy, depth, pe, step_model,
                 noise_weights, score=1.0, gc=.68, maxinsert=300, fullsearch=False):
        self.tred = tred
        self.period = period
        self.readlen = readlen
        self.t1 = readlen - FLANKMATCH
        self.t2 = readlen - 2 * FLANKMATCH
        self.t3 = readlen - 3 * FLANKMATCH
        self.max_partial = self.t2
        self.score, self.gc = score, gc
        self.counts, self.rept, self.ploidy = counts, rept, ploidy
        self.half_depth = depth / 2
        self.maxinsert, self.fullsearch = maxinsert, fullsearch
        self.step_model, self.noise_weights = step_model, noise_weights
        gl = list(pe.global_lens) if pe is not None else []
        tl = list(pe.target_lens) if pe is not None else []
        self.pemodel = PEModelOracle(gl, tl, pe.ref, pe.MINPE) \
            if (len(gl) >= 100 and len(tl) >= MIN_SPANNING_PAIRS) else None
        self.PEDP = len(tl)
        self.PEG, self.PET = mean_std(gl), mean_std(tl)
        self.P_PEG, self.P_PET = histogram(gl), histogram(tl)
        self.spanning_db, self.partial_db = {}, {}
        self.P_h1 = self.P_h2 = self.P_h1h2 = ""
        self.surface = []           # [(ml1, ml2, ml3, ml4, ml, h1, h2)] in evaluation order

    # pdfs ---------------------------------------------------------------------------------------
    def pdf_spanning(self, h):
        if h in self.spanning_db:
            return self.spanning_db[h]
        a = np.zeros(SPAN)
        stutter = noise_predict(self.noise_weights, (self.period, h // self.period, self.gc, self.score))
        p = self.step_model[self.period] * stutter
        lp = len(p)
        dev = lp // 2
        p[dev] = 1 - stutter
        start, end = max(h - dev, 0), min(h + dev + 1, SPAN)
        a[start:end] = p[lp - end + start:lp]
        self.spanning_db[h] = a
        return a

    def pdf_partial(self, h):
        if h in self.partial_db:
            return self.partial_db[h]
        if h > self.max_partial:
            h = self.max_partial
        a = np.zeros(SPAN)
        c = 1. / (h + 1)
        a[:h] = c
        a += c * self.pdf_spanning(h)
        self.partial_db[h] = a
        return a

    def get_alpha(self, h1, h2, mode):
        if mode == 0:
            s1, s2 = max(0, self.t2 - h1), max(0, self.t2 - h2)
        else:
            s1, s2 = min(h1, self.t1), min(h2, self.t1)
        return s1 * 1. / (s1 + s2) if (s1 + s2) else .5

    # terms --------------------------------------------------------------------------------------
    def evaluate_spanning(self, obs, h1, h2):
        alpha = self.get_alpha(h1, h2, 0)
        ls = safe_log(alpha * self.pdf_spanning(h1) + (1 - alpha) * self.pdf_spanning(h2))
        return sum(ls[k] * c for k, c in obs.items())

    def evaluate_partial(self, obs, h1, h2):
        alpha = self.get_alpha(h1, h2, 1)
        lp = safe_log(alpha * self.pdf_partial(h1) + (1 - alpha) * self.pdf_partial(h2))
        return sum(lp[k] * c for k, c in obs.items())

    def evaluate_rept(self, n_obs_rept, h1, h2):
        d1 = max(h1 - self.readlen, 1)
        d2 = max(h2 - self.readlen, 1)
        mu = (d1 + d2) * self.half_depth / self.readlen
        return np.log(max(poisson.pmf(n_obs_rept, mu), REALLY_SMALL_VALUE))

    # grid ---------------------------------------------------------------------------------------
    def candidate_ranges(self, obs_spanning, obs_partial, n_obs_rept):
        """-> (h1range, h2range, run_pe) or None when there is no evidence (models.py:224-257)."""
        period = self.period
        max_full = max(obs_spanning.keys()) if obs_spanning else 0
        max_partial = max(obs_partial.keys()) if obs_partial else 0
        reads_above_full = sum(c for k, c in obs_partial.items() if k > max_full + period)
        run_pe = max_partial >= self.t3 and reads_above_full > 1 and (self.pemodel is not None)
        possible = set(obs_spanning.keys())
        if obs_partial:
            if max_partial > self.max_partial:
                self.max_partial = max_partial
            possible.add(max_partial)
        if not possible:
            return None
        base = sorted(possible)
        extended = base + list(range(max_partial + period, period * self.maxinsert + 1, period))
        if self.fullsearch:
            h1range = h2range = list(range(period, period * self.maxinsert + 1, period))
        else:
            h1range = base if max_full else extended
            h2range = extended if (n_obs_rept or run_pe) else base
        return h1range, h2range, run_pe

    def evaluate(self, obs_spanning, obs_partial, n_obs_rept):
        rng = self.candidate_ranges(obs_spanning, obs_partial, n_obs_rept)
        if rng is None:
            return None, None, None, None
        h1range, h2range, run_pe = rng
        self.h1range, self.h2range, self.run_pe = h1range, h2range, run_pe
        mls = []
        self.surface = []
        for h1 in h1range:
            for h2 in ([h1] if self.ploidy == 1 else h2range):
                if h1 > h2:
                    continue
                ml1 = self.evaluate_spanning(obs_spanning, h1, h2) if obs_spanning else 0
                ml2 = self.evaluate_partial(obs_partial, h1, h2) if obs_partial else 0
                ml3 = self.evaluate_rept(n_obs_rept, h1, h2)
                ml4 = self.pemodel.evaluate(h1, h2) if run_pe else 0
                ml = ml1 + ml2 + ml3 + ml4
                mls.append((ml, (h1, h2)))
                self.surface.append((ml1, ml2, ml3, ml4, ml, h1, h2))

        P_h1, P_h2, P_h1h2 = defaultdict(float), defaultdict(float), {}
        max_ml = max(mls)[0]
        for ml, (h1, h2) in mls:
            w = exp(ml - max_ml)
            P_h1[h1] += w
            P_h2[h2] += w
            P_h1h2[(h1, h2)] = w
        h1_lo, h1_hi = self.calc_CI(P_h1)
        h2_lo, h2_hi = self.calc_CI(P_h2)
        p = self.period
        CIs = (h1_lo // p, h1_hi // p, h2_lo // p, h2_hi // p)
        self.P_h1 = self.sparsify(P_h1)
        self.P_h2 = self.sparsify(P_h2)
        self.P_h1h2 = self.sparsify(P_h1h2)
        lik, alleles = max(mls, key=lambda x: (x[0], -x[1][0]))
        PP = self.calc_PP(lik, np.array([x[0] for x in mls]), mls)
        return alleles, lik, PP, CIs

    def sparsify(self, P):
        Z = {}
        total = sum(v for v in P.values())
        for k, v in P.items():
            if v < SMALL_VALUE:
                continue
            ks = k if isinstance(k, (list, tuple)) else [k]
            Z[",".join(str(x // self.period) for x in ks)] = v / total
        return Z

    @staticmethod
    def calc_CI(P):
        cum, lo, hi, in_range = 0, 0, 0, False
        total = sum(P.values())
        k = 0
        for k, v in sorted(P.items()):
            cum += v
            if (not in_range) and cum > .025 * total:
                in_range, lo = True, k
            if cum > .975 * total:
                break
        hi = k
        return lo, hi

    def calc_PP(self, lik, all_liks, mls):
        t, p = self.tred, self.period
        if t.is_expansion:
            if not t.is_recessive:
                sel = [x[0] for x in mls if max(x[1]) // p >= t.cutoff_risk]
            else:
                sel = [x[0] for x in mls if min(x[1]) // p >= t.cutoff_risk]
        else:
            if not t.is_recessive:
                sel = [x[0] for x in mls if min(x[1]) // p <= t.cutoff_risk]
            else:
                sel = [x[0] for x in mls if max(x[1]) // p <= t.cutoff_risk]
        sel = np.array(sel)
        return min(1, np.exp(sel - lik).sum() / np.exp(all_liks - lik).sum())

    def calc_label(self, alleles):
        t = self.tred
        a, b = sorted(alleles)
        label = "ok" if a != -1 else "missing"
        pre, risk = t.cutoff_prerisk, t.cutoff_risk
        if t.is_expansion:
            crit = a if t.is_recessive else b
            if pre <= crit < risk:
                label = "prerisk"
            elif crit >= risk:
                label = "risk"
        else:
            crit = b if t.is_recessive else a
            if pre <= crit < risk:
                label = "prerisk"
            elif 0 < crit <= risk:
                label = "risk"
        return label

    def call(self):
        obs_spanning = dict((k * self.period, v) for k, v in self.counts["FULL"].items())
        obs_partial = dict((k * self.period, v) for k, v in self.counts["PREF"].items())
        alleles, lik, PP, CIs = self.evaluate(obs_spanning, obs_partial, self.rept)
        if not alleles:
            alleles = (-1, -1)
            lik = PP = -1
        # (-1 // period == -1 for every period >= 1, matching Python 2's floor division)
        self.alleles = sorted(x // self.period for x in alleles)
        self.lik = lik
        self.label = self.calc_label(self.alleles)
        self.CI = "{}-{}|{}-{}".format(*CIs) if CIs else ""
        self.PP = PP
        return self
