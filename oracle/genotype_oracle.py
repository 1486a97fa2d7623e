"""
ORACLE — TEST INFRASTRUCTURE ONLY.

One (sample, locus) problem end to end on the CPU, i.e. the body of the reference's per-sample loop
``tredparse/tred.py:225-275`` (depth -> BamParser.parse -> IntegratedCaller.call -> flattening into
``tredCalls``) built from evidence_oracle + likelihood_oracle.  ``engine='ref'`` aligns with the
reference's own ssw.c through the reference's per-call pattern; ``engine='oracle'`` with sw_oracle.c.

``samfile`` arguments are anything with pysam's fetch/getrname interface; model tables are passed in
(they live in the product's data directory; tests/test_data_tables.py pins them to the reference's files).
"""
from . import evidence_oracle as evo
from . import likelihood_oracle as lko

SPAN = 1000


def counter_s(c):
    return ";".join("{}|{}".format(k, int(v)) for k, v in sorted(c.items()))


def genotype_locus(samfile, tred, READLEN, step_model, noise_weights, gender="Unknown", clip=False,
                   alts=True, repeatpairs=True, ref="hg38", maxinsert=300, fullsearch=False,
                   engine="oracle", depth=None):
    """-> (calls dict with the '<T>.xxx' keys of tred.py:251-275 minus the prefix, evidence, caller)"""
    if depth is None:
        try:
            depth = evo.region_depth(samfile, tred.chr, max(0, tred.repeat_start - SPAN), tred.repeat_end + SPAN)
        except Exception:
            depth = 30
    ev = evo.EvidenceOracle(tred, READLEN, gender=gender, depth=depth, clip=clip, alts=alts,
                            repeatpairs=repeatpairs, ref=ref, engine=engine)
    ev.parse(samfile)
    pe = evo.PEOracle(samfile, tred.chr, tred.repeat_start, tred.repeat_end)
    counts = {"FULL": dict(ev.counts["FULL"]), "PREF": dict(ev.counts["PREF"])}
    lk = lko.LikelihoodOracle(tred, ev.period, READLEN, counts, ev.rept, ev.ploidy, depth, pe, step_model,
                              noise_weights, maxinsert=maxinsert, fullsearch=fullsearch)
    lk.call()
    calls = {
        "1": lk.alleles[0], "2": lk.alleles[1],
        "FR": counter_s(ev.counts["FULL"]), "PR": counter_s(ev.counts["PREF"]), "RR": counter_s(ev.counts["REPT"]),
        "DP": depth, "FDP": sum(ev.counts["FULL"].values()), "PDP": sum(ev.counts["PREF"].values()),
        "RDP": ev.rept, "PEDP": lk.PEDP, "PEG": lk.PEG, "PET": lk.PET, "CI": lk.CI, "PP": lk.PP,
        "label": lk.label, "details": ev.details, "P_h1": lk.P_h1, "P_h2": lk.P_h2, "P_h1h2": lk.P_h1h2,
        "P_PEG": lk.P_PEG, "P_PET": lk.P_PET,
    }
    return calls, ev, lk
