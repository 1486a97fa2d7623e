"""Oracle (CPU) evidence extraction — the bam_parser.py restatement — pinned to what the reference itself
publishes (README.md:77-86: FR / PR / RR strings of t001-HD and t002-DM1) and to the per-read golden
vectors generated with the reference's unmodified ssw.c (tests/golden/make_fixtures.py).  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import evidence_oracle as evo
from tredparse_b200 import bamio
from tredparse_b200.meta import TREDsRepo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# README.md:77-86 of the reference (evidence columns; PR is truncated there, the tail is what it prints)
README = {
    ("t001", "HD"): {"FR": "15|4", "RR": "",
                     "PR": "6|1;7|2;8|2;9|2;11|1;14|2;15|4;20|1;21|1;24|2;29|1;34|1;41|1"},
    ("t002", "DM1"): {"FR": "5|24", "RR": "49|3;50|8",
                      "PR": "4|5;5|2;6|2;7|1;8|1;12|2;13|1;14|1;16|2;18|1;19|1;21|2;23|1;24|2;27|1;28|1;29|1;"
                            "30|3;31|1;33|1;36|1;38|1;39|1;40|1;42|1;43|1;46|2"},
}
TAGCODE = {"FULL": 1, "PREF": 2, "POST": 3, "REPT": 4, "HANG": 5}


def counter_s(c):
    return ";".join("{}|{}".format(k, int(v)) for k, v in sorted(c.items()))


@pytest.fixture(scope="module")
def repo():
    return TREDsRepo()


@pytest.mark.parametrize("sample,tredname", sorted(README))
def test_evidence_strings_match_reference_readme(sample, tredname, repo):
    tred = repo[tredname]
    sam = bamio.AlignmentFile(os.path.join(GOLDEN, sample + ".mini.bam"))
    READLEN = evo.read_length(sam)
    assert READLEN == 150
    ev = evo.EvidenceOracle(tred, READLEN, alts=True, repeatpairs=True)
    ev.parse(sam)
    want = README[(sample, tredname)]
    assert counter_s(ev.counts["FULL"]) == want["FR"]
    assert counter_s(ev.counts["PREF"]) == want["PR"]
    assert counter_s(ev.counts["REPT"]) == want["RR"]
    assert ev.counts["PREF"] is ev.counts["POST"]          # quirk Q12: one shared histogram
    # golden per-read outcome (score, h, tag) produced with the reference's own ssw.c
    d = np.load(os.path.join(GOLDEN, "sw_pairs_{}_{}.npz".format(sample, tredname)))
    best = np.array([[-1, 0, 0] if b is None else [b[0], b[1], TAGCODE[b[2]]] for (_, _, b) in ev.read_log],
                    dtype=np.int32)
    assert [n for (n, _, _) in ev.read_log] == list(d["names"])
    assert np.array_equal(best, d["best"])
    assert str(d["FR"]) == want["FR"] and str(d["PR"]) == want["PR"] and str(d["RR"]) == want["RR"]


def test_template_family_order_and_rc(repo):
    t = repo["HD"]
    db = evo.template_family(t.prefix, t.repeat, t.suffix, 50)
    assert len(db) == 100 and [u for u, _ in db[:4]] == [1, 1, 2, 2]
    assert db[0][1] == t.prefix + t.repeat + t.suffix and db[1][1] == evo.rc(db[0][1])
    assert len(db[-1][1]) == 36 + 3 * 50
    assert evo.rc("ACGTN") == "NACGT"


@pytest.mark.parametrize("args,tag", [
    # score, rb, re, qb, qe, m, n, units, period, max_units
    ((100, 0, 99, 50, 149, 150, 100, 21, 3, 50), "FULL"),      # both template ends reached
    ((60, 0, 59, 90, 149, 150, 120, 28, 3, 50), "PREF"),       # starts in the prefix, read ends inside
    ((60, 60, 119, 0, 59, 150, 120, 28, 3, 50), "POST"),
    ((150, 20, 169, 0, 149, 150, 186, 50, 3, 50), "REPT"),     # buried in the repeat, longest templates only
    ((150, 20, 169, 0, 149, 150, 183, 48, 3, 50), None),       # same alignment on a shorter template: dropped
    ((70, 30, 99, 40, 109, 150, 120, 28, 3, 50), "HANG"),       # >= 9 unaligned bases on both sides
    ((29, 0, 28, 0, 28, 150, 39, 1, 3, 50), None),             # below min_score = max(min_len, 30)
    ((60, 0, 59, 0, 20, 150, 120, 28, 3, 50), None),           # aligned query span < min_len
])
def test_classification_rules(args, tag):
    assert evo.classify_alignment(*args) == tag


def test_missing_locus_yields_no_evidence(repo):
    """BASELINE config 2: only HD (t001) and DM1 (t002) have reads; every other locus parses to nothing."""
    sam = bamio.AlignmentFile(os.path.join(GOLDEN, "t001.mini.bam"))
    for name in ("DM1", "FXS", "SCA1"):
        ev = evo.EvidenceOracle(repo[name], 150, alts=True, repeatpairs=True)
        ev.parse(sam)
        assert not ev.counts["FULL"] and not ev.counts["PREF"] and ev.rept == 0
