"""The all-device cohort pipeline (tredsw_genotype_batch) against (i) the CPU oracle of the whole loop
and (ii) the per-problem BamParser/IntegratedCaller-shaped host path.  Needs a GPU: run with -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TAGS = {1: "FULL", 2: "PREF", 3: "PREF", 4: "REPT"}


def _models():
    md = json.load(open(os.path.join(os.path.dirname(GOLDEN), "..", "tredparse_b200", "data", "models.json")))
    step = {int(k): np.array(v) for k, v in md["step_size_by_period"].items()}
    for i in range(6, 18):
        step[i] = step[6]
    return step, md["stutter_weights"]


class _PE:
    def __init__(self, pr, ref, minpe):
        self.global_lens, self.target_lens = [int(x) for x in pr.global_lens], [int(x) for x in pr.target_lens]
        self.ref, self.MINPE = ref, minpe


def _oracle_call(pr, read_rows, step, w, maxinsert=300, fullsearch=False):
    """Likelihood oracle on the tallies implied by per-read (tag, h) rows."""
    from oracle import likelihood_oracle as lko
    counts = {"FULL": {}, "PREF": {}}
    rept = 0
    for tag, h in read_rows:
        name = TAGS.get(int(tag))
        if name == "REPT":
            rept += 1
        elif name:
            counts[name][int(h)] = counts[name].get(int(h), 0) + 1
    t = pr.tred
    ref = t.repeat_end - t.repeat_start + 1
    lk = lko.LikelihoodOracle(t, len(t.repeat), pr.readlen, counts, rept, pr.ploidy, pr.depth,
                              _PE(pr, ref, ref - 1 + 20), step, w, maxinsert=maxinsert, fullsearch=fullsearch)
    lk.call()
    return lk, counts, rept


def _expected_reads(pr):
    """Oracle classification of every read of a problem -> [(tagcode, h)]"""
    from oracle import evidence_oracle as evo
    t = pr.tred
    ev = evo.EvidenceOracle(t, pr.readlen, repeatpairs=True, engine="oracle")
    code = {"FULL": 1, "PREF": 2, "POST": 3, "REPT": 4, "HANG": 5}
    out = []
    for s in pr.read_strings():
        b = ev.classify_read(s)
        out.append((0, 0) if b is None else (code[b[2]], b[1]))
    return out


@pytest.fixture(scope="module")
def problems():
    from tredparse_b200 import simulate
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    ps = []
    # paper-style designs + expansions + haploid + other periods / N motifs (small coverage keeps the oracle quick)
    spec = [("HD", (15, 41)), ("HD", (17, 20)), ("HD", (20, 120)), ("HD", (40, 200)), ("DM1", (5, 62)),
            ("DM1", (12, 500)), ("FXS", (30,)), ("FXS", (250,)), ("FXS", (30, 250)), ("DM2", (15, 16)),
            ("SCA10", (13, 14)), ("SCA36", (5, 8)), ("ULD", (2, 3)), ("OPMD", (10, 11)), ("AR", (7, 22)),
            ("FRDA", (9, 70)), ("SCA8", (25, 90)), ("BPES", (14, 14))]
    for i, (name, alleles) in enumerate(spec):
        ps.append(simulate.simulate_problem(repo[name], alleles, readlen=150, cov_per_hap=15, seed=0xB200 + i))
    ps.append(simulate.simulate_problem(repo["DM1"], (13, 1000), readlen=250, cov_per_hap=15, seed=99))
    return ps


def test_cohort_pipeline_matches_oracle(problems):
    from tredparse_b200 import cohort
    step, w = _models()
    batch = cohort.CohortBatch(problems)
    out = batch.run_host(want_reads=True, want_hist=True, want_stats=True)
    calls, reads, hist, stats = out["calls"], out["reads"], out["hist"], out["stats"]
    assert stats[5] == 0
    r0 = 0
    for i, pr in enumerate(problems):
        rows = reads[r0:r0 + pr.nreads]
        r0 += pr.nreads
        exp = _expected_reads(pr)
        got = [(int(a), int(b)) if a not in (0,) else (0, 0) for a, b in rows[:, :2]]
        assert got == exp, (i, pr.tred.name)
        lk, counts, rept = _oracle_call(pr, [(a, b) for a, b in exp if a not in (0, 5)], step, w)
        c = cohort.decode_call(calls[i])
        # tallies
        for which, name in ((0, "FULL"), (1, "PREF")):
            assert {k: int(v) for k, v in enumerate(hist[i, which]) if v} == counts[name]
        assert c["RDP"] == rept and c["FDP"] == sum(counts["FULL"].values()) and c["PDP"] == sum(counts["PREF"].values())
        # call
        assert c["alleles"] == lk.alleles, (i, pr.tred.name, pr.alleles, c, lk.alleles)
        assert c["CI"] == lk.CI and c["label"] == lk.label
        assert abs(c["PP"] - lk.PP) <= 1e-9
        if lk.alleles[0] >= 0:
            assert abs(c["lik"] - lk.lik) <= 1e-9 * abs(lk.lik)
            assert c["n_points"] == len(lk.surface)
    assert stats[4] == sum(cohort.decode_call(c)["n_points"] for c in calls)


def test_cohort_pipeline_equals_per_problem_host_path(problems):
    """Same numbers whether problems go through the fused device pipeline or one by one through
    ssw.classify_reads + models.GridBatch (the BamParser / IntegratedCaller building blocks)."""
    from tredparse_b200 import cohort, ssw, models
    batch = cohort.CohortBatch(problems[:8])
    calls = batch.run_host()["calls"]
    for i, pr in enumerate(problems[:8]):
        t = pr.tred
        P = len(t.repeat)
        fam = ssw.make_family(t.prefix, t.repeat, t.suffix, -(-pr.readlen // P))
        out = ssw.classify_reads((pr.reads, pr.roff), np.zeros(pr.nreads, np.int32), fam)
        full, pref, rept = {}, {}, 0
        for tag, h in out[:, :2]:
            if tag == 1: full[int(h) * P] = full.get(int(h) * P, 0) + 1
            elif tag in (2, 3): pref[int(h) * P] = pref.get(int(h) * P, 0) + 1
            elif tag == 4: rept += 1
        has_pe = len(pr.global_lens) >= 100 and len(pr.target_lens) >= 5
        pdf = models.pe_kde([pr.global_lens])[0] if has_pe else None
        ref = t.repeat_end - t.repeat_start + 1
        gb = models.GridBatch()
        gi = gb.add(t, P, pr.readlen, dict(sorted(full.items())), dict(sorted(pref.items())), rept, pr.ploidy,
                    pr.depth, pdf, pr.target_lens, ref, ref - 1 + 20)
        gb.run()
        s = gb.summarize(gi)
        c = cohort.decode_call(calls[i])
        assert c["alleles"] == sorted(x // P for x in s["alleles"])
        assert c["CI"] == "{}-{}|{}-{}".format(*s["CIs"])
        assert c["lik"] == s["lik"] and abs(c["PP"] - s["PP"]) < 1e-12


def test_cohort_fullsearch_and_missing(problems):
    from tredparse_b200 import cohort, simulate
    step, w = _models()
    ps = [problems[0], problems[6]]
    empty = simulate.simulate_problem(problems[0].tred, (15, 41), cov_per_hap=15, seed=1)
    empty.reads = np.zeros(0, np.int8)
    empty.roff = np.zeros(1, np.int64)
    ps.append(empty)
    batch = cohort.CohortBatch(ps, maxinsert=60, fullsearch=True)
    out = batch.run_host(want_reads=True)
    calls = out["calls"]
    r0 = 0
    for i, pr in enumerate(ps[:2]):
        rows = out["reads"][r0:r0 + pr.nreads]
        r0 += pr.nreads
        lk, _, _ = _oracle_call(pr, [(a, b) for a, b in rows[:, :2] if a not in (0, 5)], step, w, maxinsert=60, fullsearch=True)
        c = cohort.decode_call(calls[i])
        assert c["alleles"] == lk.alleles and c["CI"] == lk.CI and abs(c["PP"] - lk.PP) < 1e-9
        assert c["n_points"] == len(lk.surface)
    c = cohort.decode_call(calls[2])
    assert c["alleles"] == [-1, -1] and c["PP"] == -1 and c["CI"] == "" and c["label"] == "missing"


def test_device_resident_path_equals_host_path(problems):
    import torch
    from tredparse_b200 import cohort, _lib
    batch = cohort.CohortBatch(problems)
    host = batch.run_host()["calls"]
    stream = torch.cuda.Stream()
    ctx = _lib.Context(0, stream=stream.cuda_stream)
    batch.to_device(0)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        batch.run_device(ctx)
        batch.run_device(ctx)          # idempotent
    stream.synchronize()
    dev = batch.calls_from_device()
    assert dev.tobytes() == host.tobytes()


def test_packed_transfer_formats_equal_plain_inputs(problems):
    """TREDSW_IN_READS_PACKED4 | TREDSW_IN_PE_LENS_I16: half the host->device bytes, byte-identical results
    (calls and per-read records), odd total base counts included."""
    from tredparse_b200 import cohort
    for ps in (problems, problems[:1], problems[3:8]):
        batch = cohort.CohortBatch(ps)
        a = batch.run_host(want_reads=True, want_hist=True)
        b = batch.run_host(want_reads=True, want_hist=True, packed=True)
        assert a["calls"].tobytes() == b["calls"].tobytes()
        assert np.array_equal(a["reads"], b["reads"]) and np.array_equal(a["hist"], b["hist"])
    # a read set whose base count is odd / not a multiple of 8
    import copy
    pr = copy.copy(problems[0])
    pr.reads = problems[0].reads[:-3].copy()
    pr.roff = problems[0].roff.copy()
    pr.roff[-1] -= 3
    batch = cohort.CohortBatch([pr, problems[1]])
    a = batch.run_host(want_reads=True)
    b = batch.run_host(want_reads=True, packed=True)
    assert a["calls"].tobytes() == b["calls"].tobytes() and np.array_equal(a["reads"], b["reads"])


def test_host_pipeline_keeps_order_and_results(problems):
    """cohort.HostPipeline: several host-buffer calls in flight on one GPU (one context / stream / thread per
    slot) return, in submission order, exactly what sequential calls return — plain and packed inputs."""
    from tredparse_b200 import cohort
    batches = [cohort.CohortBatch(problems[i::3]) for i in range(3)] * 2
    want = [b.run_host()["calls"].tobytes() for b in batches]
    with cohort.HostPipeline(0, depth=3) as pipe:
        got = [o["calls"].tobytes() for o in pipe.map(batches)]
        got_packed = [o["calls"].tobytes() for o in pipe.map(batches, packed=True)]
        assert pipe.launches > 0
    assert got == want and got_packed == want


def test_a_read_longer_than_max_read_len_is_an_error_not_lost_evidence(problems):
    """host buffers: tredsw_genotype_batch checks roff against max_read_len instead of skipping the read"""
    from tredparse_b200 import cohort, _lib
    batch = cohort.CohortBatch(problems[:2])
    batch.max_read_len = 100                       # the reads are 150 bp
    with pytest.raises(_lib.TredswError, match="max_read_len"):
        batch.run_host()
