#!/usr/bin/env python
"""What tools/sanitize.sh runs under compute-sanitizer: every kernel of the hot path once, at sizes the
sanitizer finishes in minutes — smoke() (pairs kernel, fused pipeline vs the oracle), one small cohort batch
through the host-buffer and the device-resident paths (persistent single-warp Smith-Waterman CTAs with their
per-CTA global scratch, atomic work queues and arena cursors), under --useclippedreads / --norepeatpairs, a
--fullsearch long-expansion problem (row-structured grid with its cluster / DSMEM reductions), and the GPU BAM ingest
on the two fixtures (warp-per-block inflate, windowed record walk, pairing tables with atomics)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import __graft_entry__ as ge
    from tredparse_b200 import _lib, cohort, simulate
    from tredparse_b200.meta import TREDsRepo
    which = sys.argv[1:] or ["smoke", "cohort", "flags", "stress", "ingest"]
    repo = TREDsRepo()
    ctx = _lib.default_context(0)
    if "smoke" in which:
        ge.smoke()
    names = list(repo.names)[:30]
    if "cohort" in which:
        probs = [simulate.simulate_problem(repo[n], (12 + i % 7, 30 + 3 * i) if repo[n].ploidy == 2 else (20 + i,), seed=7 + i)
                 for i, n in enumerate(names)]
        b = cohort.CohortBatch(probs)
        a = b.run_host(ctx=ctx, want_reads=True, want_hist=True, want_post=True)
        b.to_device(0)
        import torch
        torch.cuda.synchronize()
        b.run_device(ctx)
        torch.cuda.synchronize()                 # (ctx owns its stream)
        got = b.calls_from_device()
        assert np.asarray(a["calls"]).tobytes() == got.tobytes(), "device-resident != host-buffer calls"
    if "flags" in which:
        probs = [simulate.simulate_problem(repo[n], (15, 60), seed=70 + i) for i, n in enumerate(["HD", "DM1", "SCA1"])]
        cohort.CohortBatch(probs, clip=True).run_host(ctx=ctx)
        cohort.CohortBatch(probs, repeatpairs=False).run_host(ctx=ctx)
    if "stress" in which:
        probs = [simulate.simulate_problem(repo["DM1"], (13, 400), readlen=150, seed=100),
                 simulate.simulate_problem(repo["FXS"], (30, 300), readlen=150, seed=101)]
        cohort.CohortBatch(probs, maxinsert=500, fullsearch=True).run_host(ctx=ctx, want_post=True)
    if "ingest" in which:
        # GPU BAM ingest: inflate (warp per BGZF block), record walk, selection, pairing tables, scans, scatter
        from tredparse_b200 import ingest
        gold = os.path.join(ROOT, "tests", "golden")
        hs = [ingest.BamIngest(os.path.join(gold, "t001.mini.bam")), ingest.BamIngest(os.path.join(gold, "t002.mini.bam"))]
        qs, so, keep = [], [], []
        for si, h in enumerate(hs):
            for n in ("HD", "DM1", "SCA17", "FXS"):
                q = ingest.locus_query(h, repo[n], 150, alts=repo[n].alt)
                if q is not None:
                    qs.append(q[0]); keep.append(q[1]); so.append(si)
        with ingest.IngestBatch(ctx, hs, so, qs, keep=keep) as b:
            assert not b.status.any() and b.nreads > 50
            ref = hs[0].extract_locus(repo["HD"], 150, alts=repo["HD"].alt, want_names=True)
            ev = b.evidence(0)
            assert np.array_equal(ev.reads, ref.reads) and ev.names == ref.names and np.array_equal(ev.global_lens, ref.global_lens)
    print("sanitize_target: ok", which)


if __name__ == "__main__":
    main()
