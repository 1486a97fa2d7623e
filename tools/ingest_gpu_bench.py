#!/usr/bin/env python
"""GPU BAM ingest (csrc/bgzf_gpu.cu) on the synthetic whole-sample BAMs of bench.py: time per batch, bytes, and a
problem-by-problem comparison with the host reader.

    python tools/ingest_gpu_bench.py [--samples 8] [--reps 5] [--no-crc]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-crc", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import bench
    from tredparse_b200 import ingest, _lib
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    names = bench.distinct_loci(repo)
    bams = bench.make_bams(a.samples, max(1, min(16, os.cpu_count() or 1)))
    ctx = _lib.default_context(0)
    t = time.perf_counter()
    hs = [ingest.BamIngest(b) for b in bams]
    t_open = time.perf_counter() - t
    qs, so, keep, key = [], [], [], []
    t = time.perf_counter()
    for si, h in enumerate(hs):
        for n in names:
            q = ingest.locus_query(h, repo[n], 150, alts=repo[n].alt)
            qs.append(q[0]); keep.append(q[1]); so.append(si); key.append((si, n))
    t_q = time.perf_counter() - t
    times, stats = [], None
    for _ in range(a.reps):
        t = time.perf_counter()
        b = ingest.IngestBatch(ctx, hs, so, qs, keep=keep, check_crc=not a.no_crc)
        times.append(time.perf_counter() - t)
        stats = b.stats()
        last = b
        if _ < a.reps - 1:
            b.close()
    out = {"samples": a.samples, "problems": len(key), "open_s": t_open, "queries_s": t_q, "batch_s": times,
           "stats": stats, "status_bad": int((last.status != 0).sum()), "reads": int(last.nreads)}
    if not a.no_check:
        bad = 0
        t = time.perf_counter()
        for i, (si, n) in enumerate(key):
            ref = hs[si].extract_locus(repo[n], 150, alts=repo[n].alt, want_names=True)
            ev = last.evidence(i)
            ok = (np.array_equal(ev.reads, ref.reads) and np.array_equal(ev.roff, ref.roff)
                  and np.array_equal(ev.global_lens, ref.global_lens) and np.array_equal(ev.target_lens, ref.target_lens)
                  and ev.depth == ref.depth and ev.n_unmapped == ref.n_unmapped and ev.names == ref.names)
            bad += not ok
        out["host_reader_s_one_thread"] = time.perf_counter() - t
        out["problems_differing_from_host_reader"] = bad
    best = min(times)
    out["loci_per_s_ingest_only"] = len(key) / best
    out["inflated_GBps"] = stats["inflated_bytes"] / best / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
