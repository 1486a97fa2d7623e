#!/usr/bin/env python
"""Where the from-BAM path (tred.run_chunk on whole-sample BAMs) spends its time: wall clock of a few chunks and a
cProfile of the last one.   python tools/from_bam_probe.py [--samples 8] [--reps 3] [--host-ingest]"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--host-ingest", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--chunk", type=int, default=8)
    a = ap.parse_args()
    if a.host_ingest:
        os.environ["TREDSW_GPU_INGEST"] = "0"
    import bench
    from tredparse_b200 import tred as T
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    names = bench.distinct_loci(repo)
    bams = bench.make_bams(a.samples, max(1, min(16, os.cpu_count() or 1)))
    tasks = [("s{:04d}".format(i), p, repo, list(names), 300, False, False, True, True, "INFO") for i, p in enumerate(bams)]
    T.run_chunk(tasks[:1])
    times = []
    for _ in range(a.reps):
        t = time.perf_counter()
        list(T.run_chunks(tasks, chunk=a.chunk))
        times.append(time.perf_counter() - t)
    print(json.dumps({"samples": a.samples, "loci": len(names), "seconds": times,
                      "loci_per_s": a.samples * len(names) / min(times), "gpu_ingest": T.GPU_INGEST}))
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        list(T.run_chunks(tasks, chunk=a.chunk))
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
        print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
