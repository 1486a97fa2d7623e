#!/usr/bin/env python
"""Does a large pinned H2D copy overlap the persistent Smith-Waterman kernel?  Runs device-resident steps on
one stream while a second host thread streams H2D copies of the step's input size on another stream, and
reports the step time and the copy time alone and together.

    python tools/copy_overlap_probe.py [--samples 384] [--steps 16] [--chunk-mb 0]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=384)
    ap.add_argument("--steps", type=int, default=16)
    a = ap.parse_args()
    import torch
    from tredparse_b200 import _lib, cohort, simulate
    from tredparse_b200.meta import TREDsRepo
    import bench
    repo = TREDsRepo()
    problems = simulate.simulate_cohort(repo, bench.distinct_loci(repo), a.samples, readlen=150, seed=20240000)
    batch = cohort.CohortBatch(problems, maxinsert=300, fullsearch=False)
    batch.to_device(0)
    s1 = torch.cuda.Stream()
    ctx = _lib.Context(0, stream=s1.cuda_stream)
    for _ in range(3):
        batch.run_device(ctx)
    ctx.synchronize()
    nbytes = 146 << 20
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    out = {}

    def steps():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            batch.run_device(ctx)
        ctx.synchronize()
        return 1e3 * (time.perf_counter() - t0) / a.steps

    def copies(n, chunk):
        t0 = time.perf_counter()
        with torch.cuda.stream(s2):
            for _ in range(n):
                if chunk:
                    for o in range(0, nbytes, chunk):
                        dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
                else:
                    dst.copy_(src, non_blocking=True)
                s2.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n

    out["step_alone_ms"] = steps()
    out["copy_alone_ms"] = copies(a.steps, 0)
    for chunk in (0, 8 << 20, 1 << 20):
        res = {}
        stop = threading.Event()
        cnt = [0, 0.0]

        def pump():
            while not stop.is_set():
                cnt[1] += copies(1, chunk)
                cnt[0] += 1
        th = threading.Thread(target=pump)
        th.start()
        res["step_ms"] = steps()
        stop.set()
        th.join()
        res["copy_ms"] = cnt[1] / max(cnt[0], 1)
        res["copies"] = cnt[0]
        out["together_chunk_{}MB".format(chunk >> 20)] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
