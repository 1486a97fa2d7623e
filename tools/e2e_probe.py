#!/usr/bin/env python
"""Where does the end-to-end path lose time against the device-resident one?  Times the same cohort batch
through variants of the host pipeline (wall clock over `--steps` calls, after warm-up):

  dev/D      D host threads, device-resident inputs (no copies), one context each, synchronize per call
  host/D     D host threads, pinned host buffers through tredsw_genotype_batch (the bench's e2e leg)
  host+/D    the same with the small per-call tables and the output in pinned memory too

    python tools/e2e_probe.py [--samples 384] [--steps 24]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=384)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--only", default="", help="comma-separated variant labels to run (default: all)")
    ap.add_argument("--timeline", type=int, default=0, help="print the device-side timeline of this many calls in flight")
    a = ap.parse_args()
    import torch
    from tredparse_b200 import _lib, cohort, simulate
    from tredparse_b200.meta import TREDsRepo
    sys.path.insert(0, ROOT)
    import bench
    repo = TREDsRepo()
    names = bench.distinct_loci(repo)
    problems = simulate.simulate_cohort(repo, names, a.samples, readlen=150, seed=20240000)
    batch = cohort.CohortBatch(problems, maxinsert=300, fullsearch=False)
    batch.to_device(0)
    batch.pack_inputs()

    def pin(x):
        t = torch.from_numpy(x.view(np.uint8) if x.dtype.fields else x).pin_memory()
        return t.numpy().view(x.dtype) if x.dtype.fields else t.numpy()
    for name in ("roff", "read_problem", "problems"):
        setattr(batch, name, pin(getattr(batch, name)))
    batch._packed = {k: pin(v) for k, v in batch._packed.items()}
    out = {}

    only = set(x for x in a.only.split(",") if x)

    def timed(label, depth, fn):
        if only and label not in only:
            return
        ctxs = [_lib.Context(0) for _ in range(depth)]
        import queue
        free = queue.Queue()
        for c in ctxs:
            free.put(c)

        def run(_):
            c = free.get()
            try:
                fn(c)
            finally:
                free.put(c)
        with ThreadPoolExecutor(max_workers=depth) as pool:
            list(pool.map(run, range(2 * depth)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            list(pool.map(run, range(a.steps)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        for c in ctxs:
            c.close()
        out[label] = {"ms_per_step": 1e3 * dt / a.steps, "loci_per_s": batch.nproblems * a.steps / dt}
        sys.stderr.write("{} {}\n".format(label, out[label]))

    def dev_call(c):
        batch.run_device(c)
        c.synchronize()

    def host_call(c):
        batch.run_host(ctx=c, packed=True)

    if a.timeline:
        # device-side timeline of `depth` host-buffer calls in flight: one row per call, ms since a common reference
        import queue
        depth = a.timeline
        ctxs = [_lib.Context(0) for _ in range(depth)]
        for c in ctxs:
            c.enable_timing(True)
        free = queue.Queue()
        for i, c in enumerate(ctxs):
            free.put((i, c))
        rows = []

        def run(_):
            i, c = free.get()
            try:
                t0 = time.perf_counter()
                batch.run_host(ctx=c, packed=True)
                t1 = time.perf_counter()
                tl = c.timeline()
                rows.append(dict(ctx=i, host0=t0, host1=t1, **tl))
            finally:
                free.put((i, c))
        with ThreadPoolExecutor(max_workers=depth) as pool:
            list(pool.map(run, range(2 * depth)))
            del rows[:]
            list(pool.map(run, range(4 * depth)))
        rows.sort(key=lambda r: r["start"])
        base, hbase = rows[0]["start"], min(r["host0"] for r in rows)
        keys = ("start", "inputs", "sw0", "sw1", "kde0", "kde1", "grid0", "grid1", "final", "copied")
        sys.stderr.write("ctx  host_call_ms(begin,end)   " + " ".join("{:>8}".format(k) for k in keys) + "\n")
        for r in rows:
            sys.stderr.write("{:3d}  {:8.2f} {:8.2f}     ".format(r["ctx"], 1e3 * (r["host0"] - hbase), 1e3 * (r["host1"] - hbase)) +
                             " ".join("{:8.2f}".format(r[k] - base) for k in keys) + "\n")
        print(json.dumps({"timeline_depth": depth, "rows": [{k: (r[k] - base) for k in keys} | {"ctx": r["ctx"]} for r in rows]}))
        return
    for d in (1, 2, 4, 8):
        timed("dev/{}".format(d), d, dev_call)
    for d in (1, 2, 4, 8):
        timed("host/{}".format(d), d, host_call)
    for name in ("families", "loci", "step_pmf"):
        setattr(batch, name, pin(getattr(batch, name)))
    for d in (4, 8):
        timed("host+/{}".format(d), d, host_call)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
