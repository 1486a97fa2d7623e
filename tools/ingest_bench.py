#!/usr/bin/env python
"""Time the native one-pass BAM ingest (csrc/ingest.cpp) beside the Python reader that makes the reference's
three passes per locus (select_reads + PEextractor + region_depth).  Host code only.

    python tools/ingest_bench.py [bam] [tred] [--reps 5]
"""
import argparse
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bam", nargs="?", default=os.path.join(ROOT, "tests", "golden", "t001.mini.bam"))
    ap.add_argument("tred", nargs="?", default="HD")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from tredparse_b200 import bamio, ingest
    from tredparse_b200.bam_parser import BamParser, PEextractor, BamDepth
    from tredparse_b200.meta import TREDsRepo
    from tredparse_b200.utils import InputParams
    repo = TREDsRepo()
    t = repo[a.tred]

    def python_path():
        ip = InputParams(bam=a.bam, READLEN=150, tredName=a.tred, repo=repo, maxinsert=300, fullsearch=False,
                         gender="Unknown", depth=30, clip=False, alts=True, repeatpairs=True, log="INFO")
        bp = BamParser(ip)
        sam = bamio.AlignmentFile(a.bam)
        reads = bp.select_reads(sam)
        sam.close()
        pe = PEextractor(bp)
        d = BamDepth(a.bam, "hg38", logging.getLogger()).region_depth(t.chr, max(0, t.repeat_start - 1000), t.repeat_end + 1000)
        return len(reads), len(pe.global_lens), d

    def native_path():
        with ingest.BamIngest(a.bam) as ing:
            ev = ing.extract_locus(t, 150, alts=t.alt, want_names=True)
        return ev.nreads, len(ev.global_lens), ev.depth

    for name, fn in (("python (3 passes)", python_path), ("native (1 pass)", native_path)):
        fn()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            r = fn()
        dt = (time.perf_counter() - t0) / a.reps
        print("{:18s} {:8.2f} ms / locus  ({:.1f} loci/s/core)  reads={} pairs={} depth={:.2f}".format(name, dt * 1e3, 1 / dt, *r))


if __name__ == "__main__":
    main()
