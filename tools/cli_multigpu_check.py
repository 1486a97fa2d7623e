#!/usr/bin/env python
"""The CLI on a small cohort of whole-sample BAMs with --gpus 1 and --gpus N: the per-sample JSON files must be
identical (the multi-GPU run deals (sample, locus) problems to one process per GPU and merges the parts).

    python tools/cli_multigpu_check.py [--gpus 2] [--samples 6]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--samples", type=int, default=6)
    a = ap.parse_args()
    import bench
    from tredparse_b200 import tred as T
    bams = bench.make_bams(a.samples, max(1, min(16, os.cpu_count() or 1)))
    tmp = tempfile.mkdtemp(prefix="tredsw_cli_")
    csv = os.path.join(tmp, "samples.csv")
    with open(csv, "w") as fp:
        fp.write("#SampleKey,BAM\n" + "".join("s{:04d},{}\n".format(i, b) for i, b in enumerate(bams)))
    cwd, out = os.getcwd(), {}
    for n in (1, a.gpus):
        work = os.path.join(tmp, "work{}".format(n))
        t = time.perf_counter()
        try:
            T.main([csv, "--workdir", work, "--gpus", str(n)])
        finally:
            os.chdir(cwd)
        out[n] = (time.perf_counter() - t, {f: json.load(open(os.path.join(work, f))) for f in sorted(os.listdir(work)) if f.endswith(".json")})
    one, many = out[1][1], out[a.gpus][1]
    assert sorted(one) == sorted(many) and len(one) == a.samples, (sorted(one), sorted(many))
    bad = [f for f in one if one[f] != many[f]]
    nkeys = sum(len(v["tredCalls"]) for v in one.values())
    print(json.dumps({"samples": a.samples, "gpus": a.gpus, "files": len(one), "tredCalls_keys": nkeys,
                      "differing_files": bad, "seconds_1gpu": out[1][0], "seconds_ngpu": out[a.gpus][0]}))
    assert not bad


if __name__ == "__main__":
    main()
