#!/usr/bin/env python
"""
Derive the product's compact data tables from the reference's data directory (run in the build
container, where /root/reference is mounted; the outputs are committed).

    python tools/make_data.py [/root/reference]

Writes into tredparse_b200/data/:
  loci.tsv     one row per TRED with only the columns the hot path, JSON/VCF writers and the cohort
               simulator consume (source: tredparse/data/TREDs.meta.csv, read by meta.py:36-42)
  alts.tsv     alternative (mis-mapping) regions per TRED, hg38 and hg19 (TREDs.alts.csv, meta.py:80-94)
  chrY.tsv     the first 20 unique chrY regions per build (chrY.<build>.unique_ccn.gc; gender inference,
               bam_parser.py:413-429)
  models.json  the 37-bin step-size PMF per period and the 5 logistic stutter weights
               (illumina_v3.pcrfree.stepmodel / .stuttermodel, parsed as models.py:46-77 does)
"""
import json
import os
import sys

import pandas as pd

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = os.path.join(ref, "tredparse", "data")
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tredparse_b200", "data")
os.makedirs(dst, exist_ok=True)

df = pd.read_csv(os.path.join(src, "TREDs.meta.csv"), index_col=0)
cols = ["title", "gene_name", "id", "motif", "repeat", "repeat_location", "repeat_location.hg19",
        "prefix", "suffix", "inheritance", "mutation_nature", "cutoff_prerisk", "cutoff_risk",
        "allele_freq"]
out = df[cols].copy()
out.index.name = "name"
out.to_csv(os.path.join(dst, "loci.tsv"), sep="\t")

alts = pd.read_csv(os.path.join(src, "TREDs.alts.csv"), index_col=0)
alts.index.name = "name"
alts.to_csv(os.path.join(dst, "alts.tsv"), sep="\t")

MAX_PERIOD = 6
with open(os.path.join(src, "illumina_v3.pcrfree.stepmodel")) as fp:
    lines = fp.read().split("\n")
non_unit = [float(lines[i].strip()) for i in range(MAX_PERIOD)]
prob_increase = float(lines[MAX_PERIOD].split("=")[1])
step = {str(i + 1): [float(x) for x in lines[MAX_PERIOD + 1 + i].split()[1:]] for i in range(MAX_PERIOD)}
with open(os.path.join(src, "illumina_v3.pcrfree.stuttermodel")) as fp:
    rows = fp.read().split("\n")[MAX_PERIOD:]
weights = [float(r) for r in (x.strip() for x in rows) if r]
with open(os.path.join(dst, "models.json"), "w") as fw:
    json.dump({"model": "illumina_v3.pcrfree", "non_unit_step_by_period": non_unit,
               "prob_increase": prob_increase, "step_size_by_period": step,
               "stutter_weights": weights}, fw, indent=1)
# unique chrY regions used for the gender inference (bam_parser.py:413-429 reads the first rows of
# chrY.<build>.unique_ccn.gc, skipping ten listed row numbers, until it has five): the first 20 rows per build
with open(os.path.join(dst, "chrY.tsv"), "w") as fw:
    fw.write("build\trow\tchr\tstart\tend\tgc\n")
    for build in ("hg38", "hg19"):
        with open(os.path.join(src, "chrY.{}.unique_ccn.gc".format(build))) as fp:
            for i, row in enumerate(fp):
                if i >= 20:
                    break
                c, start, end, gc = row.split()
                fw.write("\t".join([build, str(i), c, start, end, gc]) + "\n")
print("wrote", os.listdir(dst))
