"""
Cohort report over the caller's outputs — the consumer side of SURVEY.md §8(f) rank 3.

Mirrors ``tredparse/tredreport.py`` (reference file:line in the docstrings): the per-sample JSON (or VCF) files
written by ``tred.py`` are compiled into one TSV with a ``<TRED>.calls`` / ``<TRED>.label`` column pair per
locus (+ requested extra columns), and the TSV is summarised into

    <tsv>.report.txt    one row per locus: n_prerisk / n_risk / n_carrier and the allele-frequency spectrum
    <tsv>.cases.txt     the risk calls (PP > minPP) with their read evidence, per locus
    <tsv>.details.txt   read counts (FDP / PDP / RDP / PEDP) of every risk call

    python -m tredparse_b200.tredreport work/*.json --tsv work.tsv

Host-side table logic only (pandas); nothing here touches the GPU path.  The VCF reader parses the stanza
``tred.to_vcf`` writes (the reference uses PyVCF for the same fields, tredreport.py:146-161).
"""
import argparse
import gzip
import json
import math
import os.path as op
import sys
from collections import Counter
from multiprocessing import Pool, cpu_count

import pandas as pd

from . import __version__
from .meta import TREDsRepo
from .utils import DefaultHelpParser


def left_truncate_text(a, maxcol=30):
    """Long evidence strings keep their tail (tredreport.py:29-32)."""
    return [("..." + t[-(maxcol - 3):]) if isinstance(t, str) and len(t) > maxcol else t for t in list(a)]


def counts_to_af(counts):
    """{allele: count} -> '{a:n,b:m}' sorted by allele, NaN / '.' dropped (tredreport.py:107-109)."""
    keep = [(k, v) for k, v in counts.items() if not (k == "." or (isinstance(k, float) and math.isnan(k)))]
    return "{" + ",".join("{}:{}".format(k, v) for k, v in sorted(keep)) + "}"


# ---- per-file readers -------------------------------------------------------------------------------------------
def samplekey_of(path):
    """Sample key = file name up to the first dot (tredreport.py:185-187)."""
    return op.basename(path).split(".")[0]


def json_to_df_worker(jsonfile):
    """One JSON -> flat record: SampleKey + every key of tredCalls (tredreport.py:180-191)."""
    with open(jsonfile) as fp:
        js = json.load(fp)
    d = {"SampleKey": samplekey_of(jsonfile)}
    d.update(js["tredCalls"])
    return d


def vcf_to_df_worker(vcffile):
    """One VCF -> flat record with <TRED>.1 / .2 / .PP / .FR / .PR / .label (tredreport.py:143-161)."""
    d = {"SampleKey": samplekey_of(vcffile)}
    opener = gzip.open if vcffile.endswith(".gz") else open
    with opener(vcffile, "rt") as fp:
        for line in fp:
            if line.startswith("#") or not line.strip():
                continue
            atoms = line.rstrip("\n").split("\t")
            tr = atoms[2]
            sample = dict(zip(atoms[8].split(":"), atoms[9].split(":")))
            a, b = sample["GB"].split("/")
            d[tr + ".1"], d[tr + ".2"] = int(a), int(b)
            d[tr + ".PP"] = float(sample["PP"])
            d[tr + ".FR"], d[tr + ".PR"] = sample["FR"], sample["PR"]
            d[tr + ".label"] = sample["LABEL"]
    return d


def _collect(worker, files, cpus):
    cpus = max(1, min(int(cpus), len(files)))
    if cpus == 1:
        rows = [worker(f) for f in files]
    else:
        with Pool(processes=cpus) as p:
            rows = p.map(worker, files)
    return pd.DataFrame(rows)


def json_to_df(jsonfiles, tsvfile=None, cpus=1):
    """Compile JSON files into one frame, one row per sample (tredreport.py:194-208)."""
    return _collect(json_to_df_worker, list(jsonfiles), cpus)


def vcf_to_df(vcffiles, tsvfile=None, cpus=1):
    """Compile VCF files into one frame (tredreport.py:164-177)."""
    return _collect(vcf_to_df_worker, list(vcffiles), cpus)


# ---- frame -> TSV -----------------------------------------------------------------------------------------------
def df_to_tsv(df, tsvfile, extra_columns=(), jsonformat=True, ref="hg38"):
    """Add <TRED>.1_ / .2_ / .calls, write SampleKey (+ inferredGender) and every ``.calls`` / ``.label`` /
    extra column sorted by name (tredreport.py:112-140).  Missing values become -1; the second allele of an
    X-linked locus reads '.' for males."""
    df = df.fillna(-1)
    dd = ["SampleKey"]
    if jsonformat:
        dd += ["inferredGender"]
        if "inferredGender" not in df.columns:
            df["inferredGender"] = "Unknown"
    repo = TREDsRepo(ref)
    new = {}
    for tred in repo.names:
        tr = repo[tred]
        if tred + ".1" not in df.columns:
            continue
        a = df[tred + ".1"].astype("int")
        b = df[tred + ".2"].astype("int").astype(object)
        if jsonformat and tr.is_xlinked:
            b = b.where(df["inferredGender"] != "Male", ".")
        new[tred + ".1_"], new[tred + ".2_"] = a, b
        new[tred + ".calls"] = ["{}|{}".format(x, y) for x, y in zip(a, b)]
    df = pd.concat([df, pd.DataFrame(new, index=df.index)], axis=1)
    wanted = ["calls", "label"] + list(extra_columns)
    columns = dd + sorted(x for x in df.columns if x not in dd and any(x.endswith("." + z) for z in wanted))
    tf = df.reindex(columns=columns)
    tf.to_csv(tsvfile, sep="\t", index=False)
    print("TSV output written to `{}` (# samples={})".format(tsvfile, tf.shape[0]), file=sys.stderr)
    return df


# ---- per-locus summary --------------------------------------------------------------------------------------------
def get_tred_summary(df, tred, repo, minPP=.5, casesfw=None, detailsfw=None):
    """Counts of prerisk / risk (PP > minPP) / carrier samples of one locus, the case listing and the allele
    spectrum (tredreport.py:35-104).  carrier = not called 'risk' yet the longer allele is in the risk range."""
    pf2 = tred + ".2"
    tr = repo[tred]
    label, pp = tred + ".label", tred + ".PP"
    cutoff_risk = tr.cutoff_risk
    prerisk = df[df[label] == "prerisk"]
    risk = df[(df[label] == "risk") & (df[pp] > minPP)].copy()
    if tr.is_expansion:
        carrier = df[(df[label] != "risk") & (df[pf2] >= cutoff_risk)]
    else:
        carrier = df[(df[label] != "risk") & (df[pf2] <= cutoff_risk) & (df[pf2] > 0)]
    n_prerisk, n_risk, n_carrier = prerisk.shape[0], risk.shape[0], carrier.shape[0]
    calls = tred + ".calls"
    core = ["SampleKey", "inferredGender", calls]
    columns = core + [tred + ".FR", tred + ".PR", tred + ".RR", pp]
    for k in (".FR", ".PR", ".RR"):
        if tred + k in risk.columns:
            risk[tred + k] = left_truncate_text(risk[tred + k])
    if detailsfw is not None and tred != "AR":             # (the reference leaves AR out of the details, :78-79)
        dcols = core + [tred + ".FDP", tred + ".PDP", tred + ".RDP", tred + ".PEDP"]
        if all(c in risk.columns for c in dcols):
            for _, row in risk[dcols].iterrows():
                samplekey, sex, c, fdp, pdp, rdp, pedp = row
                atoms = [tred, tr.inheritance, samplekey, sex, c, int(fdp), int(pdp), int(rdp), int(pedp)]
                print("\t".join(str(x) for x in atoms), file=detailsfw)
    if n_risk and casesfw is not None:
        print("[{}] - {}".format(tred, tr.row["title"]), file=casesfw)
        print("rep={}".format(tr.repeat), "inherit={}".format(tr.inheritance), "cutoff={}".format(cutoff_risk),
              "n_risk={}".format(n_risk), "n_carrier={}".format(n_carrier), "loc={}".format(tr.row["repeat_location"]),
              file=casesfw)
        print(risk.reindex(columns=columns).to_string(index=False), file=casesfw)
        print(file=casesfw)
    cnt = Counter()
    cnt.update(df[tred + ".1_"])
    cnt.update(df[tred + ".2_"])
    cnt.pop(-1, None)
    return tr, n_prerisk, n_risk, n_carrier, counts_to_af(cnt)


def summarize(df, tsvfile, ref="hg38", minPP=.5):
    """Write <tsv>.report.txt / .cases.txt / .details.txt (tredreport.py:262-316); -> (summary frame, totals)."""
    repo = TREDsRepo(ref)
    rows = []
    total_prerisk = total_risk = total_carrier = total_loci = 0
    header = "Locus,Inheritance,SampleKey,Sex,Calls,FullReads,PartialReads,RepeatReads,PairedReads"
    with open(tsvfile + ".cases.txt", "w") as casesfw, open(tsvfile + ".details.txt", "w") as detailsfw:
        print("\t".join(header.split(",")), file=detailsfw)
        for tred in repo.names:
            if tred + ".label" not in df.columns or tred + ".1_" not in df.columns:
                continue
            tr, n_prerisk, n_risk, n_carrier, af = get_tred_summary(df, tred, repo, minPP=minPP, casesfw=casesfw,
                                                                    detailsfw=detailsfw)
            total_prerisk += n_prerisk
            total_risk += n_risk
            total_carrier += n_carrier
            total_loci += 1 if n_risk else 0
            rows.append({"abbreviation": tred, "title": tr.row["title"], "motif": tr.row["motif"], "inheritance": tr.inheritance,
                         "cutoff_prerisk": tr.cutoff_prerisk, "cutoff_risk": tr.cutoff_risk, "n_prerisk": n_prerisk,
                         "n_risk": n_risk, "n_carrier": n_carrier, "allele_freq": af})
    summary = pd.DataFrame(rows)
    summary.to_csv(tsvfile + ".report.txt", sep="\t", index=False, float_format="%d")
    totals = {"n_prerisk": total_prerisk, "n_risk": total_risk, "n_carrier": total_carrier,
              "n_affected_loci": total_loci}
    print("Summary report written to `{}` (# loci={})".format(tsvfile + ".report.txt", summary.shape[0]), file=sys.stderr)
    print("Summary: n_prerisk={n_prerisk}, n_risk={n_risk}, n_carrier={n_carrier}, "
          "n_affected_loci={n_affected_loci}".format(**totals), file=sys.stderr)
    return summary, totals


def main(args=None):
    p = DefaultHelpParser(description=__doc__, prog=op.basename(__file__),
                          formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("files", nargs="*")
    p.add_argument("--ref", help="Reference genome version",
                   choices=("hg38", "hg38_nochr", "hg19", "hg19_nochr"), default="hg38")
    p.add_argument("--tsv", default="out.tsv", help="Path to the tsv file")
    p.add_argument("--columns", help="Columns to extract, use comma to separate")
    p.add_argument("--minPP", default=.5, type=float, help="Minimum Prob(pathological) to report cases")
    p.add_argument("--cpus", default=cpu_count(), type=int, help="Number of threads")
    p.add_argument("--version", action="version", version="%(prog)s " + __version__)
    a = p.parse_args(args)
    columns = a.columns.split(",") if a.columns else []
    if a.files:
        jsonformat = a.files[0].endswith(".json")
        print("Using {} cpus to parse {} {} files".format(min(len(a.files), a.cpus), len(a.files),
                                                          "JSON" if jsonformat else "VCF"), file=sys.stderr)
        df = (json_to_df if jsonformat else vcf_to_df)(a.files, a.tsv, a.cpus)
        df = df_to_tsv(df, a.tsv, extra_columns=columns, jsonformat=jsonformat, ref=a.ref)
    elif op.exists(a.tsv):
        df = pd.read_csv(a.tsv, sep="\t")
    else:
        p.print_help()
        return 1
    if df.empty:
        print("Dataframe empty - check input files", file=sys.stderr)
        return 1
    if not a.files:
        # a TSV read back has only .calls: recover the integer allele columns the summary counts
        for c in [c for c in df.columns if c.endswith(".calls")]:
            t = c[:-len(".calls")]
            parts = df[c].astype(str).str.split("|", expand=True)
            df[t + ".1_"] = parts[0].astype(int)
            df[t + ".2_"] = [int(x) if x != "." else "." for x in parts[1]]
            df[t + ".2"] = [int(x) if x != "." else -1 for x in parts[1]]
            if t + ".PP" not in df.columns:
                df[t + ".PP"] = 1.0
    summarize(df, a.tsv, ref=a.ref, minPP=a.minPP)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
