"""tredreport: JSON / VCF outputs of the caller -> cohort TSV -> per-locus summary (tredparse/tredreport.py;
SURVEY.md §8f rank 3).  Host-side table logic, no GPU."""
import json
import os

import pandas as pd
import pytest

from tredparse_b200 import tredreport, tred as tredmod
from tredparse_b200.meta import TREDsRepo


def _calls(a, b, label, pp, gender="Female", tred="HD", **kw):
    d = {"inferredGender": gender, "depthY": 0.1, "readLen": 150,
         tred + ".1": a, tred + ".2": b, tred + ".label": label, tred + ".PP": pp, tred + ".FR": "15|4",
         tred + ".PR": ";".join("{}|1".format(i) for i in range(5, 40)), tred + ".RR": "", tred + ".FDP": 4,
         tred + ".PDP": 35, tred + ".RDP": 0, tred + ".PEDP": 11, tred + ".DP": 30, tred + ".CI": "15-15|40-42"}
    d.update(kw)
    return d


@pytest.fixture()
def cohort_dir(tmp_path):
    samples = {
        "s1": _calls(15, 41, "risk", 0.999),
        "s2": _calls(17, 20, "ok", 0.0),
        "s3": _calls(19, 37, "prerisk", 0.1),
        "s4": _calls(17, 44, "risk", 0.3),                       # risk label but PP below minPP: not a case
        "s5": {**_calls(17, 18, "ok", 0.0, gender="Male"), **_calls(30, 30, "ok", 0.0, gender="Male", tred="FXS")},
        "s6": _calls(-1, -1, "missing", -1),
    }
    files = []
    for k, calls in samples.items():
        f = tmp_path / (k + ".json")
        f.write_text(json.dumps({"samplekey": "ignored", "bam": k + ".bam", "tredCalls": calls}, sort_keys=True, indent=4))
        files.append(str(f))
    return tmp_path, files


def test_json_to_tsv_and_summary(cohort_dir):
    tmp, files = cohort_dir
    tsv = str(tmp / "out.tsv")
    df = tredreport.json_to_df(files, tsv, cpus=1)
    assert list(df["SampleKey"]) == ["s1", "s2", "s3", "s4", "s5", "s6"]          # from the file names, in order
    df = tredreport.df_to_tsv(df, tsv, extra_columns=["PP"], jsonformat=True)
    t = pd.read_csv(tsv, sep="\t")
    assert list(t.columns) == ["SampleKey", "inferredGender", "FXS.PP", "FXS.calls", "FXS.label", "HD.PP", "HD.calls", "HD.label"]
    assert list(t["HD.calls"]) == ["15|41", "17|20", "19|37", "17|44", "17|18", "-1|-1"]
    assert t.loc[4, "FXS.calls"] == "30|." and t.loc[0, "FXS.calls"] == "-1|-1"    # X-linked male: second allele '.'
    summary, totals = tredreport.summarize(df, tsv, minPP=.5)
    hd = summary.set_index("abbreviation").loc["HD"]
    assert (hd["n_prerisk"], hd["n_risk"], hd["n_carrier"]) == (1, 1, 0)           # s3 | s1 | s4: labelled risk, PP too low
    assert hd["allele_freq"] == "{15:1,17:3,18:1,19:1,20:1,37:1,41:1,44:1}"
    assert totals == {"n_prerisk": 1, "n_risk": 1, "n_carrier": 0, "n_affected_loci": 1}
    cases = open(tsv + ".cases.txt").read()
    assert "[HD] - " in cases and "s1" in cases and "s4" not in cases and "cutoff=40" in cases
    assert "..." in cases                                                           # long PR strings keep their tail
    details = open(tsv + ".details.txt").read().splitlines()
    assert details[0].split("\t")[:5] == ["Locus", "Inheritance", "SampleKey", "Sex", "Calls"]
    assert details[1].split("\t") == ["HD", "AD", "s1", "Female", "15|41", "4", "35", "0", "11"]
    rep = pd.read_csv(tsv + ".report.txt", sep="\t")
    assert {"abbreviation", "title", "motif", "inheritance", "cutoff_prerisk", "cutoff_risk", "n_prerisk", "n_risk",
            "n_carrier", "allele_freq"} <= set(rep.columns)


def test_carrier_definition():
    """carrier: label != risk but the longer allele is in the risk range (tredreport.py:51-55) — a recessive locus
    with one expanded allele (FRDA cutoff 66, AR inheritance)."""
    repo = TREDsRepo()
    rows = [{"SampleKey": "a", "inferredGender": "Female", **_calls(9, 70, "ok", 0.0, tred="FRDA")},
            {"SampleKey": "b", "inferredGender": "Female", **_calls(70, 90, "risk", 1.0, tred="FRDA")}]
    df = pd.DataFrame(rows)
    df["FRDA.1_"], df["FRDA.2_"] = df["FRDA.1"], df["FRDA.2"]
    df["FRDA.calls"] = ["9|70", "70|90"]
    tr, n_pre, n_risk, n_car, af = tredreport.get_tred_summary(df, "FRDA", repo)
    assert (n_pre, n_risk, n_car) == (0, 1, 1) and af == "{9:1,70:2,90:1}"


def test_vcf_roundtrip_through_the_callers_writer(tmp_path, monkeypatch):
    repo = TREDsRepo()
    monkeypatch.chdir(tmp_path)
    calls = _calls(15, 41, "risk", 0.99951)
    tredmod.to_vcf({"samplekey": "v1", "bam": "v1.bam", "tredCalls": calls}, "hg38", repo, treds=["HD"])
    d = tredreport.vcf_to_df_worker(str(tmp_path / "v1.tred.vcf.gz"))
    assert d["SampleKey"] == "v1" and (d["HD.1"], d["HD.2"]) == (15, 41) and d["HD.label"] == "risk"
    assert abs(d["HD.PP"] - 0.9995) < 1e-12 and d["HD.FR"] == "15|4"
    df = tredreport.df_to_tsv(tredreport.vcf_to_df([str(tmp_path / "v1.tred.vcf.gz")]), str(tmp_path / "v.tsv"),
                              jsonformat=False)
    assert list(pd.read_csv(str(tmp_path / "v.tsv"), sep="\t").columns) == ["SampleKey", "HD.calls", "HD.label"]


def test_cli(cohort_dir, capsys):
    tmp, files = cohort_dir
    tsv = str(tmp / "cli.tsv")
    assert tredreport.main(files + ["--tsv", tsv, "--cpus", "1"]) == 0
    assert os.path.exists(tsv + ".report.txt")
    assert tredreport.main(["--tsv", tsv]) == 0                    # summarise an existing TSV
    assert tredreport.main(["--tsv", str(tmp / "absent.tsv")]) == 1
