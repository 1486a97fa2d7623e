#!/usr/bin/env python
"""
Generate the committed golden fixtures (run in the build container where /root/reference is mounted):

    python tests/golden/make_fixtures.py

Outputs (all under tests/golden/):

  t001.mini.bam(.bai), t002.mini.bam(.bai)
        The reference's two test BAMs (tests/t001.bam = chr4 around HD, tests/t002.bam = chr19 around
        DM1) re-encoded by tredparse_b200.bamio.write_bam: records overlapping the +-10 kb paired-end
        window of the locus, base qualities and aux tags dropped, header contigs kept.  Every fetch the
        hot path makes (parse window, PE window, depth window) returns the same records as on the
        original files (asserted below).

  sw_pairs_<sample>_<tred>.npz
        For every read the reference's BamParser would hand to Smith-Waterman, every template of the
        locus (DB order): score, ref_begin, ref_end, query_begin, query_end, score2, ref_end2 and the
        CIGAR — produced by the REFERENCE's own src/ssw.c (oracle/_ref/libssw_ref.so, unmodified), i.e.
        genuine reference outputs.  Plus per-read (score, h, tag) and the FR/PR/RR strings.

  sw_pairs_synthetic.npz
        Random / simulated / adversarial (read, template) pairs (150 & 250 bp, indels, N, >=250 scores,
        pure-repeat ties, tiny and ragged lengths) with the reference library's outputs.


The likelihood / pipeline goldens (ref_*.json) are made by tests/golden/make_ref_fixtures.py from the
reference's own Python code.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import sw, evidence_oracle as evo  # noqa: E402
from tredparse_b200 import bamio  # noqa: E402
from tredparse_b200.meta import TREDsRepo  # noqa: E402

REFERENCE = "/root/reference"
CASES = [("t001", "HD"), ("t002", "DM1")]


def trim_bam(sample, tred):
    src = os.path.join(REFERENCE, "tests", sample + ".bam")
    dst = os.path.join(HERE, sample + ".mini.bam")
    sam = bamio.AlignmentFile(src)
    lo = max(tred.repeat_start - evo.DNAPE_ELONGATE - 300, 0)
    hi = tred.repeat_end + evo.DNAPE_ELONGATE + 300
    recs = [r for r in sam.fetch() if r.reference_id == sam.get_tid(tred.chr)
            and r.reference_start < hi and (r.reference_end or r.reference_start + 1) > lo]
    for r in recs:
        r._qual = None
    refs = list(zip(sam.references, sam.lengths))
    bamio.write_bam(dst, refs, recs, level=9)
    print("{}: kept {} of the records -> {} bytes".format(sample, len(recs), os.path.getsize(dst)))
    # the three windows the hot path fetches must see identical records
    new = bamio.AlignmentFile(dst)
    for (a, b) in ((tred.repeat_start - 1000, tred.repeat_end + 1000),
                   (tred.repeat_start - 10000, tred.repeat_end + 10000)):
        x = [(r.query_name, r.flag, r.reference_start, r.query_sequence, tuple(r.cigartuples))
             for r in sam.fetch(tred.chr, a, b)]
        y = [(r.query_name, r.flag, r.reference_start, r.query_sequence, tuple(r.cigartuples))
             for r in new.fetch(tred.chr, a, b)]
        assert x == y, "trimmed BAM differs in window {}-{}".format(a, b)
    assert evo.read_length(sam) == evo.read_length(new)
    return dst


def counter_s(c):
    return ";".join("{}|{}".format(k, int(v)) for k, v in sorted(c.items()))


def evidence_case(sample, tredname, repo, bam):
    tred = repo[tredname]
    sam = bamio.AlignmentFile(bam)
    READLEN = evo.read_length(sam)
    depth = evo.region_depth(sam, tred.chr, max(0, tred.repeat_start - 1000), tred.repeat_end + 1000)
    ev = evo.EvidenceOracle(tred, READLEN, depth=depth, alts=True, repeatpairs=True, engine="ref")
    ev.parse(sam, keep_pairs=True)
    # CIGARs + score2 for every pair straight from the reference library
    reads = [s for (_, s, _) in ev.read_log]
    names = [n for (n, _, _) in ev.read_log]
    tseqs = [t for _, t in ev.db]
    nq, nt = len(reads), len(tseqs)
    qidx = np.repeat(np.arange(nq, dtype=np.int32), nt)
    tidx = np.tile(np.arange(nt, dtype=np.int32), nq)
    out, cig, clen = sw.ref_align_pairs(reads, tseqs, qidx, tidx, cigar_cap=64)
    assert clen.max() <= 64
    assert np.array_equal(out[:, :5].reshape(nq, nt, 5), np.stack(ev.pair_log))
    best = np.array([[-1, 0, 0] if b is None else
                     [b[0], b[1], {"FULL": 1, "PREF": 2, "POST": 3, "REPT": 4, "HANG": 5}[b[2]]]
                     for (_, _, b) in ev.read_log], dtype=np.int32)
    FR, PR, RR = (counter_s(ev.counts[t]) for t in ("FULL", "PREF", "REPT"))
    print(sample, tredname, "reads", nq, "READLEN", READLEN, "depth", depth)
    print("  FR", FR, "\n  PR", PR, "\n  RR", RR)
    np.savez_compressed(os.path.join(HERE, "sw_pairs_{}_{}.npz".format(sample, tredname)),
                        reads=np.array(reads), names=np.array(names), templates=np.array(tseqs),
                        units=np.array([u for u, _ in ev.db], dtype=np.int32),
                        pairs=out.reshape(nq, nt, 7).astype(np.int16),
                        cigar=cig.reshape(nq, nt, 64)[:, :, :int(clen.max())].astype(np.uint16),
                        cigar_len=clen.reshape(nq, nt).astype(np.int8),
                        best=best, FR=FR, PR=PR, RR=RR, READLEN=READLEN, depth=depth,
                        period=len(tred.repeat))
    pe = evo.PEOracle(sam, tred.chr, tred.repeat_start, tred.repeat_end)
    return ev, pe, READLEN, depth


def synthetic_pairs():
    rng = np.random.default_rng(20261017)
    B = "ACGT"

    def rnd(n, pn=0.0):
        s = rng.integers(0, 4, n)
        out = np.array(list(B))[s]
        if pn:
            out[rng.random(n) < pn] = "N"
        return "".join(out)

    def mutate(s, sub=0.01, indel=0.003, maxindel=12):
        out = []
        i = 0
        while i < len(s):
            r = rng.random()
            if r < sub:
                out.append(B[rng.integers(0, 4)])
                i += 1
            elif r < sub + indel:
                L = int(rng.integers(1, maxindel + 1))
                if rng.random() < 0.5:
                    out.append(rnd(L))
                else:
                    i += L
            else:
                out.append(s[i])
                i += 1
        return "".join(out)

    repo = TREDsRepo()
    queries, templates = [], []
    # (1) reads simulated from templates of several loci (periods 3,4,5,6,12; N motifs), 150 and 250 bp
    for name in ("HD", "DM1", "FXS", "DM2", "SCA10", "SCA36", "ULD", "OPMD", "BPES", "FRDA", "SCA8"):
        t = repo[name]
        for readlen in (150, 250):
            mu = -(-readlen // len(t.repeat))
            for _ in range(30):
                u = int(rng.integers(1, mu + 1))
                h = int(rng.integers(1, mu + 30))
                hap = rnd(300) + t.prefix + t.repeat.replace("N", B[rng.integers(0, 4)]) * h + t.suffix + rnd(300)
                st = int(rng.integers(0, max(1, len(hap) - readlen)))
                read = mutate(hap[st:st + readlen])
                if rng.random() < 0.5:
                    read = evo.rc(read)
                tpl = t.prefix + t.repeat * u + t.suffix
                if rng.random() < 0.5:
                    tpl = evo.rc(tpl)
                queries.append(read)
                templates.append(tpl)
    # (2) pure-repeat ties and perfect long matches (score >= 250 -> the reference's 16-bit kernel)
    for motif in ("CAG", "CGG", "GAA", "CAGG", "ATTCT", "GGCCTG"):
        for readlen in (150, 250, 300):
            for u in (10, 50, 84, 120):
                off = int(rng.integers(0, len(motif)))
                queries.append((motif * 200)[off:off + readlen])
                templates.append("ACGTACGTACGTACGTAC" + motif * u + "TGCATGCATGCATGCATG")
    for n in (250, 255, 256, 260, 300, 400):
        s = rnd(n)
        queries.append(s)
        templates.append(rnd(20) + s + rnd(20))
        queries.append(mutate(s, 0.02, 0.004))
        templates.append(rnd(5) + s + rnd(7))
    # (3) adjacent insertion/deletion, gaps at the edges, N-rich, tiny and ragged lengths
    for _ in range(300):
        n = int(rng.integers(20, 200))
        s = rnd(n)
        cut = int(rng.integers(5, n - 5))
        L1, L2 = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        q = s[:cut] + rnd(L1) + s[cut + L2:]
        queries.append(q)
        templates.append(rnd(int(rng.integers(0, 10))) + s + rnd(int(rng.integers(0, 10))))
    for _ in range(200):
        queries.append(rnd(int(rng.integers(1, 40)), pn=0.1))
        templates.append(rnd(int(rng.integers(1, 60)), pn=0.1))
    for _ in range(200):
        queries.append(rnd(int(rng.integers(31, 151)), pn=0.05))
        templates.append(rnd(int(rng.integers(39, 190)), pn=0.05))
    queries += ["A", "ACGT", "N" * 40 + "ACGTACGTAAGGCCTTAGCATCGATCGATCGACTAGCTAGCTAC"]
    templates += ["A", "TTTT", "ACGTACGTAAGGCCTTAGCATCGATCGATCGACTAGCTAGCTAC"]
    # every output must be a genuine alignment for the reference to be well defined (score > 0)
    n = len(queries)
    idx = np.arange(n, dtype=np.int32)
    out, cig, clen = sw.ref_align_pairs(queries, templates, idx, idx, cigar_cap=128)
    keep = out[:, 0] > 0
    assert clen.max() <= 128
    print("synthetic pairs:", n, "kept", int(keep.sum()), "score>=250:", int((out[:, 0] >= 250).sum()))
    np.savez_compressed(os.path.join(HERE, "sw_pairs_synthetic.npz"),
                        queries=np.array(queries)[keep], templates=np.array(templates)[keep],
                        pairs=out[keep], cigar=cig[keep][:, :int(clen.max())], cigar_len=clen[keep])


def main():
    sw.build(REFERENCE)
    repo = TREDsRepo()
    for sample, tredname in CASES:
        tred = repo[tredname]
        bam = trim_bam(sample, tred)
        evidence_case(sample, tredname, repo, bam)
    synthetic_pairs()


if __name__ == "__main__":
    main()
