"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the evidence-extraction half of the reference's hot path,
``tredparse/bam_parser.py`` (citations are into /root/reference/):

* template family (DB) of a locus ............ bam_parser.py:84-100, rc at :448-450
* per-read SW + classification + arg-max ...... bam_parser.py:102-182 (post-filter src/ssw_wrap.py:213-220)
* read selection, alts, tallies, REPT pairs ... bam_parser.py:184-287
* paired-end distances ........................ bam_parser.py:316-369
* depth, read length .......................... bam_parser.py:372-411

The alignments themselves come from ``oracle.sw`` (engine="oracle": our C restatement; engine="ref":
the reference's ssw.c compiled unmodified).  ``samfile`` arguments are anything with pysam's
``fetch`` / ``getrname`` interface.

Parity status: PINNED to the reference's own code — tests/test_reference_pinned.py compares this module with
the outputs of the reference's BamParser / PEextractor / tred.run (tests/golden/ref_tred_*.json,
ref_problems.json.gz, made through oracle/refshim.py), defaults and --useclippedreads / --norepeatpairs /
--noalts; the README.md:77-86 evidence strings reproduce exactly.  pysam itself is unavailable: fetch / pileup
semantics are restated (oracle/pysam_stub.py lists the pileup readings; tests/golden/ref_depth_dm1.json).
"""
import math
from collections import defaultdict

import numpy as np

from . import sw

SPAN = 1000
FLANKMATCH = 9
DNAPE_ELONGATE = SPAN * 10

_COMPLEMENT = str.maketrans("ATCGatcgNnXx", "TAGCtagcNnXx")


def rc(s):
    return s.translate(_COMPLEMENT)[::-1]


def template_family(prefix, repeat, suffix, max_units):
    """[(units, sequence)] in DB order: units ascending, forward template then its reverse complement."""
    db = []
    for units in range(1, max_units + 1):
        fwd = prefix + repeat * units + suffix
        db.append((units, fwd))
        db.append((units, rc(fwd)))
    return db


def get_hangs(rb, re, qb, qe, n, m):
    aL, aR = rb, n - re - 1
    bL, bR = qb, m - qe - 1
    return min(aR + bL, aL + bR, aL + aR, bL + bR)


def classify_alignment(score, rb, re, qb, qe, m, n, units, period, max_units_eff):
    """Python twin of tro_classify (kept separate on purpose: the two are cross-checked)."""
    min_len = min(m, n) // 2
    min_score = max(min_len, 30)
    if not (score >= min_score and (qe - qb + 1) >= min_len):
        return None
    prefix_read = rb < FLANKMATCH
    suffix_read = re > n - FLANKMATCH - 1
    hang_read = get_hangs(rb, re, qb, qe, n, m) >= FLANKMATCH
    if hang_read:
        return "HANG"
    if prefix_read:
        return "FULL" if suffix_read else "PREF"
    if suffix_read:
        return "POST"
    if units >= max_units_eff - 1 and units * period <= m:
        return "REPT"
    return None


class EvidenceOracle:
    """BamParser restated.  Attributes mirror the reference: counts, details, rept, ploidy, ..."""

    def __init__(self, tred, READLEN, gender="Unknown", depth=30, clip=False, alts=True,
                 repeatpairs=False, ref="hg38", engine="oracle"):
        self.tred = tred
        self.READLEN = READLEN
        self.gender = gender
        self.depth = depth
        self.clip = clip
        self.alts = alts
        self.repeatpairs = repeatpairs
        self.ref = ref
        self.engine = engine
        self.repeat = tred.repeat
        self.period = self.repeatSize = len(tred.repeat)
        self.chr = tred.chr
        self.ploidy = 1 if (gender == "Male" and tred.is_xlinked) else tred.ploidy
        self.startRepeat, self.endRepeat = tred.repeat_start, tred.repeat_end
        self.referenceLen = tred.repeat_end - tred.repeat_start + 1
        self.max_units = int(math.ceil(READLEN * 1.0 / self.period))
        shared = defaultdict(int)
        self.counts = {"PREF": shared, "POST": shared, "FULL": defaultdict(int),
                       "REPT": defaultdict(int), "HANG": defaultdict(int)}
        self.details = []
        self.rept = 0
        self.db = template_family(tred.prefix, tred.repeat, tred.suffix, self.max_units)
        self._tseqs = [t for _, t in self.db]
        self.pair_log = []          # optional: every (read, template) alignment, for golden vectors
        self.read_log = []          # every read handed to the SW classifier, with its outcome

    # --- one read ---------------------------------------------------------------------------------
    def align_all(self, seq):
        fn = sw.ref_align_pairs if self.engine == "ref" else sw.oracle_align_pairs
        k = len(self._tseqs)
        return fn([seq], self._tseqs, np.zeros(k, dtype=np.int32), np.arange(k, dtype=np.int32))

    def classify_read(self, seq, keep_pairs=False):
        res = []
        m = len(seq)
        al = self.align_all(seq)
        if keep_pairs:
            self.pair_log.append(al[:, :5].copy())
        max_units_eff = int(math.ceil(m * 1.0 / self.period)) if self.clip else self.max_units
        for (units, target), row in zip(self.db, al):
            score, rb, re, qb, qe = (int(x) for x in row[:5])
            tag = classify_alignment(score, rb, re, qb, qe, m, len(target), units, self.period,
                                     max_units_eff)
            if tag is None:
                continue
            res.append((score, units, tag))
        if not res:
            return None
        return max(res, key=lambda x: (x[0], -x[1]))

    def _parse_read(self, read, keep_pairs=False):
        seq = read.query_sequence
        best = self.classify_read(seq, keep_pairs=keep_pairs)
        self.read_log.append((read.query_name, seq, best))
        if best is None:
            return
        score, h, tag = best
        self.counts["HANG"][h] += 1
        if tag == "HANG":
            return
        self.details.append({"tag": tag, "h": h, "id": read.query_name, "seq": seq})

    # --- whole locus ------------------------------------------------------------------------------
    def parse(self, samfile, pad=SPAN, keep_pairs=False):
        WINDOW_START = max(0, self.startRepeat - pad)
        WINDOW_END = self.endRepeat + pad
        READ_START = max(0, self.startRepeat - self.READLEN)
        READ_END = self.endRepeat + self.READLEN
        chr = self.chr
        try:
            primary = samfile.fetch(chr, WINDOW_START, WINDOW_END)
            ok = True
        except ValueError:
            ok = False
        if ok:
            for read in primary:
                if not read.is_unmapped:
                    if read.reference_start < READ_START or read.reference_start > READ_END:
                        continue
                self._parse_read(read, keep_pairs)
            if self.alts and not self.clip:
                for c, s, e in self.tred.alt:
                    try:
                        if "nochr" in self.ref:
                            c = c[3:]
                        for read in samfile.fetch(c, s, e):
                            rid = read.next_reference_id
                            if rid == -1:
                                continue
                            if samfile.getrname(rid) != chr:
                                continue
                            rstart = read.next_reference_start
                            if rstart < WINDOW_START or rstart > WINDOW_END:
                                continue
                            self._parse_read(read, keep_pairs)
                    except Exception:
                        continue
        return self._finish()

    def _finish(self):
        """Tail of BamParser.parse (bam_parser.py:248-258): REPT-pair removal, tallies, rept."""
        if not (self.repeatpairs or self.clip):
            self.remove_pairs_of_rept()
        for x in self.details:
            self.counts[x["tag"]][x["h"]] += 1
        self.rept = sum(self.counts["REPT"].values()) if self.counts["REPT"] else 0
        return self

    def parse_reads(self, seqs, names):
        """The same for reads given as strings (synthetic problems: no BAM in between)."""
        class _R(object):
            __slots__ = ("query_sequence", "query_name")
        for s, n in zip(seqs, names):
            r = _R()
            r.query_sequence, r.query_name = s, n
            self._parse_read(r)
        return self._finish()

    def remove_pairs_of_rept(self):
        seen = defaultdict(int)
        for x in self.details:
            if x["tag"] == "REPT":
                seen[x["id"]] += 1
        drop = set(k for k, v in seen.items() if v > 1)
        self.details = [x for x in self.details if x["id"] not in drop]


class PEOracle:
    """PEextractor restated (bam_parser.py:316-369)."""

    def __init__(self, samfile, chr, start, end):
        self.ref = end - start + 1
        pstart = max(start - DNAPE_ELONGATE, 0)
        pend = end + DNAPE_ELONGATE
        cache = {}
        try:
            it = samfile.fetch(chr, pstart, pend)
            cache = defaultdict(list)
            for x in it:
                if not x.is_paired or x.is_unmapped or x.is_duplicate:
                    continue
                cache[x.query_name].append(x)
        except ValueError:
            cache = {}
        self.global_lens, self.target_lens = [], []
        tstart = start - FLANKMATCH
        tend = end + FLANKMATCH
        for name, reads in cache.items():
            if len(reads) < 2:
                continue
            a, b = reads[:2]
            if not ((not a.is_reverse) and b.is_reverse):
                continue
            tlen = self.target_length(a, b)
            if tlen >= SPAN:
                continue
            if a.reference_start < tstart and b.reference_end > tend:
                self.target_lens.append(tlen)
            else:
                self.global_lens.append(tlen)
        self.MINPE = end - start + 2 * FLANKMATCH + 2

    @staticmethod
    def target_length(a, b):
        start, end = a.reference_start, b.reference_end
        if a.query_alignment_start > 0:
            start -= a.query_alignment_start
        if b.query_alignment_end < b.query_length:
            end += b.query_length - b.query_alignment_end
        return end - start


def region_depth(samfile, chr, start, end):
    """BamDepth.region_depth restated (bam_parser.py:404-411) under pysam's default pileup semantics:
    every column of every qualifying read overlapping the window is counted (truncate=False), reads
    flagged UNMAP / SECONDARY / QCFAIL / DUP are skipped (stepper='all'). UNPINNED — see module doc."""
    total = 0
    for r in samfile.fetch(chr, start, end):
        if r.flag & (0x4 | 0x100 | 0x200 | 0x400):
            continue
        if not r.cigartuples:
            continue
        total += sum(l for op, l in r.cigartuples if op in (0, 2, 3, 7, 8))
    return total * 1.0 / (end - start + 1)


def read_length(samfile, firstN=100):
    """BamReadLen.readlen restated (bam_parser.py:381-391): max query_length of the first ~100 reads."""
    rls = []
    for read in samfile.fetch():
        rls.append(read.query_length)
        if len(rls) > firstN:
            break
    return max(rls)
