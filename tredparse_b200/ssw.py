"""
``ssw``-compatible module on top of libtredsw.so.

Mirrors the reference's ``src/ssw_wrap.py`` (package ``ssw``): ``Aligner(ref_seq, match, mismatch,
gap_open, gap_extend, report_secondary, report_cigar)`` (:110-117) with ``align(query_seq, min_score,
min_len) -> PyAlignRes | None`` (:177-227) and ``PyAlignRes`` (:259-383) — same argument meaning,
same filtering rule, same coordinates — plus the batched entry points the hot path uses
(:func:`align_pairs`, :func:`classify_reads`), which replace thousands of per-call
``ssw_init/ssw_align`` round trips by one kernel launch.
"""
import numpy as np

from . import _lib

_CODE = np.full(256, 4, dtype=np.int8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq):
    """DNA string -> int8 codes, A,C,G,T,N -> 0..4, anything else 4 (ssw_wrap.py:61,229-244)."""
    return _CODE[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]


def score_matrix(match, mismatch):
    """5x5 matrix of ssw_wrap.py:154-167: +match on the diagonal, -mismatch elsewhere, N scores 0."""
    m = np.full((5, 5), -mismatch, dtype=np.int8)
    np.fill_diagonal(m, match)
    m[4, :] = 0
    m[:, 4] = 0
    return np.ascontiguousarray(m)


def flatten(seqs):
    """list of str / code arrays -> (flat int8 buffer, int64 offsets[n+1])"""
    arrs = [encode(s) if isinstance(s, str) else np.asarray(s, dtype=np.int8) for s in seqs]
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs])
    buf = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.int8)
    return np.ascontiguousarray(buf, dtype=np.int8), off


def align_pairs(queries, templates, qidx, tidx, match=1, mismatch=5, gap_open=7, gap_extend=2,
                score2=False, cigar_cap=0, ctx=None):
    """Batch of independent alignments.  Returns int32 [npairs, 8] = score, ref_begin, ref_end,
    query_begin, query_end, score2, ref_end2, cigar_len (and the CIGAR words when cigar_cap > 0)."""
    ctx = ctx or _lib.default_context()
    qbuf, qoff = flatten(queries)
    tbuf, toff = flatten(templates)
    qidx = np.ascontiguousarray(qidx, dtype=np.int32)
    tidx = np.ascontiguousarray(tidx, dtype=np.int32)
    n = len(qidx)
    out = np.zeros((n, 8), dtype=np.int32)
    flags = (_lib.SCORE2 if score2 else 0) | (_lib.CIGAR if cigar_cap else 0)
    cig = np.zeros((n, cigar_cap), dtype=np.uint32) if cigar_cap else None
    mat = score_matrix(match, mismatch)
    rc = ctx.lib.tredsw_align_pairs(ctx.handle, _lib.ptr(qbuf), _lib.ptr(qoff), len(qoff) - 1,
                                    _lib.ptr(tbuf), _lib.ptr(toff), len(toff) - 1, _lib.ptr(qidx),
                                    _lib.ptr(tidx), n, _lib.ptr(mat), gap_open, gap_extend, flags,
                                    _lib.ptr(out), _lib.ptr(cig), cigar_cap)
    _lib.check(rc, "tredsw_align_pairs")
    return (out, cig) if cigar_cap else out


def make_family(prefix, repeat, suffix, max_units, clip=False):
    """Template family of a locus (bam_parser.py:84-100) as a tredsw_family record."""
    fam = np.zeros(1, dtype=_lib.FAMILY_DTYPE)
    p, s, r = encode(prefix), encode(suffix), encode(repeat)
    if len(p) > 32 or len(s) > 32 or len(r) > 32:
        raise _lib.TredswError("flank / motif longer than 32 bp is not supported")
    fam["prefix"][0, :len(p)] = p
    fam["suffix"][0, :len(s)] = s
    fam["repeat"][0, :len(r)] = r
    fam["prefix_len"], fam["suffix_len"], fam["period"] = len(p), len(s), len(r)
    fam["max_units"], fam["clip"] = max_units, int(bool(clip))
    return fam


def classify_reads(reads, read_family, families, match=1, mismatch=5, gap_open=7, gap_extend=2,
                   ctx=None, want_stats=False):
    """Fused per-read Smith-Waterman + classification (BamParser._parseReadSW for a batch).

    reads: list of str / code arrays, or a (flat int8 buffer, int64 offsets) tuple;
    read_family[r]: index into `families` (array of FAMILY_DTYPE).
    Returns int32 [nreads, 8] = tag, h, score, ref_begin, ref_end, query_begin, query_end, rank."""
    ctx = ctx or _lib.default_context()
    rbuf, roff = reads if isinstance(reads, tuple) else flatten(reads)
    n = len(roff) - 1
    read_family = np.ascontiguousarray(read_family, dtype=np.int32)
    families = np.ascontiguousarray(families, dtype=_lib.FAMILY_DTYPE)
    out = np.zeros((n, 8), dtype=np.int32)
    stats = np.zeros(4, dtype=np.int64)
    mat = score_matrix(match, mismatch)
    rc = ctx.lib.tredsw_classify_reads(ctx.handle, _lib.ptr(rbuf), _lib.ptr(roff), n,
                                       _lib.ptr(read_family), _lib.ptr(families), len(families),
                                       _lib.ptr(mat), gap_open, gap_extend, 0, _lib.ptr(out),
                                       _lib.ptr(stats))
    _lib.check(rc, "tredsw_classify_reads")
    return (out, stats) if want_stats else out


class PyAlignRes(object):
    """Alignment result, attribute-compatible with ssw_wrap.PyAlignRes (:259-383)."""
    _OPS = "MIDNSHP=X"

    def __init__(self, row, cigar_words, query_seq, ref_seq):
        self.score = int(row[0])
        self.score2 = None
        self.ref_seq = ref_seq
        self.ref_begin = int(row[1])
        self.ref_end = int(row[2])
        self.query_seq = query_seq
        self.query_begin = int(row[3])
        self.query_end = int(row[4])
        self._cigar_string = [int(x) for x in cigar_words]

    @property
    def iter_cigar(self):
        for val in self._cigar_string:
            yield (val >> 4, self._OPS[val & 0xF] if (val & 0xF) < len(self._OPS) else "M")

    @property
    def cigar_string(self):
        """CIGAR with soft clips for the unaligned query ends (ssw_wrap.py:320-340)."""
        cig = ""
        if self.query_begin > 0:
            cig += "{}S".format(self.query_begin)
        cig += "".join("{}{}".format(l, op) for l, op in self.iter_cigar)
        end_len = len(self.query_seq) - self.query_end - 1
        if end_len != 0:
            cig += "{}S".format(end_len)
        return cig

    @property
    def alignment(self):
        """(reference line, match line, query line) of the alignment (ssw_wrap.py:342-383)."""
        r_index, q_index = 0, 0
        r_seq = self.ref_seq[self.ref_begin:self.ref_end + 1]
        q_seq = self.query_seq[self.query_begin:self.query_end + 1]
        r_line = m_line = q_line = ""
        for op_len, op_char in self.iter_cigar:
            if op_char.upper() == "M":
                for (r, q) in zip(r_seq[r_index:r_index + op_len], q_seq[q_index:q_index + op_len]):
                    r_line += r
                    q_line += q
                    m_line += "|" if r == q else "*"
                r_index += op_len
                q_index += op_len
            elif op_char.upper() == "I":
                r_line += "-" * op_len
                m_line += " " * op_len
                q_line += q_seq[q_index:q_index + op_len]
                q_index += op_len
            elif op_char.upper() == "D":
                r_line += r_seq[r_index:r_index + op_len]
                m_line += " " * op_len
                q_line += "-" * op_len
                r_index += op_len
        return r_line, m_line, q_line

    def __str__(self):
        msg = "OPTIMAL MATCH\n"
        msg += "Score            {}\n".format(self.score)
        msg += "Reference begin  {}\n".format(self.ref_begin)
        msg += "Reference end    {}\n".format(self.ref_end)
        msg += "Query begin      {}\n".format(self.query_begin)
        msg += "Query end        {}\n".format(self.query_end)
        if self.cigar_string:
            msg += "Cigar_string     {}\n".format(self.cigar_string)
        return msg


class Aligner(object):
    """Drop-in for ssw_wrap.Aligner (:54-256): one reference sequence, many queries."""

    def __init__(self, ref_seq="", match=2, mismatch=2, gap_open=3, gap_extend=1,
                 report_secondary=False, report_cigar=False):
        self.report_secondary = report_secondary
        self.report_cigar = report_cigar
        self.set_gap(gap_open, gap_extend)
        self.set_mat(match, mismatch)
        self.reference = ref_seq

    def set_gap(self, gap_open=3, gap_extend=1):
        self.gap_open = gap_open
        self.gap_extend = gap_extend

    def set_mat(self, match=2, mismatch=2):
        self.match = match
        self.mismatch = mismatch
        self.mat = score_matrix(match, mismatch)

    def get_reference(self):
        return self.ref_seq

    def set_reference(self, ref_seq):
        self.ref_seq = ref_seq
        self._ref_seq = encode(ref_seq)

    reference = property(get_reference, set_reference)

    def align(self, query_seq, min_score=0, min_len=0):
        """One alignment; None when filtered out (score < min_score or aligned query span < min_len,
        ssw_wrap.py:213-220)."""
        cap = 2 * (len(query_seq) + len(self.ref_seq)) + 8
        out, cig = align_pairs([query_seq], [self._ref_seq], [0], [0], self.match, self.mismatch,
                               self.gap_open, self.gap_extend, score2=True, cigar_cap=cap)
        row = out[0]
        match_len = int(row[4]) - int(row[3]) + 1
        if row[0] >= min_score and match_len >= min_len:
            return PyAlignRes(row, cig[0, :max(int(row[7]), 0)], query_seq, self.ref_seq)
        return None

    def __str__(self):
        return "\n<Instance of {} from {} >\n".format(self.__class__.__name__, self.__module__)
