"""
ORACLE / CPU BASELINE — TEST INFRASTRUCTURE ONLY.

The reference's hot path on the CPU with the reference's own call pattern, for ``bench.py``'s
``cpu_baseline`` and ``--impl reference`` legs:

* every (read, template) alignment is one Python-level call that re-encodes the query base by base and
  goes ``ssw_init -> ssw_align(flag=1) -> read fields -> init_destroy -> align_destroy`` through ctypes —
  the pattern of src/ssw_wrap.py:177-244 — against ``oracle/_ref/libssw_ref.so`` (the reference's
  src/ssw.c compiled unmodified) when present, else against our C restatement's batch entry;
* classification / tallies / PE model / grid are the Python restatements (the reference's Python layer is
  Python-2 only), with scipy's gaussian_kde and poisson.pmf exactly where the reference uses them.

``genotype_problem`` is one (sample, locus) unit of work; ``run_pool`` maps it over a process pool like
the reference's ``multiprocessing.Pool(cpus).imap(run, ...)`` (tredparse/tred.py:528-532).
"""
import ctypes
import math
import os
from ctypes import POINTER, Structure, c_int8, c_int32, c_uint8, c_uint16, c_uint32, c_void_p

from . import sw, evidence_oracle as evo, likelihood_oracle as lko


class CAlignRes(Structure):
    _fields_ = [("score", c_uint16), ("score2", c_uint16), ("ref_begin", c_int32), ("ref_end", c_int32),
                ("query_begin", c_int32), ("query_end", c_int32), ("ref_end2", c_int32),
                ("cigar", POINTER(c_uint32)), ("cigarLen", c_int32)]


_BASE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "a": 0, "c": 1, "g": 2, "t": 3, "n": 4}
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(sw.REF_SO)
        lib.ssw_init.restype = c_void_p
        lib.ssw_init.argtypes = [POINTER(c_int8), c_int32, POINTER(c_int8), c_int32, c_int8]
        lib.init_destroy.restype = None
        lib.init_destroy.argtypes = [c_void_p]
        lib.ssw_align.restype = POINTER(CAlignRes)
        lib.ssw_align.argtypes = [c_void_p, POINTER(c_int8), c_int32, c_uint8, c_uint8, c_uint8, c_uint16,
                                  c_int32, c_int32]
        lib.align_destroy.restype = None
        lib.align_destroy.argtypes = [POINTER(CAlignRes)]
        _lib = lib
    return _lib


def dna_to_int_mat(seq):
    arr = (c_int8 * len(seq))()
    for idx, base in enumerate(seq):
        arr[idx] = _BASE.get(base, 4)
    return arr


class RefAligner(object):
    """One template, aligned against many queries one call at a time (ssw_wrap.Aligner's behaviour)."""

    def __init__(self, ref_seq, match=1, mismatch=5, gap_open=7, gap_extend=2):
        self.lib = _load()
        self.ref_seq = ref_seq
        self._ref = dna_to_int_mat(ref_seq)
        m, x = match, -mismatch
        self.mat = (c_int8 * 25)(m, x, x, x, 0, x, m, x, x, 0, x, x, m, x, 0, x, x, x, m, 0, 0, 0, 0, 0, 0)
        self.go, self.ge = gap_open, gap_extend

    def align(self, query_seq, min_score=0, min_len=0):
        q = dna_to_int_mat(query_seq)
        prof = self.lib.ssw_init(q, c_int32(len(query_seq)), self.mat, 5, 2)
        mask_len = len(query_seq) // 2 if len(query_seq) > 30 else 15
        res = self.lib.ssw_align(prof, self._ref, c_int32(len(self.ref_seq)), self.go, self.ge, 1, 0, 0, mask_len)
        c = res.contents
        out = None
        if c.score >= min_score and (c.query_end - c.query_begin + 1) >= min_len:
            out = (c.score, c.ref_begin, c.ref_end, c.query_begin, c.query_end,
                   [c.cigar[i] for i in range(c.cigarLen)])
        self.lib.init_destroy(prof)
        self.lib.align_destroy(res)
        return out


def classify_reads_percall(tred, readlen, reads):
    """BamParser._buildDB + _parseReadSW for a list of read strings -> counts (FULL/PREF/REPT), cells"""
    period = len(tred.repeat)
    max_units = int(math.ceil(readlen * 1. / period))
    db = [(u, t, RefAligner(t)) for u, t in evo.template_family(tred.prefix, tred.repeat, tred.suffix, max_units)]
    counts = {"FULL": {}, "PREF": {}, "REPT": {}}
    cells = 0
    for seq in reads:
        res = []
        for units, target, al in db:
            min_len = min(len(seq), len(target)) // 2
            min_score = max(min_len, 30)
            cells += len(seq) * len(target)
            r = al.align(seq, min_score=min_score, min_len=min_len)
            if not r:
                continue
            tag = evo.classify_alignment(r[0], r[1], r[2], r[3], r[4], len(seq), len(target), units, period,
                                         max_units) if True else None
            # classify_alignment re-applies the (already passed) filter; harmless
            if tag is None:
                continue
            res.append((r[0], units, tag))
        if not res:
            continue
        score, h, tag = max(res, key=lambda x: (x[0], -x[1]))
        if tag == "HANG":
            continue
        tag = "PREF" if tag == "POST" else tag
        counts[tag][h] = counts[tag].get(h, 0) + 1
    return counts, cells


class _PE:
    pass


def genotype_problem(args):
    """One (sample, locus) unit through the reference-shaped CPU path.
    args = (tred, readlen, ploidy, depth, read strings, global_lens, target_lens, step_model, weights)"""
    tred, readlen, ploidy, depth, reads, global_lens, target_lens, step, weights = args
    counts, cells = classify_reads_percall(tred, readlen, reads)
    pe = _PE()
    pe.global_lens, pe.target_lens = list(global_lens), list(target_lens)
    pe.ref = tred.repeat_end - tred.repeat_start + 1
    pe.MINPE = pe.ref - 1 + 2 * 9 + 2
    rept = sum(counts["REPT"].values())
    lk = lko.LikelihoodOracle(tred, len(tred.repeat), readlen, counts, rept, ploidy, depth, pe, step, weights)
    lk.call()
    return (lk.alleles, lk.CI, lk.PP, lk.label, len(reads), cells, len(lk.surface))


def run_pool(tasks, cores=None):
    """Map genotype_problem over a process pool (fork), like the reference's Pool over samples."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    cores = max(1, min(cores, len(tasks)))
    if cores == 1:
        return [genotype_problem(t) for t in tasks], 1
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=cores) as pool:
        return pool.map(genotype_problem, tasks, chunksize=1), cores
