"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes access to the two CPU Smith-Waterman checkers.

* ``oracle/_build/libsw_oracle.so``  — sw_oracle.c, our scalar restatement (always buildable);
* ``oracle/_ref/libssw_ref.so``      — the reference's src/ssw.c compiled unmodified + ref_batch.c
  (only where /root/reference was present at build time; it is git-ignored but travels to the GPU box).

Encoding follows src/ssw_wrap.py:61,229-244 (A,C,G,T,N -> 0..4, anything else 4) and the 5x5 matrix
of src/ssw_wrap.py:154-167.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libsw_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libssw_ref.so")

TAGS = {0: None, 1: "FULL", 2: "PREF", 3: "POST", 4: "REPT", 5: "HANG"}

_CODE = np.full(256, 4, dtype=np.int8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq: str) -> np.ndarray:
    return _CODE[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]


def score_matrix(match=1, mismatch=5) -> np.ndarray:
    m = np.full((5, 5), -mismatch, dtype=np.int8)
    np.fill_diagonal(m, match)
    m[4, :] = 0
    m[:, 4] = 0
    return m


def build(reference="/root/reference"):
    """Compile the oracle (always) and the reference library (when the reference tree is here)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if os.path.exists(os.path.join(reference, "src", "ssw.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REFERENCE=" + reference])


def _load(path):
    if not os.path.exists(path):
        raise OSError("{} not built; run `make -C oracle` (and `make -C oracle ref`)".format(path))
    return ctypes.CDLL(path)


_P8 = ctypes.POINTER(ctypes.c_int8)
_P32 = ctypes.POINTER(ctypes.c_int32)
_P64 = ctypes.POINTER(ctypes.c_int64)
_PU32 = ctypes.POINTER(ctypes.c_uint32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def flatten(seqs):
    """list of str or int8 arrays -> (flat int8 buffer, int64 offsets[n+1])"""
    arrs = [encode(s) if isinstance(s, str) else np.asarray(s, dtype=np.int8) for s in seqs]
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs])
    buf = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.int8)
    return np.ascontiguousarray(buf, dtype=np.int8), off


class _Lib:
    def __init__(self, path, fn_name, with_cigar):
        self.lib = _load(path)
        self.fn = getattr(self.lib, fn_name)
        self.with_cigar = with_cigar

    def align_pairs(self, queries, templates, qidx, tidx, match=1, mismatch=5, go=7, ge=2,
                    cigar_cap=0):
        qbuf, qoff = flatten(queries)
        tbuf, toff = flatten(templates)
        qidx = np.ascontiguousarray(qidx, dtype=np.int32)
        tidx = np.ascontiguousarray(tidx, dtype=np.int32)
        n = len(qidx)
        out = np.zeros((n, 7), dtype=np.int32)
        mat = score_matrix(match, mismatch)
        args = [_ptr(qbuf, _P8), _ptr(qoff, _P64), _ptr(tbuf, _P8), _ptr(toff, _P64),
                _ptr(qidx, _P32), _ptr(tidx, _P32), ctypes.c_int64(n), _ptr(mat, _P8), 5, go, ge,
                _ptr(out, _P32)]
        cig = clen = None
        if self.with_cigar:
            if cigar_cap:
                cig = np.zeros((n, cigar_cap), dtype=np.uint32)
                clen = np.zeros(n, dtype=np.int32)
                args += [_ptr(cig, _PU32), _ptr(clen, _P32), cigar_cap]
            else:
                args += [None, None, 0]
        self.fn(*args)
        if cigar_cap:
            return out, cig, clen
        return out


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = _Lib(ORACLE_SO, "tro_align_batch", False)
    return _oracle


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = _Lib(REF_SO, "ref_align_batch", True)
    return _ref


def oracle_align_pairs(queries, templates, qidx, tidx, **kw):
    """-> int32 [npairs, 7]: score, ref_begin, ref_end, query_begin, query_end, score2, ref_end2"""
    return oracle_lib().align_pairs(queries, templates, qidx, tidx, **kw)


def ref_align_pairs(queries, templates, qidx, tidx, **kw):
    return ref_lib().align_pairs(queries, templates, qidx, tidx, **kw)


def oracle_classify(score, rb, re, qb, qe, m, n, u, period, max_units_eff):
    lib = oracle_lib().lib
    return int(lib.tro_classify(int(score), int(rb), int(re), int(qb), int(qe), int(m), int(n),
                                int(u), int(period), int(max_units_eff)))


def oracle_cigar(ref_codes, read_codes, score, go=7, ge=2, match=1, mismatch=5, cap=512):
    lib = oracle_lib().lib
    ref_codes = np.ascontiguousarray(ref_codes, dtype=np.int8)
    read_codes = np.ascontiguousarray(read_codes, dtype=np.int8)
    out = np.zeros(cap, dtype=np.uint32)
    mat = score_matrix(match, mismatch)
    n = lib.tro_cigar(_ptr(ref_codes, _P8), _ptr(read_codes, _P8), len(ref_codes), len(read_codes),
                      int(score), go, ge, _ptr(mat, _P8), 5, _ptr(out, _PU32), cap)
    return None if n < 0 else out[:n].copy()
