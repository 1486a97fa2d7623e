"""
Native BAM ingest (``tredsw_bam_*`` in ``include/tredsw.h``, ``csrc/ingest.cpp``): one indexed pass per locus
instead of the reference's three pysam passes — read selection (``tredparse/bam_parser.py:194-243``),
``PEextractor`` (``:316-369``) and ``BamDepth.region_depth`` (``:404-411``) — straight into flat buffers in the
layout ``cohort.CohortBatch`` / ``tredsw_genotype_batch`` consume.  Host code only: no GPU is needed here.

    with BamIngest("sample.bam") as bam:
        ev = bam.extract_locus(repo["HD"], readlen=150)          # reads (codes), pair lengths, depth
        problem = bam.problem(repo["HD"], readlen=150)           # ready for cohort.CohortBatch([...])
"""
import ctypes

import numpy as np

from . import _lib
from .meta import TREDsRepo  # noqa: F401  (documentation cross-reference)

SPAN = 1000
FLANKMATCH = 9
DNAPE_ELONGATE = SPAN * 10


class LocusQuery(ctypes.Structure):
    """tredsw_locus_query (include/tredsw.h)"""
    _fields_ = [(n, ctypes.c_int32) for n in ("tid", "repeat_start", "repeat_end", "readlen", "pad", "pe_window",
                                              "flankmatch", "span", "n_alts", "reserved_")] + [("alts", ctypes.c_void_p)]


class LocusSummary(ctypes.Structure):
    """tredsw_locus_summary (include/tredsw.h)"""
    _fields_ = [("nreads", ctypes.c_int32), ("n_unmapped", ctypes.c_int32), ("n_global", ctypes.c_int32),
                ("n_target", ctypes.c_int32), ("nbases", ctypes.c_int64), ("name_bytes", ctypes.c_int64),
                ("depth", ctypes.c_double), ("overflow", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


def _bind(lib):
    if getattr(lib, "_ingest_bound", False):
        return
    lib.tredsw_bam_open.restype = ctypes.c_void_p
    lib.tredsw_bam_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    lib.tredsw_bam_inflate_stats.restype = None
    lib.tredsw_bam_inflate_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.tredsw_inflate_raw.restype = ctypes.c_int
    lib.tredsw_inflate_raw.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
    lib.tredsw_bam_clone.restype = ctypes.c_void_p
    lib.tredsw_bam_clone.argtypes = [ctypes.c_void_p]
    lib.tredsw_bam_close.restype = None
    lib.tredsw_bam_close.argtypes = [ctypes.c_void_p]
    lib.tredsw_bam_nref.restype = ctypes.c_int32
    lib.tredsw_bam_nref.argtypes = [ctypes.c_void_p]
    lib.tredsw_bam_tid.restype = ctypes.c_int32
    lib.tredsw_bam_tid.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.tredsw_bam_extract_locus.restype = ctypes.c_int
    lib.tredsw_bam_extract_locus.argtypes = [ctypes.c_void_p, ctypes.POINTER(LocusQuery), ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                             ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.POINTER(LocusSummary)]
    lib.tredsw_bam_region_depth.restype = ctypes.c_int
    lib.tredsw_bam_region_depth.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_double)]
    lib.tredsw_bam_read_length.restype = ctypes.c_int
    lib.tredsw_bam_read_length.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                           ctypes.POINTER(ctypes.c_int32)]
    lib._ingest_bound = True


class LocusEvidence:
    """What one locus of one BAM contributes to the hot path."""
    __slots__ = ("reads", "roff", "names", "global_lens", "target_lens", "depth", "n_unmapped")

    @property
    def nreads(self):
        return len(self.roff) - 1

    def read_strings(self):
        lut = np.array(list("ACGTN"))
        return ["".join(lut[self.reads[self.roff[i]:self.roff[i + 1]]]) for i in range(self.nreads)]


class BamIngest:
    def __init__(self, path, index=None):
        self.lib = _lib.load()
        _bind(self.lib)
        self.path = path
        self.handle = self.lib.tredsw_bam_open(path.encode(), index.encode() if index else None)
        if not self.handle:
            raise IOError("tredsw_bam_open({}): {}".format(path, _lib.last_error()))
        self._caps = dict(reads=512, bases=512 * 256, pairs=8192, names=512 * 48)

    def inflate_stats(self):
        """(blocks inflated by the library's own decoder, blocks that fell back to zlib) for this handle."""
        a, b = ctypes.c_int64(), ctypes.c_int64()
        self.lib.tredsw_bam_inflate_stats(self.handle, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def clone(self):
        """An independent handle on the same BAM for another host thread (shares the parsed index)."""
        other = object.__new__(BamIngest)
        other.lib, other.path = self.lib, self.path
        other.handle = self.lib.tredsw_bam_clone(self.handle)
        if not other.handle:
            raise IOError("tredsw_bam_clone({}): {}".format(self.path, _lib.last_error()))
        other._caps = dict(self._caps)
        return other

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tredsw_bam_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tid(self, contig):
        return int(self.lib.tredsw_bam_tid(self.handle, contig.encode()))

    def region_depth(self, contig, start, end):
        """``BamDepth.region_depth`` (bam_parser.py:404-411) on the native reader."""
        tid = self.tid(contig)
        if tid < 0:
            raise ValueError("invalid contig `{}`".format(contig))
        d = ctypes.c_double()
        _lib.check(self.lib.tredsw_bam_region_depth(self.handle, tid, int(start), int(end), ctypes.byref(d)),
                   "tredsw_bam_region_depth")
        return float(d.value)

    def read_length(self, firstN=100):
        """``BamReadLen.readlen`` (bam_parser.py:372-391) -> (max, min) over the first firstN + 1 records."""
        mx, mn = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self.lib.tredsw_bam_read_length(self.handle, int(firstN), ctypes.byref(mx), ctypes.byref(mn)),
                   "tredsw_bam_read_length")
        return int(mx.value), int(mn.value)

    def extract_locus(self, tred, readlen, alts=(), want_names=False, pad=SPAN):
        """alts: iterable of (contig, start, end) mis-mapping regions (``TREDsRepo.get_alts``)."""
        tid = self.tid(tred.chr)
        if tid < 0:
            raise ValueError("invalid contig `{}`".format(tred.chr))
        alt_rows = []
        for (c, s, e) in alts:
            t = self.tid(c)
            if t >= 0:
                alt_rows.append((t, int(s), int(e)))
        alt_arr = np.ascontiguousarray(np.array(alt_rows, dtype=np.int32).reshape(-1, 3))
        q = LocusQuery(tid=tid, repeat_start=tred.repeat_start, repeat_end=tred.repeat_end, readlen=readlen, pad=pad,
                       pe_window=DNAPE_ELONGATE, flankmatch=FLANKMATCH, span=SPAN, n_alts=len(alt_rows),
                       alts=alt_arr.ctypes.data if len(alt_rows) else None)
        summ = LocusSummary()
        while True:
            c = self._caps
            rbuf = np.empty(c["bases"], dtype=np.int8)
            roff = np.zeros(c["reads"] + 1, dtype=np.int64)
            gl = np.empty(c["pairs"], dtype=np.int32)
            tl = np.empty(c["pairs"], dtype=np.int32)
            names = np.empty(c["names"], dtype=np.uint8) if want_names else None
            rc = self.lib.tredsw_bam_extract_locus(
                self.handle, ctypes.byref(q), rbuf.ctypes.data, len(rbuf), roff.ctypes.data, c["reads"],
                gl.ctypes.data, len(gl), tl.ctypes.data, len(tl), names.ctypes.data if want_names else None,
                len(names) if want_names else 0, ctypes.byref(summ))
            _lib.check(rc, "tredsw_bam_extract_locus")
            if not summ.overflow:
                break
            c["reads"] = max(c["reads"], 2 * summ.nreads)
            c["bases"] = max(c["bases"], 2 * int(summ.nbases))
            c["pairs"] = max(c["pairs"], 2 * max(summ.n_global, summ.n_target))
            c["names"] = max(c["names"], 2 * int(summ.name_bytes))
        ev = LocusEvidence()
        ev.reads = rbuf[:summ.nbases].copy()
        ev.roff = roff[:summ.nreads + 1].copy()
        ev.global_lens = gl[:summ.n_global].copy()
        ev.target_lens = tl[:summ.n_target].copy()
        ev.depth = float(summ.depth)
        ev.n_unmapped = int(summ.n_unmapped)
        ev.names = (bytes(names[:summ.name_bytes]).decode("latin-1").split("\0")[:-1] if want_names else None)
        return ev

    def problem(self, tred, readlen, gender="Unknown", alts=(), depth=None):
        """A (sample, locus) problem for ``cohort.CohortBatch`` straight from the BAM."""
        from .simulate import Problem
        ev = self.extract_locus(tred, readlen, alts=alts)
        pr = Problem()
        pr.tred, pr.readlen = tred, readlen
        pr.ploidy = 1 if (gender == "Male" and tred.is_xlinked) else tred.ploidy
        pr.depth = ev.depth if depth is None else depth
        pr.reads, pr.roff = ev.reads, ev.roff
        pr.global_lens, pr.target_lens = ev.global_lens, ev.target_lens
        pr.alleles, pr.names = None, None
        return pr
