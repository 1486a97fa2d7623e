"""
TEST INFRASTRUCTURE — the slice of pysam the reference touches, over the repo's pure-Python BAM reader.

``oracle/refshim.py`` redirects the reference's ``import pysam`` (bam_parser.py:22) here.  ``fetch``,
``getrname`` and the read attributes come from ``tredparse_b200.bamio`` (pysam semantics restated there);
this module adds ``pileup`` (bam_parser.py:406: ``[c.n for c in sam.pileup(chr, start, end)]``).

pysam itself (0.9.1 + samtools 1.3.1 in the reference's docker image, docker/tredparse.dockerfile:24)
is not available, so ``pileup`` is a restatement with the plausible readings selectable through
``PILEUP_MODE`` — the candidates the depth question of DESIGN.md §2 enumerates:

  "all"        pysam's default: stepper="all" (skip UNMAP|SECONDARY|QCFAIL|DUP), truncate=False → every
               column covered by a read overlapping [start, end), max_depth 8000   [default]
  "truncate"   the same, columns restricted to [start, end)
  "nofilter"   stepper="nofilter": every read with a CIGAR counts (only UNMAP skipped by the pileup engine)
  "nodel"      "all", but reads are not counted in columns where they show a deletion / reference skip
  "overlap"    "all" with samtools' overlapping-mate removal (a column covered by both mates counts once)
"""
from tredparse_b200 import bamio

PILEUP_MODE = "all"
MAX_DEPTH = 8000

_MASK_ALL = bamio.FUNMAP | bamio.FSECONDARY | bamio.FQCFAIL | bamio.FDUP


class PileupColumn(object):
    __slots__ = ("reference_pos", "pos", "n", "nsegments")

    def __init__(self, pos, n):
        self.reference_pos = self.pos = pos
        self.n = self.nsegments = n


def column_depths(sam, chr, start, end, mode=None):
    """{reference position: depth} over the columns pysam's pileup would visit under ``mode``."""
    mode = mode or PILEUP_MODE
    cols = {}
    seen = {}
    for r in sam.fetch(chr, start, end):
        if not r.cigartuples or r.is_unmapped:
            continue
        if mode != "nofilter" and (r.flag & _MASK_ALL):
            continue
        pos = r.reference_start
        covered = []
        for op, ln in r.cigartuples:
            if op in (0, 7, 8):          # M = X
                covered.extend(range(pos, pos + ln))
                pos += ln
            elif op in (2, 3):           # D N
                if mode != "nodel":
                    covered.extend(range(pos, pos + ln))
                pos += ln
        if mode == "overlap":
            mate = seen.setdefault(r.query_name, set())
            covered = [p for p in covered if p not in mate]
            mate.update(covered)
        for p in covered:
            if mode == "truncate" and not (start <= p < end):
                continue
            cols[p] = min(cols.get(p, 0) + 1, MAX_DEPTH)
    return cols


class AlignmentFile(bamio.AlignmentFile):
    def pileup(self, contig=None, start=None, end=None, **kwargs):
        cols = column_depths(self, contig, start, end)
        for p in sorted(cols):
            yield PileupColumn(p, cols[p])
