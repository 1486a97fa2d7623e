"""
TEST INFRASTRUCTURE / CPU BASELINE — drives the reference's own code (loaded by ``oracle/refshim.py``) on problems
that do not come from a BAM: synthetic (sample, locus) problems of ``tredparse_b200.simulate``.

The reads go through the reference's unmodified ``BamParser.parse`` (a fake samfile yields them as the
'unmapped, anchored by the mate' reads of the window, bam_parser.py:206-208), the pair lengths through a stub
``PEextractor``, the call through ``IntegratedCaller.call`` — i.e. ``tred.runBam`` (tred.py:153-169) with the
BAM replaced by in-memory reads.  Used by tests/golden/make_ref_fixtures.py (golden vectors) and by bench.py's
``cpu_baseline`` / ``--impl reference`` legs (the reference CPU path timed on the box's host cores).
"""
import os

from . import refshim

_ref = None


def reference():
    global _ref
    if _ref is None:
        _ref = refshim.load()
    return _ref


class StubPE(object):
    """Stands in for bam_parser.PEextractor when the pair lengths are given."""

    def __init__(self, global_lens, target_lens, ref, minpe):
        self.global_lens, self.target_lens, self.ref, self.MINPE = list(global_lens), list(target_lens), ref, minpe


class FakeRead(object):
    __slots__ = ("query_sequence", "query_name", "is_unmapped", "reference_start")

    def __init__(self, seq, name):
        self.query_sequence, self.query_name, self.is_unmapped, self.reference_start = seq, name, True, -1


class FakeSam(object):
    """A samfile whose main-window fetch yields the given reads; alt regions yield nothing."""

    def __init__(self, reads, chr):
        self.reads, self.chr = reads, chr
        self.calls = 0

    def fetch(self, c=None, s=None, e=None):
        self.calls += 1
        if c == self.chr and self.calls == 2:       # call 1 is test_fetch(), call 2 the window, then the alts
            return iter(self.reads)
        return iter(())

    def getrname(self, rid):
        return self.chr


def input_params(ref, repo, bam, tred, READLEN, depth, gender="Unknown", clip=False, alts=True, repeatpairs=True,
                 maxinsert=300, fullsearch=False):
    return ref.utils.InputParams(bam=bam, READLEN=READLEN, tredName=tred, repo=repo, maxinsert=maxinsert,
                                 fullsearch=fullsearch, gender=gender, depth=depth, clip=clip, alts=alts,
                                 repeatpairs=repeatpairs, log="INFO")


def parse_reads(ref, repo, tredname, readlen, ploidy, depth, reads, names, clip=False, repeatpairs=True,
                maxinsert=300, fullsearch=False):
    """The reference's BamParser.parse over in-memory reads -> parsed BamParser."""
    tred = repo[tredname]
    gender = "Male" if (ploidy == 1 and tred.is_xlinked) else "Unknown"
    ip = input_params(ref, repo, "synthetic.bam", tredname, readlen, depth, gender=gender, clip=clip,
                      repeatpairs=repeatpairs, maxinsert=maxinsert, fullsearch=fullsearch)
    bp = ref.bam_parser.BamParser(ip)
    assert bp.ploidy == ploidy, "ploidy is decided by gender and locus (bam_parser.py:58-61)"
    fake = FakeSam([FakeRead(s, n) for s, n in zip(reads, names)], tred.chr)
    saved = ref.bam_parser.read_alignment
    ref.bam_parser.read_alignment = lambda _bam: fake
    try:
        bp.parse()
    finally:
        ref.bam_parser.read_alignment = saved
    return bp


def stub_pe(ref, tred, global_lens, target_lens):
    return StubPE(global_lens, target_lens, tred.repeat_end - tred.repeat_start + 1,
                  tred.repeat_end - tred.repeat_start + 2 * ref.bam_parser.FLANKMATCH + 2)


def call(ref, bp, pe, maxinsert=300, fullsearch=False):
    """The reference's IntegratedCaller on a parsed BamParser with the given PE object -> caller."""
    saved = ref.models.PEextractor
    ref.models.PEextractor = lambda _bp: pe
    try:
        caller = ref.models.IntegratedCaller(bp, maxinsert=maxinsert, fullsearch=fullsearch)
        caller.call()
    finally:
        ref.models.PEextractor = saved
    return caller


_repo = None


def genotype_problem(args):
    """One (sample, locus) unit through the reference's own BamParser + IntegratedCaller.
    args = (tred name, readlen, ploidy, depth, read strings, read names, global_lens, target_lens)
    -> (alleles, CI, PP, label, n reads, forward cells, FDP, PDP, RDP)"""
    global _repo
    ref = reference()
    if _repo is None:
        _repo = ref.meta.TREDsRepo()
    name, readlen, ploidy, depth, reads, names, global_lens, target_lens = args
    bp = parse_reads(ref, _repo, name, readlen, ploidy, depth, reads, names)
    tred = _repo[name]
    caller = call(ref, bp, stub_pe(ref, tred, [int(x) for x in global_lens], [int(x) for x in target_lens]))
    period, P = len(tred.repeat), len(tred.prefix) + len(tred.suffix)
    mu = -(-readlen // period)
    tcells = 2 * sum(P + period * u for u in range(1, mu + 1))
    cells = sum(len(r) for r in reads) * tcells
    return ([int(x) for x in caller.alleles], caller.CI, float(caller.PP), caller.label, len(reads), cells,
            sum(bp.counts["FULL"].values()), sum(bp.counts["PREF"].values()), int(bp.rept))


def usable():
    from . import sw
    return refshim.usable() and os.path.exists(sw.REF_SO)


def run_bam(args):
    """The reference's own ``tred.run`` (tred.py:180-278) on a BAM file through the pysam stand-in
    (oracle/pysam_stub.py over the repo's Python BAM reader) — the from-BAM CPU arm.
    args = (samplekey, bam path, [tred names]) -> tredCalls"""
    import tempfile
    global _repo
    ref = reference()
    if _repo is None:
        _repo = ref.meta.TREDsRepo()
    samplekey, bam, treds = args
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="reftred_"))
    try:
        out = ref.tred.run((samplekey, bam, _repo, list(treds), 300, False, False, True, True, "INFO"))
    finally:
        os.chdir(cwd)
    return out["tredCalls"]
