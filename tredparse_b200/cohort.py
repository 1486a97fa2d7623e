"""
Cohort path: many (sample, locus) problems through ONE device pipeline call
(``tredsw_genotype_batch``: Smith-Waterman + classification -> tallies -> candidate ranges -> KDE ->
likelihood grid -> call / CI / PP / label), sharded by (sample, locus) over GPUs with no collective.

This is the batched equivalent of looping ``tred.runBam`` (``tredparse/tred.py:153-169,225-249``) and is
what ``bench.py`` times.  ``CohortBatch`` packs problems (from ``simulate`` or from BAM-derived read
sets) into flat buffers; ``run_host`` goes through the C ABI with host buffers (H2D / D2H inside the
call); ``to_device`` + ``run_device`` keep every buffer resident in HBM (torch tensors used purely as
device-memory plumbing) and only enqueue kernels.
"""
import ctypes

import numpy as np

from . import _lib, ssw
from .models import NoiseModel, StepModel

LABELS = {0: "ok", 1: "prerisk", 2: "risk", 3: "missing"}

PROBLEM_DTYPE = np.dtype([("family", "<i4"), ("ploidy", "<i4"), ("n_global", "<i4"), ("n_target", "<i4"),
                          ("off_global", "<i8"), ("off_target", "<i8"), ("depth", "<f8")])
LOCUS_DTYPE = np.dtype([(n, "<i4") for n in ("period", "readlen", "pe_ref", "pe_minpe", "expansion",
                                              "recessive", "cutoff_prerisk", "cutoff_risk")])
CALL_DTYPE = np.dtype([("allele1", "<i4"), ("allele2", "<i4"), ("ci", "<i4", 4), ("label", "<i4"),
                       ("n_points", "<i4"), ("fdp", "<i4"), ("pdp", "<i4"), ("rdp", "<i4"),
                       ("run_pe", "<i4"), ("pp", "<f8"), ("lik", "<f8")])


class Cohort(ctypes.Structure):
    """tredsw_cohort (include/tredsw.h)"""
    _fields_ = [("rbuf", ctypes.c_void_p), ("roff", ctypes.c_void_p), ("read_problem", ctypes.c_void_p),
                ("problems", ctypes.c_void_p), ("pe_lens", ctypes.c_void_p), ("n_pe_lens", ctypes.c_int64),
                ("nreads", ctypes.c_int32), ("nproblems", ctypes.c_int32), ("max_read_len", ctypes.c_int32),
                ("nfamilies", ctypes.c_int32), ("families", ctypes.c_void_p), ("loci", ctypes.c_void_p),
                ("step_pmf", ctypes.c_void_p), ("stutter_w", ctypes.c_double * 5), ("gc", ctypes.c_double),
                ("score", ctypes.c_double), ("maxinsert", ctypes.c_int32), ("fullsearch", ctypes.c_int32),
                ("mat25", ctypes.c_int8 * 25), ("pad_", ctypes.c_int8 * 3), ("gap_open", ctypes.c_int32),
                ("gap_extend", ctypes.c_int32), ("input_flags", ctypes.c_uint32), ("norepeatpairs", ctypes.c_int32),
                ("n_bases", ctypes.c_int64), ("read_name", ctypes.c_void_p)]


IN_READS_PACKED4, IN_PE_LENS_I16 = 1, 2          # tredsw_cohort.input_flags


assert PROBLEM_DTYPE.itemsize == 40 and LOCUS_DTYPE.itemsize == 32 and CALL_DTYPE.itemsize == 64


def _bind(lib):
    if not getattr(lib, "_cohort_bound", False):
        lib.tredsw_genotype_batch.restype = ctypes.c_int
        lib.tredsw_genotype_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(Cohort), ctypes.c_uint32,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int32, ctypes.c_void_p]
        lib.tredsw_genotype_batch_ex.restype = ctypes.c_int
        lib.tredsw_genotype_batch_ex.argtypes = lib.tredsw_genotype_batch.argtypes + [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        lib._cohort_bound = True


class CohortBatch:
    """Flat buffers of a list of problems (objects with tred, readlen, ploidy, depth, reads, roff,
    global_lens, target_lens — see simulate.Problem)."""

    def __init__(self, problems, maxinsert=300, fullsearch=False, score=1.0, gc=.68, match=1, mismatch=5,
                 gap_open=7, gap_extend=2, clip=False, repeatpairs=True, family_keys=None):
        """clip / repeatpairs: --useclippedreads / not --norepeatpairs (tred.py:75-81; note that InputParams
        defaults repeatpairs to False while the CLI passes True).  Without repeatpairs every problem needs
        ``names`` (read names or ids: mates share one)."""
        self.maxinsert, self.fullsearch = maxinsert, fullsearch
        self.clip, self.repeatpairs = bool(clip), bool(repeatpairs)
        self.score, self.gc = score, gc
        self.match, self.mismatch, self.gap_open, self.gap_extend = match, mismatch, gap_open, gap_extend
        fam_index, fams, loci, step_rows = {}, [], [], []
        step = StepModel()

        def add_family(t, readlen):
            period = len(t.repeat)
            fam_index[(t.name, readlen)] = len(fams)
            fams.append(ssw.make_family(t.prefix, t.repeat, t.suffix, -(-readlen // period), clip=self.clip))
            L = np.zeros(1, dtype=LOCUS_DTYPE)
            ref = t.repeat_end - t.repeat_start + 1
            L["period"], L["readlen"], L["pe_ref"], L["pe_minpe"] = period, readlen, ref, ref - 1 + 2 * 9 + 2
            L["expansion"], L["recessive"] = int(t.is_expansion), int(t.is_recessive)
            L["cutoff_prerisk"], L["cutoff_risk"] = int(t.cutoff_prerisk), int(t.cutoff_risk)
            loci.append(L)
            step_rows.append(step.step_size_by_period[period])

        # family_keys: [(tred, readlen)] — a fixed family table for every batch of a cohort (otherwise families are
        # numbered in order of first appearance in this batch)
        for t, readlen in (family_keys or []):
            add_family(t, readlen)
        n = len(problems)
        P = np.zeros(n, dtype=PROBLEM_DTYPE)
        rbufs, roffs, rprob, pe, rname = [], [np.zeros(1, dtype=np.int64)], [], [], []
        rbase, pebase, max_len = 0, 0, 1
        for i, pr in enumerate(problems):
            t = pr.tred
            key = (t.name, pr.readlen)
            if key not in fam_index:
                add_family(t, pr.readlen)
            f = fam_index[key]
            reads = np.asarray(pr.reads, dtype=np.int8)
            roff = np.asarray(pr.roff, dtype=np.int64)
            nr = len(roff) - 1
            rbufs.append(reads)
            roffs.append(roff[1:] + rbase)
            rbase += int(roff[-1])
            rprob.append(np.full(nr, i, dtype=np.int32))
            if not self.repeatpairs and not self.clip:
                names = getattr(pr, "names", None)
                if names is None or len(names) != nr:
                    raise ValueError("repeatpairs=False needs the read names of every problem")
                ids = {}
                rname.append(np.array([ids.setdefault(x, len(ids)) for x in
                                       (names.tolist() if hasattr(names, "tolist") else names)], dtype=np.int32))
            if nr:
                max_len = max(max_len, int(np.max(np.diff(roff))))
            g = np.asarray(pr.global_lens, dtype=np.int32)
            tl = np.asarray(pr.target_lens, dtype=np.int32)
            P[i] = (f, pr.ploidy, len(g), len(tl), pebase, pebase + len(g), pr.depth)
            pe += [g, tl]
            pebase += len(g) + len(tl)
        self.nproblems = n
        self.problems = P
        self.rbuf = np.ascontiguousarray(np.concatenate(rbufs) if rbufs else np.zeros(0, np.int8), dtype=np.int8)
        self.roff = np.ascontiguousarray(np.concatenate(roffs), dtype=np.int64)
        self.read_problem = np.ascontiguousarray(np.concatenate(rprob) if rprob else np.zeros(0, np.int32), dtype=np.int32)
        self.pe_lens = np.ascontiguousarray(np.concatenate(pe) if pe else np.zeros(0, np.int32), dtype=np.int32)
        self.read_name = np.ascontiguousarray(np.concatenate(rname), dtype=np.int32) if rname else None
        self.nreads = len(self.roff) - 1
        self.max_read_len = max_len
        self.families = np.ascontiguousarray(np.concatenate(fams))
        self.loci = np.ascontiguousarray(np.concatenate(loci))
        self.step_pmf = np.ascontiguousarray(np.stack(step_rows), dtype=np.float64)
        self.hist_units = int(self.families["max_units"].max())
        self.objects = problems
        self._dev = None
        self._packed = None

    def pack_inputs(self, out=None, threads=4, keep=True):
        """Compact transfer formats (TREDSW_IN_READS_PACKED4 | TREDSW_IN_PE_LENS_I16): two base codes per byte
        and int16 pair lengths — half the host->device bytes; the library expands them on the device.
        Native host code (tredsw_pack_reads4 / tredsw_narrow_i16; ctypes releases the GIL): the per-batch cost of a
        cohort pipeline between ingest and the copy.  `out`: optional dict of preallocated (e.g. pinned) arrays
        {"rbuf": uint8[(n+7)//8*4...], "pe_lens": int16[n]} to pack into; run_host(packed=True) sends them."""
        lib = _lib.load()
        n = len(self.rbuf)
        nbytes = ((n + 7) // 8) * 4                      # whole 32-bit words of 8 bases
        if out is None:
            out = {"rbuf": np.zeros(nbytes, dtype=np.uint8), "pe_lens": np.zeros(len(self.pe_lens), dtype=np.int16)}
        if len(out["rbuf"]) < nbytes or len(out["pe_lens"]) < len(self.pe_lens):
            raise ValueError("pack_inputs: output buffers too small")
        _lib.check(lib.tredsw_pack_reads4(_lib.ptr(self.rbuf), n, _lib.ptr(out["rbuf"]), int(threads)), "tredsw_pack_reads4")
        _lib.check(lib.tredsw_narrow_i16(_lib.ptr(self.pe_lens), len(self.pe_lens), _lib.ptr(out["pe_lens"]), int(threads)),
                   "tredsw_narrow_i16")
        if keep:
            self._packed = out
            return self
        return out                                      # (several threads may pack one batch into different buffers)

    def pack_reads4(self, out, threads=4):
        """Only the reads (for producers that already write int16 pair lengths): base codes -> `out`, 4 bit/base."""
        n = len(self.rbuf)
        if len(out) < ((n + 7) // 8) * 4:
            raise ValueError("pack_reads4: output buffer too small")
        _lib.check(_lib.load().tredsw_pack_reads4(_lib.ptr(self.rbuf), n, _lib.ptr(out), int(threads)), "tredsw_pack_reads4")
        return out

    # ---- descriptor -----------------------------------------------------------------------------------
    def _descriptor(self, rbuf, roff, rprob, problems, pe_lens, input_flags=0, read_name=None):
        c = Cohort()
        c.input_flags, c.n_bases = input_flags, int(len(self.rbuf))
        c.rbuf, c.roff, c.read_problem, c.problems, c.pe_lens = rbuf, roff, rprob, problems, pe_lens
        c.n_pe_lens, c.nreads, c.nproblems = len(self.pe_lens), self.nreads, self.nproblems
        c.max_read_len, c.nfamilies = self.max_read_len, len(self.families)
        c.families, c.loci, c.step_pmf = self.families.ctypes.data, self.loci.ctypes.data, self.step_pmf.ctypes.data
        for i, w in enumerate(NoiseModel().weights):
            c.stutter_w[i] = w
        c.gc, c.score = self.gc, self.score
        c.maxinsert, c.fullsearch = self.maxinsert, int(self.fullsearch)
        mat = ssw.score_matrix(self.match, self.mismatch).ravel()
        for i in range(25):
            c.mat25[i] = int(mat[i])
        c.gap_open, c.gap_extend = self.gap_open, self.gap_extend
        if self.read_name is not None:
            c.norepeatpairs = 1
            c.read_name = read_name if read_name is not None else self.read_name.ctypes.data
        return c

    # ---- host buffers through the C ABI (H2D + kernels + D2H inside the call) -------------------------
    def run_host(self, ctx=None, want_reads=False, want_hist=False, want_stats=False, packed=False, want_post=False,
                 device_view=None):
        """want_post: also return the sparsified posteriors (``posteriors()`` turns them into the reference's
        P_h1 / P_h2 / P_h1h2 dicts)."""
        ctx = ctx or _lib.default_context()
        _bind(ctx.lib)
        calls = np.zeros(self.nproblems, dtype=CALL_DTYPE)
        read_out = np.zeros((self.nreads, 8), dtype=np.int32) if want_reads else None
        hist = np.zeros((self.nproblems, 3, self.hist_units + 1), dtype=np.int32) if want_hist else None
        stats = np.zeros(8, dtype=np.int64) if want_stats else None
        if packed:
            if isinstance(packed, dict):
                bufs = packed                                # buffers filled by pack_inputs(out=..., keep=False)
            else:
                if self._packed is None:
                    self.pack_inputs()
                bufs = self._packed
            # (a dict may carry only one of the two compact formats: e.g. int16 pair lengths, reads one byte per base)
            rbuf, pe = bufs.get("rbuf", self.rbuf), bufs.get("pe_lens", self.pe_lens)
            flags_in = (IN_READS_PACKED4 if "rbuf" in bufs else 0) | (IN_PE_LENS_I16 if "pe_lens" in bufs else 0)
        else:
            rbuf, pe, flags_in = self.rbuf, self.pe_lens, 0
        call_flags = 0
        if device_view is not None:
            # the bulk evidence is an ingest.IngestBatch's device buffers (TREDSW_DEVICE_INPUTS)
            v, call_flags = device_view, _lib.DEVICE_INPUTS
            c = self._descriptor(v.d_rbuf, v.d_roff, v.d_read_problem, self.problems.ctypes.data, v.d_pe_lens, 0)
        else:
            c = self._descriptor(rbuf.ctypes.data, self.roff.ctypes.data, self.read_problem.ctypes.data,
                                 self.problems.ctypes.data, pe.ctypes.data, flags_in)
        post_cap = max(4096, 64 * self.nproblems) if want_post else 0
        n_post = np.zeros(1, dtype=np.int64)
        for attempt in range(4):
            post = np.zeros(post_cap, dtype=_lib.POSTERIOR_DTYPE) if want_post else None
            rc = ctx.lib.tredsw_genotype_batch_ex(ctx.handle, ctypes.byref(c), call_flags, _lib.ptr(calls), _lib.ptr(read_out),
                                                  _lib.ptr(hist), self.hist_units, _lib.ptr(stats), _lib.ptr(post),
                                                  post_cap, _lib.ptr(n_post) if want_post else None)
            if rc == 0 and want_post and int(n_post[0]) > post_cap:
                post_cap = int(n_post[0]) + 1024          # the entry list was truncated: once more with room for it
                continue
            if rc == 0 or "arena overflow" not in _lib.last_error():
                break
        _lib.check(rc, "tredsw_genotype_batch")
        n_rbuf = (len(self.rbuf) + 1) // 2 if (flags_in & IN_READS_PACKED4) else self.rbuf.nbytes   # bytes the call copies
        n_pe = len(self.pe_lens) * (2 if (flags_in & IN_PE_LENS_I16) else 4)
        bulk = 0 if device_view is not None else n_rbuf + self.roff.nbytes + self.read_problem.nbytes + n_pe
        self.h2d_bytes = bulk + self.problems.nbytes + self.families.nbytes + self.loci.nbytes + self.step_pmf.nbytes
        self.d2h_bytes = calls.nbytes + (read_out.nbytes if want_reads else 0) + (hist.nbytes if want_hist else 0)
        out = {"calls": calls}
        if want_reads:
            out["reads"] = read_out
        if want_hist:
            out["hist"] = hist
        if want_stats:
            out["stats"] = stats
        if want_post:
            out["post"] = post[:int(n_post[0])]
        return out

    # ---- device-resident buffers (torch = memory + stream plumbing only) ------------------------------
    def to_device(self, device):
        import torch
        dev = torch.device("cuda", device)
        t = lambda a, dt: torch.from_numpy(a.view(dt) if a.dtype.fields else a).to(dev)
        self._dev = {
            "rbuf": t(self.rbuf, np.int8), "roff": t(self.roff, np.int64), "rprob": t(self.read_problem, np.int32),
            "problems": torch.from_numpy(self.problems.view(np.uint8)).to(dev),
            "pe_lens": t(self.pe_lens, np.int32),
            "calls": torch.zeros(self.nproblems * CALL_DTYPE.itemsize, dtype=torch.uint8, device=dev),
        }
        if self.read_name is not None:
            self._dev["read_name"] = t(self.read_name, np.int32)
        return self

    def run_device(self, ctx):
        """Enqueue the pipeline on ctx's stream; results stay on the device (``calls_from_device``)."""
        _bind(ctx.lib)
        d = self._dev
        c = self._descriptor(d["rbuf"].data_ptr(), d["roff"].data_ptr(), d["rprob"].data_ptr(),
                             d["problems"].data_ptr(), d["pe_lens"].data_ptr(),
                             read_name=d["read_name"].data_ptr() if "read_name" in d else None)
        rc = ctx.lib.tredsw_genotype_batch(ctx.handle, ctypes.byref(c), _lib.DEVICE_PTRS, d["calls"].data_ptr(),
                                           None, None, 0, None)
        _lib.check(rc, "tredsw_genotype_batch")

    # ---- evidence that is already in device memory (the GPU ingest's buffers) ------------------------------
    @classmethod
    def from_ingest(cls, ing, family_of, ploidy, depth, family_keys, names=None, **kw):
        """A batch over the flat buffers an ``ingest.IngestBatch`` left in HBM (problem i of the batch = problem i of
        the ingest): nothing is concatenated or copied, only the 40-byte problem records are built.
        family_of: index into family_keys per problem; names: per problem read names (only without repeatpairs)."""
        self = cls([], family_keys=family_keys, **kw)
        n = ing.nproblems
        summ, span = ing.summary_table, ing.span_table
        P = np.zeros(n, dtype=PROBLEM_DTYPE)
        P["family"], P["ploidy"], P["depth"] = family_of, ploidy, depth
        P["n_global"], P["n_target"] = summ["n_global"], summ["n_target"]
        P["off_global"], P["off_target"] = span["off_global"], span["off_target"]
        self.nproblems, self.problems = n, P
        self.rbuf, self.roff, self.pe_lens = ing.h_rbuf, ing.h_roff, ing.h_pe_lens        # (host copies: sizes only)
        self.read_problem = None
        self.nreads = ing.nreads
        self.max_read_len = max(1, int(np.diff(ing.h_roff).max())) if ing.nreads else 1
        self.read_name = None
        if not self.repeatpairs and not self.clip:
            if names is None:
                raise ValueError("repeatpairs=False needs the read names of every problem")
            ids = []
            for nm in names:
                seen = {}
                ids.append(np.array([seen.setdefault(x, len(seen)) for x in nm], dtype=np.int32))
            self.read_name = np.ascontiguousarray(np.concatenate(ids) if ids else np.zeros(0, np.int32), dtype=np.int32)
        self._ingest = ing
        return self

    def run_ingest(self, ctx, want_reads=False, want_hist=False, want_post=False):
        """``tredsw_genotype_batch_ex`` with TREDSW_DEVICE_INPUTS on the ingest's device buffers: the evidence stays
        where the ingest left it, the problem records and the outputs are host buffers.  Same dict as ``run_host``."""
        return self.run_host(ctx=ctx, want_reads=want_reads, want_hist=want_hist, want_post=want_post,
                             device_view=self._ingest.view)

    def calls_from_device(self):
        calls = self._dev["calls"].cpu().numpy().view(CALL_DTYPE)
        if len(calls) and int(calls["n_points"].min()) < 0:
            # a likelihood arena was too small for some problem of this batch (the library has grown it on the
            # way when it could see the counters); the caller re-runs the batch
            raise _lib.TredswError("likelihood arena overflow in a device-resident batch; run it again")
        return calls


class HostPipeline:
    """Keeps `depth` host-buffer calls in flight on one GPU: every slot owns a context (its own stream and
    staging buffers) and a host thread, so the H2D copy of one batch overlaps the kernels of the previous
    one and the tail of one step is filled by the head of the next.  This is the cohort-streaming pattern
    (read BAM windows of batch k+1 while batch k is on the GPU); ctypes releases the GIL during the call.

        with HostPipeline(device=0, depth=2) as pipe:
            for out in pipe.map(batches):        # results in submission order
                ...
    """

    def __init__(self, device=0, depth=2):
        from concurrent.futures import ThreadPoolExecutor
        import queue
        self.contexts = [_lib.Context(device) for _ in range(depth)]
        self._free = queue.Queue()
        for c in self.contexts:
            self._free.put(c)
        self._pool = ThreadPoolExecutor(max_workers=depth)

    def _run(self, batch, kwargs):
        ctx = self._free.get()
        try:
            return batch.run_host(ctx=ctx, **kwargs)
        finally:
            self._free.put(ctx)

    def submit(self, batch, **kwargs):
        return self._pool.submit(self._run, batch, kwargs)

    def map(self, batches, **kwargs):
        futures = [self.submit(b, **kwargs) for b in batches]
        for f in futures:
            yield f.result()

    @property
    def launches(self):
        return sum(c.launches for c in self.contexts)

    def close(self):
        self._pool.shutdown(wait=True)
        for c in self.contexts:
            c.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def posteriors(post, nproblems):
    """Entry list of tredsw_genotype_batch_ex -> per problem {"P_h1": {...}, "P_h2": {...}, "P_h1h2": {...}} with the
    reference's keys ("15", "15,41") and values (models.py:304-317)."""
    out = [{"P_h1": {}, "P_h2": {}, "P_h1h2": {}} for _ in range(nproblems)]
    if not len(post):
        return out
    kinds, probs = post["kind"].tolist(), post["problem"].tolist()
    a, b, p = post["a"].tolist(), post["b"].tolist(), post["p"].tolist()
    totals = {pr: v for k, pr, v in zip(kinds, probs, p) if k == _lib.POST_JOINT_TOTAL}
    for k, pr, x, y, v in zip(kinds, probs, a, b, p):
        if k == _lib.POST_H1:
            out[pr]["P_h1"][str(x)] = v
        elif k == _lib.POST_H2:
            out[pr]["P_h2"][str(x)] = v
        elif k == _lib.POST_JOINT:
            out[pr]["P_h1h2"]["{},{}".format(x, y)] = v / totals[pr]
    return out


def decode_call(call, period=None):
    """tredsw_call record -> dict with the reference's field names."""
    a1, a2, ci, label, n_points, fdp, pdp, rdp, run_pe, pp, lik = call.tolist() if hasattr(call, "tolist") else call
    missing = a1 < 0
    return {"alleles": [a1, a2],
            "CI": "" if missing else "{}-{}|{}-{}".format(*ci),
            "PP": -1 if missing else pp, "label": LABELS.get(label, "error"),
            "FDP": fdp, "PDP": pdp, "RDP": rdp, "lik": lik, "n_points": n_points, "run_pe": bool(run_pe)}


def shard(n_items, rank, world):
    """Static round-robin partition of (sample, locus) problems over GPUs (no collective); see dist.py."""
    from .dist import shard_indices
    return shard_indices(n_items, rank, world)
