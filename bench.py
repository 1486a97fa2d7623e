#!/usr/bin/env python
"""
bench.py — the genotyping hot path on a synthetic cohort (BASELINE.json configs[3]: samples x 30 TREDs,
sharded by (sample, locus), one process per GPU, no collective on the data path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl reference]

A *step* is one pass of the whole hot path — Smith-Waterman of every read against its locus' template
family + classification, tallies, candidate ranges, KDE, likelihood grid, call / CI / PP / label — over one
fixed batch of S samples x 30 loci per GPU (weak scaling: the per-GPU batch is constant as N grows).

Rank 0 prints ONE JSON line:
  value        loci genotyped / s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H per step
  roofline     dominant kernel = sw_family classify kernels (integer/DPX-pipe bound);
               roofline_grid = likelihood grid vs the measured HBM copy bandwidth
  cpu_baseline the reference's CPU path (ssw.c compiled unmodified + per-call ctypes pattern + numpy/scipy
               grid) on a bounded sample of the same workload, all host cores

`--impl reference` times only that CPU path (rank 0 alone when launched under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loci genotyped/sec"
UNIT = "loci/s"
READLEN = 150


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--samples", type=int, default=384, help="samples per GPU per step (x 30 loci)")
    p.add_argument("--no-grid-stress", action="store_true", help="skip the long-expansion grid measurement "
                   "(BASELINE configs[4]) reported as roofline_grid_stress")
    p.add_argument("--depth", type=int, default=2, help="host-buffer calls kept in flight by the e2e pipeline")
    p.add_argument("--streams", type=int, default=3, help="streams the device-resident steps alternate over "
                   "(2: the tail of one step's persistent SW kernel overlaps the head of the next step)")
    p.add_argument("--impl", default="tredsw", choices=("tredsw", "reference"))
    p.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def distinct_loci(repo):
    """30 distinct locations among the 32 catalogue names (FXS == FXTAS, SBMA == AR coordinates)."""
    seen, names = set(), []
    for n in repo.names:
        t = repo[n]
        key = (t.chr, t.repeat_start, t.repeat_end)
        if key in seen:
            continue
        seen.add(key)
        names.append(n)
    return names


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (every 100 ms: faster polling
    contends for the driver and measurably slows the host-buffer calls of the e2e leg)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def model_tables():
    with open(os.path.join(ROOT, "tredparse_b200", "data", "models.json")) as fp:
        md = json.load(fp)
    step = {int(k): np.array(v) for k, v in md["step_size_by_period"].items()}
    for i in range(6, 18):
        step[i] = step[6]
    return step, md["stutter_weights"]


def cpu_sample_size(args, cores, nloci, passes=1):
    """Bounded CPU sample: whole samples (30 loci each) worth about 12 s of wall time per pass at the
    ~1.5 loci/s/core the reference path reaches (unmodified ssw.c behind the per-call ctypes pattern +
    numpy/scipy grid) — less per pass when many passes are requested, so that the run ends within minutes."""
    if args.cpu_sample:
        return max(1, args.cpu_sample // nloci)
    seconds = min(12.0, 150.0 / max(1, passes))
    return max(1, min(16, int(round(seconds * 1.5 * cores / nloci))))


def cpu_reference_run(problems, cores=None):
    """Time the reference-shaped CPU path on `problems`; returns (loci/s, seconds, cores, reads, cells, kind)."""
    from oracle import ref_percall, sw
    kind = "reference" if sw.ref_available() else "port"
    if kind == "port":
        raise RuntimeError("oracle/_ref/libssw_ref.so missing (build it where /root/reference is mounted)")
    step, w = model_tables()
    tasks = [(p.tred, p.readlen, p.ploidy, p.depth, p.read_strings(), p.global_lens, p.target_lens, step, w)
             for p in problems]
    t0 = time.perf_counter()
    res, used = ref_percall.run_pool(tasks, cores)
    dt = time.perf_counter() - t0
    reads = sum(r[4] for r in res)
    cells = sum(r[5] for r in res)
    return len(problems) / dt, dt, used, reads, cells, kind, res


def ssw_c_loop_ceiling(problems, max_pairs=60000):
    """SURVEY 8(d), CPU baseline (2): the reference's own ssw.c in a plain C loop (ssw_init -> ssw_align(flag=1) ->
    destroy per pair, oracle/ref_batch.c), one core, no Python in the loop — the ceiling of any CPU path built
    on ssw.c.  Returns microseconds per alignment and forward-matrix GCUPS per core."""
    from oracle import sw, evidence_oracle as evo
    queries, templates, qidx, tidx = [], [], [], []
    for p in problems:
        t = p.tred
        mu = -(-p.readlen // len(t.repeat))
        db = [x for _, x in evo.template_family(t.prefix, t.repeat, t.suffix, mu)]
        t0 = len(templates)
        templates += db
        for r in p.read_strings()[:8]:
            q0 = len(queries)
            queries.append(r)
            qidx += [q0] * len(db)
            tidx += list(range(t0, t0 + len(db)))
        if len(qidx) >= max_pairs:
            break
    qidx, tidx = np.array(qidx, dtype=np.int32), np.array(tidx, dtype=np.int32)
    sw.ref_align_pairs(queries[:1], templates[:1], qidx[:1] * 0, tidx[:1] * 0)       # load / warm
    t0 = time.perf_counter()
    sw.ref_align_pairs(queries, templates, qidx, tidx)
    dt = time.perf_counter() - t0
    cells = sum(len(queries[a]) * len(templates[b]) for a, b in zip(qidx, tidx))
    return {"pairs": int(len(qidx)), "us_per_alignment": 1e6 * dt / len(qidx), "gcups_per_core": cells / dt / 1e9}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path, rank 0 only."""
    if rank != 0:
        return
    from tredparse_b200 import simulate
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    names = distinct_loci(repo)
    cores = os.cpu_count() or 1
    nsamp = cpu_sample_size(args, cores, len(names), passes=args.warmup + args.steps)
    problems = simulate.simulate_cohort(repo, names, nsamp, readlen=READLEN)
    times = []
    for it in range(args.warmup + args.steps):
        v, dt, used, reads, cells, kind, _ = cpu_reference_run(problems, cores)
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(problems) * len(times) / total
    sample = "{} samples x {} loci = {} problems ({} reads) per step".format(nsamp, len(names), len(problems), reads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/f64",
        "data": "synthetic",
        "config": {"workload": "synthetic cohort x 30 TREDs (BASELINE configs[3]), bounded sample: " + sample,
                   "readlen": READLEN, "maxinsert": 300, "parallelism": "multiprocessing.Pool({})".format(used)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample,
                         "reads_per_s": reads * len(times) / total,
                         "sw_gcups": cells * len(times) / total / 1e9},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner, ...) are sent
    # to stderr for the whole run, the line itself goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
        for ln in _LINES:
            print(ln, flush=True)


_LINES = []


def emit(line):
    _LINES.append(json.dumps(line))


def _main(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from tredparse_b200 import _lib, cohort, simulate, dist as tdist
    from tredparse_b200.meta import TREDsRepo

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (tredparse_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        tdist.init("nccl", device_id=torch.device("cuda", local_rank))

    repo = TREDsRepo()
    names = distinct_loci(repo)
    # this rank's shard of the (sample, locus) list: samples [rank*S, (rank+1)*S) x all loci
    t_gen = time.perf_counter()
    problems = simulate.simulate_cohort(repo, names, args.samples, readlen=READLEN,
                                        seed=20240000 + rank * args.samples)
    batch = cohort.CohortBatch(problems, maxinsert=300, fullsearch=False)
    t_gen = time.perf_counter() - t_gen

    stream = torch.cuda.Stream(device=local_rank)
    ctx = _lib.Context(local_rank, stream=stream.cuda_stream)
    int_peak = ctx.int_pipe_peak()
    batch.to_device(local_rank)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            batch.run_device(ctx)
    stream.synchronize()
    # stage timing of one extra (untimed) step, used for the rooflines
    ctx.enable_timing(True)
    with torch.cuda.stream(stream):
        batch.run_device(ctx)
    stage = ctx.timing()
    ctx.enable_timing(False)
    st = batch.run_host(ctx=ctx, want_stats=True)["stats"]      # cell / point counts of this batch

    # the timed steps alternate over `--streams` contexts (each with its own stream and scratch; inputs are
    # shared, read-only): consecutive steps are independent batches of a cohort, so nothing orders them
    lanes = [(ctx, stream)]
    for _ in range(max(1, args.streams) - 1):
        s2 = torch.cuda.Stream(device=local_rank)
        lanes.append((_lib.Context(local_rank, stream=s2.cuda_stream), s2))
    for c2, s2 in lanes[1:]:
        for _ in range(2):
            batch.run_device(c2)
        s2.synchronize()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = sum(c.launches for c, _ in lanes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _, s2 in lanes[1:]:
        s2.wait_event(e0)
    for i in range(args.steps):
        batch.run_device(lanes[i % len(lanes)][0])
    for _, s2 in lanes[1:]:
        ev = torch.cuda.Event()
        ev.record(s2)
        stream.wait_event(ev)
    e1.record(stream)
    stream.synchronize()
    barrier()
    launches = sum(c.launches for c, _ in lanes) - launches0
    ms = e0.elapsed_time(e1)
    calls_dev = batch.calls_from_device()

    # ---- end to end through the C ABI with pinned host buffers -------------------------------------
    # Every step copies its inputs H2D from pinned memory and its calls D2H inside the timed region;
    # `depth` calls are kept in flight (cohort streaming: the copy of batch k+1 overlaps the kernels of k).
    def pin(a):
        t = torch.from_numpy(a.view(np.uint8) if a.dtype.fields else a).pin_memory()
        return t.numpy().view(a.dtype) if a.dtype.fields else t.numpy()
    # compact transfer formats (two base codes per byte, int16 pair lengths: half the bytes), packed once per
    # batch like the base encoding itself; the library expands them on the device
    batch.pack_inputs()
    for name in ("roff", "read_problem", "problems"):
        setattr(batch, name, pin(getattr(batch, name)))
    batch._packed = {k: pin(v) for k, v in batch._packed.items()}
    pipe = cohort.HostPipeline(local_rank, depth=max(1, args.depth))
    for host in pipe.map([batch] * (2 * max(1, args.depth)), packed=True):
        pass
    barrier()
    e2e_launch0 = pipe.launches
    t0 = time.perf_counter()
    for host in pipe.map([batch] * args.steps, packed=True):
        pass
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    e2e_launches = pipe.launches - e2e_launch0
    e2e_kernel_ms = None
    if os.environ.get("TREDSW_E2E_TIMING"):
        for c in pipe.contexts:
            c.enable_timing(True)
        for host in pipe.map([batch] * (2 * max(1, args.depth)), packed=True):
            pass
        e2e_kernel_ms = [c.timing() for c in pipe.contexts]
        sys.stderr.write("e2e per-call device stage times: {}\n".format(e2e_kernel_ms))
    pipe.close()
    assert host["calls"].tobytes() == calls_dev.tobytes(), "device-resident and host paths disagree"

    # ---- reduce over ranks: timings MAX, unit counters SUM (no data-path collective anywhere) --------
    (ms, t_e2e_ms), cnt = tdist.reduce_max_sum(
        [ms, t_e2e * 1e3],
        [batch.nproblems, batch.nreads, int(st[0]), int(st[1]), int(st[2]), int(st[4]), launches],
        device="cuda" if world > 1 else "cpu")
    t_e2e = t_e2e_ms / 1e3
    nprob, nreads, alg_cells, ex1, ex2, points, launches = cnt

    if rank == 0:
        sec = ms / 1e3
        value = nprob * args.steps / sec
        # roofline of the dominant kernel (this rank's launch): integer (ALU) pipe.
        # executed ALU lane-instructions = executed DP cells x ALU_PER_CELL.  ALU_PER_CELL is calibrated on the
        # committed ncu capture of this kernel and workload (profiles/r1_classify_full.txt:
        # sm__inst_executed_pipe_alu 75.0 % of 14.84 M active cycles x 64 lanes x 148 SMs = 1.055e11 lane-instr
        # over 3.446e10 executed cells): 6.1 ALU instructions per packed cell pair (4 for the cell, 0.5 running
        # max, 0.67 suffix hooks, byte (un)packing, phase 2) -> 3.06 per cell.
        ALU_PER_CELL, ALU_PER_SCALAR_CELL = 3.06, 3.06
        sw_s = stage["sw"] / 1e3
        lane_instr = int(st[1]) * ALU_PER_CELL + int(st[2]) * ALU_PER_SCALAR_CELL
        achieved = lane_instr / sw_s / 1e9
        hbm, hbm_src = peaks()
        grid_bytes = 8.0 * int(st[4]) * 2            # surface written once, read once by the reduction
        grid_s = max(stage["grid"], 1e-6) / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16 (SW, DPX s16x2) / f64 (likelihood)",
            "data": "synthetic",
            "config": {"workload": "synthetic cohort x 30 TREDs (BASELINE configs[3]): {} samples x {} loci = {} "
                                   "problems, {} reads per GPU per step".format(args.samples, len(names), batch.nproblems, batch.nreads),
                       "readlen": READLEN, "maxinsert": 300, "fullsearch": False,
                       "parallelism": "(sample, locus) shards over {} GPU(s), no collective; {} stream(s) per GPU".format(world, len(lanes)),
                       "l2": "inputs + scratch per step ({:.0f} MB) exceed the 126 MB L2".format(
                           (batch.rbuf.nbytes + batch.pe_lens.nbytes + 8 * batch.nproblems * 1000) / 1e6)},
            "reads_per_s": nreads * args.steps / sec,
            "sw_gcups_algorithmic": alg_cells * args.steps / sec / 1e9,
            "sw_gcups_executed": (ex1 + ex2) * args.steps / sec / 1e9,
            "e2e": {"value": nprob * args.steps / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(batch.h2d_bytes) * world, "d2h_bytes_per_step": int(batch.d2h_bytes) * world,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "calls_in_flight": max(1, args.depth),
                    "transfer_format": "reads 4 bit/base, pair lengths int16 (expanded on the device)"},
            "gpu_launches": int(launches) + int(e2e_launches),
            "clocks": clocks,
            "roofline": {"kernel": "classify_kernel<P> (sw_family.cu), all period instantiations of one step",
                         "bound": "int", "achieved": achieved, "peak": int_peak, "unit": "G lane-instr/s",
                         "frac": achieved / int_peak,
                         # dram__bytes_read + dram__bytes_write of one launch at the default workload (ncu capture
                         # in profiles/): scratch write-back, ~8 % of HBM bandwidth; algorithmic input is 0.17 GB
                         "traffic": 6.53e9 if (args.samples == 384 and world == 1) else None,
                         "peak_source": "tredsw_int_pipe_peak (VIADDMNMX.S16x2 register loop, 64 lanes/clk/SM) measured in this run; "
                                        "achieved = executed DP cells x 3.06 ALU instr/cell (calibrated on the ncu capture in profiles/)",
                         "algorithmic_int_ops_per_s": 10.0 * int(st[0]) / sw_s,
                         "kernel_ms": stage["sw"], "share_of_step": stage["sw"] / max(stage["total"], 1e-9),
                         "executed_gcups": (int(st[1]) + int(st[2])) / sw_s / 1e9,
                         "algorithmic_gcups": int(st[0]) / sw_s / 1e9},
            "roofline_grid": {"kernel": "grid_tiles/setup/fill + grid_surface_{points,tiles} + grid_reduce_{warp,<1>,<8>} (default search)", "bound": "hbm",
                              "achieved": grid_bytes / grid_s / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": grid_bytes / grid_s / 1e9 / hbm, "peak_source": hbm_src + " copy bandwidth",
                              "kernel_ms": stage["grid"], "points": int(st[4]),
                              "note": "FP64-log bound, not HBM bound: ~30-150 logs per 8-byte point (SURVEY 8d)"},
            "stage_ms": stage, "setup_s": {"simulate": t_gen},
        }
        if not args.no_grid_stress:
            # the grid kernels at the configuration their HBM roofline is quoted on (BASELINE configs[4]: 64
            # problems x 500,500 points, --fullsearch, maxinsert 1000) — the cohort's default-search surfaces
            # above are a few dozen points each and only measure launch latency
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import grid_stress
                gs = grid_stress.run(problems=64, maxinsert=1000, readlen=150, reps=5, device=local_rank)
                line["roofline_grid_stress"] = {
                    "kernel": "grid_setup/fill + grid_surface_tiles + grid_reduce<8> (cluster of 8 CTAs per surface)",
                    "workload": "long-expansion stress (BASELINE configs[4]): 64 problems x 500,500 points, fullsearch, maxinsert 1000, 150 bp",
                    "bound": "hbm", "achieved": gs["algorithmic_GBps"], "peak": hbm, "unit": "GB/s",
                    "frac": gs["algorithmic_GBps"] / hbm, "kernel_ms": gs["grid_ms"], "points": gs["points"],
                    "bytes_per_point": 16, "traffic": 832e6 if gs["points"] == 32032000 else None,
                    "note": "16 B/point algorithmic (8 written + 8 read by the reduction); traffic = dram bytes of the "
                            "tiles + reduce kernels in profiles/r1_grid_stress_full.txt; FP64 log/exp bound (DESIGN 4.2)"}
            except Exception as e:
                line["roofline_grid_stress"] = {"error": str(e)}
        if not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                nsamp = min(cpu_sample_size(args, cores, len(names)), args.samples)
                sample_problems = problems[:nsamp * len(names)]
                v, dt, used, reads, cells, kind, res = cpu_reference_run(sample_problems, cores)
                # the sample doubles as a parity check of the GPU calls
                agree = sum(1 for r, c in zip(res, calls_dev[:len(res)])
                            if r[0] == [int(c["allele1"]), int(c["allele2"])])
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": kind,
                                        "sample": "{} samples x {} loci = {} problems, {} reads, {:.1f} s wall".format(
                                            nsamp, len(names), len(sample_problems), reads, dt),
                                        "reads_per_s": reads / dt, "sw_gcups": cells / dt / 1e9,
                                        "calls_identical_to_gpu": "{}/{}".format(agree, len(res)),
                                        "ssw_c_loop_one_core": ssw_c_loop_ceiling(sample_problems)}
            except Exception as e:  # the GPU line must still be printed
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
