"""
Multi-GPU plumbing: (sample, locus) problems are independent (tredparse/tred.py:225-249 reads nothing but
its own BAM window), so the path shards with NO data-path collective — the only exchanges are the final
host gather of per-problem results (the reference's analogue: ``Pool.imap`` returning dicts,
tred.py:528-532) and the max / sum of timings and counters that ``bench.py`` reports.

One process per GPU under ``torch.distributed`` (backend ``nccl`` on the GPU box, ``gloo`` in the CPU
tests); every helper degrades to the single-process case when the process group is not initialised.
"""
import os

import numpy as np


def bind_to_gpu_numa(device_index):
    """Pin this process (and the threads / processes it starts from now on) to the CPU cores next to GPU
    `device_index`, so that the page-locked staging buffers it allocates are first-touched on that GPU's NUMA node:
    with one process per GPU and 8 GPUs pulling ~200 MB per step each from host memory, buffers that all live on one
    socket make every second GPU copy across the inter-socket link.  Uses NVML's CPU affinity of the device (what
    `nvidia-smi topo -m` prints); returns a short description, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # (NVML is asked directly — no CUDA context is created here, the caller may still fork helper processes)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x.strip() for x in visible.split(",") if x.strip()]
        index = device_index
        if ids and all(x.isdigit() for x in ids) and device_index < len(ids):
            index = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(ncpu, 1) + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return "{} cpus next to GPU {} ({}..{})".format(len(target), device_index, min(target), max(target))
    except Exception:
        return None


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend, device_id=None):
    """Initialise the default process group from the environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if not dist.is_initialized():
        kw = {"device_id": device_id} if device_id is not None else {}
        dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def _active():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def shard_indices(n_items, rank, world, costs=None):
    """Indices of the problems rank `rank` owns.

    Without costs: round-robin (problem i -> rank i % world), which keeps every sample's loci spread over
    all GPUs.  With per-problem costs (e.g. reads x template cells): longest-processing-time greedy, ties
    broken by index so that every rank computes the same partition without communicating."""
    if costs is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def shard_by_cost(costs, world):
    """Owner rank of every item, int32 [n]: items sorted by cost (descending, ties by index) are dealt to the ranks
    in serpentine order (0..w-1, w-1..0, ...) — the vectorised stand-in for the greedy LPT of shard_indices for
    lists of 10^5 - 10^7 problems; every rank computes the same partition without communicating."""
    costs = np.asarray(costs, dtype=np.float64).ravel()
    n = len(costs)
    order = np.lexsort((np.arange(n), -costs))
    k = np.arange(n)
    pos = k % (2 * world)
    lane = np.where(pos < world, pos, 2 * world - 1 - pos)
    owner = np.empty(n, dtype=np.int32)
    owner[order] = lane.astype(np.int32)
    return owner


def gather_records(local, indices, n_total, dst=0, all_counts=None, require_all=True, return_seen=False):
    """Host gather of per-problem result records (a numpy structured/plain array, one row per owned
    problem) into problem order on rank `dst`; other ranks get None.  `indices` = shard_indices(...)."""
    local = np.ascontiguousarray(local)
    if not _active():
        out = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
        idx = np.asarray(indices, dtype=np.int64)
        out[idx] = local
        seen = np.zeros(n_total, dtype=bool)
        seen[idx] = True
        return (out, seen) if return_seen else out
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if all_counts is not None:
        return _gather_records_tensor(local, indices, n_total, dst, all_counts, require_all, return_seen)
    payload = (np.asarray(indices, dtype=np.int64), local.view(np.uint8).reshape(len(local), -1) if len(local) else
               np.zeros((0, local.dtype.itemsize), np.uint8))
    bucket = [None] * world if rank == dst else None
    dist.gather_object(payload, bucket, dst=dst)
    if rank != dst:
        return (None, None) if return_seen else None
    out = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
    flat = out.view(np.uint8).reshape(n_total, -1)
    seen = np.zeros(n_total, dtype=bool)
    for idx, rows in bucket:
        if len(idx):
            assert not seen[idx].any(), "a problem was computed by two ranks"
            flat[idx] = rows
            seen[idx] = True
    assert seen.all() or not require_all, "some problems were computed by no rank"
    return (out, seen) if return_seen else out


def _gather_records_tensor(local, indices, n_total, dst, all_counts, require_all=True, return_seen=False):
    """gather_records without pickling: `all_counts[r]` = number of records rank r holds (every rank can compute
    it from the deterministic partition).  Index and record bytes travel as one padded uint8 tensor per rank —
    on the GPU with the nccl backend (one small H2D / D2H), on the host with gloo."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    width = 8 + local.dtype.itemsize * int(np.prod(local.shape[1:], dtype=np.int64))
    cap = int(max(all_counts))
    buf = np.zeros((cap, width), dtype=np.uint8)
    if len(local):
        buf[:len(local), :8] = np.asarray(indices, dtype=np.int64).view(np.uint8).reshape(-1, 8)
        buf[:len(local), 8:] = local.view(np.uint8).reshape(len(local), -1)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(buf).to(dev)
    bucket = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bucket, dst=dst)
    if rank != dst:
        return (None, None) if return_seen else None
    out = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
    flat = out.view(np.uint8).reshape(n_total, -1)
    seen = np.zeros(n_total, dtype=bool)
    for r, b in enumerate(bucket):
        rows = b.cpu().numpy()[:int(all_counts[r])]
        idx = rows[:, :8].copy().view(np.int64).ravel()
        assert not seen[idx].any(), "a problem was computed by two ranks"
        flat[idx] = rows[:, 8:]
        seen[idx] = True
    assert seen.all() or not require_all, "some problems were computed by no rank"
    return (out, seen) if return_seen else out


def reduce_max_sum(maxima, sums, device="cpu"):
    """(max over ranks of each entry of `maxima`, sum over ranks of each entry of `sums`) as float lists —
    the timing (max) and unit counters (sum) of bench.py.  Identity without a process group."""
    if not _active():
        return [float(x) for x in maxima], [float(x) for x in sums]
    import torch
    import torch.distributed as dist
    a = torch.tensor([float(x) for x in maxima], dtype=torch.float64, device=device)
    b = torch.tensor([float(x) for x in sums], dtype=torch.float64, device=device)
    if len(maxima):
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
    if len(sums):
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return [float(x) for x in a.cpu()], [float(x) for x in b.cpu()]


def barrier():
    if _active():
        import torch.distributed as dist
        dist.barrier()
