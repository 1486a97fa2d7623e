"""The N>1 path on CPU: world_size-2 `gloo` process group, (sample, locus) sharding with no data-path
collective, host gather of the per-problem records, max / sum reduction of timings and counters."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as tdist
    from tredparse_b200 import dist, simulate, cohort
    from tredparse_b200.meta import TREDsRepo
    try:
        r, w = dist.init("gloo")
        assert (r, w) == (rank, world) and dist.env_world() == (rank, rank, world)
        repo = TREDsRepo()
        names = ["HD", "DM1", "FXS", "SCA1", "FRDA"]
        problems = simulate.simulate_cohort(repo, names, 3, readlen=150)          # 15 (sample, locus) problems
        n = len(problems)
        costs = [p.nreads * (p.readlen + 36) for p in problems]
        mine = dist.shard_indices(n, rank, world, costs)
        # this rank packs ONLY its shard, exactly like bench.py / tred.py do before the GPU call
        batch = cohort.CohortBatch([problems[i] for i in mine])
        assert batch.nproblems == len(mine) and batch.nreads == sum(problems[i].nreads for i in mine)
        # stand-in for the device results (no GPU here): a record that identifies the problem
        rec = np.zeros(len(mine), dtype=cohort.CALL_DTYPE)
        rec["allele1"] = [problems[i].alleles[0] for i in mine]
        rec["n_points"] = mine
        rec["fdp"] = [problems[i].nreads for i in mine]
        allrec = dist.gather_records(rec, mine, n)
        mx, sm = dist.reduce_max_sum([10.0 + rank, 5.0 - rank], [len(mine), batch.nreads])
        dist.barrier()
        if rank == 0:
            ok = (allrec is not None and list(allrec["n_points"]) == list(range(n)) and
                  list(allrec["fdp"]) == [p.nreads for p in problems] and
                  list(allrec["allele1"]) == [p.alleles[0] for p in problems])
            q.put(("rank0", ok, mx, sm, n, sum(p.nreads for p in problems)))
        else:
            q.put(("rank1", allrec is None, mx, sm, len(mine), 0))
        tdist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put(("error", repr(e), None, None, 0, 0))


def test_two_rank_gloo_shard_gather_reduce():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    by = {g[0]: g for g in got}
    assert "error" not in by, by.get("error")
    tag, ok, mx, sm, n, nreads = by["rank0"]
    assert ok, "gathered records are not in problem order"
    assert mx == [11.0, 5.0]                               # max over ranks (timings)
    assert sm == [float(n), float(nreads)]                 # sum over ranks (units processed)
    assert by["rank1"][1] is True and by["rank1"][2] == mx and by["rank1"][3] == sm


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shards_partition_every_problem_exactly_once(world):
    sys.path.insert(0, ROOT)
    from tredparse_b200 import dist
    n = 301
    rng = np.random.default_rng(7)
    costs = rng.integers(1, 1000, n)
    for c in (None, costs):
        parts = [dist.shard_indices(n, r, world, c) for r in range(world)]
        allidx = sorted(i for p in parts for i in p)
        assert allidx == list(range(n))
        if c is not None and world > 1:
            loads = [int(costs[p].sum()) for p in parts]
            assert max(loads) - min(loads) <= int(costs.max())      # LPT balance bound


def test_single_process_helpers_are_identities():
    sys.path.insert(0, ROOT)
    from tredparse_b200 import dist
    rec = np.arange(6, dtype=np.int32).reshape(3, 2)
    out = dist.gather_records(rec, [4, 0, 2], 5)
    assert out.tolist() == [[2, 3], [0, 0], [4, 5], [0, 0], [0, 1]]
    assert dist.reduce_max_sum([1.5], [2, 3]) == ([1.5], [2.0, 3.0])
    dist.barrier()


def test_cost_sorted_deal_balances_and_partitions():
    sys.path.insert(0, ROOT)
    from tredparse_b200 import dist
    rng = np.random.default_rng(3)
    costs = rng.gamma(2.0, 100.0, 30011)
    for world in (1, 2, 4, 8):
        owner = dist.shard_by_cost(costs, world)
        assert owner.min() == 0 and owner.max() == world - 1 and len(owner) == len(costs)
        loads = np.array([costs[owner == r].sum() for r in range(world)])
        assert loads.max() - loads.min() <= costs.max()                     # serpentine deal of the sorted costs
        assert np.array_equal(owner, dist.shard_by_cost(costs.copy(), world))  # deterministic: no communication needed


def _tensor_gather_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        from tredparse_b200 import dist
        dist.init("gloo")
        n = 1001
        costs = np.arange(n, dtype=float) % 17 + 1
        owner = dist.shard_by_cost(costs, world)
        mine = np.nonzero(owner == rank)[0]
        rec = np.zeros(len(mine), dtype=[("a", "<i4"), ("b", "<f8")])
        rec["a"], rec["b"] = mine * 3, mine / 7.0
        counts = np.bincount(owner, minlength=world)
        out, seen = dist.gather_records(rec, mine, n, all_counts=counts, return_seen=True)
        ok = True
        if rank == 0:
            ok = bool(seen.all() and np.array_equal(out["a"], np.arange(n) * 3) and np.allclose(out["b"], np.arange(n) / 7.0))
        else:
            ok = out is None
        q.put(("rank%d" % rank, ok))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("error", traceback.format_exc() + repr(e)))


def test_two_rank_gloo_tensor_gather_of_records():
    """bench.py's final gather: padded uint8 tensors, counts known to every rank from the deterministic partition."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tensor_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert "error" not in got, got.get("error")
    assert got == {"rank0": True, "rank1": True}
