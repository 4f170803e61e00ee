"""
world_size-2 `gloo` test (CPU) of the multi-GPU plumbing bench.py uses: Machine partition without
overlap, barrier, max-over-ranks timing reduction -- the only collectives on this path.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gordo_b200.partition import lpt, machine_cost, round_robin


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_machines, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = round_robin(n_machines, world, rank)
    owned = torch.zeros(n_machines, dtype=torch.int64)
    owned[mine] = 1
    dist.barrier()
    dist.all_reduce(owned)                                   # every Machine owned exactly once
    t = torch.tensor([1.0 + rank], dtype=torch.float64)     # fake per-rank step time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    windows = torch.tensor([len(mine) * 1000], dtype=torch.int64)
    dist.all_reduce(windows)
    if rank == 0:
        out.put((owned.tolist(), float(t), int(windows)))
    dist.destroy_process_group()


def test_round_robin_partition_over_gloo():
    world, n = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    owned, tmax, windows = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert owned == [1] * n and tmax == 2.0 and windows == n * 1000


def test_lpt_balances_heterogeneous_machines():
    costs = [machine_cost(100_000, f) for f in (210, 734, 2964, 18494, 73844, 18494, 2964, 210)]
    parts = lpt(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) == pytest.approx(costs[4])             # the heaviest Machine alone bounds the makespan
    with pytest.raises(ValueError):
        round_robin(4, 2, 2)
