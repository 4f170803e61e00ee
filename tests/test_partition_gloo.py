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


# ----------------------------------------------------------------------------- round 2: sharded builds, row-split LSTM
def _shard_worker(rank, world, port, out):
    """FleetModelBuilder.build_sharded over gloo with a stand-in build function (the real one needs a GPU):
    every Machine built exactly once, by the rank LPT gave it to, results gathered on rank 0 in Machine order."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gordo_b200.builder import FleetMachine, FleetModelBuilder
    lstm = {"gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo_b200.machine.model.models.KerasLSTMAutoEncoder": {"kind": "lstm_hourglass", "lookback_window": 16}}]}}}}
    machines = [FleetMachine(f"m{i}", np.zeros((1000 + 100 * i, 20 + 7 * i), np.float32), model=lstm if i % 4 == 3 else None)
                for i in range(9)]
    fake_build = lambda ms: [(f"model-of-{m.name}", {"rank": rank, "name": m.name}) for m in ms]
    res = FleetModelBuilder.build_sharded(machines, rank, world, gather=True, build_fn=fake_build)
    if rank == 0:
        out.put({k: v for k, v in res.items()})
    dist.destroy_process_group()


def test_build_sharded_gathers_every_machine_once_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == list(range(9))
    assert all(res[i][0] == f"model-of-m{i}" and res[i][1]["name"] == f"m{i}" for i in range(9))
    ranks = [res[i][1]["rank"] for i in range(9)]
    assert set(ranks) == {0, 1}
    # the two LSTM Machines (16x the per-row cost) must not share a rank
    assert ranks[3] != ranks[7]


def test_shard_indices_and_lstm_row_split_are_exact_partitions():
    import numpy as np
    import bench
    from gordo_b200.builder import FleetMachine, FleetModelBuilder
    ms = [FleetMachine(f"m{i}", np.zeros((500, 10 + i), np.float32)) for i in range(11)]
    for world in (1, 2, 4, 8):
        parts = FleetModelBuilder.shard_indices(ms, world)
        assert sorted(i for p in parts for i in p) == list(range(11)) and len(parts) == world
    # c4: the window range of 8 Machines cut into N contiguous pieces; a Machine may straddle ranks
    rows, L, M = 100_000, 128, 8
    per = rows - L + 1
    for world in (1, 2, 3, 4, 8):
        shards = bench.lstm_shards(M, rows, L, world)
        seen = np.zeros(M * per, np.int32)
        for parts in shards:
            for m, w0, w1 in parts:
                assert 0 <= w0 < w1 <= per
                seen[m * per + w0:m * per + w1] += 1
        assert (seen == 1).all()                                   # every window scored exactly once
        sizes = [sum(w1 - w0 for _, w0, w1 in parts) for parts in shards]
        assert max(sizes) - min(sizes) <= 1                        # evenly
