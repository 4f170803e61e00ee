"""
Partition a project's Machines over the GPUs of one box.  Machines are
independent (own weights, scalers, thresholds, optimizer state), so there is NO data-path
collective: each rank builds / scores its own shard and results are gathered on the host.
Replaces the one-pod-per-Machine fan-out of argo-workflow.yml.template:1543-1557.
"""
from typing import List, Sequence


def round_robin(n_machines: int, world_size: int, rank: int) -> List[int]:
    """Machine ids rank, rank + world_size, ...: the bench's layout (equal-cost Machines)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_machines, world_size))


def machine_cost(n_rows: int, flops_per_row: float, epochs: int = 1, lookback_window: int = 1,
                 full_build: bool = True) -> float:
    """F * (inference rows + 3 * training rows); default full build = 2.5N of each."""
    rows = 2.5 * n_rows if full_build else float(n_rows)
    return flops_per_row * lookback_window * (rows + 3.0 * rows * epochs)


def lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of heterogeneous Machines to ranks."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i); loads[r] += costs[i]
    return [sorted(p) for p in parts]
