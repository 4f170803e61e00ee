#!/usr/bin/env python
"""
Secondary measurement (SURVEY.md §8 f-4 upstream): raw tag series -> resampled, interpolated, joined grid for a fleet
(default 32 Machines x 50 tags x 1 day of 1-second samples = 1.4e8 points, 10-minute bins), points/s and effective HBM
bandwidth of gb200_resample (16 B per point: timestamp + value), with gordo-core's pandas path (oracle/dataset.py) on
one host core beside it.

  python tools/bench_resample.py [--machines 32] [--tags 50] [--seconds 86400] [--resolution 10T]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--machines", type=int, default=32)
    ap.add_argument("--tags", type=int, default=50)
    ap.add_argument("--seconds", type=int, default=86400)
    ap.add_argument("--resolution", default="10T")
    ap.add_argument("--cpu-machines", type=int, default=1)
    a = ap.parse_args()
    import pandas as pd
    import torch
    from gordo_b200 import dataset as ds, _native as N
    from oracle import dataset as ods
    start = pd.Timestamp("2022-01-01 00:00:00+00:00"); end = start + pd.Timedelta(seconds=a.seconds)
    idx = pd.date_range(start, periods=a.seconds, freq="s")
    rng = np.random.default_rng(0)
    machines = [ds.MachineSeries([pd.Series(rng.normal(j, 1, a.seconds), index=idx, name=f"m{m}-t{j}") for j in range(a.tags)],
                                 start, end) for m in range(a.machines)]
    n_points = a.machines * a.tags * a.seconds
    fleet = ds.FleetTimeSeries("cuda:0")
    res = {"machines": a.machines, "tags": a.tags, "points": n_points, "resolution": a.resolution}
    # whole call, host series in -> device grid (host descriptor work + H2D of 16 B/point + kernels)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        joined = fleet.join(machines, a.resolution)
        torch.cuda.synchronize(); wall = time.time() - t0
    res["join_call"] = {"s": wall, "points_per_s": n_points / wall, "rows_out": int(sum(len(j) for j in joined))}
    # the resample kernel alone on resident points
    dev = torch.device("cuda:0")
    step = ds._step_ns(a.resolution)
    S = a.machines * a.tags
    t = torch.as_tensor(np.tile(idx.as_unit("ns").asi8, S), device=dev)
    v = torch.randn(S * a.seconds, dtype=torch.float64, device=dev)
    poff = torch.arange(S + 1, device=dev, dtype=torch.int64) * a.seconds
    nb = a.seconds * 10 ** 9 // step + 1
    bin0 = torch.full((S,), int(start.value), dtype=torch.int64, device=dev); nbins = torch.full((S,), nb, dtype=torch.int64, device=dev)
    off = (torch.arange(S, device=dev, dtype=torch.int64) // a.tags) * (nb * a.tags) + torch.arange(S, device=dev, dtype=torch.int64) % a.tags
    stride = torch.full((S,), a.tags, dtype=torch.int64, device=dev)
    out = torch.empty(S * nb, dtype=torch.float64, device=dev)

    def run():
        N.check(N.lib().gb200_resample(S, N.ptr(poff), N.ptr(t), N.ptr(v), N.ptr(bin0), N.ptr(nbins), N.ptr(off), N.ptr(stride),
                                       step, 0, nb, S * a.seconds, S * nb, N.ptr(out), torch.cuda.current_stream().cuda_stream), "resample")
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["resample_kernel"] = {"ms": ms, "points_per_s": n_points / (ms * 1e-3), "GBps_16B_per_point": n_points * 16 / (ms * 1e-3) / 1e9}
    t0 = time.time()
    for m in range(a.cpu_machines):
        ods.join_timeseries(machines[m].series, start, end, a.resolution)
    res["pandas_points_per_s_1core"] = a.cpu_machines * a.tags * a.seconds / (time.time() - t0)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
