#!/usr/bin/env python
"""
Secondary measurement (not the driver's bench): a full build -- TimeSeriesSplit(3) CV + final fit +
fold scoring + thresholds -- of M Machines x T tags x N rows through the batched kernels, the way
FleetModelBuilder drives them.  Reports training rows/s (2.5 N per Machine per epoch), the fit
kernel's share, and the CPU oracle on a bounded sample beside it.

  python tools/bench_build.py [--machines 128] [--tags 50] [--rows 100000] [--cpu-rows 20000]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--machines", type=int, default=128)
    ap.add_argument("--tags", type=int, default=50)
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--cpu-rows", type=int, default=20_000)
    a = ap.parse_args()
    import torch
    from gordo_b200.builder import segmented_randperm
    from gordo_b200.fleet import FFFleet, Schedule, time_series_split_bounds
    from gordo_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
    dev = torch.device("cuda:0")
    M, T, N = a.machines, a.tags, a.rows
    topo = feedforward_hourglass(T)
    g = torch.Generator(device=dev); g.manual_seed(0)
    X = torch.rand((M * N, T), generator=g, device=dev)
    off = np.arange(M + 1, dtype=np.int64) * N
    lo, hi, tlo, thi = [], [], [], []
    for m in range(M):
        for s, e in time_series_split_bounds(N, 3):
            lo.append(off[m]); hi.append(off[m] + s); tlo.append(off[m] + s); thi.append(off[m] + e)
        lo.append(off[m]); hi.append(off[m] + N)
    J = len(lo)
    lo_t = torch.as_tensor(np.asarray(lo), device=dev); hi_t = torch.as_tensor(np.asarray(hi), device=dev)
    fleet = FFFleet(topo, M, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    torch.cuda.synchronize(); t0 = time.time()
    ev[0].record()
    in_scale, in_min = FFFleet.minmax_fit(X, lo_t, hi_t)
    err_scale, _ = FFFleet.minmax_fit(X, lo_t, hi_t)
    params = topo.glorot_init(J, g, dev)
    n_job = np.asarray(hi) - np.asarray(lo)
    pool = segmented_randperm(np.repeat(n_job, a.epochs), g, dev)
    poff = torch.as_tensor(np.concatenate([[0], np.cumsum(n_job * a.epochs)[:-1]]).astype(np.int64), device=dev)
    ev[1].record()
    hl, ha, _, _ = fleet.fit_jobs(X, None, lo_t, hi_t, params, in_scale=in_scale, in_min=in_min, epochs=a.epochs,
                                  batch_size=32, perm_pool=pool, perm_off=poff)
    ev[2].record()
    fold_jobs = torch.as_tensor(np.array([m * 4 + i for m in range(M) for i in range(3)]), device=dev)
    vf = FFFleet(topo, M * 3, dev)
    vf.set_params(params[fold_jobs]); vf.in_scale = in_scale[fold_jobs].contiguous(); vf.in_min = in_min[fold_jobs].contiguous()
    vf.err_scale = err_scale[fold_jobs].contiguous()
    vs = Schedule(rows_lo=tlo, rows_hi=thi, rows_total=M * N)
    res = vf.score(vs, X, precision="f32", columns=("tag-anomaly-unscaled", "total-anomaly-scaled"))
    tl = torch.as_tensor(np.asarray(tlo), device=dev); th = torch.as_tensor(np.asarray(thi), device=dev)
    FFFleet.rolling_min_max(res["tag-anomaly-unscaled"], tl, th, 6)
    FFFleet.rolling_min_max(res["total-anomaly-scaled"], tl, th, 6)
    ev[3].record()
    torch.cuda.synchronize(); wall = time.time() - t0
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    train_rows = float(n_job.sum()) * a.epochs
    out = {"machines": M, "tags": T, "rows": N, "fit_jobs": J, "epochs": a.epochs,
           "ms_setup": ms[0], "ms_fit": ms[1], "ms_cv_score_thresholds": ms[2], "wall_s": wall,
           "train_rows_per_s": train_rows / (ms[1] * 1e-3), "build_windows_per_s": M * N / wall,
           "final_loss_mean": float(hl[3::4, -1].mean())}
    # CPU oracle beside it: one fit of --cpu-rows rows (batch 32, 1 epoch), scaled per row
    from oracle import dense, factories
    spec = factories.feedforward_hourglass(T)
    rng = np.random.default_rng(0)
    Xc = rng.random((a.cpu_rows, T), dtype=np.float32)
    p = dense.ff_init(spec, rng)
    t1 = time.time(); dense.ff_fit(spec, p, Xc, Xc, epochs=1, batch_size=32, perms=[rng.permutation(a.cpu_rows)]); dt = time.time() - t1
    out["cpu_oracle_train_rows_per_s_1core"] = a.cpu_rows / dt
    out["speedup_fit_vs_1core"] = out["train_rows_per_s"] / out["cpu_oracle_train_rows_per_s_1core"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
