#!/usr/bin/env python
"""
Secondary measurement: the fused scorer (gb200_ff_score, bf16 tensor-core path) on the other fleet
shapes BASELINE.json names -- tag counts 5 / 20 / 50 / 100 and 128 / 1 024 / 10 000 Machines -- as
windows/s and fraction of the measured HBM bandwidth (algorithmic bytes per window from BASELINE.md §3,
confidences included).  Rows per Machine are reduced where the full 1e5 would not fit one GPU; the
kernel's schedule only sees tiles, so throughput does not depend on the split between Machines and rows
once every Machine has many tiles.

  python tools/bench_shapes.py [--quick]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    import torch
    from gordo_b200.fleet import FFFleet, Schedule
    from gordo_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6570.3))
    shapes = [  # (label, machines, tags, rows per Machine)
        ("c2 128x50x1e5", 128, 50, 100_000),
        ("1024x50x1e5", 1024, 50, 100_000),
        ("c5 10000x5 (rows 2e4)", 10_000, 5, 20_000),
        ("c3 T=20 256x20x1e5", 256, 20, 100_000),
        ("c3 T=100 128x100x1e5", 128, 100, 100_000),
        ("c1 1x10x1000", 1, 10, 1000),
    ]
    if a.quick:
        shapes = [s for s in shapes if s[1] * s[2] * s[3] <= 128 * 50 * 100_000 * 2]
    free = torch.cuda.mem_get_info()[0]
    out = []
    for label, M, T, N in shapes:
        need = M * N * (T * 4 * 5 + 16)
        if need > 0.9 * free:
            out.append({"shape": label, "skipped": f"needs {need / 1e9:.0f} GB"})
            continue
        topo = feedforward_hourglass(T)
        g = torch.Generator(device=dev); g.manual_seed(1)
        fl = FFFleet(topo, M, dev)
        fl.set_params(topo.glorot_init(M, g, dev))
        X = torch.rand((M * N, T), generator=g, device=dev)
        sched = Schedule([N] * M)
        lo = torch.arange(M, device=dev, dtype=torch.int64) * N
        fl.in_scale, fl.in_min = FFFleet.minmax_fit(X, lo, lo + N)
        fl.err_scale = fl.in_scale.clone()
        fl.feat_thr = torch.rand((M, T), generator=g, device=dev) * 0.4 + 0.1
        fl.agg_thr = torch.rand((M,), generator=g, device=dev) * 0.09 + 0.01
        prec = a.precision if fl.tc_eligible(a.precision) else "f32"
        res = fl.score(sched, X, precision=prec)
        for _ in range(2):
            fl.score(sched, X, precision=prec, out=res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            fl.score(sched, X, precision=prec, out=res)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bpw = 4 * T + 4 * (4 * T + 3)
        out.append({"shape": label, "machines": M, "tags": T, "rows": N, "precision": prec, "ms": ms,
                    "windows_per_s": M * N / (ms * 1e-3), "bytes_per_window": bpw,
                    "GBps": M * N * bpw / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": M * N * bpw / (ms * 1e-3) / 1e9 / peak})
        del X, res, fl
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
