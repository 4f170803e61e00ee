#!/usr/bin/env python
"""
Secondary measurement: smoothing + percentile thresholds (gb200_smooth / gb200_quantile) on the
score columns of a fleet (default: the c2 shape, 128 Machines x 50 tags x 100 000 rows, window 144),
elements/s and effective HBM bandwidth (8 B per element: one float32 read, one written), with
pandas (what the reference calls, diff.py:302-308 / 631-635) timed on a sample beside it.

  python tools/bench_smooth.py [--machines 128] [--tags 50] [--rows 100000] [--window 144]
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
    ap.add_argument("--window", type=int, default=144)
    ap.add_argument("--cpu-machines", type=int, default=1)
    a = ap.parse_args()
    import pandas as pd
    import torch
    from gordo_b200.fleet import FFFleet
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(0)
    R = a.machines * a.rows
    v = torch.rand((R, a.tags), generator=g, device=dev)
    lo = torch.arange(a.machines, device=dev, dtype=torch.int64) * a.rows
    hi = lo + a.rows
    out = torch.empty_like(v)
    res = {"machines": a.machines, "tags": a.tags, "rows": a.rows, "window": a.window, "elements": R * a.tags}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for method in ("smm", "sma", "ewma"):
        ms = timed(lambda: FFFleet.smooth(v, lo, hi, method, a.window, out=out))
        res[method] = {"ms": ms, "elements_per_s": R * a.tags / (ms * 1e-3), "GBps_8B_per_element": R * a.tags * 8 / (ms * 1e-3) / 1e9}
    ms = timed(lambda: FFFleet.quantile(out, lo, hi, 0.99))
    res["quantile"] = {"ms": ms, "elements_per_s": R * a.tags / (ms * 1e-3)}
    # pandas on the host, one Machine at a time as the reference does
    x = v[: a.cpu_machines * a.rows].cpu().numpy().astype(np.float64)
    cpu = {}
    for method, f in (("smm", lambda d: d.rolling(a.window).median()), ("sma", lambda d: d.rolling(a.window).mean()),
                      ("ewma", lambda d: d.ewm(span=a.window).mean())):
        t0 = time.time()
        for m in range(a.cpu_machines):
            sm = f(pd.DataFrame(x[m * a.rows:(m + 1) * a.rows]))
        cpu[method] = a.cpu_machines * a.rows * a.tags / (time.time() - t0)
    t0 = time.time(); sm.quantile(0.99); cpu["quantile"] = a.rows * a.tags / (time.time() - t0)
    res["pandas_elements_per_s_1core"] = cpu
    print(json.dumps(res))


if __name__ == "__main__":
    main()
