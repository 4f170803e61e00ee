#!/usr/bin/env python
"""
Secondary measurement: LSTM autoencoder predict (BASELINE.json configs[3] shape: 200 tags,
lookback 128, hourglass units 167-133-100-100-133-167) on one Machine, windows/s and the
fraction of the bf16 tensor roofline (2.31e8 FLOP/window), with the CPU oracle on a small sample.

  python tools/bench_lstm.py [--rows 20000] [--tags 200] [--lookback 128] [--cpu-windows 64]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--tags", type=int, default=200)
    ap.add_argument("--lookback", type=int, default=128)
    ap.add_argument("--cpu-windows", type=int, default=64)
    ap.add_argument("--max-windows", type=int, default=18944)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--fit-jobs", type=int, default=0, help="also time KerasLSTMBaseEstimator.fit (1 epoch) for J jobs")
    ap.add_argument("--fit-rows", type=int, default=1152)
    a = ap.parse_args()
    import torch
    from gordo_b200.fleet import Schedule
    from gordo_b200.lstm import LSTMFleet
    from gordo_b200.machine.model.factories.lstm_autoencoder import lstm_hourglass
    dev = torch.device("cuda:0")
    topo = lstm_hourglass(a.tags, lookback_window=a.lookback)
    g = torch.Generator(device=dev); g.manual_seed(0)
    fl = LSTMFleet(topo, 1, 0, dev)
    fl.set_params(topo.init_params(1, g, dev))
    X = torch.rand((a.rows, a.tags), generator=g, device=dev)
    sched = Schedule([a.rows])
    fl.predict(sched, X, max_windows=a.max_windows, precision=a.precision); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, off = fl.predict(sched, X, max_windows=a.max_windows, precision=a.precision); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_win = int(off[-1])
    flops_per_window = a.lookback * sum(8 * u * (i + u) for i, u in zip([a.tags] + topo.units[:-1], topo.units)) \
        + 2 * topo.units[-1] * a.tags
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    res = {"windows": n_win, "ms": ms, "windows_per_s": n_win / (ms * 1e-3), "flops_per_window": flops_per_window,
           "achieved_tflops": n_win * flops_per_window / (ms * 1e-3) / 1e12, "peak_tflops": peak,
           "frac_of_bf16_tensor_peak": n_win * flops_per_window / (ms * 1e-3) / 1e12 / peak, "units": topo.units, "precision": a.precision}
    from oracle import factories, lstm as olstm
    spec = factories.lstm_hourglass(a.tags, lookback_window=a.lookback)
    p = olstm.lstm_unflatten(fl.params[0].cpu().numpy(), spec)
    Xc = X[: a.cpu_windows + a.lookback - 1].cpu().numpy()
    t0 = time.time(); want = olstm.lstm_predict(spec, p, Xc, a.lookback, 0); dt = time.time() - t0
    res["cpu_oracle_windows_per_s_1core"] = a.cpu_windows / dt
    res["max_abs_err_vs_oracle"] = float(np.abs(out[: a.cpu_windows].cpu().numpy() - want).max())
    if a.fit_jobs > 0:
        # training: J independent jobs of fit_rows rows, 1 epoch, batch 32 (models.py:557-616);
        # FLOPs ~ 3 x forward per window
        J, n = a.fit_jobs, a.fit_rows
        flt = LSTMFleet(topo, J, 0, dev)
        P = topo.init_params(J, g, dev)
        Xf = torch.rand((J * n, a.tags), generator=g, device=dev)
        lo = np.arange(J, dtype=np.int64) * n; hi = lo + n
        flt.fit_jobs(Xf, Xf, lo, hi, P.clone(), epochs=1); torch.cuda.synchronize()
        Pw = P.clone()
        e0.record(); hl, pl = flt.fit_jobs(Xf, Xf, lo, hi, Pw, epochs=1); e1.record(); torch.cuda.synchronize()
        fms = e0.elapsed_time(e1)
        wins = J * (n - a.lookback + 1)
        res["fit"] = {"jobs": J, "rows_per_job": n, "ms": fms, "train_windows_per_s": wins / (fms * 1e-3),
                      "achieved_tflops_fp32": wins * 3 * flops_per_window / (fms * 1e-3) / 1e12,
                      "optimizer_steps": -(-(n - a.lookback + 1) // 32) + 1, "loss": float(hl[0, 0])}
        cw = 40                                   # CPU oracle: one job, a few windows
        pc = olstm.lstm_unflatten(P[0].cpu().numpy(), spec)
        Xo = Xf[: cw + a.lookback - 1].cpu().numpy()
        t0 = time.time(); olstm.lstm_fit(spec, pc, Xo, Xo, lookback_window=a.lookback, lookahead=0); dt = time.time() - t0
        res["fit"]["cpu_oracle_train_windows_per_s_1core"] = (cw + 1) / dt
    print(json.dumps(res))


if __name__ == "__main__":
    main()
