#!/usr/bin/env python
"""Where a c3-style build spends its time: one FF bucket and one LSTM bucket alone, phase by phase (torch profiler-free:
wall clock around the library calls with synchronisation).  python tools/prof_build.py [--tags 60] [--rows 100000]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tags", type=int, default=60)
    ap.add_argument("--rows", type=int, default=100000)
    a = ap.parse_args()
    import torch
    import bench
    from gordo_b200.builder import FleetBuild, FleetMachine
    res = {}
    for kind in ("ff", "lstm"):
        est = ({"gordo_b200.machine.model.models.KerasLSTMAutoEncoder": {"kind": "lstm_hourglass", "lookback_window": 16, "precision": "bf16"}}
               if kind == "lstm" else {"gordo_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass"}})
        model = {"gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {
            "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler", est]}}}}
        for n_m in (1, 4):
            ms = [FleetMachine(f"{kind}{i}", bench.machine_data(i, a.rows, a.tags)[1], model=model, evaluation={"seed": i}) for i in range(n_m)]
            FleetBuild(ms[:1] if kind == "ff" else ms[:1], device="cuda:0").build() if n_m == 1 else None
            torch.cuda.synchronize(); t0 = time.time()
            out = FleetBuild(ms, device="cuda:0").build()
            torch.cuda.synchronize()
            res[f"{kind}_T{a.tags}_M{n_m}"] = {"wall_s": round(time.time() - t0, 3), "fleet": out[0][1]["fleet"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
