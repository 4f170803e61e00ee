#!/usr/bin/env python
"""
Secondary measurement: request latency of the drop-in estimator surface, the shape the reference's own
benchmark uses (benchmarks/test_ml_server.py:21-44: POST 100 rows x 4 tags to /anomaly/prediction,
which calls `model.anomaly(X, y, frequency)`, server/blueprints/anomaly.py:50).  Times
`DiffBasedAnomalyDetector.anomaly()` and `.predict()` on a fitted Machine, host DataFrame in,
host DataFrame out, and the same request on the CPU oracle.

  python tools/bench_serve.py [--rows 100] [--tags 4] [--calls 200]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100)
    ap.add_argument("--tags", type=int, default=4)
    ap.add_argument("--calls", type=int, default=200)
    a = ap.parse_args()
    import pandas as pd
    import torch
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import MinMaxScaler
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from gordo_b200.machine.model.models import KerasAutoEncoder
    rng = np.random.default_rng(0)
    tags = [f"tag-{i}" for i in range(a.tags)]
    Xtrain = pd.DataFrame(rng.random((1000, a.tags)), columns=tags,
                          index=pd.date_range("2019-01-01", periods=1000, freq="10min"))
    det = DiffBasedAnomalyDetector(base_estimator=Pipeline([("s", MinMaxScaler()),
                                                            ("m", KerasAutoEncoder(kind="feedforward_hourglass"))]))
    det.cross_validate(X=Xtrain, y=Xtrain)
    det.fit(Xtrain, Xtrain)
    X = Xtrain.iloc[:a.rows]
    res = {"rows": a.rows, "tags": a.tags, "calls": a.calls}
    for name, fn in (("anomaly", lambda: det.anomaly(X, X)), ("predict", lambda: det.predict(X))):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.calls):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        res[name] = {"ms_median": float(np.median(ts)), "ms_p95": float(np.percentile(ts, 95)), "ms_min": float(ts.min())}
    # the same request on the CPU oracle (numpy restatement of the reference path)
    from oracle import dense, factories
    from oracle.scaler import MinMaxScaler as OMM
    spec = factories.feedforward_hourglass(a.tags)
    est = det.base_estimator.steps[1][1]
    params = dense.ff_unflatten(est.model.params, spec["widths"])
    sx = OMM().fit(Xtrain.to_numpy()); sy = OMM().fit(Xtrain.to_numpy())
    Xv = X.to_numpy()
    ts = []
    for _ in range(a.calls):
        t0 = time.perf_counter()
        yhat = dense.ff_predict(spec, params, sx.transform(Xv).astype(np.float32), batch_size=32)
        d_s = np.abs(sy.transform(yhat) - sy.transform(Xv)); d_u = np.abs(yhat - Xv)
        _ = (np.square(d_s).mean(axis=1), np.square(d_u).mean(axis=1))
        ts.append(time.perf_counter() - t0)
    res["cpu_oracle_arithmetic_only_ms_median"] = float(np.median(np.array(ts) * 1e3))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
