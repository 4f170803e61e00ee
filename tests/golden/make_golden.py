"""
Generate tests/golden/*.npz|json by running the REAL reference code in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (not on the GPU box)

What is real here: gordo/machine/model/factories/utils.py, gordo/machine/model/utils.py,
gordo/machine/model/anomaly/base.py and gordo/machine/model/anomaly/diff.py are loaded from the
reference checkout and executed unmodified.  What is stubbed: the packages the reference
imports at module import time but that are absent from this container (tensorflow, scikeras,
xarray, gordo_core) and the package ``__init__``s that would pull them in; none of the stubbed
names is touched by the code paths exercised.  ``DataFrame.append`` (removed in pandas 2; the
reference pins pandas 1.5.3) is shimmed with the pandas-1 behaviour for a named Series.

The base estimator is scikit-learn's MultiOutputRegressor(LinearRegression()) -- the one the
reference's own detector tests use (tests/gordo/machine/model/anomaly/test_anomaly_detectors.py:45)
-- plus an offset variant that drops leading prediction rows the way an LSTM does.
"""
import importlib.util
import json
import os
import sys
import types
from datetime import timedelta

import numpy as np
import pandas as pd

REF = os.environ.get("GORDO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, **_):
            return pd.concat([self, other.to_frame().T])
        pd.DataFrame.append = _append

    class _DataArray:       # xarray.DataArray / Dataset are only used in annotations / isinstance
        pass
    _stub("xarray", DataArray=_DataArray, Dataset=_DataArray)

    class SensorTag:
        def __init__(self, name):
            self.name = name
    _stub("gordo_core"); _stub("gordo_core.sensor_tag", SensorTag=SensorTag)

    g = os.path.join(REF, "gordo")
    _pkg("gordo", g)
    _pkg("gordo.machine", os.path.join(g, "machine"))
    _pkg("gordo.machine.model", os.path.join(g, "machine", "model"))
    _pkg("gordo.machine.model.anomaly", os.path.join(g, "machine", "model", "anomaly"))
    _pkg("gordo.machine.model.factories", os.path.join(g, "machine", "model", "factories"))

    class KerasAutoEncoder:     # diff.py:23 only builds it as a default argument
        def __init__(self, kind=None, **kw):
            self.kind = kind
    _stub("gordo.machine.model.models", KerasAutoEncoder=KerasAutoEncoder)

    mm = os.path.join(g, "machine", "model")
    _load("gordo.machine.model.base", os.path.join(mm, "base.py"))
    utils = _load("gordo.machine.model.utils", os.path.join(mm, "utils.py"))
    sys.modules["gordo.machine.model"].utils = utils
    _load("gordo.machine.model.anomaly.base", os.path.join(mm, "anomaly", "base.py"))
    diff = _load("gordo.machine.model.anomaly.diff", os.path.join(mm, "anomaly", "diff.py"))
    futils = _load("gordo.machine.model.factories.utils", os.path.join(mm, "factories", "utils.py"))
    return diff, utils, futils


def main():
    from sklearn.base import BaseEstimator, RegressorMixin
    from sklearn.linear_model import LinearRegression
    from sklearn.multioutput import MultiOutputRegressor
    from sklearn.preprocessing import MinMaxScaler
    from sklearn.model_selection import TimeSeriesSplit

    diff, utils, futils = load_reference()

    class OffsetLinear(BaseEstimator, RegressorMixin):
        """Linear model whose predict() drops the first ``offset`` rows (LSTM-like)."""

        def __init__(self, offset=0):
            self.offset = offset

        def fit(self, X, y):
            self.m_ = MultiOutputRegressor(LinearRegression()).fit(X, y)
            return self

        def predict(self, X):
            return self.m_.predict(X)[self.offset:]

    out = {}
    # ---- hourglass dims from the real factories/utils.py
    grid = [(cf, el, n) for cf in (0.0, 0.2, 0.3, 0.5, 0.75, 1.0) for el in (1, 2, 3, 4, 5)
            for n in (1, 3, 5, 10, 20, 37, 50, 100, 200, 1000)]
    hg = {f"{cf}|{el}|{n}": list(futils.hourglass_calc_dims(cf, el, n)) for cf, el, n in grid}

    # ---- sklearn pieces the path leans on
    rng = np.random.default_rng(7)
    A = rng.random((57, 4)) * np.array([1.0, 10.0, 0.01, 5.0]) + np.array([0, -3, 2, 0])
    A[:, 3] = 2.5                                   # constant column -> scale 1
    sc = MinMaxScaler().fit(A)
    out["mm_in"] = A; out["mm_scale"] = sc.scale_; out["mm_min"] = sc.min_
    out["mm_out"] = sc.transform(A)
    tss = {}
    for n in (10, 11, 100, 1003, 100000):
        for k in (2, 3, 5):
            tss[f"{n}|{k}"] = [[int(tr[0]), int(tr[-1]), int(te[0]), int(te[-1])]
                               for tr, te in TimeSeriesSplit(n_splits=k).split(np.zeros((n, 1)))]

    # ---- the detector itself (diff.py) on seeded data
    cases = []
    for ci, (n, t, offset, window, method, idx_kind) in enumerate([
            (300, 3, 0, None, None, "range"),
            (300, 3, 0, None, None, "dates"),
            (257, 5, 4, None, None, "range"),
            (300, 4, 0, 12, "smm", "dates"),
            (300, 4, 2, 12, "sma", "range"),
            (300, 4, 0, 12, "ewma", "range"),
            (64, 2, 0, None, None, "range")]):
        r = np.random.default_rng(100 + ci)
        Xa = r.random((n, t)); ya = Xa * r.random(t) + 0.1 * r.random((n, t))
        tags = [f"tag-{j}" for j in range(t)]
        index = (pd.date_range("2019-01-01", periods=n, freq="10min") if idx_kind == "dates"
                 else pd.RangeIndex(n))
        X = pd.DataFrame(Xa, columns=tags, index=index)
        y = pd.DataFrame(ya, columns=tags, index=index)
        model = diff.DiffBasedAnomalyDetector(base_estimator=OffsetLinear(offset),
                                              scaler=MinMaxScaler(), window=window,
                                              smoothing_method=method)
        model.cross_validate(X=X, y=y)
        model.fit(X, y)
        freq = timedelta(minutes=10)
        frame = model.anomaly(X, y, frequency=freq)
        pre = f"c{ci}_"
        out[pre + "X"] = Xa; out[pre + "y"] = ya
        out[pre + "pred"] = model.predict(X)
        out[pre + "feature_thresholds"] = np.asarray(model.feature_thresholds_, float)
        out[pre + "aggregate_threshold"] = float(model.aggregate_threshold_)
        out[pre + "feature_thresholds_per_fold"] = model.feature_thresholds_per_fold_.to_numpy(float)
        out[pre + "aggregate_thresholds_per_fold"] = np.array(
            [model.aggregate_thresholds_per_fold_[f"fold-{i}"] for i in range(3)], float)
        if window is not None:
            out[pre + "smooth_feature_thresholds"] = np.asarray(model.smooth_feature_thresholds_, float)
            out[pre + "smooth_aggregate_threshold"] = float(model.smooth_aggregate_threshold_)
        groups = []
        for top in dict.fromkeys(c[0] for c in frame.columns):
            if top in ("start", "end"):
                continue
            v = frame[top].to_numpy(float)
            out[pre + "col_" + top] = v
            groups.append(top)
        cases.append({
            "n": n, "t": t, "offset": offset, "window": window, "method": method,
            "index": idx_kind, "columns": [list(map(str, c)) for c in frame.columns],
            "groups": groups, "n_rows": int(len(frame)),
            "start": [None if s is None else str(s) for s in frame["start"].iloc[:3, 0].tolist()]
            if hasattr(frame["start"], "iloc") and frame["start"].ndim == 2 else
            [None if s is None else str(s) for s in frame["start"].iloc[:3].tolist()],
            "end": [None if s is None else str(s) for s in np.asarray(frame["end"]).ravel()[:3].tolist()],
            "metadata_keys": sorted(model.get_metadata().keys()),
        })

    # require_thresholds without cross_validate -> AttributeError (diff.py:448-456)
    r = np.random.default_rng(5)
    Xd = pd.DataFrame(r.random((50, 2))); yd = Xd.copy()
    m = diff.DiffBasedAnomalyDetector(base_estimator=OffsetLinear(0), scaler=MinMaxScaler())
    m.fit(Xd, yd)
    try:
        m.anomaly(Xd, yd)
        raised = False
    except AttributeError:
        raised = True

    # ---- DiffBasedKFCVAnomalyDetector (diff.py:461-635): KFold(5, shuffle, seed 0) predictions, smoothed
    # validation errors, percentile thresholds, and the anomaly frame with its smooth-* columns
    kf_out, kf_cases = {}, []
    for ci, (n, t, window, method, pct) in enumerate([
            (1000, 4, 144, "smm", 0.99),
            (600, 3, 24, "sma", 0.95),
            (500, 5, 12, "ewma", 0.99),
            (400, 2, 7, "smm", 0.9)]):
        r = np.random.default_rng(200 + ci)
        Xa = r.random((n, t)); ya = Xa * r.random(t) + 0.1 * r.random((n, t))
        tags = [f"tag-{j}" for j in range(t)]
        X = pd.DataFrame(Xa, columns=tags); y = pd.DataFrame(ya, columns=tags)
        model = diff.DiffBasedKFCVAnomalyDetector(base_estimator=OffsetLinear(0), scaler=MinMaxScaler(),
                                                  window=window, smoothing_method=method,
                                                  threshold_percentile=pct)
        model.cross_validate(X=X, y=y)
        model.fit(X, y)
        frame = model.anomaly(X, y, frequency=timedelta(minutes=10))
        pre = f"k{ci}_"
        kf_out[pre + "X"] = Xa; kf_out[pre + "y"] = ya
        kf_out[pre + "feature_thresholds"] = np.asarray(model.feature_thresholds_, float)
        kf_out[pre + "aggregate_threshold"] = float(model.aggregate_threshold_)
        groups = []
        for top in dict.fromkeys(c[0] for c in frame.columns):
            if top in ("start", "end"):
                continue
            kf_out[pre + "col_" + top] = frame[top].to_numpy(float)
            groups.append(top)
        kf_cases.append({"n": n, "t": t, "window": window, "method": method, "percentile": pct,
                         "groups": groups, "metadata_keys": sorted(model.get_metadata().keys())})
    np.savez_compressed(os.path.join(HERE, "kfcv_golden.npz"), **kf_out)

    np.savez_compressed(os.path.join(HERE, "detector_golden.npz"), **out)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump({"hourglass": hg, "time_series_split": tss, "cases": cases, "kfcv_cases": kf_cases,
                   "require_thresholds_raises": raised,
                   "versions": {"numpy": np.__version__, "pandas": pd.__version__,
                                "sklearn": __import__("sklearn").__version__},
                   "reference": "equinor/gordo @ 99a4819d"}, f, indent=1)
    print("wrote", os.path.join(HERE, "detector_golden.npz"), "and golden_meta.json;",
          len(cases), "detector cases, require_thresholds raises:", raised)


if __name__ == "__main__":
    main()
