"""
-m gpu parity tests of the smoothing and percentile kernels (gb200_smooth, gb200_quantile) through
the C-ABI against the oracle (pandas semantics are the reference's definition: diff.py:302-308,
631-635), on seeded inputs with ragged row ranges, NaNs, ties, short ranges and long windows.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle.anomaly import smoothing

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _i64(a):
    return torch.tensor(np.asarray(a, np.int64), device=DEV)


def _data(rng, n, c, ties=False, nans=0):
    V = rng.random((n, c)).astype(np.float32) * rng.uniform(0.01, 20, c).astype(np.float32)
    if ties:
        V = np.round(V * 8) / 8          # many equal values, including +0.0
        V[rng.random((n, c)) < 0.05] *= -1
    V = V.astype(np.float32)
    for _ in range(nans):
        V[rng.integers(n), rng.integers(c)] = np.nan
    return V


@pytest.mark.parametrize("method", ["smm", "sma", "ewma"])
@pytest.mark.parametrize("window,c", [(1, 3), (2, 1), (5, 7), (12, 50), (144, 50), (145, 4), (600, 2), (144, 300)])
def test_smooth_matches_pandas(method, window, c):
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(1000 * window + c)
    n = 9000 if c <= 50 else 1500
    V = _data(rng, n, c, ties=(window in (5, 145)), nans=(6 if window in (12, 144) and c == 50 else 0))
    lo, hi = [0, n // 3, n - 700], [n // 3, n - 700, n]
    got = FFFleet.smooth(torch.from_numpy(V).to(DEV), _i64(lo), _i64(hi), method, window).cpu().numpy()
    for a, b in zip(lo, hi):
        want = smoothing(V[a:b], method, window)
        if method == "smm":
            np.testing.assert_array_equal(got[a:b], want.astype(np.float32))
        else:
            np.testing.assert_allclose(got[a:b], want, rtol=2e-6, atol=1e-7, equal_nan=True)


def test_smooth_short_job_and_series_and_errors():
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(5)
    v = rng.random(500).astype(np.float32)
    for method in ("smm", "sma", "ewma"):
        got = FFFleet.smooth(torch.from_numpy(v).to(DEV), _i64([0, 100]), _i64([100, 143]), method, 144).cpu().numpy()
        assert got.shape == (500,)
        for a, b in ((0, 100), (100, 143)):
            np.testing.assert_allclose(got[a:b], smoothing(v[a:b], method, 144), rtol=2e-6, equal_nan=True)
        assert np.isnan(got[143:]).all()            # rows outside every job keep the NaN fill
    with pytest.raises(ValueError):
        FFFleet.smooth(torch.from_numpy(v).to(DEV), _i64([0]), _i64([500]), "median", 3)
    with pytest.raises(RuntimeError, match="window"):
        FFFleet.smooth(torch.from_numpy(v).to(DEV), _i64([0]), _i64([500]), "smm", 0)
    with pytest.raises(RuntimeError, match="too long"):
        FFFleet.smooth(torch.from_numpy(v).to(DEV), _i64([0]), _i64([500]), "smm", 5000)


def test_smooth_full_size_properties():
    """c2-sized column set: constant input is a fixed point; sma of a ramp is the ramp shifted by (w-1)/2."""
    from gordo_b200.fleet import FFFleet
    n, c, w = 400_000, 50, 144
    ramp = torch.arange(n, device=DEV, dtype=torch.float32)[:, None].repeat(1, c).contiguous() * (1.0 / 1024)
    lo, hi = _i64([0, 100_000]), _i64([100_000, n])
    for method in ("smm", "sma"):
        got = FFFleet.smooth(ramp, lo, hi, method, w)
        for a, b in ((0, 100_000), (100_000, n)):
            assert torch.isnan(got[a:a + w - 1]).all()
            want = ramp[a + w - 1:b] - (w - 1) / 2 / 1024
            assert (got[a + w - 1:b] - want).abs().max().item() <= 1e-3     # float32 values up to ~390
    const = torch.full((n, c), 0.375, device=DEV)
    for method in ("smm", "sma", "ewma"):
        got = FFFleet.smooth(const, lo, hi, method, w)
        ok = got[~torch.isnan(got)]
        assert (ok == 0.375).all()


@pytest.mark.parametrize("q", [0.0, 0.5, 0.99, 0.999, 1.0, 0.37])
def test_quantile_matches_pandas(q):
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(int(q * 1000) + 3)
    V = _data(rng, 20000, 9, ties=False, nans=40)
    V[:, 2] = np.round(V[:, 2] * 4) / 4                  # heavy ties
    V[:, 5] = -V[:, 5]                                   # negative values
    V[100:200, 7] = np.nan                               # a job whose column is all NaN
    lo, hi = [0, 100, 5000, 19999], [100, 200, 19999, 20000]
    got = FFFleet.quantile(torch.from_numpy(V).to(DEV), _i64(lo), _i64(hi), q).cpu().numpy()
    assert got.dtype == np.float64 and got.shape == (4, 9)
    for j, (a, b) in enumerate(zip(lo, hi)):
        want = pd.DataFrame(V[a:b].astype(np.float64)).quantile(q).to_numpy()
        np.testing.assert_allclose(got[j], want, rtol=1e-12, atol=0, equal_nan=True)


def test_kfcv_threshold_pipeline_matches_pandas():
    """smooth -> quantile, as DiffBasedKFCVAnomalyDetector._calculate_threshold (diff.py:631-635)."""
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(11)
    err = np.abs(rng.normal(size=(50_000, 20))).astype(np.float32)
    lo, hi = _i64([0, 20_000]), _i64([20_000, 50_000])
    for method in ("smm", "sma", "ewma"):
        sm = FFFleet.smooth(torch.from_numpy(err).to(DEV), lo, hi, method, 144)
        thr = FFFleet.quantile(sm, lo, hi, 0.99).cpu().numpy()
        for j, (a, b) in enumerate(((0, 20_000), (20_000, 50_000))):
            want = pd.DataFrame(smoothing(err[a:b], method, 144)).quantile(0.99).to_numpy()
            np.testing.assert_allclose(thr[j], want, rtol=2e-6)


def _golden():
    import json, os
    g = os.path.join(os.path.dirname(__file__), "golden")
    return json.load(open(os.path.join(g, "golden_meta.json"))), np.load(os.path.join(g, "kfcv_golden.npz")), \
        np.load(os.path.join(g, "detector_golden.npz"))


@pytest.mark.parametrize("ci", range(4))
def test_kfcv_detector_matches_real_reference_golden(ci):
    """
    gordo_b200's DiffBasedKFCVAnomalyDetector (smoothing + percentile thresholds on the device) on the
    inputs of tests/golden/kfcv_golden.npz, against what the REAL reference produced for them.
    Tolerance: the device path holds the validation errors in float32 (reference: float64).
    """
    from sklearn.linear_model import LinearRegression
    from sklearn.multioutput import MultiOutputRegressor
    from sklearn.preprocessing import MinMaxScaler
    from gordo_b200.machine.model.anomaly.diff import DiffBasedKFCVAnomalyDetector
    meta, K, _ = _golden()
    case = meta["kfcv_cases"][ci]
    pre = f"k{ci}_"
    tags = [f"tag-{j}" for j in range(case["t"])]
    X = pd.DataFrame(K[pre + "X"], columns=tags); y = pd.DataFrame(K[pre + "y"], columns=tags)
    det = DiffBasedKFCVAnomalyDetector(base_estimator=MultiOutputRegressor(LinearRegression()), scaler=MinMaxScaler(),
                                       window=case["window"], smoothing_method=case["method"],
                                       threshold_percentile=case["percentile"])
    det.cross_validate(X=X, y=y)
    det.fit(X, y)
    np.testing.assert_allclose(np.asarray(det.feature_thresholds_, float), K[pre + "feature_thresholds"], rtol=2e-6)
    np.testing.assert_allclose(det.aggregate_threshold_, K[pre + "aggregate_threshold"], rtol=2e-6)
    assert sorted(det.get_metadata().keys()) == case["metadata_keys"]
    frame = det.anomaly(X, y)
    assert list(dict.fromkeys(c[0] for c in frame.columns if c[0] not in ("start", "end"))) == case["groups"]
    for grp in case["groups"]:
        np.testing.assert_allclose(frame[grp].to_numpy(float), K[pre + "col_" + grp].reshape(frame[grp].shape),
                                   rtol=5e-6, atol=1e-9, equal_nan=True)


# The offset stand-in estimator (an LSTM-like model whose output is shorter than its input) inherits sklearn's
# RegressorMixin.score = r2_score(y, predict(X)): with offset > 0 the lengths differ, sklearn's cross_validate catches the
# ValueError, warns "Scoring failed" and records NaN -- in the reference run that produced the golden as well
# (tests/golden/make_golden.py uses the same stand-in; the default `score` is never read by diff.py:176-266).  Expected,
# hence filtered here rather than left to look like an accident.
@pytest.mark.filterwarnings("ignore:Scoring failed:UserWarning")
@pytest.mark.parametrize("ci", [3, 4, 5])
def test_smooth_columns_match_real_reference_golden(ci):
    """smooth-* columns of DiffBasedAnomalyDetector.anomaly (window 12: smm / sma / ewma) vs the real reference."""
    from sklearn.base import BaseEstimator, RegressorMixin
    from sklearn.linear_model import LinearRegression
    from sklearn.multioutput import MultiOutputRegressor
    from sklearn.preprocessing import MinMaxScaler
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector

    class OffsetLinear(BaseEstimator, RegressorMixin):
        def __init__(self, offset=0):
            self.offset = offset

        def fit(self, X, y):
            self.m_ = MultiOutputRegressor(LinearRegression()).fit(X, y)
            return self

        def predict(self, X):
            return self.m_.predict(X)[self.offset:]

    meta, _, D = _golden()
    case = meta["cases"][ci]
    pre = f"c{ci}_"
    tags = [f"tag-{j}" for j in range(case["t"])]
    X = pd.DataFrame(D[pre + "X"], columns=tags); y = pd.DataFrame(D[pre + "y"], columns=tags)
    det = DiffBasedAnomalyDetector(base_estimator=OffsetLinear(case["offset"]), scaler=MinMaxScaler(),
                                   window=case["window"], smoothing_method=case["method"])
    det.cross_validate(X=X, y=y)
    det.fit(X, y)
    frame = det.anomaly(X, y)
    for grp in case["groups"]:
        np.testing.assert_allclose(frame[grp].to_numpy(float), D[pre + "col_" + grp].reshape(frame[grp].shape),
                                   rtol=5e-6, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("window", [3, 8, 144, 145])
def test_moving_median_rank_tracking_edge_cases(window):
    """The rank-tracking moving median (one ring pass per output) on the inputs that stress its bookkeeping: heavy ties
    (a handful of distinct values, signed zeros), bursts of missing values that force a re-selection, +-inf (pandas
    turns them into missing values for every rolling / ewm function), runs long enough to cross several thread runs."""
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(window)
    n, c = 40_000, 9
    V = rng.standard_normal((n, c)).astype(np.float32)
    V[:, 1] = rng.integers(-2, 3, n)                       # five distinct values
    V[:, 2] = np.where(rng.random(n) < 0.5, 0.0, -0.0)     # signed zeros only
    V[:, 3] = 1.25                                         # constant
    V[:, 4] = np.where(rng.random(n) < 0.01, np.nan, rng.integers(0, 3, n))
    V[rng.integers(0, n, 40), 5] = np.inf
    V[rng.integers(0, n, 40), 5] = -np.inf
    V[5000:5400, 6] = np.nan                               # a burst longer than the window
    V[:, 7] = np.sort(V[:, 7])                             # monotone: the median moves every step, same direction
    V[:, 8] = -np.sort(V[:, 8])
    lo, hi = [0, 25_000], [25_000, n]
    for method in ("smm", "sma", "ewma"):
        got = FFFleet.smooth(torch.from_numpy(V).to(DEV), _i64(lo), _i64(hi), method, window).cpu().numpy()
        for a, b in zip(lo, hi):
            want = smoothing(V[a:b], method, window)
            if method == "smm":
                np.testing.assert_array_equal(got[a:b], want.astype(np.float32))
            else:
                np.testing.assert_allclose(got[a:b], want, rtol=2e-6, atol=1e-6, equal_nan=True)
