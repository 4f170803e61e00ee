"""
Oracle: DiffBasedAnomalyDetector arithmetic (test infrastructure, see oracle/__init__.py).

Restates gordo/machine/model/anomaly/diff.py:166-458 (fit, cross_validate, thresholds,
anomaly columns, smoothing) and gordo/machine/model/utils.py:49-165 (frame layout) on numpy
arrays.  PINNED: tests/golden/make_golden.py runs the REAL reference diff.py / utils.py (imported
from the reference checkout with TensorFlow stubbed out) and tests/test_oracle_golden.py checks
this file against those outputs.
"""
import numpy as np
import pandas as pd

from .scaler import MinMaxScaler, time_series_split, kfold_split, shuffle_rows
from . import dense as _dense
from . import lstm as _lstm


def rolling_min_max(x, window=6):
    """
    pandas ``Series.rolling(window).min().max()`` (diff.py:229-233): max over t >= window-1 of
    min(x[t-window+1 .. t]); NaN when fewer than ``window`` rows.  x: [n] or [n, k] -> scalar / [k].
    """
    x = np.asarray(x, np.float64)
    one_d = x.ndim == 1
    if one_d:
        x = x[:, None]
    if x.shape[0] < window:
        out = np.full(x.shape[1], np.nan)
    else:
        v = np.lib.stride_tricks.sliding_window_view(x, window, axis=0)   # [n-w+1, k, w]
        mins = v.min(axis=2)            # NaN in a window propagates, as pandas min_periods=window
        with np.errstate(all="ignore"):
            out = np.nanmax(mins, axis=0) if not np.all(np.isnan(mins)) else np.full(x.shape[1], np.nan)
    return float(out[0]) if one_d else out


def scaled_mse_per_timestep(scaler, y_true, y_pred):
    # diff.py:268-293
    d = scaler.transform(y_pred) - scaler.transform(y_true)
    return (d ** 2).mean(axis=1)


def smoothing(metric, method, window):
    # diff.py:302-308 (pandas semantics are the definition)
    m = pd.DataFrame(np.asarray(metric, np.float64))
    if method == "smm":
        r = m.rolling(window).median()
    elif method == "sma":
        r = m.rolling(window).mean()
    elif method == "ewma":
        r = m.ewm(span=window).mean()
    else:
        raise ValueError(method)
    r = r.to_numpy()
    return r[:, 0] if np.asarray(metric).ndim == 1 else r


class FFBase:
    """Pipeline[MinMaxScaler, KerasAutoEncoder] of examples/config.yaml:75-82, on the oracle."""

    def __init__(self, spec, params, *, epochs=1, batch_size=32, perms=None, l1_mode="sum",
                 scale_input=True):
        self.spec, self.params = spec, params
        self.epochs, self.batch_size, self.perms, self.l1_mode = epochs, batch_size, perms, l1_mode
        self.scale_input = scale_input
        self.offset = 0

    def _sx(self, X):
        return self.xscaler.transform(X).astype(np.float32) if self.scale_input else np.asarray(X, np.float32)

    def fit(self, X, y):
        if self.scale_input:
            self.xscaler = MinMaxScaler().fit(X)
        self.history, _ = _dense.ff_fit(self.spec, self.params, self._sx(X), y, epochs=self.epochs,
                                        batch_size=self.batch_size, perms=self.perms,
                                        l1_mode=self.l1_mode)
        return self

    def predict(self, X):
        return _dense.ff_predict(self.spec, self.params, self._sx(X))


class LSTMBase:
    """Pipeline[MinMaxScaler, KerasLSTMAutoEncoder / KerasLSTMForecast] on the oracle."""

    def __init__(self, spec, params, *, lookback_window, lookahead=0, epochs=1, batch_size=32,
                 scale_input=True):
        self.spec, self.params = spec, params
        self.L, self.lookahead = lookback_window, lookahead
        self.epochs, self.batch_size, self.scale_input = epochs, batch_size, scale_input

    def _sx(self, X):
        return self.xscaler.transform(X).astype(np.float32) if self.scale_input else np.asarray(X, np.float32)

    def fit(self, X, y):
        if self.scale_input:
            self.xscaler = MinMaxScaler().fit(X)
        self.history_primer, self.history, _ = _lstm.lstm_fit(
            self.spec, self.params, self._sx(X), y, lookback_window=self.L,
            lookahead=self.lookahead, batch_size=self.batch_size, epochs=self.epochs)
        return self

    def predict(self, X):
        return _lstm.lstm_predict(self.spec, self.params, self._sx(X), self.L, self.lookahead)


class DiffDetector:
    """
    diff.py:21-458 over numpy arrays.  ``make_base(tag)`` returns a fresh base estimator for
    ``tag`` in ("fold-0", "fold-1", ..., "final") -- the analogue of sklearn.clone per CV fold,
    with the initial weights / permutations chosen by the caller.
    """

    def __init__(self, make_base, *, require_thresholds=True, window=None, smoothing_method=None):
        self.make_base = make_base
        self.require_thresholds = require_thresholds
        self.window = window
        self.smoothing_method = smoothing_method
        if window is not None and smoothing_method is None:
            self.smoothing_method = "smm"

    def fit(self, X, y):
        # diff.py:166-174 (shuffle=False)
        self.base = self.make_base("final").fit(X, y)
        self.scaler = MinMaxScaler().fit(y)          # fitted on UNSCALED y, after training
        return self

    def predict(self, X):
        return self.base.predict(X)

    def cross_validate(self, X, y, n_splits=3):
        # diff.py:176-266: TimeSeriesSplit folds; per fold a fresh detector (base + scaler)
        X = np.asarray(X); y = np.asarray(y)
        self.aggregate_thresholds_per_fold_ = {}
        self.feature_thresholds_per_fold_ = {}
        self.smooth_aggregate_thresholds_per_fold_ = {}
        self.smooth_feature_thresholds_per_fold_ = {}
        self.fold_predictions_ = []
        agg = tag = sagg = stag = None
        for i, (tr, te) in enumerate(time_series_split(len(X), n_splits)):
            base = self.make_base(f"fold-{i}").fit(X[tr], y[tr])
            fold_scaler = MinMaxScaler().fit(y[tr])
            y_pred = base.predict(X[te])
            te_adj = te[-len(y_pred):]
            y_true = y[te_adj]
            smse = scaled_mse_per_timestep(fold_scaler, y_true, y_pred)
            mae = np.abs(np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64))
            agg = rolling_min_max(smse, 6)
            tag = rolling_min_max(mae, 6)
            self.aggregate_thresholds_per_fold_[f"fold-{i}"] = agg
            self.feature_thresholds_per_fold_[f"fold-{i}"] = tag
            if self.window is not None:
                sagg = rolling_min_max(smse, self.window)
                stag = rolling_min_max(mae, self.window)
                self.smooth_aggregate_thresholds_per_fold_[f"fold-{i}"] = sagg
                self.smooth_feature_thresholds_per_fold_[f"fold-{i}"] = stag
            self.fold_predictions_.append((te_adj, y_pred))
        # final thresholds = the LAST fold's (diff.py:256-264)
        self.feature_thresholds_ = tag
        self.aggregate_threshold_ = agg
        self.smooth_aggregate_threshold_ = sagg
        self.smooth_feature_thresholds_ = stag
        return self

    def anomaly(self, X, y):
        """diff.py:310-458 -> dict of float64 arrays keyed by the reference's column groups."""
        X = np.asarray(X); y = np.asarray(y)
        out = self.predict(X)
        n = len(out)
        res = {"model-input": X[-n:], "model-output": out}
        d_scaled = np.abs(self.scaler.transform(out) - self.scaler.transform(y)[-n:])
        res["tag-anomaly-scaled"] = d_scaled
        res["total-anomaly-scaled"] = np.square(d_scaled).mean(axis=1)
        d_un = np.abs(np.asarray(out, np.float64) - np.asarray(y, np.float64)[-n:])
        res["tag-anomaly-unscaled"] = d_un
        res["total-anomaly-unscaled"] = np.square(d_un).mean(axis=1)
        if self.window is not None and self.smoothing_method is not None:
            res["smooth-tag-anomaly-scaled"] = smoothing(d_scaled, self.smoothing_method, self.window)
            res["smooth-total-anomaly-scaled"] = smoothing(res["total-anomaly-scaled"], self.smoothing_method, self.window)
            res["smooth-tag-anomaly-unscaled"] = smoothing(d_un, self.smoothing_method, self.window)
            res["smooth-total-anomaly-unscaled"] = smoothing(res["total-anomaly-unscaled"], self.smoothing_method, self.window)
        if hasattr(self, "feature_thresholds_"):
            res["anomaly-confidence"] = d_un / np.asarray(self.feature_thresholds_)
        if hasattr(self, "aggregate_threshold_"):
            res["total-anomaly-confidence"] = res["total-anomaly-scaled"] / self.aggregate_threshold_
        if self.require_thresholds and not (hasattr(self, "feature_thresholds_")
                                            or hasattr(self, "aggregate_threshold_")):
            raise AttributeError("`require_thresholds=True` however `.cross_validate` needs to be "
                                 "called in order to calculate these thresholds before calling `.anomaly`")
        return res


class KFCVDetector(DiffDetector):
    """
    DiffBasedKFCVAnomalyDetector (diff.py:461-635): KFold(5, shuffle, seed 0) out-of-fold predictions
    for every row, thresholds = ``threshold_percentile`` quantile of the smoothed validation errors.
    ``shuffle`` mirrors the detector's fit-time row shuffle (diff.py:166-174; on by default here as
    in the reference's constructor).  PINNED by tests/golden/kfcv_golden.npz (real reference run).
    """

    def __init__(self, make_base, *, require_thresholds=True, window=144, smoothing_method="smm",
                 threshold_percentile=0.99, shuffle=True):
        super().__init__(make_base, require_thresholds=require_thresholds, window=window,
                         smoothing_method=smoothing_method)
        self.threshold_percentile = threshold_percentile
        self.shuffle = shuffle

    def _fit_base(self, tag, X, y):
        if self.shuffle:
            order = shuffle_rows(len(X), 0)
            return self.make_base(tag).fit(X[order], y[order])
        return self.make_base(tag).fit(X, y)

    def fit(self, X, y):
        X = np.asarray(X); y = np.asarray(y)
        self.base = self._fit_base("final", X, y)
        self.scaler = MinMaxScaler().fit(y)
        return self

    def _threshold(self, metric):
        # diff.py:631-635: smoothing, then pandas quantile (linear interpolation, NaN skipped)
        sm = smoothing(metric, self.smoothing_method, self.window)
        q = pd.DataFrame(sm).quantile(self.threshold_percentile).to_numpy()
        return float(q[0]) if np.asarray(metric).ndim == 1 else q

    def cross_validate(self, X, y, n_splits=5, seed=0, splits=None):
        """``splits``: explicit (train, test) index pairs instead of the default KFold -- gordo's builder
        passes its own ``cv`` (TimeSeriesSplit by default, build_model.py:239-262); rows no test fold
        covers keep a zero prediction and a NaN validation error, exactly as diff.py:580-612 leaves them."""
        X = np.asarray(X); y = np.asarray(y)
        y_pred = np.zeros_like(np.asarray(y, np.float64))
        y_val_mse = np.full(len(y), np.nan)
        for i, (tr, te) in enumerate(splits if splits is not None else kfold_split(len(X), n_splits, seed)):
            base = self._fit_base(f"fold-{i}", X[tr], y[tr])
            fold_scaler = MinMaxScaler().fit(y[tr])
            y_pred[te] = base.predict(X[te])
            y_val_mse[te] = scaled_mse_per_timestep(fold_scaler, y[te], y_pred[te])
        self.cv_predictions_ = y_pred
        self.aggregate_threshold_ = self._threshold(y_val_mse)
        self.feature_thresholds_ = self._threshold(np.abs(np.asarray(y, np.float64) - y_pred))
        return self


COLUMN_ORDER = ("model-input", "model-output", "tag-anomaly-scaled", "total-anomaly-scaled",
                "tag-anomaly-unscaled", "total-anomaly-unscaled",
                "smooth-tag-anomaly-scaled", "smooth-total-anomaly-scaled",
                "smooth-tag-anomaly-unscaled", "smooth-total-anomaly-unscaled",
                "anomaly-confidence", "total-anomaly-confidence")


def anomaly_frame(res, tags, target_tags=None, index=None, frequency=None):
    """model/utils.py:49-165 + the joins of diff.py:341-444: the MultiIndex-column frame."""
    target_tags = list(target_tags) if target_tags is not None else list(tags)
    n = len(res["model-output"])
    idx = index[-n:] if index is not None else pd.RangeIndex(n)
    if isinstance(idx, pd.DatetimeIndex):
        start = [t.isoformat() for t in idx]
        end = [(t + frequency).isoformat() if frequency is not None else None for t in idx]
    else:
        start = [None] * n; end = [None] * n
    cols = {("start", ""): pd.Series(start, index=idx, dtype=object),
            ("end", ""): pd.Series(end, index=idx, dtype=object)}
    for key in COLUMN_ORDER:
        if key not in res:
            continue
        v = np.asarray(res[key])
        if v.ndim == 1:
            cols[(key, "")] = pd.Series(v, index=idx)
        else:
            names = tags if key == "model-input" else target_tags
            if v.shape[1] != len(names):
                names = [str(i) for i in range(v.shape[1])]
            for j, nm in enumerate(names):
                cols[(key, str(nm))] = pd.Series(v[:, j], index=idx)
    df = pd.DataFrame(cols, index=idx)
    df.columns = pd.MultiIndex.from_tuples(list(cols.keys()))
    return df
