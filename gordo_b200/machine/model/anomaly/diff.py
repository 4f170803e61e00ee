"""
DiffBasedAnomalyDetector / DiffBasedKFCVAnomalyDetector with the surface and semantics of
gordo/machine/model/anomaly/diff.py:21-635: same constructor, ``get_params`` / ``get_metadata`` /
``score`` / ``fit`` / ``cross_validate`` / ``anomaly``, same threshold attributes, same output
frame (columns, order, row alignment), same ``AttributeError`` when thresholds are required but
absent.

``anomaly()`` takes the fused GPU path -- MinMax input scaling, Dense stack, |yhat - y| columns and
confidences in ONE launch of libgordo_b200.so -- whenever the detector is the standard
``Pipeline[MinMaxScaler, KerasAutoEncoder]`` (or a bare KerasAutoEncoder) with a MinMaxScaler error
scaler; for the LSTM estimators the GPU predict is followed by ``gb200_score_outputs`` on the offset
output.  Any other composition (RobustScaler, extra transformers) predicts on the GPU through the
base estimator and derives the columns on the host exactly as the reference does.
"""
from datetime import timedelta
from typing import Optional, Union

import numpy as np
import pandas as pd
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.exceptions import NotFittedError
from sklearn.model_selection import KFold, TimeSeriesSplit, cross_validate as c_val
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler
from sklearn.utils import shuffle

from gordo_b200.machine.model import utils as model_utils
from gordo_b200.machine.model.anomaly.base import AnomalyDetectorBase
from gordo_b200.machine.model.base import GordoBase
from gordo_b200.machine.model.models import KerasAutoEncoder, KerasLSTMBaseEstimator


import threading

# gordo.server shares ONE model object between its gthread workers (server/utils.py:334-335): creating the
# device-side serving state is serialised; scoring itself runs concurrently (stream-ordered, per-call buffers)
_SERVING_LOCK = threading.RLock()


def _rows(a, idx):
    return a.iloc[idx] if isinstance(a, pd.DataFrame) else a[idx]


def _rolling_min_max(values: np.ndarray, window: int):
    """`.rolling(window).min().max()` per column on the GPU (gb200_rolling_min_max)."""
    import torch
    from gordo_b200.fleet import FFFleet
    v = np.array(values, dtype=np.float32, order="C")          # private writable copy for torch
    one_d = v.ndim == 1
    if one_d:
        v = v[:, None]
    if len(v) == 0:
        out = np.full(v.shape[1], np.nan)
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        lo = torch.zeros(1, dtype=torch.int64, device=dev)
        hi = torch.full((1,), len(v), dtype=torch.int64, device=dev)
        out = FFFleet.rolling_min_max(torch.as_tensor(v, device=dev), lo, hi, window)[0].double().cpu().numpy()
    return float(out[0]) if one_d else out


class DiffBasedAnomalyDetector(AnomalyDetectorBase):
    def __init__(self, base_estimator: BaseEstimator = None, scaler: TransformerMixin = None,
                 require_thresholds: bool = True, shuffle: bool = False, window: Optional[int] = None,
                 smoothing_method: Optional[str] = None):
        """
        Wraps ``base_estimator`` and scores by reconstruction error.  ``scaler`` is fitted on the
        target AFTER training, purely for the error columns; thresholds are the rolling-min-max of
        the validation errors of the last cross-validation fold (diff.py:30-76).
        """
        self.base_estimator = base_estimator if base_estimator is not None else KerasAutoEncoder(kind="feedforward_hourglass")
        self.scaler = scaler if scaler is not None else MinMaxScaler()
        self.require_thresholds = require_thresholds
        self.shuffle = shuffle
        self.window = window
        self.smoothing_method = smoothing_method
        if self.window is not None and self.smoothing_method is None:
            self.smoothing_method = "smm"

    def __getattr__(self, item):
        # transparent into base_estimator for anything this object does not own (diff.py:78-86)
        if item in ("base_estimator", "__setstate__", "__getstate__") or item.startswith("__"):
            raise AttributeError(item)
        if item in self.__dict__:
            return self.__dict__[item]
        return getattr(self.__dict__["base_estimator"], item) if "base_estimator" in self.__dict__ else \
            object.__getattribute__(self, item)

    # ------------------------------------------------------------------ metadata / params
    def get_metadata(self):
        metadata = dict()
        d = self.__dict__
        if "feature_thresholds_" in d:
            metadata["feature-thresholds"] = np.asarray(self.feature_thresholds_).tolist()
        if "aggregate_threshold_" in d:
            metadata["aggregate-threshold"] = self.aggregate_threshold_
        if "feature_thresholds_per_fold_" in d:
            metadata["feature-thresholds-per-fold"] = self.feature_thresholds_per_fold_.to_dict()
        if "aggregate_thresholds_per_fold_" in d:
            metadata["aggregate-thresholds-per-fold"] = self.aggregate_thresholds_per_fold_
        metadata["window"] = self.window
        metadata["smoothing-method"] = self.smoothing_method
        if d.get("smooth_feature_thresholds_") is not None and d.get("smooth_aggregate_threshold_") is not None:
            metadata["smooth-feature-thresholds"] = np.asarray(self.smooth_feature_thresholds_).tolist()
        if d.get("smooth_aggregate_threshold_") is not None:
            metadata["smooth-aggregate-threshold"] = self.smooth_aggregate_threshold_
        if "smooth_feature_thresholds_per_fold_" in d:
            metadata["smooth-feature-thresholds-per-fold"] = self.smooth_feature_thresholds_per_fold_.to_dict()
        if "smooth_aggregate_thresholds_per_fold_" in d:
            metadata["smooth-aggregate-thresholds-per-fold"] = self.smooth_aggregate_thresholds_per_fold_
        if isinstance(self.base_estimator, GordoBase):
            metadata.update(self.base_estimator.get_metadata())
        else:
            metadata.update({"scaler": str(self.scaler), "base_estimator": str(self.base_estimator),
                             "shuffle": self.shuffle})
        return metadata

    def score(self, X, y, sample_weight: Optional[np.ndarray] = None) -> float:
        return self.base_estimator.score(X, y)

    def get_params(self, deep=True):
        params = {"base_estimator": self.base_estimator, "scaler": self.scaler, "shuffle": self.shuffle}
        if self.window is not None:
            params["window"] = self.window
            params["smoothing_method"] = self.smoothing_method
        return params

    # ------------------------------------------------------------------ fit / cross_validate
    def fit(self, X, y):
        if self.shuffle:
            X_shuff, y_shuff = shuffle(X, y, random_state=0)
            self.base_estimator.fit(X_shuff, y_shuff)
        else:
            self.base_estimator.fit(X, y)
        self.scaler.fit(y)              # used for the error columns of .anomaly()
        return self

    def cross_validate(self, *, X, y, cv=TimeSeriesSplit(n_splits=3), **kwargs):
        """
        sklearn ``cross_validate`` over ``cv`` (fold models are returned), then per fold the
        validation errors -> ``rolling(6).min().max()`` thresholds; the detector keeps the LAST
        fold's (diff.py:176-266).
        """
        kwargs.update(dict(return_estimator=True, cv=cv))
        cv_output = c_val(self, X=X, y=y, **kwargs)

        feature_rows, smooth_rows = {}, {}
        self.aggregate_thresholds_per_fold_ = {}
        self.smooth_aggregate_thresholds_per_fold_ = {}
        aggregate_threshold_fold = tag_thresholds_fold = None
        smooth_aggregate_threshold_fold = smooth_tag_thresholds_fold = None
        columns = None
        for i, ((_, test_idxs), split_model) in enumerate(zip(kwargs["cv"].split(X, y), cv_output["estimator"])):
            y_pred = split_model.predict(_rows(X, test_idxs))
            test_idxs = test_idxs[-len(y_pred):]            # model offset (LSTM lookback)
            y_true = _rows(y, test_idxs)
            columns = list(y_true.columns) if isinstance(y_true, pd.DataFrame) else list(range(np.shape(y_true)[1]))
            scaled_mse = self._scaled_mse_per_timestep(split_model, y_true, y_pred)
            mae = self._absolute_error(y_true, y_pred)
            aggregate_threshold_fold = _rolling_min_max(scaled_mse.to_numpy(), 6)
            self.aggregate_thresholds_per_fold_[f"fold-{i}"] = aggregate_threshold_fold
            tag_thresholds_fold = pd.Series(_rolling_min_max(mae.to_numpy(), 6), index=columns, name=f"fold-{i}")
            feature_rows[f"fold-{i}"] = tag_thresholds_fold
            if self.window is not None:
                smooth_aggregate_threshold_fold = _rolling_min_max(scaled_mse.to_numpy(), self.window)
                self.smooth_aggregate_thresholds_per_fold_[f"fold-{i}"] = smooth_aggregate_threshold_fold
                smooth_tag_thresholds_fold = pd.Series(_rolling_min_max(mae.to_numpy(), self.window),
                                                       index=columns, name=f"fold-{i}")
                smooth_rows[f"fold-{i}"] = smooth_tag_thresholds_fold
        self.feature_thresholds_per_fold_ = pd.DataFrame(feature_rows).T if feature_rows else pd.DataFrame()
        self.smooth_feature_thresholds_per_fold_ = pd.DataFrame(smooth_rows).T if smooth_rows else pd.DataFrame()
        self.feature_thresholds_ = tag_thresholds_fold
        self.aggregate_threshold_ = aggregate_threshold_fold
        self.smooth_aggregate_threshold_ = smooth_aggregate_threshold_fold
        self.smooth_feature_thresholds_ = smooth_tag_thresholds_fold
        return cv_output

    @staticmethod
    def _scaled_mse_per_timestep(model, y_true, y_pred) -> pd.Series:
        try:
            scaled_y_true = model.scaler.transform(y_true)
        except (NotFittedError, ValueError):
            scaled_y_true = model.scaler.fit_transform(y_true)
        scaled_y_pred = model.scaler.transform(_as_like(y_pred, y_true))
        return pd.Series(((np.asarray(scaled_y_pred) - np.asarray(scaled_y_true)) ** 2).mean(axis=1))

    @staticmethod
    def _absolute_error(y_true, y_pred) -> pd.DataFrame:
        return pd.DataFrame(np.abs(np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64)))

    def _smoothing(self, metric: Union[pd.DataFrame, pd.Series]):
        """'smm' / 'sma' / 'ewma' of every column (diff.py:302-308), one ``gb200_smooth`` launch."""
        if self.smoothing_method not in ("smm", "sma", "ewma"):
            # the reference falls through its if-chain, returns None and dies on `None.columns`
            raise AttributeError(f"smoothing_method must be 'smm', 'sma' or 'ewma', got {self.smoothing_method!r}")
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        v = torch.as_tensor(np.ascontiguousarray(np.asarray(metric, np.float32)), device=dev)
        sm = self._smooth_device(v).cpu().numpy().astype(np.float64)
        if isinstance(metric, pd.Series):
            return pd.Series(sm, index=metric.index, name=metric.name)
        return pd.DataFrame(sm, index=metric.index, columns=metric.columns)

    def _smooth_device(self, v):
        """v: [n] or [n, C] float32 CUDA, one series per column, all rows one range."""
        import torch
        from gordo_b200.fleet import FFFleet
        lo = torch.zeros(1, dtype=torch.int64, device=v.device)
        hi = torch.full((1,), v.shape[0], dtype=torch.int64, device=v.device)
        return FFFleet.smooth(v.contiguous(), lo, hi, self.smoothing_method, self.window)

    def _add_smooth_columns(self, res):
        """smooth-* columns of .anomaly() (diff.py:387-415) from the device-resident score columns."""
        if self.window is not None and self.smoothing_method in ("smm", "sma", "ewma"):
            for k in ("tag-anomaly-scaled", "total-anomaly-scaled", "tag-anomaly-unscaled", "total-anomaly-unscaled"):
                res["smooth-" + k] = self._smooth_device(res[k])
        return res

    # ------------------------------------------------------------------ anomaly
    def _fused_plan(self):
        """(input MinMaxScaler or None, KerasAutoEncoder) when the fused launch applies, else None."""
        if not isinstance(self.scaler, MinMaxScaler) or not hasattr(self.scaler, "scale_"):
            return None
        if tuple(getattr(self.scaler, "feature_range", (0, 1))) != (0, 1) or getattr(self.scaler, "clip", False):
            return None         # the fused |d| * scale_ identity only holds for an unclipped (0, 1) scaler
        be = self.base_estimator
        ours = lambda e: (type(e) is KerasAutoEncoder or isinstance(e, KerasLSTMBaseEstimator)) and e.model is not None
        if ours(be):
            return (None, be)
        if isinstance(be, Pipeline) and len(be.steps) == 2:
            sc, est = be.steps[0][1], be.steps[1][1]
            if (isinstance(sc, MinMaxScaler) and hasattr(sc, "scale_") and ours(est)
                    and tuple(sc.feature_range) == (0, 1) and not getattr(sc, "clip", False)):
                return (sc, est)
        return None

    def _fused_columns(self, plan, Xv: np.ndarray, yv: np.ndarray):
        import torch
        from gordo_b200.fleet import FFFleet, Schedule
        sc, est = plan
        topo = est.model.topology
        dev = torch.device("cuda", torch.cuda.current_device())
        if isinstance(est, KerasLSTMBaseEstimator):
            return self._fused_columns_lstm(sc, est, Xv, yv, dev)
        fleet = self._serving_fleet(sc, est, topo, dev)
        n, To = len(Xv), topo.n_out
        xd = torch.as_tensor(np.ascontiguousarray(Xv, np.float32), device=dev)
        yd = None if (Xv.shape == yv.shape and np.array_equal(Xv, yv)) else \
            torch.as_tensor(np.ascontiguousarray(yv, np.float32), device=dev)
        prec = fleet.auto_precision(est._precision)
        # every column is a slice of ONE device buffer: the kernel writes them in place, the host reads them with one copy
        mats = ["model-output", "tag-anomaly-scaled", "tag-anomaly-unscaled"] + (["anomaly-confidence"] if fleet.feat_thr is not None else [])
        vecs = ["total-anomaly-scaled", "total-anomaly-unscaled"] + (["total-anomaly-confidence"] if fleet.agg_thr is not None else [])
        flat = torch.empty(n * (len(mats) * To + len(vecs)), dtype=torch.float32, device=dev)
        out, o = {}, 0
        for k in mats:
            out[k] = flat[o:o + n * To].view(n, To); o += n * To
        for k in vecs:
            out[k] = flat[o:o + n]; o += n
        res = fleet.score(Schedule.single(n), xd, yd, precision=prec, columns=tuple(mats + vecs), out=out)
        if self.window is not None and self.smoothing_method in ("smm", "sma", "ewma"):
            return {k: v.cpu().numpy() for k, v in self._add_smooth_columns(res).items()}
        host = flat.cpu().numpy()
        cols, o = {}, 0
        for k in mats:
            cols[k] = host[o:o + n * To].reshape(n, To); o += n * To
        for k in vecs:
            cols[k] = host[o:o + n]; o += n
        return cols

    def _serving_fleet(self, sc, est, topo, dev):
        """
        The one-Machine device fleet of this detector (weights, bf16 operand image, scalers, thresholds),
        kept between calls: a server scores the same fitted model request after request
        (server/blueprints/anomaly.py:50).  Rebuilt when any of its sources changed.
        """
        import torch
        from gordo_b200.fleet import FFFleet
        feat = self.__dict__.get("feature_thresholds_"); agg = self.__dict__.get("aggregate_threshold_")
        parts = [est.model.params, np.asarray(self.scaler.scale_)]
        if sc is not None:
            parts += [np.asarray(sc.scale_), np.asarray(sc.min_)]
        if feat is not None:
            parts.append(np.asarray(feat, np.float64))
        key = (dev.index, id(est.model), agg, sc is None, feat is None,
               hash(b"".join(np.ascontiguousarray(a).tobytes() for a in parts)))
        cached = self.__dict__.get("_gb200_serving")
        if cached is not None and cached[0] == key:
            return cached[1]
        with _SERVING_LOCK:
            return self._serving_fleet_build(key, sc, est, topo, dev, feat, agg)

    def _serving_fleet_build(self, key, sc, est, topo, dev, feat, agg):
        import torch
        from gordo_b200.fleet import FFFleet
        cached = self.__dict__.get("_gb200_serving")
        if cached is not None and cached[0] == key:             # another thread built it while this one waited
            return cached[1]
        fleet = FFFleet(topo, 1, dev)
        fleet.set_params(torch.as_tensor(est.model.params[None], device=dev))
        f32 = lambda a: torch.as_tensor(np.array(a, np.float32)[None], device=dev)
        if sc is not None:
            fleet.in_scale, fleet.in_min = f32(sc.scale_), f32(sc.min_)
        fleet.err_scale = f32(self.scaler.scale_)
        if feat is not None:
            fleet.feat_thr = f32(np.asarray(feat, np.float64))
        if agg is not None:
            fleet.agg_thr = torch.as_tensor(np.array([agg], np.float32), device=dev)
        self.__dict__["_gb200_serving"] = (key, fleet)
        return fleet

    def __getstate__(self):
        # device-side caches never travel with the pickled model (serializer.dump, copy.deepcopy)
        state = BaseEstimator.__getstate__(self)
        return {k: v for k, v in state.items() if not k.startswith("_gb200_")}

    def _fused_columns_lstm(self, sc, est, Xv, yv, dev):
        """LSTM base: predict on the GPU (windows never materialised), then score the offset output."""
        import torch
        from gordo_b200.fleet import FFFleet, Schedule
        from gordo_b200.lstm import LSTMFleet
        topo = est.model.topology
        f32 = lambda a: torch.as_tensor(np.array(a, np.float32)[None], device=dev)
        parts = [est.model.params] + ([np.asarray(sc.scale_), np.asarray(sc.min_)] if sc is not None else [])
        key = (dev.index, id(est.model), est.lookahead, sc is None,
               hash(b"".join(np.ascontiguousarray(a).tobytes() for a in parts)))
        with _SERVING_LOCK:
            cached = self.__dict__.get("_gb200_serving_lstm")
            if cached is not None and cached[0] == key:
                fleet = cached[1]
            else:
                fleet = LSTMFleet(topo, 1, est.lookahead, dev)
                fleet.set_params(torch.as_tensor(est.model.params[None], device=dev))
                if sc is not None:
                    fleet.in_scale, fleet.in_min = f32(sc.scale_), f32(sc.min_)
                self.__dict__["_gb200_serving_lstm"] = (key, fleet)
        xd = torch.as_tensor(np.ascontiguousarray(Xv, np.float32), device=dev)
        yd = torch.as_tensor(np.ascontiguousarray(yv, np.float32), device=dev)
        prec = "bf16" if (est._precision == "bf16" and fleet.tc_eligible()) else "f32"
        out, off = fleet.predict(Schedule.single(len(Xv)), xd, precision=prec)
        n_out = int(off[-1])
        feat = self.__dict__.get("feature_thresholds_"); agg = self.__dict__.get("aggregate_threshold_")
        res = FFFleet.score_outputs(
            out, yd, off, [len(yv) - n_out], err_scale=f32(self.scaler.scale_),
            feat_thr=None if feat is None else f32(np.asarray(feat, np.float64)),
            agg_thr=None if agg is None else torch.as_tensor(np.array([agg], np.float32), device=dev))
        res["model-output"] = out
        return {k: v.cpu().numpy() for k, v in self._add_smooth_columns(res).items()}

    def anomaly(self, X: pd.DataFrame, y: pd.DataFrame, frequency: Optional[timedelta] = None) -> pd.DataFrame:
        """
        The anomaly frame: ``model-input``, ``model-output``, ``tag-anomaly-scaled``,
        ``total-anomaly-scaled``, ``tag-anomaly-unscaled``, ``total-anomaly-unscaled``, optional
        ``smooth-*``, ``anomaly-confidence``, ``total-anomaly-confidence`` (diff.py:310-458).
        """
        return model_utils.assemble_frame(self.anomaly_groups(X, y), getattr(X, "index", None), frequency)

    def anomaly_response(self, X: pd.DataFrame, y: pd.DataFrame, frequency: Optional[timedelta] = None,
                         fmt: str = "parquet", all_columns: bool = False):
        """
        The body of the server's anomaly response (gordo/server/blueprints/anomaly.py:57-72) straight from the column
        groups, without the DataFrame pivot: ``fmt="parquet"`` -> the bytes of ``dataframe_into_parquet_bytes(frame)``,
        ``fmt="json"`` -> the dict of ``dataframe_to_dict(frame)`` (the ``data`` member).  ``all_columns=False`` drops the
        smooth-* groups as the view does (anomaly.py:17-22, 57-62).
        """
        from gordo_b200.server import utils as server_utils
        groups = self.anomaly_groups(X, y)
        if not all_columns:
            groups = [g for g in groups if not g[0].startswith("smooth-")]
        index = getattr(X, "index", None)
        if fmt == "parquet":
            return server_utils.columns_into_parquet_bytes(groups, index, frequency)
        if fmt == "json":
            return server_utils.columns_to_dict(groups, index, frequency)
        raise ValueError("fmt must be 'parquet' or 'json'")

    def anomaly_groups(self, X: pd.DataFrame, y: pd.DataFrame):
        """The frame's column groups [(top-level name, values, second-level names)], in the reference's order."""
        if not hasattr(X, "values"):
            raise ValueError("Unable to find X.values property")
        has_feat = self.__dict__.get("feature_thresholds_") is not None
        has_agg = self.__dict__.get("aggregate_threshold_") is not None
        if self.require_thresholds and not (has_feat or has_agg):
            # the reference raises after computing the frame (diff.py:448-456); same error, sooner
            raise AttributeError(
                f"`require_thresholds={self.require_thresholds}` however `.cross_validate` needs to be called "
                f"in order to calculate thesethresholds before calling `.anomaly`")
        Xv = np.asarray(X.values)
        yv = np.asarray(y.values if hasattr(y, "values") else y)
        x_tags = list(X.columns)
        y_tags = list(y.columns) if hasattr(y, "columns") else x_tags
        plan = self._fused_plan()
        if plan is not None and Xv.shape[1] == _n_in(plan[1]) and yv.shape[1] == _n_out(plan[1]):
            cols = self._fused_columns(plan, Xv, yv)
            out = cols["model-output"]
            d_scaled, tot_scaled = cols["tag-anomaly-scaled"], cols["total-anomaly-scaled"]
            d_un, tot_un = cols["tag-anomaly-unscaled"], cols["total-anomaly-unscaled"]
            conf, tconf = cols.get("anomaly-confidence"), cols.get("total-anomaly-confidence")
            smooth = {k: v for k, v in cols.items() if k.startswith("smooth-")}
        else:
            smooth = None
            out = np.asarray(self.predict(X) if hasattr(self, "predict") else self.transform(X))
            n = len(out)
            out_names = y_tags if out.shape[1] == len(y_tags) else [str(i) for i in range(out.shape[1])]
            sc_out = np.asarray(self.scaler.transform(pd.DataFrame(out, columns=out_names)
                                                      if hasattr(self.scaler, "feature_names_in_") else out))
            d_scaled = np.abs(sc_out - np.asarray(self.scaler.transform(y))[-n:, :])
            tot_scaled = np.square(d_scaled).mean(axis=1)
            d_un = np.abs(out.astype(np.float64) - yv.astype(np.float64)[-n:, :])
            tot_un = np.square(d_un).mean(axis=1)
            conf = d_un / np.asarray(self.feature_thresholds_, np.float64) if has_feat else None
            tconf = tot_scaled / self.aggregate_threshold_ if has_agg else None
        n = len(out)
        in_names = model_utils._second_level(Xv, x_tags)
        out_names = model_utils._second_level(out, y_tags)
        groups = [("model-input", Xv[-n:], in_names), ("model-output", out, out_names),
                  ("tag-anomaly-scaled", d_scaled, out_names), ("total-anomaly-scaled", tot_scaled, None),
                  ("tag-anomaly-unscaled", d_un, [model_utils._tag_name(t) for t in y_tags]),
                  ("total-anomaly-unscaled", tot_un, None)]
        if self.window is not None and self.smoothing_method is not None:
            if not smooth:
                sm = lambda a: np.asarray(self._smoothing(pd.DataFrame(np.asarray(a, np.float64))))
                smooth = {"smooth-tag-anomaly-scaled": sm(d_scaled), "smooth-total-anomaly-scaled": sm(tot_scaled)[:, 0],
                          "smooth-tag-anomaly-unscaled": sm(d_un), "smooth-total-anomaly-unscaled": sm(tot_un)[:, 0]}
            groups += [("smooth-tag-anomaly-scaled", smooth["smooth-tag-anomaly-scaled"], out_names),
                       ("smooth-total-anomaly-scaled", smooth["smooth-total-anomaly-scaled"], None),
                       ("smooth-tag-anomaly-unscaled", smooth["smooth-tag-anomaly-unscaled"],
                        [model_utils._tag_name(t) for t in y_tags]),
                       ("smooth-total-anomaly-unscaled", smooth["smooth-total-anomaly-unscaled"], None)]
        if conf is not None:
            groups.append(("anomaly-confidence", conf, out_names))
        if tconf is not None:
            groups.append(("total-anomaly-confidence", tconf, None))
        return groups


def _n_in(est):
    t = est.model.topology
    return getattr(t, "n_in", None) or t.n_features


def _n_out(est):
    t = est.model.topology
    return getattr(t, "n_out", None) or t.n_features_out


def _as_like(y_pred, y_true):
    """Give predictions the target's column labels so a scaler fitted on a DataFrame does not warn."""
    if isinstance(y_true, pd.DataFrame) and not isinstance(y_pred, pd.DataFrame) and np.shape(y_pred)[1] == y_true.shape[1]:
        return pd.DataFrame(np.asarray(y_pred), columns=y_true.columns)
    return y_pred


class DiffBasedKFCVAnomalyDetector(DiffBasedAnomalyDetector):
    def __init__(self, base_estimator: BaseEstimator = None, scaler: TransformerMixin = None,
                 require_thresholds: bool = True, shuffle: bool = True, window: int = 144,
                 smoothing_method: str = "smm", threshold_percentile: float = 0.99):
        """
        K-fold variant: thresholds are a percentile of the smoothed validation errors assembled
        from all folds (diff.py:461-635).
        """
        self.base_estimator = base_estimator if base_estimator is not None else KerasAutoEncoder(kind="feedforward_hourglass")
        self.scaler = scaler if scaler is not None else MinMaxScaler()
        self.require_thresholds = require_thresholds
        self.window = window
        self.shuffle = shuffle
        self.smoothing_method = smoothing_method
        self.threshold_percentile = threshold_percentile

    def get_params(self, deep=True):
        return {"base_estimator": self.base_estimator, "scaler": self.scaler, "window": self.window,
                "smoothing_method": self.smoothing_method, "shuffle": self.shuffle,
                "threshold_percentile": self.threshold_percentile}

    def get_metadata(self):
        metadata = dict()
        if "feature_thresholds_" in self.__dict__:
            metadata["feature-thresholds"] = np.asarray(self.feature_thresholds_).tolist()
        if "aggregate_threshold_" in self.__dict__:
            metadata["aggregate-threshold"] = self.aggregate_threshold_
        if isinstance(self.base_estimator, GordoBase):
            metadata.update(self.base_estimator.get_metadata())
        else:
            metadata.update({"scaler": str(self.scaler), "base_estimator": str(self.base_estimator),
                             "shuffle": self.shuffle, "window": self.window,
                             "smoothing-method": self.smoothing_method,
                             "threshold-percentile": self.threshold_percentile})
        return metadata

    def cross_validate(self, *, X, y, cv=KFold(n_splits=5, shuffle=True, random_state=0), **kwargs):
        kwargs.update(dict(return_estimator=True, cv=cv))
        cv_output = c_val(self, X=X, y=y, **kwargs)
        y = pd.DataFrame(y)
        y_pred = pd.DataFrame(np.zeros_like(np.asarray(y, np.float64)), index=y.index, columns=y.columns)
        y_val_mse = pd.Series(np.nan, index=y.index, dtype=np.float64)
        for i, ((_, test_idxs), split_model) in enumerate(zip(kwargs["cv"].split(X, y), cv_output["estimator"])):
            y_pred.iloc[test_idxs] = split_model.predict(
                X.iloc[test_idxs].to_numpy() if isinstance(X, pd.DataFrame) else X[test_idxs])
            y_val_mse.iloc[test_idxs] = self._scaled_mse_per_timestep(
                split_model, y.iloc[test_idxs], y_pred.iloc[test_idxs]).to_numpy()
        self.aggregate_threshold_ = self._calculate_threshold(y_val_mse)
        self.feature_thresholds_ = self._calculate_feature_thresholds(y, y_pred)
        return cv_output

    def _calculate_feature_thresholds(self, y_true: pd.DataFrame, y_pred: pd.DataFrame):
        absolute_error = self._absolute_error(y_true, y_pred)
        return self._calculate_threshold(absolute_error)

    def _calculate_threshold(self, validation_metric):
        """percentile of the smoothed validation metric (diff.py:631-635): gb200_smooth -> gb200_quantile."""
        import torch
        from gordo_b200.fleet import FFFleet
        dev = torch.device("cuda", torch.cuda.current_device())
        v = torch.as_tensor(np.ascontiguousarray(np.asarray(validation_metric, np.float32)), device=dev)
        sm = self._smooth_device(v)
        lo = torch.zeros(1, dtype=torch.int64, device=dev)
        hi = torch.full((1,), v.shape[0], dtype=torch.int64, device=dev)
        q = FFFleet.quantile(sm, lo, hi, self.threshold_percentile)[0].cpu().numpy()
        if isinstance(validation_metric, pd.DataFrame):
            return pd.Series(q, index=validation_metric.columns, name=self.threshold_percentile)
        return float(q[0])
