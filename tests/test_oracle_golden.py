"""
The oracle against the reference's own golden vectors and against outputs of the REAL
reference code (tests/golden/make_golden.py).  CPU only.
"""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import factories, lstm
from oracle.scaler import MinMaxScaler, time_series_split
from oracle.anomaly import DiffDetector, anomaly_frame, rolling_min_max

G = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(G, "golden_meta.json")))
NPZ = np.load(os.path.join(G, "detector_golden.npz"))


# reference golden vectors: tests/gordo/machine/model/test_factories_utils.py:8-24
@pytest.mark.parametrize("args,expected", [
    ((0.2, 4, 5), (4, 3, 2, 1)), ((0.5, 3, 10), (8, 7, 5)), ((0.5, 3, 3), (3, 2, 2)),
    ((0.3, 3, 10), (8, 5, 3)), ((1, 3, 10), (10, 10, 10)), ((0, 3, 100000), (66667, 33334, 1))])
def test_hourglass_reference_vectors(args, expected):
    assert factories.hourglass_calc_dims(*args) == expected


def test_hourglass_against_real_reference_grid():
    for key, dims in META["hourglass"].items():
        cf, el, n = key.split("|")
        assert list(factories.hourglass_calc_dims(float(cf), int(el), int(n))) == dims, key


def test_hourglass_docstring_topologies():
    # feedforward_autoencoder.py:225-238
    assert factories.feedforward_hourglass(10)["widths"] == [10, 8, 7, 5, 5, 7, 8, 10]
    assert factories.feedforward_hourglass(5)["widths"] == [5, 4, 4, 3, 3, 4, 4, 5]
    assert factories.feedforward_hourglass(10, compression_factor=0.2)["widths"] == [10, 7, 5, 2, 2, 5, 7, 10]
    assert factories.feedforward_hourglass(10, encoding_layers=1)["widths"] == [10, 5, 5, 10]
    # l1 activity regulariser on encoder layers i >= 1 only (:78-81)
    assert factories.feedforward_hourglass(10)["l1"] == [0.0, 10e-5, 10e-5, 0.0, 0.0, 0.0, 0.0]
    with pytest.raises(ValueError):
        factories.hourglass_calc_dims(1.5, 3, 10)
    with pytest.raises(ValueError):
        factories.hourglass_calc_dims(0.5, 0, 10)
    with pytest.raises(ValueError):
        factories.feedforward_model(4, encoding_dim=(3, 2), encoding_func=("tanh",))


def test_minmax_scaler_against_sklearn():
    sc = MinMaxScaler().fit(NPZ["mm_in"])
    np.testing.assert_allclose(sc.scale_, NPZ["mm_scale"], rtol=1e-13)
    np.testing.assert_allclose(sc.min_, NPZ["mm_min"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(sc.transform(NPZ["mm_in"]), NPZ["mm_out"], rtol=1e-12, atol=1e-14)


def test_time_series_split_against_sklearn():
    for key, folds in META["time_series_split"].items():
        n, k = map(int, key.split("|"))
        got = [[int(tr[0]), int(tr[-1]), int(te[0]), int(te[-1])] for tr, te in time_series_split(n, k)]
        assert got == folds, key


# reference golden vectors: tests/gordo/machine/model/test_model.py:239-311
def test_windowing_reference_vectors():
    X = np.array([[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]); y = X.copy()
    b = lstm.timeseries_batches(X, y, 2, 3, 0)
    assert b[0][0].tolist() == [[[0, 1], [2, 3], [4, 5]], [[2, 3], [4, 5], [6, 7]]]
    assert b[0][1].tolist() == [[4, 5], [6, 7]]
    assert b[1][0].tolist() == [[[4, 5], [6, 7], [8, 9]]] and b[1][1].tolist() == [[8, 9]]
    b = lstm.timeseries_batches(X, y, 2, 2, 1)
    assert b[0][0].tolist() == [[[0, 1], [2, 3]], [[2, 3], [4, 5]]] and b[0][1].tolist() == [[4, 5], [6, 7]]
    assert b[1][0].tolist() == [[[4, 5], [6, 7]]] and b[1][1].tolist() == [[8, 9]]
    b = lstm.timeseries_batches(X, y, 2, 2, 2)
    assert b[0][0].tolist() == [[[0, 1], [2, 3]], [[2, 3], [4, 5]]] and b[0][1].tolist() == [[6, 7], [8, 9]]
    assert len(b) == 1                                   # "No more elements left"
    with pytest.raises(ValueError):
        lstm.window_index(1, 2, -1)
    # docstring models.py:753-768: len(gen) == 9 for 100 rows, lookback 20, batch 10
    assert len(lstm.timeseries_batches(np.zeros((100, 2)), np.zeros((100, 2)), 10, 20, 0)) == 9
    # output length: tests/gordo/machine/model/test_model.py:324-338, builder offsets
    # tests/gordo/builder/test_builder.py:99-115 (LSTM-AE L=10 -> 9, Forecast L=13 -> 13)
    assert lstm.window_count(4, 3, 0) == 2
    assert 100 - lstm.window_count(100, 10, 0) == 9 and 100 - lstm.window_count(100, 13, 1) == 13


class _OffsetLinear:
    """numpy twin of make_golden.OffsetLinear (OLS with intercept, drops leading rows)."""

    def __init__(self, offset):
        self.offset = offset

    def fit(self, X, y):
        A = np.hstack([np.asarray(X, float), np.ones((len(X), 1))])
        self.coef, *_ = np.linalg.lstsq(A, np.asarray(y, float), rcond=None)
        return self

    def predict(self, X):
        A = np.hstack([np.asarray(X, float), np.ones((len(X), 1))])
        return (A @ self.coef)[self.offset:]


@pytest.mark.parametrize("ci", range(len(META["cases"])))
def test_detector_against_real_reference(ci):
    case = META["cases"][ci]
    pre = f"c{ci}_"
    X, y = NPZ[pre + "X"], NPZ[pre + "y"]
    det = DiffDetector(lambda tag: _OffsetLinear(case["offset"]), window=case["window"],
                       smoothing_method=case["method"])
    det.cross_validate(X, y)
    det.fit(X, y)
    res = det.anomaly(X, y)
    tol = dict(rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(det.feature_thresholds_, NPZ[pre + "feature_thresholds"], **tol)
    np.testing.assert_allclose(det.aggregate_threshold_, NPZ[pre + "aggregate_threshold"], **tol)
    np.testing.assert_allclose(np.stack([det.feature_thresholds_per_fold_[f"fold-{i}"] for i in range(3)]),
                               NPZ[pre + "feature_thresholds_per_fold"], **tol)
    np.testing.assert_allclose([det.aggregate_thresholds_per_fold_[f"fold-{i}"] for i in range(3)],
                               NPZ[pre + "aggregate_thresholds_per_fold"], **tol)
    if case["window"] is not None:
        np.testing.assert_allclose(det.smooth_feature_thresholds_, NPZ[pre + "smooth_feature_thresholds"], **tol)
        np.testing.assert_allclose(det.smooth_aggregate_threshold_, NPZ[pre + "smooth_aggregate_threshold"], **tol)
    assert case["n_rows"] == len(res["model-output"]) == case["n"] - case["offset"]
    for grp in case["groups"]:
        want = NPZ[pre + "col_" + grp]
        got = np.asarray(res[grp], float)
        if want.ndim == 2 and want.shape[1] == 1 and got.ndim == 1:
            want = want[:, 0]
        np.testing.assert_allclose(got, want, equal_nan=True, **tol)
    # frame layout (model/utils.py:49-165 + the joins of diff.py)
    tags = [f"tag-{j}" for j in range(case["t"])]
    index = (pd.date_range("2019-01-01", periods=case["n"], freq="10min")
             if case["index"] == "dates" else pd.RangeIndex(case["n"]))
    frame = anomaly_frame(res, tags, index=index, frequency=pd.Timedelta(minutes=10))
    assert [list(map(str, c)) for c in frame.columns] == case["columns"]
    assert [None if s is None else str(s) for s in frame["start"].iloc[:3, 0].tolist()] == case["start"] \
        if frame["start"].ndim == 2 else True
    if case["index"] == "dates":
        assert frame[("end", "")].iloc[:3].tolist() == case["end"]
        assert frame[("start", "")].iloc[:3].tolist() == case["start"]


def test_require_thresholds_error():
    assert META["require_thresholds_raises"] is True
    r = np.random.default_rng(5); X = r.random((50, 2))
    det = DiffDetector(lambda tag: _OffsetLinear(0)).fit(X, X)
    with pytest.raises(AttributeError):
        det.anomaly(X, X)
    DiffDetector(lambda tag: _OffsetLinear(0), require_thresholds=False).fit(X, X).anomaly(X, X)


def test_rolling_min_max_matches_pandas():
    r = np.random.default_rng(0)
    for n in (3, 6, 7, 100):
        x = r.random((n, 3))
        want = pd.DataFrame(x).rolling(6).min().max().to_numpy()
        np.testing.assert_allclose(rolling_min_max(x, 6), want, equal_nan=True)
        w1 = pd.Series(x[:, 0]).rolling(6).min().max()
        got = rolling_min_max(x[:, 0], 6)
        assert (np.isnan(w1) and np.isnan(got)) or np.isclose(w1, got)


def test_kfold_and_shuffle_against_sklearn():
    from sklearn.model_selection import KFold
    from sklearn.utils import shuffle
    from oracle.scaler import kfold_split, shuffle_rows
    for n in (10, 11, 503, 1000):
        for k in (2, 5, 7):
            want = list(KFold(n_splits=k, shuffle=True, random_state=0).split(np.zeros((n, 1))))
            got = list(kfold_split(n, k, 0))
            assert len(want) == len(got) == k
            for (wtr, wte), (gtr, gte) in zip(want, got):
                np.testing.assert_array_equal(wtr, gtr)
                np.testing.assert_array_equal(wte, gte)
        np.testing.assert_array_equal(shuffle(np.arange(n), random_state=0), shuffle_rows(n, 0))


KNPZ = np.load(os.path.join(G, "kfcv_golden.npz"))


@pytest.mark.parametrize("ci", range(len(META["kfcv_cases"])))
def test_kfcv_detector_against_real_reference(ci):
    """DiffBasedKFCVAnomalyDetector of the real reference (diff.py:461-635) vs oracle.KFCVDetector."""
    from oracle.anomaly import KFCVDetector
    case = META["kfcv_cases"][ci]
    pre = f"k{ci}_"
    X, y = KNPZ[pre + "X"], KNPZ[pre + "y"]
    det = KFCVDetector(lambda tag: _OffsetLinear(0), window=case["window"], smoothing_method=case["method"],
                       threshold_percentile=case["percentile"])
    det.cross_validate(X, y)
    det.fit(X, y)
    res = det.anomaly(X, y)
    tol = dict(rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(det.feature_thresholds_, KNPZ[pre + "feature_thresholds"], **tol)
    np.testing.assert_allclose(det.aggregate_threshold_, KNPZ[pre + "aggregate_threshold"], **tol)
    assert set(case["groups"]) == set(res)
    for grp in case["groups"]:
        want = KNPZ[pre + "col_" + grp]
        got = np.asarray(res[grp], float)
        if want.ndim == 2 and want.shape[1] == 1 and got.ndim == 1:
            want = want[:, 0]
        np.testing.assert_allclose(got, want, equal_nan=True, **tol)


def test_builder_metric_scorers_match_the_real_reference():
    """`_metrics_dict` / `metric_wrapper` (the mirror of build_model.py:377-446 and model/utils.py:18-46) against values
    the REAL reference's `ModelBuilder.build_metrics_dict` produced (tests/golden/make_metrics_golden.py): same 20
    scorer names (' ' -> '-'), same values, with a model whose output is shorter than y (the LSTM offset)."""
    import json
    import os
    import warnings
    import pandas as pd
    from sklearn.base import BaseEstimator, RegressorMixin
    from gordo_b200.builder import _metrics_dict
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "builder_metrics_golden.json")))
    y = pd.DataFrame(np.asarray(g["y"]), columns=g["columns"])
    y_pred = np.asarray(g["y_pred"])

    class Dummy(RegressorMixin, BaseEstimator):
        def fit(self, X, y=None):
            return self

        def predict(self, X):
            return y_pred

    scorers = _metrics_dict(y)
    assert set(scorers) == set(g["values"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, s in scorers.items():
            np.testing.assert_allclose(float(s(Dummy(), y, y)), g["values"][k], rtol=1e-12, err_msg=k)


def test_build_split_dict_matches_the_real_reference():
    """Split metadata of the build (build_model.py:347-375) for TimeSeriesSplit(3) / KFold(4), time-indexed and
    range-indexed frames, and the (train_end, test_end) form the batched build uses."""
    import json
    import os
    import pandas as pd
    from sklearn.model_selection import KFold, TimeSeriesSplit
    from gordo_b200.builder import build_split_dict
    from gordo_b200.fleet import time_series_split_bounds
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "builder_metrics_golden.json")))
    y = pd.DataFrame(np.asarray(g["y"]), columns=g["columns"])
    Xdt = y.set_axis(pd.date_range("2020-01-01", periods=len(y), freq="10min", tz="UTC"))
    s = lambda d: {k: str(v) for k, v in d.items()}
    assert s(build_split_dict(Xdt, TimeSeriesSplit(n_splits=3))) == g["splits"]["tss3"]
    assert s(build_split_dict(Xdt, KFold(n_splits=4))) == g["splits"]["kfold4"]
    assert s(build_split_dict(y, TimeSeriesSplit(n_splits=3))) == g["splits"]["tss3_rangeindex"]
    assert s(build_split_dict(Xdt, bounds=time_series_split_bounds(len(y), 3))) == g["splits"]["tss3"]
    assert s(build_split_dict(y.to_numpy(), bounds=time_series_split_bounds(len(y), 3))) == g["splits"]["tss3_rangeindex"]


def _topology_cases():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "topology_golden.json")))


@pytest.mark.parametrize("ci", range(17))
def test_factories_build_the_topologies_the_real_reference_asks_keras_for(ci):
    """tests/golden/make_topology_golden.py ran the reference's own factories (feedforward_autoencoder.py:15-251,
    lstm_autoencoder.py:15-263) against recording Keras stand-ins; the mirror's factories must describe the same stack:
    layer widths, activations, which layers carry the L1 activity regulariser and its strength, the LSTM
    return_sequences pattern / input shapes, the output layer, Adam's configuration, the loss."""
    from gordo_b200.machine.model import factories as F      # noqa: F401  (registers the factories)
    from gordo_b200.machine.model.factories import feedforward_autoencoder as ffa, lstm_autoencoder as lsa
    case = _topology_cases()["cases"][ci]
    kw = {k: (tuple(v) if isinstance(v, list) else (dict(v) if isinstance(v, dict) else v)) for k, v in case["kwargs"].items()}
    fn = getattr(ffa, case["factory"], None) or getattr(lsa, case["factory"])
    loss = case["compile_kwargs"]["loss"]
    if loss not in ("mse", "mean_squared_error"):
        with pytest.raises(ValueError, match="not supported"):           # stated limitation: MSE only
            fn(**kw)
        kw.pop("compile_kwargs")
    topo = fn(**kw)
    layers = case["layers"]
    opt = case["compile_kwargs"]["optimizer"]["optimizers.get"]
    assert opt["class_name"] == "Adam"
    want_adam = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    for k, v in opt["config"].items():
        want_adam["lr" if k == "learning_rate" else k] = v
    assert topo.adam == pytest.approx(want_adam)
    if layers[0]["kind"] == "Dense":
        assert topo.widths == [layers[0]["input_dim"]] + [l["units"] for l in layers]
        assert topo.acts == [l["activation"] for l in layers]
        assert topo.l1 == pytest.approx([(l["activity_regularizer"] or {"l1": 0.0})["l1"] for l in layers])
        assert all(l["kind"] == "Dense" for l in layers)
    else:
        lstm = [l for l in layers if l["kind"] == "LSTM"]
        assert layers[-1]["kind"] == "Dense" and len(lstm) == len(layers) - 1
        assert topo.units == [l["units"] for l in lstm] and topo.acts == [l["activation"] for l in lstm]
        assert topo.out_func == layers[-1]["activation"] and topo.n_features_out == layers[-1]["units"]
        assert [topo.lookback_window, topo.n_features] == lstm[0]["input_shape"]
        # every LSTM layer returns sequences except the last one, whose last step feeds the Dense layer
        assert [l["return_sequences"] for l in lstm] == [True] * (len(lstm) - 1) + [False]


def test_factory_registry_matches_the_real_reference():
    from gordo_b200.machine.model import factories as F      # noqa: F401
    from gordo_b200.machine.model.register import register_model_builder
    got = {t: set(k) for t, k in register_model_builder.factories.items()}
    for typ, kinds in _topology_cases()["registered"].items():        # (other tests may have registered kinds of their own)
        assert set(kinds) <= got[typ], typ


@pytest.mark.parametrize("ci", range(17))
def test_oracle_factories_match_the_real_reference_topologies(ci):
    """The checker's own topology specs (oracle/factories.py) against the same reference-derived goldens."""
    from oracle import factories as OF
    case = _topology_cases()["cases"][ci]
    kw = {k: (tuple(v) if isinstance(v, list) else (dict(v) if isinstance(v, dict) else v)) for k, v in case["kwargs"].items()}
    spec = getattr(OF, case["factory"])(**kw)
    layers = case["layers"]
    assert spec["loss"] in (case["compile_kwargs"]["loss"],) or {spec["loss"], case["compile_kwargs"]["loss"]} <= {"mse", "mean_squared_error"}
    if layers[0]["kind"] == "Dense":
        assert spec["widths"] == [layers[0]["input_dim"]] + [l["units"] for l in layers]
        assert spec["acts"] == [l["activation"] for l in layers]
        assert spec["l1"] == pytest.approx([(l["activity_regularizer"] or {"l1": 0.0})["l1"] for l in layers])
    else:
        lstm = [l for l in layers if l["kind"] == "LSTM"]
        assert spec["units"] == [l["units"] for l in lstm] and spec["acts"] == [l["activation"] for l in lstm]
        assert spec["out_func"] == layers[-1]["activation"] and spec["n_features_out"] == layers[-1]["units"]
        assert [spec["lookback_window"], spec["n_features"]] == lstm[0]["input_shape"]


# ----------------------------------------------------------------------------- training / prediction protocol
def _protocol():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "protocol_golden.json")))


def _numbered(n, c):
    return np.arange(n, dtype=np.float64)[:, None] + np.arange(c)[None, :] / 100.0


@pytest.mark.parametrize("name", ["lstm_autoencoder_lb4_bs5_ep2", "lstm_forecast_lb4_bs5", "lstm_autoencoder_defaults",
                                  "lstm_forecast_lb1_fit_kwargs"])
def test_oracle_lstm_fit_feeds_the_batches_the_real_reference_feeds_keras(name, monkeypatch):
    """tests/golden/make_protocol_golden.py executed the reference's KerasLSTM*.fit / predict (models.py:557-660, 713-793)
    with a logging Keras model: one primer step on the first window, then `fit(generator, shuffle=False)` -- epochs x
    batches of consecutive windows -- and prediction over all windows.  The oracle's `lstm_fit` / `lstm_predict` (what
    the GPU kernels are compared with) must train on exactly those windows, in that order."""
    from oracle import factories as OF, lstm as olstm
    case = next(c for c in _protocol()["cases"] if c["name"] == name)
    primer, main, pred = case["log"]
    kw = case["estimator_kwargs"]
    L, bs = kw["lookback_window"], kw["batch_size"]
    lookahead = 1 if "forecast" in name else 0
    epochs = main["kwargs"].get("epochs", 1)
    assert primer["kwargs"]["epochs"] == 1 and main["kwargs"]["shuffle"] is False and main["generator_batch_size"] == bs
    assert not ({"validation_split", "batch_size"} & set(main["kwargs"]))     # only fit_generator_params reach the main fit
    X = _numbered(case["n_rows"], case["n_cols"]).astype(np.float32)
    spec = OF.lstm_model(case["n_cols"], lookback_window=L, encoding_dim=(3,), encoding_func=("tanh",), decoding_dim=(3,),
                         decoding_func=("tanh",))
    params = olstm.lstm_init(spec, np.random.default_rng(0))
    seen = []
    real = olstm.lstm_loss_and_grads

    def spy(spec_, params_, xw, yb):
        seen.append({"window_rows": np.rint(xw[:, :, 0]).astype(int).tolist(), "target_rows": np.rint(yb[:, 0]).astype(int).tolist()})
        return real(spec_, params_, xw, yb)
    monkeypatch.setattr(olstm, "lstm_loss_and_grads", spy)
    olstm.lstm_fit(spec, params, X, X.copy(), lookback_window=L, lookahead=lookahead, batch_size=bs, epochs=epochs)
    assert seen[0] == {"window_rows": primer["window_rows"], "target_rows": primer["target_rows"]}
    assert seen[1:] == main["batches"]
    # prediction: every window once, in order, 10 000 per batch; output rows = n - lookback + 1 - lookahead
    assert pred["generator_batch_size"] == 10000 and pred["kwargs"] == {"verbose": 0}
    starts, tgt = olstm.window_index(case["predict_rows"], L, lookahead)
    assert [w[0] for b in pred["batches"] for w in b["window_rows"]] == starts.tolist()
    assert [t for b in pred["batches"] for t in b["target_rows"]] == tgt.tolist()
    assert case["predict_out_rows"] == olstm.window_count(case["predict_rows"], L, lookahead)


def test_estimator_kwargs_routing_and_errors_match_the_real_reference():
    """What reaches Keras from the constructor / fit kwargs of the feed-forward estimator, the kwargs the LSTM estimators
    keep, and the two ValueErrors -- against the reference's own estimator classes run with logging stand-ins."""
    from gordo_b200.machine.model.models import KerasAutoEncoder, KerasLSTMAutoEncoder, KerasLSTMForecast
    g = _protocol()
    ff = next(c for c in g["cases"] if c["name"] == "ff_autoencoder")["log"]
    assert ff[0]["kwargs"] == {"epochs": 3, "batch_size": 16, "validation_split": 0.1, "shuffle": False, "verbose": 0}
    assert set(ff[0]["kwargs"]) <= set(KerasAutoEncoder.supported_fit_args)
    dflt = next(c for c in g["cases"] if c["name"] == "ff_autoencoder_defaults")["log"]
    assert dflt[0]["kwargs"] == {"epochs": 2, "verbose": 0}                      # Keras' own defaults otherwise: batch 32, shuffle
    for cls, name in ((KerasLSTMAutoEncoder, "lstm_autoencoder_lb4_bs5_ep2"), (KerasLSTMForecast, "lstm_forecast_lb4_bs5")):
        want = {k: v for k, v in next(c for c in g["cases"] if c["name"] == name)["estimator_kwargs"].items()
                if k not in ("n_features", "n_features_out")}
        est = cls(kind="lstm_hourglass", **{k: v for k, v in want.items()})
        assert {k: est.kwargs[k] for k in want} == want and est.lookback_window == want["lookback_window"]
    with pytest.raises(ValueError) as e:
        KerasAutoEncoder(kind="no_such_kind")
    assert str(e.value) == g["errors"]["unknown_kind"]
    assert g["errors"]["lookback_ge_rows"] == "For KerasLSTMForecast lookback_window must be < size of X"
    import inspect
    import gordo_b200.machine.model.models as mirror
    assert g["errors"]["lookback_ge_rows"] in inspect.getsource(mirror)           # raised by the mirror before any device call


def test_model_metadata_extraction_matches_the_real_reference():
    """`extract_model_metadata` against `ModelBuilder._extract_metadata_from_model` (build_model.py:516-570) run from
    /root/reference on the same nested structures: a bare estimator, a Pipeline (only its LAST step is looked at), a
    detector over a Pipeline / over an estimator holding another estimator / over a TransformedTargetRegressor (the
    unfitted `regressor` attribute is skipped, the fitted clone is found), inner keys overriding outer ones."""
    import json
    import os
    import sys
    from gordo_b200.builder import extract_model_metadata
    from gordo_b200.machine.model.base import GordoBase
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from metadata_structures import build_structures
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "builder_metrics_golden.json")))
    got = {name: extract_model_metadata(model) for name, model in build_structures(GordoBase).items()}
    assert got == g["extracted_metadata"]
