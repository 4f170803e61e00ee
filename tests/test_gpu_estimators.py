"""
-m gpu tests of the reference-facing surface: the estimators, the detector and the fleet builder
(all arithmetic through libgordo_b200.so) against the oracle on the same seeded inputs.
They read like the reference's own tests: test_model.py (fit/predict/pickle), test_builder.py
(seed determinism, offsets), test_anomaly_detectors.py (the scoring contract).
"""
import pickle
from datetime import timedelta

import numpy as np
import pandas as pd
import pytest
import torch
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler, RobustScaler

from oracle import dense, factories
from oracle.anomaly import DiffDetector, FFBase, rolling_min_max
from oracle.scaler import MinMaxScaler as OMinMax, time_series_split

pytestmark = pytest.mark.gpu



def _builder_init(init_fn, seed, n_machines, per):
    """The initial weights FleetBuild gives job i of every Machine: its own generator, seeded like the per-Machine path."""
    from gordo_b200.builder import job_seeds
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    rows = []
    for _ in range(n_machines):
        for sd in job_seeds(seed, per):
            gen.manual_seed(sd)
            rows.append(init_fn(1, gen, dev)[0].cpu().numpy())
    return np.stack(rows)

def _data(seed, n, T):
    rng = np.random.default_rng(seed)
    Z = np.cumsum(rng.normal(size=(n, 2)), axis=0) * 0.05
    X = (1 / (1 + np.exp(-(Z @ rng.normal(size=(2, T))))) + 0.05 * rng.random((n, T))).astype(np.float32)
    return X * rng.uniform(1, 30, T).astype(np.float32) + rng.uniform(-4, 4, T).astype(np.float32)


def test_autoencoder_fit_predict_matches_oracle_and_pickles():
    from gordo_b200.machine.model.models import KerasAutoEncoder, _Model
    X = _data(0, 640, 8); Xs = OMinMax().fit(X).transform(X).astype(np.float32)
    spec = factories.feedforward_hourglass(8)
    init = dense.ff_init(spec, np.random.default_rng(1))
    est = KerasAutoEncoder(kind="feedforward_hourglass", epochs=2, shuffle=False)
    est.kwargs.update(n_features=8, n_features_out=8)
    est.model = _Model(est._topology(), dense.ff_flatten(init))      # injected initial weights
    est.fit(Xs, Xs)
    p = dense.ff_unflatten(dense.ff_flatten(init), spec["widths"])
    hist, _ = dense.ff_fit(spec, p, Xs, Xs, epochs=2, batch_size=32, perms=None)
    np.testing.assert_allclose(est.model.params, dense.ff_flatten(p), atol=3e-4)
    md = est.get_metadata()["history"]
    np.testing.assert_allclose(md["loss"], hist["loss"], rtol=3e-4)
    assert set(md) >= {"loss", "accuracy", "params"} and md["params"]["epochs"] == 2     # test_builder.py:58-63
    out = est.predict(Xs)
    assert out.shape == (640, 8) and out.dtype == np.float32
    np.testing.assert_allclose(out, dense.ff_forward(spec, dense.ff_unflatten(est.model.params, spec["widths"]), Xs), atol=2e-5)
    # pickle round trip reproduces predictions + history (test_model.py:112-158)
    clone = pickle.loads(pickle.dumps(est))
    assert np.allclose(clone.predict(Xs), out) and clone.get_metadata() == est.get_metadata()
    # bf16 tensor-core inference is selectable per model
    est.kwargs["precision"] = "bf16"
    np.testing.assert_allclose(est.predict(Xs), out, atol=3e-2)
    assert est.score(Xs, Xs) <= 1.0
    with pytest.raises(ValueError):
        est.predict(Xs[:, :5])


def test_seed_determinism_and_dataframe_inputs():
    # tests/gordo/builder/test_builder.py:658-707: same seed => identical result
    from gordo_b200.machine.model.models import KerasAutoEncoder
    X = pd.DataFrame(_data(2, 300, 5), columns=list("abcde"))
    runs = []
    for seed in (7, 7, 8):
        np.random.seed(seed)
        runs.append(KerasAutoEncoder(kind="feedforward_hourglass", epochs=1).fit(X, X).model.params)
    assert np.array_equal(runs[0], runs[1]) and not np.array_equal(runs[0], runs[2])
    # validation_split holds out the LAST rows and reports val_loss per epoch
    np.random.seed(0)
    est = KerasAutoEncoder(kind="feedforward_hourglass", epochs=2, validation_split=0.2).fit(X, X)
    h = est.get_metadata()["history"]
    assert len(h["val_loss"]) == 2 and h["params"]["steps"] == -(-240 // 32)
    with pytest.raises(NotImplementedError):
        KerasAutoEncoder(kind="feedforward_hourglass", callbacks=[{"x": 1}]).fit(X, X)


def _oracle_columns(est_params, spec, sx, sy, X, y, feat_thr=None, agg_thr=None):
    xs = sx.transform(X).astype(np.float32)
    yhat = dense.ff_forward(spec, dense.ff_unflatten(est_params, spec["widths"]), xs)
    d = np.abs(yhat.astype(np.float64) - y)
    s = np.abs(sy.transform(yhat) - sy.transform(y))
    return yhat, d, s


@pytest.mark.parametrize("index_kind", ["range", "dates"])
def test_detector_full_flow_matches_oracle(index_kind):
    """cross_validate -> fit -> anomaly on the standard Pipeline: thresholds and every column."""
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from gordo_b200.machine.model.models import KerasAutoEncoder
    n, T = 600, 6
    Xa = _data(3, n, T).astype(np.float64)
    tags = [f"tag-{i}" for i in range(T)]
    idx = pd.date_range("2020-01-01", periods=n, freq="10min") if index_kind == "dates" else pd.RangeIndex(n)
    X = pd.DataFrame(Xa, columns=tags, index=idx); y = X.copy()
    np.random.seed(11)
    det = DiffBasedAnomalyDetector(base_estimator=Pipeline([("s", MinMaxScaler()),
                                   ("m", KerasAutoEncoder(kind="feedforward_hourglass", epochs=1))]))
    with pytest.raises(AttributeError):                      # test_anomaly_detectors.py:767-796
        DiffBasedAnomalyDetector(base_estimator=det.base_estimator).fit(X, y).anomaly(X, y)
    cv = det.cross_validate(X=X, y=y)
    assert {"fit_time", "score_time", "test_score", "estimator"} <= set(cv)     # :735-764
    det.fit(X, y)
    spec = factories.feedforward_hourglass(T)
    # thresholds recomputed by the oracle from the fold models the detector trained
    for i, ((tr, te), fold) in enumerate(zip(time_series_split(n, 3), cv["estimator"])):
        sx = OMinMax().fit(Xa[tr]); sy = OMinMax().fit(Xa[tr])
        est = fold.base_estimator.steps[1][1]
        yhat, d, s = _oracle_columns(est.model.params, spec, sx, sy, Xa[te], Xa[te])
        np.testing.assert_allclose(det.feature_thresholds_per_fold_.loc[f"fold-{i}"].to_numpy(),
                                   rolling_min_max(d, 6), rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(det.aggregate_thresholds_per_fold_[f"fold-{i}"],
                                   rolling_min_max((s ** 2).mean(axis=1), 6), rtol=1e-3)
    assert det.feature_thresholds_.name == "fold-2" and len(det.feature_thresholds_) == T
    frame = det.anomaly(X, y, frequency=timedelta(minutes=10))
    est = det.base_estimator.steps[1][1]
    sx = OMinMax().fit(Xa); sy = OMinMax().fit(Xa)
    yhat, d, s = _oracle_columns(est.model.params, spec, sx, sy, Xa, Xa)
    tol = dict(rtol=2e-4, atol=4e-5)
    np.testing.assert_allclose(frame["model-output"].to_numpy(), yhat, atol=2e-5)
    np.testing.assert_allclose(frame["tag-anomaly-unscaled"].to_numpy(), d, **tol)
    np.testing.assert_allclose(frame["tag-anomaly-scaled"].to_numpy(), s, **tol)
    np.testing.assert_allclose(frame["total-anomaly-unscaled"].to_numpy(), (d ** 2).mean(1), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(frame["total-anomaly-scaled"].to_numpy(), (s ** 2).mean(1), rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(frame["anomaly-confidence"].to_numpy(), d / det.feature_thresholds_.to_numpy(), rtol=2e-4, atol=1e-3)
    np.testing.assert_allclose(frame["total-anomaly-confidence"].to_numpy(),
                               (s ** 2).mean(1) / det.aggregate_threshold_, rtol=2e-3, atol=1e-5)
    np.testing.assert_array_equal(frame["model-input"].to_numpy(), Xa)
    top = list(dict.fromkeys(c[0] for c in frame.columns))
    assert top == ["start", "end", "model-input", "model-output", "tag-anomaly-scaled", "total-anomaly-scaled",
                   "tag-anomaly-unscaled", "total-anomaly-unscaled", "anomaly-confidence", "total-anomaly-confidence"]
    if index_kind == "dates":
        assert frame[("start", "")].iloc[0] == "2020-01-01T00:00:00" and frame[("end", "")].iloc[0] == "2020-01-01T00:10:00"
    md = det.get_metadata()
    # base_estimator is a Pipeline (not a GordoBase): the reference then reports its repr, not a history
    # (diff.py:131-141; ModelBuilder walks the pipeline steps itself, build_model.py:515-569)
    assert {"feature-thresholds", "aggregate-threshold", "feature-thresholds-per-fold",
            "aggregate-thresholds-per-fold", "scaler", "base_estimator", "shuffle"} <= set(md)
    assert "history" in det.base_estimator.steps[1][1].get_metadata()
    # the fused launch and the generic host-side path agree (RobustScaler forces the generic one
    # for the scaled columns; unscaled columns must be identical up to float32 rounding)
    det_g = DiffBasedAnomalyDetector(base_estimator=det.base_estimator, scaler=RobustScaler(), require_thresholds=False)
    det_g.scaler.fit(y)
    fg = det_g.anomaly(X, y)
    np.testing.assert_allclose(fg["tag-anomaly-unscaled"].to_numpy(), frame["tag-anomaly-unscaled"].to_numpy(), rtol=1e-5, atol=1e-6)
    assert "anomaly-confidence" not in fg.columns


def test_smoothing_columns_and_kfcv_detector():
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector
    from gordo_b200.machine.model.models import KerasAutoEncoder
    X = pd.DataFrame(_data(4, 400, 4)); y = X.copy()
    np.random.seed(0)
    det = DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass"), window=12,
                                   smoothing_method="sma")
    det.cross_validate(X=X, y=y); det.fit(X, y)
    f = det.anomaly(X, y)
    want = f["total-anomaly-scaled"].rolling(12).mean().to_numpy().ravel()
    np.testing.assert_allclose(f["smooth-total-anomaly-scaled"].to_numpy().ravel(), want, equal_nan=True, rtol=1e-6)
    assert det.smooth_aggregate_threshold_ is not None and len(det.smooth_feature_thresholds_) == 4
    k = DiffBasedKFCVAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass"), window=12)
    k.cross_validate(X=X, y=y); k.fit(X, y)
    assert np.isfinite(k.aggregate_threshold_) and len(k.feature_thresholds_) == 4
    assert "total-anomaly-confidence" in k.anomaly(X, y).columns


def test_fleet_builder_matches_oracle_full_build():
    """FleetModelBuilder (batched CV + fit + thresholds for a bucket of Machines) vs the oracle."""
    from gordo_b200.builder import FleetBuild, FleetMachine, FleetModelBuilder
    from gordo_b200.fleet import FFTopology
    T, rows = 6, [512, 480, 333]
    Xs = [_data(20 + i, n, T) for i, n in enumerate(rows)]
    defn = {"gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "shuffle": False}}]}}}}
    mcs = [FleetMachine(f"m{i}", pd.DataFrame(X, columns=[f"t{j}" for j in range(T)]), model=defn,
                        evaluation={"seed": 5}) for i, X in enumerate(Xs)]
    builder = FleetBuild(mcs)
    built = builder.build()
    spec = factories.feedforward_hourglass(T)
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    init = _builder_init(topo.glorot_init, 5, len(rows), 4)
    for m, ((model, meta), X) in enumerate(zip(built, Xs)):
        Xd = X.astype(np.float64)
        det = DiffDetector(lambda tag, m=m: FFBase(spec, dense.ff_unflatten(
            init[m * 4 + (3 if tag == "final" else int(tag[-1]))], spec["widths"]), perms=None))
        det.cross_validate(Xd, Xd); det.fit(Xd, Xd)
        est = model.base_estimator.steps[1][1]
        np.testing.assert_allclose(est.model.params, dense.ff_flatten(det.base.params), atol=3e-4)
        np.testing.assert_allclose(model.feature_thresholds_.to_numpy(), det.feature_thresholds_, rtol=5e-3, atol=1e-5)
        np.testing.assert_allclose(model.aggregate_threshold_, det.aggregate_threshold_, rtol=5e-3)
        for i in range(3):
            np.testing.assert_allclose(model.feature_thresholds_per_fold_.loc[f"fold-{i}"].to_numpy(),
                                       det.feature_thresholds_per_fold_[f"fold-{i}"], rtol=5e-3, atol=1e-5)
        np.testing.assert_allclose(model.scaler.scale_, det.scaler.scale_, rtol=1e-6)
        np.testing.assert_allclose(model.base_estimator.steps[0][1].min_, OMinMax().fit(Xd).min_, rtol=1e-5, atol=1e-6)
        assert meta["model_offset"] == 0 and meta["fleet"]["fit_jobs"] == 12          # test_builder.py:99-115 (FF -> 0)
        frame = model.anomaly(mcs[m].X, mcs[m].X)
        want = det.anomaly(Xd, Xd)
        np.testing.assert_allclose(frame["total-anomaly-confidence"].to_numpy().ravel(), want["total-anomaly-confidence"], rtol=2e-2, atol=1e-4)
        assert pickle.loads(pickle.dumps(model)).aggregate_threshold_ == model.aggregate_threshold_


def test_fleet_builder_cv_scores_match_sklearn_scorers():
    """The four builder metrics (build_model.py:377-446) per tag / averaged / per fold, from gb200_cv_sums."""
    from sklearn import metrics as skm
    from sklearn.preprocessing import MinMaxScaler as SkMinMax
    from gordo_b200.builder import FleetBuild, FleetMachine, FleetModelBuilder
    from gordo_b200.machine.model.utils import metric_wrapper
    T, n = 5, 400
    X = pd.DataFrame(_data(31, n, T), columns=[f"tag {j}" for j in range(T)])
    built = FleetBuild([FleetMachine("m0", X, evaluation={"seed": 3})]).build()
    model, meta = built[0]
    scores = meta["cross_validation"]["scores"]
    assert set(scores["r2-score"]) == {"fold-mean", "fold-std", "fold-max", "fold-min", "fold-1", "fold-2", "fold-3"}
    assert "explained-variance-score-tag-0" in scores and "mean-absolute-error-tag-4" in scores     # ' ' -> '-'
    # recompute with the reference's scorer construction on the fold models' predictions: the fold models
    # are not kept by the batched build, so check the algebra on the FINAL model over the last test fold
    # through the same kernel instead, and the fold bookkeeping separately
    import torch
    from gordo_b200.fleet import FFFleet
    Xa = X.to_numpy(np.float64)
    yhat = model.predict(X)
    scaler = SkMinMax().fit(Xa)
    lo = torch.tensor([300], device="cuda:0"); hi = torch.tensor([400], device="cuda:0")
    got = FFFleet.cv_scores(torch.as_tensor(np.ascontiguousarray(Xa, np.float32), device="cuda:0"),
                            torch.as_tensor(np.ascontiguousarray(yhat), device="cuda:0"), lo, hi, scaler.scale_[None])
    for name, fn in (("explained-variance-score", skm.explained_variance_score), ("r2-score", skm.r2_score),
                     ("mean-squared-error", skm.mean_squared_error), ("mean-absolute-error", skm.mean_absolute_error)):
        want_all = metric_wrapper(fn, scaler=scaler)(Xa[300:400], yhat[300:400].astype(np.float64))
        np.testing.assert_allclose(got[name][0].mean(), want_all, rtol=1e-5, atol=1e-9, err_msg=name)
        for j in range(T):
            want = fn(scaler.transform(Xa[300:400])[:, j], scaler.transform(yhat[300:400].astype(np.float64))[:, j])
            np.testing.assert_allclose(got[name][0, j], want, rtol=1e-5, atol=1e-9, err_msg=f"{name} tag {j}")
    m = scores["mean-squared-error"]
    np.testing.assert_allclose(m["fold-mean"], np.mean([m["fold-1"], m["fold-2"], m["fold-3"]]))
    sp = meta["cross_validation"]["splits"]                        # build_model.py:347-375: 6 entries per fold
    assert {k: sp[k] for k in sp if k.endswith("n-train")} == {"fold-1-n-train": 100, "fold-2-n-train": 200, "fold-3-n-train": 300}
    assert sp["fold-2-n-test"] == 100 and sp["fold-3-test-end"] == 399 and sp["fold-1-train-end"] == 99 and len(sp) == 18


def test_validation_split_val_loss_and_early_stopping():
    """val_loss (MSE + activity loss, batch-weighted) vs the oracle; Keras-3 EarlyStopping semantics."""
    from gordo_b200.machine.model.models import KerasAutoEncoder, _Model, EarlyStopping
    X = _data(9, 500, 6); Xs = OMinMax().fit(X).transform(X).astype(np.float32)
    spec = factories.feedforward_hourglass(6)
    init = dense.ff_flatten(dense.ff_init(spec, np.random.default_rng(2)))
    est = KerasAutoEncoder(kind="feedforward_hourglass", epochs=3, shuffle=False, validation_split=0.2)
    est.kwargs.update(n_features=6, n_features_out=6)
    est.model = _Model(est._topology(), init.copy())
    est.fit(Xs, Xs)
    p = dense.ff_unflatten(init, spec["widths"])
    hist, _ = dense.ff_fit(spec, p, Xs, Xs, epochs=3, batch_size=32, perms=None, validation_split=0.2)
    h = est.get_metadata()["history"]
    np.testing.assert_allclose(h["loss"], hist["loss"], rtol=3e-4)
    np.testing.assert_allclose(h["val_loss"], hist["val_loss"], rtol=3e-4)
    # EarlyStopping from a Machine-YAML style definition: a huge min_delta means "never an improvement"
    # after the first epoch -> stop at epoch index `patience`, restore the first epoch's weights
    cb = [{"tensorflow.keras.callbacks.EarlyStopping": {"monitor": "val_loss", "patience": 2, "min_delta": 10.0,
                                                        "restore_best_weights": True}}]
    es = KerasAutoEncoder(kind="feedforward_hourglass", epochs=50, shuffle=False, validation_split=0.2, callbacks=cb)
    es.kwargs.update(n_features=6, n_features_out=6)
    es.model = _Model(es._topology(), init.copy())
    es.fit(Xs, Xs)
    he = es.get_metadata()["history"]
    assert len(he["loss"]) == 3 and len(he["val_loss"]) == 3          # epochs 0,1,2 then wait(2) >= patience
    one = KerasAutoEncoder(kind="feedforward_hourglass", epochs=1, shuffle=False, validation_split=0.2)
    one.kwargs.update(n_features=6, n_features_out=6)
    one.model = _Model(one._topology(), init.copy())
    one.fit(Xs, Xs)
    np.testing.assert_allclose(es.model.params, one.model.params, atol=1e-7)     # best = epoch 0 restored
    assert isinstance(EarlyStopping(monitor="val_accuracy").mode, str) and EarlyStopping(monitor="val_accuracy").mode == "max"
    with pytest.raises(NotImplementedError):
        KerasAutoEncoder(kind="feedforward_hourglass", callbacks=["keras.callbacks.TensorBoard"]).fit(Xs, Xs)


def test_fleet_builder_smooth_thresholds_with_window():
    """Detectors with ``window`` stay in the batched build: smooth thresholds = rolling(window).min().max() of the fold errors."""
    from gordo_b200.builder import FleetBuild, FleetMachine, FleetModelBuilder
    from gordo_b200.fleet import FFTopology
    T, rows, W = 5, [420, 377], 12
    Xs = [_data(50 + i, n, T) for i, n in enumerate(rows)]
    defn = {"gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"window": W, "smoothing_method": "sma", "base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "shuffle": False}}]}}}}
    mcs = [FleetMachine(f"m{i}", pd.DataFrame(X, columns=[f"t{j}" for j in range(T)]), model=defn,
                        evaluation={"seed": 9}) for i, X in enumerate(Xs)]
    built = FleetBuild(mcs).build()
    spec = factories.feedforward_hourglass(T)
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    init = _builder_init(topo.glorot_init, 9, len(rows), 4)
    for m, ((model, meta), X) in enumerate(zip(built, Xs)):
        assert meta["fleet"]["machines_in_launch"] == 2               # batched, not the per-Machine fallback
        Xd = X.astype(np.float64)
        det = DiffDetector(lambda tag, m=m: FFBase(spec, dense.ff_unflatten(
            init[m * 4 + (3 if tag == "final" else int(tag[-1]))], spec["widths"]), perms=None),
            window=W, smoothing_method="sma")
        det.cross_validate(Xd, Xd); det.fit(Xd, Xd)
        np.testing.assert_allclose(model.smooth_feature_thresholds_.to_numpy(), det.smooth_feature_thresholds_, rtol=5e-3, atol=1e-5)
        np.testing.assert_allclose(model.smooth_aggregate_threshold_, det.smooth_aggregate_threshold_, rtol=5e-3)
        assert list(model.smooth_feature_thresholds_per_fold_.index) == ["fold-0", "fold-1", "fold-2"]
        assert "smooth-aggregate-threshold" in model.get_metadata()
        frame = model.anomaly(mcs[m].X, mcs[m].X)
        want = det.anomaly(Xd, Xd)
        np.testing.assert_allclose(frame["smooth-total-anomaly-scaled"].to_numpy().ravel(),
                                   want["smooth-total-anomaly-scaled"], rtol=2e-2, atol=1e-6, equal_nan=True)


def test_serving_cache_follows_the_model():
    """The device-side copy kept between .anomaly() calls is rebuilt when thresholds, scalers or weights change."""
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from gordo_b200.machine.model.models import KerasAutoEncoder
    X = pd.DataFrame(_data(60, 300, 4), columns=list("abcd")); X2 = pd.DataFrame(_data(61, 300, 4) * 3 + 1, columns=list("abcd"))
    np.random.seed(1)
    det = DiffBasedAnomalyDetector(base_estimator=Pipeline([("s", MinMaxScaler()), ("m", KerasAutoEncoder(kind="feedforward_hourglass"))]))
    det.cross_validate(X=X, y=X); det.fit(X, X)
    f1 = det.anomaly(X, X); f1b = det.anomaly(X, X)
    pd.testing.assert_frame_equal(f1, f1b)
    assert "_gb200_serving" in det.__dict__
    det.aggregate_threshold_ = det.aggregate_threshold_ * 2.0
    f2 = det.anomaly(X, X)
    np.testing.assert_allclose(f2["total-anomaly-confidence"].to_numpy(), f1["total-anomaly-confidence"].to_numpy() / 2.0, rtol=1e-6)
    np.testing.assert_array_equal(f2["model-output"].to_numpy(), f1["model-output"].to_numpy())
    det.fit(X2, X2)                                  # new weights and new scalers
    f3 = det.anomaly(X, X)
    assert np.abs(f3["model-output"].to_numpy() - f1["model-output"].to_numpy()).max() > 1e-3
    clone = pickle.loads(pickle.dumps(det))
    assert "_gb200_serving" not in clone.__dict__ and "_gb200_serving" not in clone.base_estimator.steps[1][1].__dict__
    pd.testing.assert_frame_equal(clone.anomaly(X, X), f3)
    # a different frame length reuses the cached weights with another schedule
    assert len(det.anomaly(X.iloc[:37], X.iloc[:37])) == 37


def test_fleet_builder_batches_kfcv_detectors():
    """DiffBasedKFCVAnomalyDetector Machines in the batched build: builder folds (TimeSeriesSplit), thresholds =
    percentile of the smoothed errors over all rows, uncovered rows keeping zero predictions (diff.py:580-635)."""
    from gordo_b200.builder import FleetBuild, FleetMachine, FleetModelBuilder
    from gordo_b200.fleet import FFTopology
    from gordo_b200.machine.model.anomaly.diff import DiffBasedKFCVAnomalyDetector
    from oracle.anomaly import KFCVDetector
    T, rows, W = 5, [400, 333], 12
    Xs = [_data(70 + i, n, T) for i, n in enumerate(rows)]
    defn = {"gordo_b200.machine.model.anomaly.diff.DiffBasedKFCVAnomalyDetector": {
        "window": W, "smoothing_method": "sma", "threshold_percentile": 0.95, "shuffle": False, "base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "shuffle": False}}]}}}}
    mcs = [FleetMachine(f"m{i}", pd.DataFrame(X, columns=[f"t{j}" for j in range(T)]), model=defn,
                        evaluation={"seed": 4}) for i, X in enumerate(Xs)]
    built = FleetBuild(mcs).build()
    spec = factories.feedforward_hourglass(T)
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    init = _builder_init(topo.glorot_init, 4, len(rows), 4)
    for m, ((model, meta), X) in enumerate(zip(built, Xs)):
        assert type(model) is DiffBasedKFCVAnomalyDetector and meta["fleet"]["machines_in_launch"] == 2
        Xd = X.astype(np.float64)
        det = KFCVDetector(lambda tag, m=m: FFBase(spec, dense.ff_unflatten(
            init[m * 4 + (3 if tag == "final" else int(tag[-1]))], spec["widths"]), perms=None),
            window=W, smoothing_method="sma", threshold_percentile=0.95, shuffle=False)
        det.cross_validate(Xd, Xd, splits=list(time_series_split(len(Xd), 3)))
        det.fit(Xd, Xd)
        np.testing.assert_allclose(model.feature_thresholds_.to_numpy(), det.feature_thresholds_, rtol=5e-3, atol=1e-5)
        np.testing.assert_allclose(model.aggregate_threshold_, det.aggregate_threshold_, rtol=5e-3)
        assert sorted(model.get_metadata()) == sorted(["aggregate-threshold", "feature-thresholds", "base_estimator", "scaler",
                                                       "shuffle", "smoothing-method", "threshold-percentile", "window"])
        frame = model.anomaly(mcs[m].X, mcs[m].X)
        want = det.anomaly(Xd, Xd)
        np.testing.assert_allclose(frame["total-anomaly-confidence"].to_numpy().ravel(), want["total-anomaly-confidence"], rtol=2e-2, atol=1e-5)


def test_cv_sums_kernel_matches_the_real_reference_metrics():
    """gb200_cv_sums -> explained variance / r2 / MSE / MAE per tag and averaged, against the values the REAL reference's
    scorers (build_model.py:377-446, executed from /root/reference by tests/golden/make_metrics_golden.py) gave for the
    same y and a prediction that is `offset` rows shorter."""
    import json, os
    import torch
    from gordo_b200.fleet import FFFleet
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "builder_metrics_golden.json")))
    y = np.asarray(g["y"]); yp = np.asarray(g["y_pred"]); off = g["offset"]
    full = np.zeros_like(y); full[off:] = yp                      # row-aligned with y; rows < offset are outside the job
    scale = 1.0 / (y.max(0) - y.min(0))                           # MinMaxScaler fitted on the full y (the scoring scaler)
    lo = torch.tensor([off], device="cuda:0"); hi = torch.tensor([len(y)], device="cuda:0")
    got = FFFleet.cv_scores(torch.as_tensor(y.astype(np.float32), device="cuda:0"),
                            torch.as_tensor(full.astype(np.float32), device="cuda:0"), lo, hi, scale[None])
    for name in ("explained-variance-score", "r2-score", "mean-squared-error", "mean-absolute-error"):
        for j, col in enumerate(g["columns"]):
            np.testing.assert_allclose(got[name][0, j], g["values"][f"{name}-{col.replace(' ', '-')}"], rtol=2e-5, err_msg=f"{name} {col}")
        np.testing.assert_allclose(got[name][0].mean(), g["values"][name], rtol=2e-5, err_msg=name)

