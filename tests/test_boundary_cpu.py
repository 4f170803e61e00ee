"""
CPU-only tests of the drop-in boundary: factories, `kind` registry, definition
codec hooks, get_params / clone, pickling, the output-frame contract.  They mirror the reference's
own tests (cited per test); nothing here computes on a device.
"""
import pickle

import numpy as np
import pandas as pd
import pytest
from sklearn.base import clone
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler

from gordo_b200 import serializer
from gordo_b200.machine.model import utils as model_utils
from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector
from gordo_b200.machine.model.base import GordoBase
from gordo_b200.machine.model.factories import feedforward_autoencoder as ffa, lstm_autoencoder as lstma
from gordo_b200.machine.model.factories.utils import hourglass_calc_dims, check_dim_func_len
from gordo_b200.machine.model.models import KerasAutoEncoder, KerasLSTMAutoEncoder, KerasLSTMForecast
from gordo_b200.machine.model.register import register_model_builder


# tests/gordo/machine/model/test_factories_utils.py:8-24
@pytest.mark.parametrize("args,expected", [
    ((0.2, 4, 5), (4, 3, 2, 1)), ((0.5, 3, 10), (8, 7, 5)), ((0.5, 3, 3), (3, 2, 2)),
    ((0.3, 3, 10), (8, 5, 3)), ((1, 3, 10), (10, 10, 10)), ((0, 3, 100000), (66667, 33334, 1))])
def test_hourglass_calc_dims_check_dims(args, expected):
    assert hourglass_calc_dims(*args) == expected


def test_check_dim_func_len():       # test_factories_utils.py:27-36
    with pytest.raises(ValueError):
        check_dim_func_len("test", dim=(256, 128), func=("tanh", "tanh", "tanh"))
    with pytest.raises(ValueError):
        check_dim_func_len("test", dim=(256, 128, 56), func=("tanh", "tanh"))


def test_factory_topologies():
    # docstrings feedforward_autoencoder.py:225-238, lstm_autoencoder.py:235-248
    assert ffa.feedforward_hourglass(10).widths == [10, 8, 7, 5, 5, 7, 8, 10]
    assert ffa.feedforward_hourglass(5).widths == [5, 4, 4, 3, 3, 4, 4, 5]
    assert ffa.feedforward_hourglass(10, encoding_layers=1).widths == [10, 5, 5, 10]
    t = ffa.feedforward_hourglass(10)
    assert t.acts == ["tanh"] * 6 + ["linear"] and t.l1 == [0.0, 10e-5, 10e-5, 0.0, 0.0, 0.0, 0.0]
    assert ffa.feedforward_symmetric(6, dims=(4, 2), funcs=("relu", "tanh")).acts == ["relu", "tanh", "tanh", "relu", "linear"]
    lt = lstma.lstm_hourglass(10, lookback_window=7)
    assert lt.units == [8, 7, 5, 5, 7, 8] and lt.lookback_window == 7 and lt.n_features_out == 10
    with pytest.raises(ValueError):
        ffa.feedforward_symmetric(4, dims=())
    with pytest.raises(ValueError):
        ffa.feedforward_model(4, encoding_dim=(3, 2), encoding_func=("tanh",))
    with pytest.raises(ValueError):
        ffa.feedforward_hourglass(4, optimizer="SGD")        # Adam only, loudly


def test_registry_and_kind_validation():     # models.py:109-128, register.py:48-75
    assert {"feedforward_model", "feedforward_symmetric", "feedforward_hourglass"} <= set(
        register_model_builder.factories["KerasAutoEncoder"])
    assert {"lstm_model", "lstm_symmetric", "lstm_hourglass"} <= set(register_model_builder.factories["KerasLSTMAutoEncoder"])
    assert "lstm_model" in register_model_builder.factories["KerasLSTMForecast"]
    with pytest.raises(ValueError):
        KerasAutoEncoder(kind="no_such_kind")
    with pytest.raises(ValueError):
        KerasAutoEncoder(kind="no.such.module.factory")
    with pytest.raises(ValueError):
        register_model_builder(type="KerasAutoEncoder")(lambda features: None)

    def my_factory(n_features, n_features_out=None, **kw):
        return ffa.feedforward_symmetric(n_features, n_features_out, dims=(3,), funcs=("tanh",))
    m = KerasAutoEncoder(kind=my_factory)
    assert m.kind == "my_factory" and "my_factory" in register_model_builder.factories["KerasAutoEncoder"]
    KerasAutoEncoder(kind="gordo_b200.machine.model.factories.feedforward_autoencoder.feedforward_hourglass")


def test_definition_hooks_params_clone_pickle():
    # models.py:146-171 + tests/gordo/serializer/definition_test_model.py protocol
    m = KerasAutoEncoder.from_definition({"kind": "feedforward_hourglass", "epochs": 3, "compression_factor": 0.3})
    assert isinstance(m, GordoBase)
    assert m.into_definition() == {"kind": "feedforward_hourglass", "epochs": 3, "compression_factor": 0.3}
    assert m.get_params() == {"kind": "feedforward_hourglass", "epochs": 3, "compression_factor": 0.3}
    c = clone(m)
    assert c is not m and c.get_params() == m.get_params() and c.model is None
    assert m.get_metadata() == {}
    m2 = pickle.loads(pickle.dumps(m))
    assert m2.get_params() == m.get_params()
    l = KerasLSTMAutoEncoder(kind="lstm_hourglass", lookback_window=5)
    assert l.lookahead == 0 and KerasLSTMForecast(kind="lstm_model").lookahead == 1
    assert l.get_metadata() == {"forecast_steps": 0}
    assert clone(l).lookback_window == 5
    from sklearn.exceptions import NotFittedError
    with pytest.raises(NotFittedError):
        m.score(np.zeros((3, 2)), np.zeros((3, 2)))
    with pytest.raises(ValueError):
        KerasAutoEncoder.get_n_features(np.zeros(3))


def test_example_config_model_block_loads():
    # examples/config.yaml:72-82, with gordo.* paths redirected onto gordo_b200.*
    import yaml
    block = yaml.safe_load("""
    gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector:
      base_estimator:
        sklearn.pipeline.Pipeline:
          steps:
            - sklearn.preprocessing.MinMaxScaler
            - gordo.machine.model.models.KerasAutoEncoder:
                kind: feedforward_hourglass
    """)
    model = serializer.from_definition(block, redirect_gordo=True)
    assert type(model) is DiffBasedAnomalyDetector and isinstance(model.base_estimator, Pipeline)
    assert isinstance(model.base_estimator.steps[0][1], MinMaxScaler)
    assert type(model.base_estimator.steps[1][1]) is KerasAutoEncoder
    assert model.base_estimator.steps[1][1].kind == "feedforward_hourglass"
    d = serializer.into_definition(model)
    again = serializer.from_definition(d)
    assert type(again.base_estimator.steps[1][1]) is KerasAutoEncoder
    assert clone(model).get_params().keys() == model.get_params().keys()


def test_detector_params_metadata_surface():
    # tests/gordo/machine/model/anomaly/test_anomaly_detectors.py:55-57, 675-732
    base = KerasAutoEncoder(kind="feedforward_hourglass")
    sc = MinMaxScaler()
    det = DiffBasedAnomalyDetector(base_estimator=base, scaler=sc, shuffle=True)
    assert det.get_params() == dict(base_estimator=base, scaler=sc, shuffle=True)
    det2 = DiffBasedAnomalyDetector(base_estimator=base, scaler=sc, window=144)
    assert det2.smoothing_method == "smm" and det2.get_params()["window"] == 144
    assert det.kind == "feedforward_hourglass"            # transparent into base_estimator (diff.py:78-86)
    k = DiffBasedKFCVAnomalyDetector(base_estimator=base, scaler=sc)
    assert k.get_params()["threshold_percentile"] == 0.99 and k.window == 144 and k.shuffle is True
    assert pickle.loads(pickle.dumps(det)).shuffle is True
    md = det.get_metadata()
    assert md["window"] is None and "feature-thresholds" not in md


# tests/gordo/machine/model/test_utils.py:35-106
@pytest.mark.parametrize("dates", [pd.date_range("2016-01-01", "2016-01-02", periods=10), None])
@pytest.mark.parametrize("tags", [["tag1", "tag2"], ["tag"]])
@pytest.mark.parametrize("target_tag_list", (["tag1", "tag2"], ["tag3", "tag4"], ["tagA"], ["tag1"],
                                             ["tagA", "tagB", "tagC"], None))
@pytest.mark.parametrize("output_offset", (0, 1, 2, 3))
def test_base_dataframe_creation(dates, tags, target_tag_list, output_offset):
    rng = np.random.default_rng(0)
    n = 10
    model_input = rng.random((n, len(tags)))
    model_output = rng.random((n, len(target_tag_list or list(range(20)))))[output_offset:]
    df = model_utils.make_base_dataframe(tags=tags, model_input=model_input, model_output=model_output,
                                         target_tag_list=target_tag_list, index=dates)
    assert np.array_equal(df["model-input"].values, model_input[-len(df):, :])
    assert df["model-input"].columns.tolist() == tags
    assert np.array_equal(df["model-output"].values, model_output[-len(df):, :])
    if target_tag_list is not None:
        assert target_tag_list == df["model-output"].columns.tolist()
    elif model_output.shape[1] == len(tags):
        assert tags == df["model-output"].columns.tolist()
    else:
        assert list(map(str, range(model_output.shape[1]))) == df["model-output"].columns.tolist()
    if dates is not None:
        assert np.array_equal(df.index.values, dates.values[output_offset:])
    else:
        assert np.array_equal(df.index.values, np.arange(0, len(df)))


def test_metrics_wrapper():          # tests/gordo/machine/model/test_utils.py:12-32
    from sklearn.metrics import mean_squared_error
    y = np.array([[1, 1], [2, 2], [3, 3], [4, 4], [5, 5]]) * [1, 100]
    f = model_utils.metric_wrapper(mean_squared_error)
    assert not np.isclose(f(y, y * [0.8, 1]), f(y, y * [1, 0.8]))
    scaler = MinMaxScaler().fit(y)
    g = model_utils.metric_wrapper(mean_squared_error, scaler=scaler)
    assert np.isclose(g(y, y * [0.8, 1]), g(y, y * [1, 0.8]))


def test_frame_layout_matches_reference_golden():
    """assemble_frame reproduces the column tuples / start / end the REAL reference produced."""
    import json, os
    G = os.path.join(os.path.dirname(__file__), "golden")
    meta = json.load(open(os.path.join(G, "golden_meta.json")))
    npz = np.load(os.path.join(G, "detector_golden.npz"))
    for ci, case in enumerate(meta["cases"]):
        pre = f"c{ci}_"
        tags = [f"tag-{j}" for j in range(case["t"])]
        idx = (pd.date_range("2019-01-01", periods=case["n"], freq="10min") if case["index"] == "dates"
               else pd.RangeIndex(case["n"]))
        groups = []
        for g in case["groups"]:
            v = npz[pre + "col_" + g]
            if v.ndim == 2 and v.shape[1] == 1 and g.startswith(("total", "smooth-total")):
                v = v[:, 0]
            groups.append((g, v, tags if v.ndim == 2 else None))
        df = model_utils.assemble_frame(groups, idx, pd.Timedelta(minutes=10))
        assert [list(map(str, c)) for c in df.columns] == case["columns"]
        assert len(df) == case["n_rows"]
        if case["index"] == "dates":
            assert df[("start", "")].iloc[:3].tolist() == case["start"]
            assert df[("end", "")].iloc[:3].tolist() == case["end"]
        np.testing.assert_array_equal(df["model-output"].to_numpy(), npz[pre + "col_model-output"])


def test_fleet_builder_bucketing_rules():
    """Which Machines share a batched launch (host logic only: no device needed)."""
    import pandas as pd
    from gordo_b200 import serializer
    from gordo_b200.builder import FleetBuild, FleetMachine, FleetModelBuilder

    def machine(detector="DiffBasedAnomalyDetector", det_kw=None, est="KerasAutoEncoder", est_kw=None, tags=4, evaluation=None):
        kw = {"kind": "feedforward_hourglass" if est == "KerasAutoEncoder" else "lstm_hourglass"}
        kw.update(est_kw or {})
        d = {f"gordo_b200.machine.model.anomaly.diff.{detector}": dict(det_kw or {}, base_estimator={
            "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
                                                    {f"gordo_b200.machine.model.models.{est}": kw}]}})}
        return FleetMachine("m", pd.DataFrame(np.zeros((50, tags))), model=d, evaluation=evaluation or {})

    b = FleetBuild([])
    key = lambda mc: b._bucket_key(serializer.from_definition(mc.definition()), mc)
    base = key(machine())
    assert base is not None and base == key(machine())                       # same topology + fit settings: one bucket
    assert key(machine(tags=5)) != base                                      # another topology
    assert key(machine(est_kw={"epochs": 3})) != base                        # another fit schedule
    assert key(machine(det_kw={"window": 12})) not in (None, base)           # smoothing window: batched, own bucket
    assert key(machine(evaluation={"cv_mode": "build_only"})) != base
    assert key(machine(det_kw={"shuffle": True})) is None                    # detector-level shuffle: per-Machine path
    assert key(machine(est_kw={"validation_split": 0.1})) is None            # callbacks / validation: per-Machine path
    assert key(machine(evaluation={"cv_mode": "cross_val_only"})) is None
    kf = key(machine(detector="DiffBasedKFCVAnomalyDetector"))
    assert kf not in (None, base)                                            # K-fold detector over the FF estimator: batched
    assert key(machine(detector="DiffBasedKFCVAnomalyDetector", det_kw={"threshold_percentile": 0.9})) != kf
    assert key(machine(detector="DiffBasedKFCVAnomalyDetector", det_kw={"smoothing_method": "median"})) is None
    lstm = key(machine(est="KerasLSTMAutoEncoder", est_kw={"lookback_window": 4}))
    assert lstm is not None and lstm[0] == "lstm"
    assert key(machine(detector="DiffBasedKFCVAnomalyDetector", est="KerasLSTMAutoEncoder", est_kw={"lookback_window": 4})) is None


# ----------------------------------------------------------------------------- round 2 host logic (no device needed)
def test_job_seeds_are_the_per_machine_path_draws():
    """FleetBuild seeds fit job i of a Machine with the i-th draw the per-Machine path makes after np.random.seed(seed)."""
    from gordo_b200.builder import job_seeds
    np.random.seed(7)
    want = [int(np.random.randint(0, 2 ** 31 - 1)) for _ in range(4)]
    assert job_seeds(7, 4) == want and job_seeds(7, 2) == want[:2] and job_seeds(8, 4) != want


def test_redirect_definition_and_build_metadata_layout():
    from gordo_b200.builder import build_metadata_dict, redirect_definition
    d = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass"}}]}}}}
    r = redirect_definition(d)
    assert list(r) == ["gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector"]
    steps = r[list(r)[0]]["base_estimator"]["sklearn.pipeline.Pipeline"]["steps"]
    assert steps[0] == "sklearn.preprocessing.MinMaxScaler" and list(steps[1]) == ["gordo_b200.machine.model.models.KerasAutoEncoder"]
    assert "gordo.machine" in list(d)[0]                           # the input is not modified
    assert redirect_definition(d, enable=False) is d
    bm = build_metadata_dict({"model_offset": 3, "model_training_duration_sec": 1.5, "cv_duration_sec": 2.5,
                              "cross_validation": {"scores": {"r2-score": {"fold-mean": 0.5}}, "splits": {"fold-1-n-train": 10}},
                              "model": {"history": {"loss": [1.0]}}})
    # gordo/machine/metadata/metadata.py:17-56
    assert set(bm) == {"model", "dataset"} and bm["model"]["model_offset"] == 3
    assert set(bm["model"]) == {"model_offset", "model_creation_date", "model_builder_version", "model_training_duration_sec",
                                "cross_validation", "model_meta"}
    assert bm["model"]["cross_validation"] == {"cv_duration_sec": 2.5, "scores": {"r2-score": {"fold-mean": 0.5}},
                                               "splits": {"fold-1-n-train": 10}}
    assert set(bm["dataset"]) == {"query_duration_sec", "dataset_meta"}


def test_extract_model_metadata_digs_through_pipelines_like_the_reference():
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import MinMaxScaler
    from gordo_b200.builder import extract_model_metadata
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from gordo_b200.machine.model.models import History, KerasAutoEncoder
    est = KerasAutoEncoder(kind="feedforward_hourglass")
    est._history = History({"loss": [0.5], "accuracy": [0.1]}, {"epochs": 1}, [0])
    det = DiffBasedAnomalyDetector(base_estimator=Pipeline([("s", MinMaxScaler()), ("m", est)]))
    det.aggregate_threshold_ = 0.25
    meta = extract_model_metadata(det)                             # build_model.py:518-570
    assert meta["history"]["loss"] == [0.5] and meta["aggregate-threshold"] == 0.25
    assert extract_model_metadata(Pipeline([("s", MinMaxScaler()), ("m", est)]))["history"]["params"] == {"epochs": 1}


def test_hostbind_cpu_slicing_and_quota(monkeypatch, tmp_path):
    from gordo_b200 import hostbind
    assert hostbind._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    # 8 cores with hyper-thread siblings c, c + 8: four ranks get two cores (+ siblings) each, nothing shared
    monkeypatch.setattr(hostbind, "_siblings", lambda c: [c % 8, c % 8 + 8])
    cpus = list(range(16))
    parts = [hostbind.slice_cpus(cpus, r, 4) for r in range(4)]
    assert parts[0] == [0, 1, 8, 9] and parts[3] == [6, 7, 14, 15]
    assert sorted(c for p in parts for c in p) == cpus
    assert hostbind.slice_cpus(cpus, 0, 1) == cpus
    monkeypatch.setattr(hostbind, "cpu_quota", lambda: 6.0)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    assert hostbind.effective_cpus() <= 3                          # the quota is shared by the ranks of the box
    monkeypatch.setattr(hostbind, "cpu_quota", lambda: None)
    assert hostbind.effective_cpus() >= 1


def test_fleet_anomaly_result_response_methods_without_a_device():
    """FleetAnomalyResult is host-only: frame / parquet / dict of one Machine from the column buffers."""
    import torch
    from gordo_b200.serving import FleetAnomalyResult
    from gordo_b200.server import utils as su
    rng = np.random.default_rng(0)
    R, T = 30, 3
    cols = {k: torch.from_numpy(rng.random((R, T)).astype(np.float32)) for k in
            ("model-output", "tag-anomaly-scaled", "tag-anomaly-unscaled", "anomaly-confidence")}
    cols.update({k: torch.from_numpy(rng.random(R).astype(np.float32)) for k in
                 ("total-anomaly-scaled", "total-anomaly-unscaled", "total-anomaly-confidence")})
    x = torch.from_numpy(rng.random((R, T)).astype(np.float32))
    res = FleetAnomalyResult(cols, np.array([0, 10, 30]), x, tags=[["a", "b", "c"], ["d", "e", "f"]])
    idx = pd.date_range("2021-01-01", periods=20, freq="10min", tz="UTC")
    frame = res.frame(1, idx, pd.Timedelta("10min"))
    assert len(frame) == 20 and ("anomaly-confidence", "e") in frame.columns and frame.columns[0] == ("start", "")
    np.testing.assert_allclose(frame["model-output"].to_numpy(), cols["model-output"][10:30].numpy())
    back = su.dataframe_from_parquet_bytes(res.parquet_bytes(1, idx, pd.Timedelta("10min")))
    pd.testing.assert_frame_equal(back, frame, check_freq=False)
    d = res.to_dict(0)
    assert set(d["model-input"]) == {"a", "b", "c"} and len(d["total-anomaly-scaled"]["total-anomaly-scaled"]) == 10


REFERENCE_SERIALIZABILITY_CONFIGS = ["""
    gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector:
        base_estimator:
            sklearn.pipeline.Pipeline:
                steps:
                - sklearn.preprocessing.MinMaxScaler
                - gordo.machine.model.models.KerasAutoEncoder:
                    kind: feedforward_hourglass
""", """
    gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector:
        base_estimator:
            gordo.machine.model.models.KerasAutoEncoder:
                kind: feedforward_hourglass
""", """
    gordo.machine.model.anomaly.diff.DiffBasedKFCVAnomalyDetector:
        base_estimator:
            sklearn.compose.TransformedTargetRegressor:
                transformer:
                    sklearn.preprocessing.MinMaxScaler
                regressor:
                    sklearn.pipeline.Pipeline:
                        steps:
                        - sklearn.preprocessing.MinMaxScaler
                        - gordo.machine.model.models.KerasAutoEncoder:
                            kind: feedforward_hourglass
                            batch_size: 128
                            compression_factor: 0.5
                            encoding_layers: 1
                            func: tanh
                            out_func: linear
                            optimizer: Adam
                            loss: mse
                            epochs: 1000
                            validation_split: 0.1
                            callbacks:
                                - tensorflow.keras.callbacks.EarlyStopping:
                                    monitor: val_loss
                                    patience: 10
                                    restore_best_weights: true
        scaler: sklearn.preprocessing.MinMaxScaler
        window: 144
        shuffle: true
        threshold_percentile: 0.975
"""]


@pytest.mark.parametrize("config", REFERENCE_SERIALIZABILITY_CONFIGS)
def test_reference_detector_configs_are_serializable(config):
    """The unmodified YAMLs of tests/gordo/machine/model/anomaly/test_anomaly_detectors.py:480-552 (incl. the
    TransformedTargetRegressor + EarlyStopping K-fold config): from_definition with the gordo -> gordo_b200 redirect,
    into_definition and back, pickle round trip -- "should play well with the gordo serializer"."""
    import pickle
    import yaml
    from gordo_b200 import serializer
    from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    model = serializer.from_definition(yaml.safe_load(config), redirect_gordo=True)
    assert isinstance(model, DiffBasedAnomalyDetector)
    definition = serializer.into_definition(model)
    again = serializer.from_definition(definition)
    assert type(again) is type(model) and again.get_params().keys() == model.get_params().keys()
    clone = pickle.loads(pickle.dumps(model))
    assert type(clone) is type(model)
    if "KFCV" in config:
        assert model.threshold_percentile == 0.975 and model.window == 144 and model.shuffle is True
        est = model.base_estimator.regressor.steps[1][1]
        assert est.kwargs["epochs"] == 1000 and est.kwargs["callbacks"][0] == {
            "tensorflow.keras.callbacks.EarlyStopping": {"monitor": "val_loss", "patience": 10, "restore_best_weights": True}}


RAW_KERAS_SPECS = ["""
        compile:
            loss: mse
            optimizer:
              tensorflow.keras.optimizers.SGD:
                learning_rate: 0.001
        spec:
            tensorflow.keras.models.Sequential:
                layers:
                    - tensorflow.keras.layers.Dense:
                        units: 10
                    - tensorflow.keras.layers.Dense:
                        units: 32
                        kernel_regularizer:
                            tensorflow.keras.regularizers.L1L2:
                                l1: 0.2
                    - tensorflow.keras.layers.Dense:
                        units: 1
    """, """
        compile:
            loss: mse
            optimizer: adam
        spec:
            tensorflow.keras.models.Sequential:
                layers:
                    - tensorflow.keras.layers.Input:
                        shape: [9]
                    - tensorflow.keras.layers.Reshape:
                        target_shape: [3, 3]
                    - tensorflow.keras.layers.LSTM:
                        units: 12
                    - tensorflow.keras.layers.Flatten
                    - tensorflow.keras.layers.Dense:
                        units: 1
    """]


def test_raw_keras_regressor_subset_and_its_limits():
    """KerasRawModelRegressor (models.py:401-460; the reference's tests/gordo/machine/model/test_raw_keras.py): the
    Dense / MSE / Adam subset becomes a feed-forward topology; the reference's two arbitrary-graph specs (SGD + kernel
    regulariser, Reshape / LSTM / Flatten) are refused by name instead of being approximated."""
    import yaml
    from gordo_b200 import serializer
    from gordo_b200.machine.model.models import KerasRawModelRegressor
    from sklearn.pipeline import Pipeline
    for spec_str in RAW_KERAS_SPECS:
        est = KerasRawModelRegressor(yaml.safe_load(spec_str))
        est.kwargs["n_features"] = 9
        with pytest.raises(NotImplementedError):
            est._topology()
    config = yaml.safe_load("""
    sklearn.pipeline.Pipeline:
        steps:
            - sklearn.decomposition.PCA:
                n_components: 4
            - gordo.machine.model.models.KerasRawModelRegressor:
                kind:
                    compile:
                        loss: mse
                        optimizer:
                            tensorflow.keras.optimizers.Adam:
                                learning_rate: 0.01
                    spec:
                        tensorflow.keras.models.Sequential:
                            layers:
                                - tensorflow.keras.layers.Dense:
                                    units: 4
                                    input_shape: [4]
                                    activation: tanh
                                    activity_regularizer:
                                        tensorflow.keras.regularizers.L1:
                                            l1: 0.001
                                - tensorflow.keras.layers.Dense:
                                    units: 1
    """)
    pipe = serializer.from_definition(config, redirect_gordo=True)
    assert isinstance(pipe, Pipeline) and type(pipe.steps[1][1]) is KerasRawModelRegressor
    topo = pipe.steps[1][1]._topology()
    assert topo.widths == [4, 4, 1] and topo.acts == ["tanh", "linear"] and topo.l1 == [0.001, 0.0] and topo.adam["lr"] == 0.01
    again = serializer.from_definition(serializer.into_definition(pipe))
    assert again.steps[1][1].kind == pipe.steps[1][1].kind
    with pytest.raises(ValueError, match="Expected spec to have keys"):
        KerasRawModelRegressor({"spec": {}})._topology()
    bare = KerasRawModelRegressor(yaml.safe_load("{compile: {loss: mse, optimizer: adam}, spec: {tensorflow.keras.models.Sequential: {layers: [{tensorflow.keras.layers.Dense: {units: 3}}]}}}"))
    with pytest.raises(ValueError, match="input_shape"):
        bare._topology()
    bare.kwargs["n_features"] = 7                     # what fit() records before building the model
    assert bare._topology().widths == [7, 3]
