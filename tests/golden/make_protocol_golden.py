#!/usr/bin/env python
"""
Golden TRAINING / PREDICTION PROTOCOL of the reference estimators (gordo/machine/model/models.py:35-398 KerasBaseEstimator /
KerasAutoEncoder, :463-793 KerasLSTMBaseEstimator, create_keras_timeseriesgenerator), executed unmodified from
/root/reference with recording stand-ins for what it imports:

  * `scikeras.wrappers.KerasRegressor`: keeps constructor kwargs as attributes, `fit` = `self.model.fit(X, y, **{own fit
    kwargs, call kwargs})`, as scikeras does;
  * `tensorflow.keras` layers / Sequential (tests/golden/make_topology_golden.py): `Sequential.fit / predict` only LOG
    what they are handed -- arrays (shape + kwargs) or, for a generator, every batch in the order Keras would draw them
    (epochs x batches in index order: `shuffle=False`);
  * `TimeseriesGenerator` / `pad_sequences`: the published Keras utilities, restated (the windowing they produce is
    pinned separately by the reference's own literal vectors, tests/gordo/machine/model/test_model.py:239-311).

Rows of X are numbered (X[r, c] = r + c / 100), so a logged batch says exactly which windows and targets it holds.
What this pins: the primer step on the first window, the ordered batches that follow it with the Adam state carried
over, which constructor / fit kwargs reach Keras, the 10 000-window prediction batches and the output offset.

    python tests/golden/make_protocol_golden.py      ->  tests/golden/protocol_golden.json
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_topology_golden as T          # noqa: E402  (stubs + loader of the reference factories)

LOG = []


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    out = []
    for s in sequences:
        s = np.asarray(s)
        if maxlen is not None and len(s) > maxlen:
            s = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        n = (maxlen or len(s)) - len(s)
        pad = np.full((n,) + s.shape[1:], value, dtype=dtype)
        out.append(np.concatenate([pad, s.astype(dtype)]) if padding == "pre" else np.concatenate([s.astype(dtype), pad]))
    return np.stack(out)


class TimeseriesGenerator:
    def __init__(self, data, targets, length, sampling_rate=1, stride=1, start_index=0, end_index=None, shuffle=False,
                 reverse=False, batch_size=128):
        assert len(data) == len(targets)
        self.data, self.targets, self.length, self.sampling_rate, self.stride = data, targets, length, sampling_rate, stride
        self.start_index = start_index + length
        self.end_index = len(data) - 1 if end_index is None else end_index
        self.batch_size = batch_size
        assert not shuffle and not reverse

    def __len__(self):
        return (self.end_index - self.start_index + self.batch_size * self.stride) // (self.batch_size * self.stride)

    def __getitem__(self, index):
        i = self.start_index + self.batch_size * self.stride * index
        rows = np.arange(i, min(i + self.batch_size * self.stride, self.end_index + 1), self.stride)
        samples = np.array([self.data[r - self.length:r:self.sampling_rate] for r in rows])
        targets = np.array([self.targets[r] for r in rows])
        return samples, targets


def _describe(x, y):
    x, y = np.asarray(x), np.asarray(y)
    if x.ndim == 3:      # windows: first column holds the row number of every time step
        return {"window_rows": np.rint(x[:, :, 0]).astype(int).tolist(), "target_rows": np.rint(y[:, 0]).astype(int).tolist()}
    return {"x_shape": list(x.shape), "y_shape": list(y.shape)}


class History:
    def __init__(self):
        self.history, self.params, self.epoch = {"loss": []}, {}, []


def _seq_fit(self, x, y=None, **kw):
    entry = {"call": "fit", "kwargs": {k: v for k, v in kw.items() if k != "callbacks"}}
    if kw.get("callbacks") is not None:
        entry["kwargs"]["callbacks"] = [type(c).__name__ if not isinstance(c, dict) else c for c in kw["callbacks"]]
    if hasattr(x, "__getitem__") and hasattr(x, "batch_size") and y is None:
        entry["generator_batch_size"] = x.batch_size
        entry["batches"] = [_describe(*x[i]) for _ in range(kw.get("epochs", 1)) for i in range(len(x))]
    else:
        entry.update(_describe(x, y))
    LOG.append(entry)
    self.history = History()
    return self.history


def _seq_predict(self, x, **kw):
    entry = {"call": "predict", "kwargs": dict(kw)}
    if hasattr(x, "batch_size") and not isinstance(x, np.ndarray):
        entry["generator_batch_size"] = x.batch_size
        parts = [x[i] for i in range(len(x))]
        entry["batches"] = [_describe(a, b) for a, b in parts]
        n = sum(len(a) for a, _ in parts)
    else:
        entry["x_shape"] = list(np.asarray(x).shape)
        n = len(x)
    LOG.append(entry)
    return np.zeros((n, self.layers[-1]["units"]))


class KerasRegressor:
    _fit_kwargs = {"batch_size", "epochs", "verbose", "callbacks", "validation_split", "shuffle", "class_weight",
                   "sample_weight", "initial_epoch", "validation_steps", "steps_per_epoch", "validation_batch_size",
                   "validation_freq"}
    _predict_kwargs = {"batch_size", "verbose", "steps"}
    _compile_kwargs = {"optimizer", "loss", "metrics", "loss_weights", "weighted_metrics", "run_eagerly"}

    def __init__(self, model=None, **kwargs):
        self.model = model
        self._own = dict(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def fit(self, X, y, sample_weight=None, **kwargs):
        merged = {k: v for k, v in self._own.items() if k in self._fit_kwargs}
        merged.update(kwargs)
        self.model.fit(X, y, **merged)
        return self

    def get_params(self, **params):
        return dict(self._own)


def load_models():
    ff, ls, register = T.load_factories()
    T.Sequential.fit, T.Sequential.predict = _seq_fit, _seq_predict
    kmodels = sys.modules["tensorflow.keras.models"]
    kmodels.load_model = lambda *a, **k: None
    kmodels.save_model = lambda *a, **k: None
    sys.modules["tensorflow.keras"].models = kmodels
    sys.modules["tensorflow"].keras.models = kmodels
    T._stub("tensorflow.keras.preprocessing")
    T._stub("tensorflow.keras.preprocessing.sequence", pad_sequences=pad_sequences, TimeseriesGenerator=TimeseriesGenerator)
    sys.modules["tensorflow.keras"].preprocessing = types.SimpleNamespace(sequence=sys.modules["tensorflow.keras.preprocessing.sequence"])
    T._stub("scikeras"); T._stub("scikeras.wrappers", KerasRegressor=KerasRegressor)
    T._stub("xarray", DataArray=type("DataArray", (), {}))
    ser = T._stub("gordo.serializer", load_params_from_definition=lambda d: d, build_callbacks=lambda c: c,
                  from_definition=lambda d: d)
    sys.modules["gordo"].serializer = ser
    fpkg = sys.modules["gordo.machine.model.factories"]
    fpkg.feedforward_autoencoder, fpkg.lstm_autoencoder = ff, ls
    fpkg.__all__ = ["feedforward_autoencoder", "lstm_autoencoder"]
    g = os.path.join(T.REF, "gordo", "machine", "model")
    return T._load("gordo.machine.model.models", os.path.join(g, "models.py"))


def numbered(n, c):
    return np.arange(n, dtype=np.float64)[:, None] + np.arange(c)[None, :] / 100.0


def main():
    M = load_models()
    cases = []

    def run(name, build, n_rows, n_cols, fit_kwargs=None, predict_rows=None):
        LOG.clear()
        est = build()
        X = numbered(n_rows, n_cols)
        est.fit(X, X.copy(), **(fit_kwargs or {}))
        Xp = numbered(predict_rows or n_rows, n_cols)
        out = est.predict(Xp)
        cases.append({"name": name, "n_rows": n_rows, "n_cols": n_cols, "fit_kwargs": fit_kwargs or {},
                      "predict_rows": predict_rows or n_rows, "predict_out_rows": int(len(out)),
                      "estimator_kwargs": {k: v for k, v in est.kwargs.items() if k not in ("kind",)},
                      "log": json.loads(json.dumps(LOG))})

    run("lstm_autoencoder_lb4_bs5_ep2", lambda: M.KerasLSTMAutoEncoder(kind="lstm_hourglass", lookback_window=4, batch_size=5, epochs=2), 23, 3)
    run("lstm_forecast_lb4_bs5", lambda: M.KerasLSTMForecast(kind="lstm_hourglass", lookback_window=4, batch_size=5), 23, 3)
    run("lstm_autoencoder_defaults", lambda: M.KerasLSTMAutoEncoder(kind="lstm_symmetric", dims=(4,), funcs=("tanh",)), 70, 2)
    run("lstm_forecast_lb1_fit_kwargs", lambda: M.KerasLSTMForecast(kind="lstm_model", lookback_window=1, batch_size=8,
                                                                      encoding_dim=(3,), encoding_func=("tanh",), decoding_dim=(3,),
                                                                      decoding_func=("tanh",), validation_split=0.2, shuffle=True),
        20, 2, fit_kwargs={"epochs": 3, "verbose": 1}, predict_rows=12)
    run("ff_autoencoder", lambda: M.KerasAutoEncoder(kind="feedforward_hourglass", epochs=3, batch_size=16, validation_split=0.1,
                                                      shuffle=False), 50, 6)
    run("ff_autoencoder_defaults", lambda: M.KerasAutoEncoder(kind="feedforward_model"), 40, 5, fit_kwargs={"epochs": 2})
    errors = {}
    try:
        M.KerasLSTMAutoEncoder(kind="lstm_hourglass", lookback_window=10).fit(numbered(10, 2), numbered(10, 2))
    except ValueError as e:
        errors["lookback_ge_rows"] = str(e)
    try:
        M.KerasAutoEncoder(kind="no_such_kind")
    except ValueError as e:
        errors["unknown_kind"] = str(e)
    path = os.path.join(HERE, "protocol_golden.json")
    with open(path, "w") as f:
        json.dump({"cases": cases, "errors": errors}, f)
    for c in cases:
        print(c["name"], [(e["call"], len(e.get("batches", [])) or e.get("x_shape")) for e in c["log"]], "out rows", c["predict_out_rows"])
    print(errors)


if __name__ == "__main__":
    main()
