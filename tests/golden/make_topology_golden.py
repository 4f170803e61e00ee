#!/usr/bin/env python
"""
Golden topologies from the REAL reference factories (gordo/machine/model/factories/feedforward_autoencoder.py:15-251,
lstm_autoencoder.py:15-263, register.py), executed unmodified from /root/reference with `tensorflow.keras` replaced by
RECORDING stand-ins: `Dense(...)`, `LSTM(...)`, `regularizers.l1(...)`, `Sequential.add / compile`,
`keras.optimizers.get(...)` only note their arguments.  What is recorded is therefore exactly what the reference asks
Keras to build -- layer kinds, units, activations, activity regularisers, return_sequences, input shapes, the
optimizer config and the compile kwargs -- for a matrix of factory arguments.  No Keras arithmetic is involved.

    python tests/golden/make_topology_golden.py      ->  tests/golden/topology_golden.json
"""
import importlib.util
import json
import os
import sys
import types

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


class Dense:
    def __init__(self, units=None, activation=None, **kw):
        reg = kw.pop("activity_regularizer", None)
        self.rec = {"kind": "Dense", "units": int(units), "activation": activation, "activity_regularizer": reg, **kw}


class LSTM:
    def __init__(self, units=None, activation=None, return_sequences=False, **kw):
        if "input_shape" in kw:
            kw["input_shape"] = list(kw["input_shape"])
        self.rec = {"kind": "LSTM", "units": int(units), "activation": activation, "return_sequences": bool(return_sequences), **kw}


class Sequential:
    def __init__(self):
        self.layers, self.compile_kwargs = [], None

    def add(self, layer):
        self.layers.append(layer.rec)

    def compile(self, **kw):
        self.compile_kwargs = dict(kw)


class Optimizer:
    pass


def load_factories():
    opt = types.SimpleNamespace(get=lambda cfg: {"optimizers.get": cfg}, Optimizer=Optimizer)
    reg = _stub("tensorflow.keras.regularizers", l1=lambda v: {"l1": v})
    keras = _stub("tensorflow.keras", optimizers=opt, regularizers=reg,
                  models=types.SimpleNamespace(Sequential=Sequential, Model=Sequential))
    _stub("tensorflow", keras=keras)
    _stub("tensorflow.keras.optimizers", Optimizer=Optimizer, get=opt.get)
    _stub("tensorflow.keras.layers", Dense=Dense, LSTM=LSTM)
    _stub("tensorflow.keras.models", Sequential=Sequential, Model=Sequential)
    g = os.path.join(REF, "gordo")
    _pkg("gordo", g); _pkg("gordo.machine", os.path.join(g, "machine"))
    mm = os.path.join(g, "machine", "model")
    _pkg("gordo.machine.model", mm); _pkg("gordo.machine.model.factories", os.path.join(mm, "factories"))
    _load("gordo.machine.model.base", os.path.join(mm, "base.py"))
    _load("gordo.machine.model.register", os.path.join(mm, "register.py"))
    _load("gordo.machine.model.factories.utils", os.path.join(mm, "factories", "utils.py"))
    ff = _load("gordo.machine.model.factories.feedforward_autoencoder", os.path.join(mm, "factories", "feedforward_autoencoder.py"))
    ls = _load("gordo.machine.model.factories.lstm_autoencoder", os.path.join(mm, "factories", "lstm_autoencoder.py"))
    return ff, ls, sys.modules["gordo.machine.model.register"]


CASES = [
    ("feedforward_model", dict(n_features=20)),
    ("feedforward_model", dict(n_features=7, n_features_out=3, encoding_dim=(12, 6), encoding_func=("relu", "tanh"),
                               decoding_dim=(6, 9, 12), decoding_func=("tanh", "elu", "sigmoid"), out_func="relu")),
    ("feedforward_model", dict(n_features=5, encoding_dim=(4,), encoding_func=("tanh",), decoding_dim=(4,), decoding_func=("tanh",),
                               optimizer="Adam", optimizer_kwargs={"learning_rate": 0.01, "beta_1": 0.8},
                               compile_kwargs={"loss": "mae"})),
    ("feedforward_symmetric", dict(n_features=20)),
    ("feedforward_symmetric", dict(n_features=9, n_features_out=4, dims=(8, 5, 2, 2), funcs=("tanh", "relu", "tanh", "linear"))),
    ("feedforward_hourglass", dict(n_features=10)),
    ("feedforward_hourglass", dict(n_features=10, compression_factor=0.2)),
    ("feedforward_hourglass", dict(n_features=50)),
    ("feedforward_hourglass", dict(n_features=100, encoding_layers=4, compression_factor=0.3, func="relu")),
    ("feedforward_hourglass", dict(n_features=3, encoding_layers=1, compression_factor=1.0)),
    ("feedforward_hourglass", dict(n_features=20, n_features_out=5, encoding_layers=2, compression_factor=0.5)),
    ("lstm_model", dict(n_features=4, lookback_window=5)),
    ("lstm_model", dict(n_features=6, n_features_out=2, lookback_window=3, encoding_dim=(8, 4), encoding_func=("tanh", "relu"),
                        decoding_dim=(4, 8), decoding_func=("relu", "tanh"), out_func="tanh",
                        optimizer_kwargs={"learning_rate": 0.005}, compile_kwargs={"loss": "mae"})),
    ("lstm_symmetric", dict(n_features=5, lookback_window=2, dims=(6, 3), funcs=("tanh", "tanh"))),
    ("lstm_hourglass", dict(n_features=10, lookback_window=16)),
    ("lstm_hourglass", dict(n_features=50, lookback_window=128, encoding_layers=3, compression_factor=0.5)),
    ("lstm_hourglass", dict(n_features=7, lookback_window=4, encoding_layers=1, compression_factor=0.3, func="relu", out_func="linear")),
]


def main():
    ff, ls, register = load_factories()
    out = []
    for name, kw in CASES:
        fn = getattr(ff, name, None) or getattr(ls, name)
        kw2 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}    # the factories mutate compile_kwargs
        model = fn(**kw2)
        out.append({"factory": name, "kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                    "layers": model.layers, "compile_kwargs": model.compile_kwargs})
    reg = {t: sorted(kinds) for t, kinds in register.register_model_builder.factories.items()}
    path = os.path.join(HERE, "topology_golden.json")
    with open(path, "w") as f:
        json.dump({"cases": out, "registered": reg}, f, indent=1)
    print(f"wrote {path}: {len(out)} cases; registered:", reg)


if __name__ == "__main__":
    main()
