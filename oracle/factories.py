"""
Oracle: network topologies (test infrastructure, see oracle/__init__.py).

Restates gordo/machine/model/factories/utils.py:7-41,
feedforward_autoencoder.py:15-251 and lstm_autoencoder.py:15-263.  A topology is returned as
a plain dict (no Keras object): widths, activation names, which Dense layers carry the L1
activity regulariser, optimizer / loss settings.
"""
import math

ACTIVATIONS = ("linear", "tanh", "relu", "sigmoid", "elu", "softplus")


def hourglass_calc_dims(compression_factor, encoding_layers, n_features):
    # factories/utils.py:31-41
    if not (1 >= compression_factor >= 0):
        raise ValueError("compression_factor must be 0 <= compression_factor <= 1")
    if encoding_layers < 1:
        raise ValueError("encoding_layers must be >= 1")
    smallest = max(min(math.ceil(compression_factor * n_features), n_features), 1)
    slope = (n_features - smallest) / encoding_layers
    return tuple(round(n_features - i * slope) for i in range(1, encoding_layers + 1))


def _check_len(prefix, dim, func):
    # factories/utils.py:44-63
    if len(dim) != len(func):
        raise ValueError(
            f"The length of {prefix}_dim ({len(dim)}) and {prefix}_func ({len(func)}) must be equal."
        )


def _optimizer(optimizer, optimizer_kwargs):
    if not isinstance(optimizer, str) or optimizer.lower() != "adam":
        raise ValueError("oracle restates Adam only")
    kw = dict(optimizer_kwargs or {})
    return {
        "lr": float(kw.get("learning_rate", kw.get("lr", 1e-3))),
        "beta_1": float(kw.get("beta_1", 0.9)),
        "beta_2": float(kw.get("beta_2", 0.999)),
        "epsilon": float(kw.get("epsilon", 1e-7)),
    }


def feedforward_model(n_features, n_features_out=None, encoding_dim=(256, 128, 64),
                      encoding_func=("tanh", "tanh", "tanh"), decoding_dim=(64, 128, 256),
                      decoding_func=("tanh", "tanh", "tanh"), out_func="linear",
                      optimizer="Adam", optimizer_kwargs=None, compile_kwargs=None, **_):
    # feedforward_autoencoder.py:64-103
    n_features_out = n_features_out or n_features
    _check_len("encoding", encoding_dim, encoding_func)
    _check_len("decoding", decoding_dim, decoding_func)
    widths = [n_features] + list(encoding_dim) + list(decoding_dim) + [n_features_out]
    acts = list(encoding_func) + list(decoding_func) + [out_func]
    # activity_regularizer=l1(10e-5) on encoder layers i >= 1 only (:78-81)
    l1 = [0.0] + [10e-5] * (len(encoding_dim) - 1) + [0.0] * (len(decoding_dim) + 1)
    ck = dict(compile_kwargs or {})
    return {"type": "ff", "widths": widths, "acts": acts, "l1": l1,
            "loss": ck.get("loss", "mean_squared_error"),
            "adam": _optimizer(optimizer, optimizer_kwargs)}


def feedforward_symmetric(n_features, n_features_out=None, dims=(256, 128, 64),
                          funcs=("tanh", "tanh", "tanh"), **kw):
    # feedforward_autoencoder.py:143-157
    if len(dims) == 0:
        raise ValueError("Parameter dims must have len > 0")
    return feedforward_model(n_features, n_features_out, encoding_dim=tuple(dims),
                             decoding_dim=tuple(dims)[::-1], encoding_func=tuple(funcs),
                             decoding_func=tuple(funcs)[::-1], **kw)


def feedforward_hourglass(n_features, n_features_out=None, encoding_layers=3,
                          compression_factor=0.5, func="tanh", **kw):
    # feedforward_autoencoder.py:240-251
    dims = hourglass_calc_dims(compression_factor, encoding_layers, n_features)
    return feedforward_symmetric(n_features, n_features_out, dims=dims,
                                 funcs=tuple([func] * len(dims)), **kw)


def lstm_model(n_features, n_features_out=None, lookback_window=1, encoding_dim=(256, 128, 64),
               encoding_func=("tanh", "tanh", "tanh"), decoding_dim=(64, 128, 256),
               decoding_func=("tanh", "tanh", "tanh"), out_func="linear", optimizer="Adam",
               optimizer_kwargs=None, compile_kwargs=None, **_):
    # lstm_autoencoder.py:70-103: stacked LSTMs (return_sequences except the last), Dense out,
    # loss mse, no metrics, no activity regulariser.
    n_features_out = n_features_out or n_features
    _check_len("encoding", encoding_dim, encoding_func)
    _check_len("decoding", decoding_dim, decoding_func)
    units = list(encoding_dim) + list(decoding_dim)
    acts = list(encoding_func) + list(decoding_func)
    ck = dict(compile_kwargs or {})
    return {"type": "lstm", "n_features": n_features, "n_features_out": n_features_out,
            "units": units, "acts": acts, "out_func": out_func,
            "lookback_window": lookback_window, "loss": ck.get("loss", "mse"),
            "adam": _optimizer(optimizer, optimizer_kwargs)}


def lstm_symmetric(n_features, n_features_out=None, lookback_window=1, dims=(256, 128, 64),
                   funcs=("tanh", "tanh", "tanh"), **kw):
    # lstm_autoencoder.py:160-174
    if len(dims) == 0:
        raise ValueError("Parameter dims must have len > 0")
    return lstm_model(n_features, n_features_out, lookback_window, encoding_dim=tuple(dims),
                      decoding_dim=tuple(dims)[::-1], encoding_func=tuple(funcs),
                      decoding_func=tuple(funcs)[::-1], **kw)


def lstm_hourglass(n_features, n_features_out=None, lookback_window=1, encoding_layers=3,
                   compression_factor=0.5, func="tanh", **kw):
    # lstm_autoencoder.py:250-263
    dims = hourglass_calc_dims(compression_factor, encoding_layers, n_features)
    return lstm_symmetric(n_features, n_features_out, lookback_window, dims=dims,
                          funcs=tuple([func] * len(dims)), **kw)


FACTORIES = {
    "feedforward_model": feedforward_model,
    "feedforward_symmetric": feedforward_symmetric,
    "feedforward_hourglass": feedforward_hourglass,
    "lstm_model": lstm_model,
    "lstm_symmetric": lstm_symmetric,
    "lstm_hourglass": lstm_hourglass,
}
