"""
LSTM autoencoder / forecast topologies; same names, arguments and validation as
gordo/machine/model/factories/lstm_autoencoder.py:15-263, returning an LSTMTopology.
"""
from typing import Any, Dict, Tuple

from gordo_b200.lstm import LSTMTopology
from gordo_b200.machine.model.register import register_model_builder
from gordo_b200.machine.model.factories.utils import hourglass_calc_dims, check_dim_func_len, adam_from, loss_from


@register_model_builder(type="KerasLSTMAutoEncoder")
@register_model_builder(type="KerasLSTMForecast")
def lstm_model(n_features: int, n_features_out: int = None, lookback_window: int = 1,
               encoding_dim: Tuple[int, ...] = (256, 128, 64),
               encoding_func: Tuple[str, ...] = ("tanh", "tanh", "tanh"),
               decoding_dim: Tuple[int, ...] = (64, 128, 256),
               decoding_func: Tuple[str, ...] = ("tanh", "tanh", "tanh"),
               out_func: str = "linear", optimizer="Adam", optimizer_kwargs: Dict[str, Any] = dict(),
               compile_kwargs: Dict[str, Any] = dict(), **kwargs) -> LSTMTopology:
    """
    Stacked LSTMs (all but the last return sequences) + a Dense output layer; Adam + MSE,
    no metrics, no activity regulariser.
    """
    n_features_out = n_features_out or n_features
    check_dim_func_len("encoding", encoding_dim, encoding_func)
    check_dim_func_len("decoding", decoding_dim, decoding_func)
    loss_from(compile_kwargs, "mse")
    return LSTMTopology(int(n_features), int(n_features_out), [int(u) for u in (*encoding_dim, *decoding_dim)],
                        [*encoding_func, *decoding_func], out_func, int(lookback_window),
                        adam_from(optimizer, optimizer_kwargs))


@register_model_builder(type="KerasLSTMAutoEncoder")
@register_model_builder(type="KerasLSTMForecast")
def lstm_symmetric(n_features: int, n_features_out: int = None, lookback_window: int = 1,
                   dims: Tuple[int, ...] = (256, 128, 64), funcs: Tuple[str, ...] = ("tanh", "tanh", "tanh"),
                   out_func: str = "linear", optimizer="Adam", optimizer_kwargs: Dict[str, Any] = dict(),
                   compile_kwargs: Dict[str, Any] = dict(), **kwargs) -> LSTMTopology:
    if len(dims) == 0:
        raise ValueError("Parameter dims must have len > 0")
    return lstm_model(n_features, n_features_out, lookback_window, encoding_dim=tuple(dims),
                      decoding_dim=tuple(dims)[::-1], encoding_func=tuple(funcs), decoding_func=tuple(funcs)[::-1],
                      out_func=out_func, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                      compile_kwargs=compile_kwargs, **kwargs)


@register_model_builder(type="KerasLSTMAutoEncoder")
@register_model_builder(type="KerasLSTMForecast")
def lstm_hourglass(n_features: int, n_features_out: int = None, lookback_window: int = 1,
                   encoding_layers: int = 3, compression_factor: float = 0.5, func: str = "tanh",
                   out_func: str = "linear", optimizer="Adam", optimizer_kwargs: Dict[str, Any] = dict(),
                   compile_kwargs: Dict[str, Any] = dict(), **kwargs) -> LSTMTopology:
    dims = hourglass_calc_dims(compression_factor, encoding_layers, n_features)
    return lstm_symmetric(n_features, n_features_out, lookback_window, dims=dims,
                          funcs=tuple([func] * len(dims)), out_func=out_func, optimizer=optimizer,
                          optimizer_kwargs=optimizer_kwargs, compile_kwargs=compile_kwargs, **kwargs)
