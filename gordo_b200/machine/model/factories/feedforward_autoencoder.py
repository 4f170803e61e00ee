"""
Feed-forward autoencoder topologies; same names, arguments and validation as
gordo/machine/model/factories/feedforward_autoencoder.py:15-251, returning an FFTopology.
"""
from typing import Any, Dict, Tuple

from gordo_b200.fleet import FFTopology
from gordo_b200.machine.model.register import register_model_builder
from gordo_b200.machine.model.factories.utils import hourglass_calc_dims, check_dim_func_len, adam_from, loss_from


@register_model_builder(type="KerasAutoEncoder")
def feedforward_model(n_features: int, n_features_out: int = None,
                      encoding_dim: Tuple[int, ...] = (256, 128, 64),
                      encoding_func: Tuple[str, ...] = ("tanh", "tanh", "tanh"),
                      decoding_dim: Tuple[int, ...] = (64, 128, 256),
                      decoding_func: Tuple[str, ...] = ("tanh", "tanh", "tanh"),
                      out_func: str = "linear", optimizer="Adam",
                      optimizer_kwargs: Dict[str, Any] = dict(), compile_kwargs: Dict[str, Any] = dict(),
                      **kwargs) -> FFTopology:
    """
    Dense encoder (``encoding_dim`` / ``encoding_func``), Dense decoder, linear-by-default output
    layer; L1 activity regulariser 10e-5 on every encoder layer but the first; Adam + MSE.
    """
    n_features_out = n_features_out or n_features
    check_dim_func_len("encoding", encoding_dim, encoding_func)
    check_dim_func_len("decoding", decoding_dim, decoding_func)
    widths = [n_features, *encoding_dim, *decoding_dim, n_features_out]
    acts = [*encoding_func, *decoding_func, out_func]
    l1 = [0.0] + [10e-5] * (len(encoding_dim) - 1) + [0.0] * (len(decoding_dim) + 1)
    loss_from(compile_kwargs, "mean_squared_error")
    return FFTopology([int(w) for w in widths], list(acts), l1, adam_from(optimizer, optimizer_kwargs))


@register_model_builder(type="KerasAutoEncoder")
def feedforward_symmetric(n_features: int, n_features_out: int = None, dims: Tuple[int, ...] = (256, 128, 64),
                          funcs: Tuple[str, ...] = ("tanh", "tanh", "tanh"), optimizer="Adam",
                          optimizer_kwargs: Dict[str, Any] = dict(), compile_kwargs: Dict[str, Any] = dict(),
                          **kwargs) -> FFTopology:
    if len(dims) == 0:
        raise ValueError("Parameter dims must have len > 0")
    return feedforward_model(n_features, n_features_out, encoding_dim=tuple(dims), decoding_dim=tuple(dims)[::-1],
                             encoding_func=tuple(funcs), decoding_func=tuple(funcs)[::-1], optimizer=optimizer,
                             optimizer_kwargs=optimizer_kwargs, compile_kwargs=compile_kwargs, **kwargs)


@register_model_builder(type="KerasAutoEncoder")
def feedforward_hourglass(n_features: int, n_features_out: int = None, encoding_layers: int = 3,
                          compression_factor: float = 0.5, func: str = "tanh", optimizer="Adam",
                          optimizer_kwargs: Dict[str, Any] = dict(), compile_kwargs: Dict[str, Any] = dict(),
                          **kwargs) -> FFTopology:
    """
    >>> feedforward_hourglass(10).widths
    [10, 8, 7, 5, 5, 7, 8, 10]
    >>> feedforward_hourglass(10, compression_factor=0.2).widths
    [10, 7, 5, 2, 2, 5, 7, 10]
    """
    dims = hourglass_calc_dims(compression_factor, encoding_layers, n_features)
    return feedforward_symmetric(n_features, n_features_out, dims=dims, funcs=tuple([func] * len(dims)),
                                 optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                                 compile_kwargs=compile_kwargs, **kwargs)
