"""Layer-width helpers; behaviour of gordo/machine/model/factories/utils.py:7-63."""
import math
from typing import Tuple


def hourglass_calc_dims(compression_factor: float, encoding_layers: int, n_features: int) -> Tuple[int, ...]:
    """
    Widths of the encoder layers of an hourglass network, reversed for the decoder.

    >>> hourglass_calc_dims(0.5, 3, 10)
    (8, 7, 5)
    >>> hourglass_calc_dims(0.2, 4, 5)
    (4, 3, 2, 1)
    """
    if not (1 >= compression_factor >= 0):
        raise ValueError("compression_factor must be 0 <= compression_factor <= 1")
    if encoding_layers < 1:
        raise ValueError("encoding_layers must be >= 1")
    narrowest = max(min(math.ceil(compression_factor * n_features), n_features), 1)
    step = (n_features - narrowest) / encoding_layers
    return tuple(round(n_features - step * i) for i in range(1, encoding_layers + 1))


def check_dim_func_len(prefix: str, dim: Tuple[int, ...], func: Tuple[str, ...]):
    if len(dim) != len(func):
        raise ValueError(
            f"The length (i.e. the number of network layers) of {prefix}_dim ({len(dim)}) and "
            f"{prefix}_func ({len(func)}) must be equal. If only {prefix}_dim or {prefix}_func was "
            f"passed, ensure that its length matches that of the {prefix} parameter not passed.")


def adam_from(optimizer, optimizer_kwargs):
    """Keras optimizer spec -> Adam hyper-parameters (the only optimizer the kernels implement)."""
    name = optimizer if isinstance(optimizer, str) else getattr(optimizer, "__name__", type(optimizer).__name__)
    if str(name).lower() != "adam":
        raise ValueError(f"optimizer {optimizer!r} is not supported by gordo_b200 (Adam only)")
    kw = dict(optimizer_kwargs or {})
    unknown = set(kw) - {"learning_rate", "lr", "beta_1", "beta_2", "epsilon", "name"}
    if unknown:
        raise ValueError(f"unsupported Adam arguments: {sorted(unknown)}")
    return dict(lr=float(kw.get("learning_rate", kw.get("lr", 1e-3))), beta_1=float(kw.get("beta_1", 0.9)),
                beta_2=float(kw.get("beta_2", 0.999)), epsilon=float(kw.get("epsilon", 1e-7)))


def loss_from(compile_kwargs, default):
    loss = dict(compile_kwargs or {}).get("loss", default)
    if loss not in ("mean_squared_error", "mse"):
        raise ValueError(f"loss {loss!r} is not supported by gordo_b200 (mean_squared_error only)")
    return "mean_squared_error"
