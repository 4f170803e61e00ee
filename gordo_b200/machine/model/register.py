"""
The ``kind`` registry behind ``KerasAutoEncoder(kind=...)`` (contract of
gordo/machine/model/register.py:10-75): ``register_model_builder(type=T)`` decorates a topology
factory and files it under ``factories[T][function name]``; a factory must accept ``n_features``.
Factories here return a topology (gordo_b200.fleet.FFTopology / gordo_b200.lstm.LSTMTopology)
instead of a compiled Keras model.
"""
import inspect
from collections import defaultdict
from typing import Callable, Dict

_REGISTRY: Dict[str, Dict[str, Callable]] = defaultdict(dict)


def _requires_n_features(factory: Callable) -> None:
    params = inspect.signature(factory).parameters
    if "n_features" not in params:
        raise ValueError(f"Build function: {factory.__name__} does not have 'n_features' as an argument; it should.")


class register_model_builder:
    #: {estimator type name: {kind: factory}} -- shared by every decorator instance, as in the reference
    factories = _REGISTRY

    def __init__(self, type: str):
        self.type = type

    def __call__(self, factory: Callable) -> Callable:
        _requires_n_features(factory)
        _REGISTRY[self.type][factory.__name__] = factory
        return factory
