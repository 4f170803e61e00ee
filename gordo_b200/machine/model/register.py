"""
register_model_builder -- the `kind` factory registry of gordo/machine/model/register.py:10-75:
``factories[type][kind]``, keyed by the decorated function's ``__name__``; a factory must take
``n_features``.  Factories here return a topology (gordo_b200.fleet.FFTopology or
gordo_b200.lstm.LSTMTopology) instead of a compiled Keras model.
"""
import inspect
from typing import Callable, Dict


class register_model_builder:
    factories: Dict[str, Dict[str, Callable]] = dict()

    def __init__(self, type: str):
        self.type = type

    def __call__(self, model: Callable):
        self._register(self.type, model)
        return model

    @classmethod
    def _register(cls, type: str, model: Callable):
        cls._validate_func(model)
        if type not in cls.factories:
            cls.factories[type] = dict()
        cls.factories[type][model.__name__] = model

    @staticmethod
    def _validate_func(func):
        if "n_features" not in inspect.getfullargspec(func).args:
            raise ValueError(
                f"Build function: {func.__name__} does not have 'n_features' as an argument; it should.")
