"""
Minimal definition codec for the model block of a Machine YAML -- the subset of
gordo/serializer/from_definition.py:23-373 / into_definition.py:12-190 the hot path needs:
dotted class paths, keyword parameters, nested definitions, sklearn Pipeline ``steps``, and the
``from_definition`` / ``into_definition`` class hooks.  When gordo itself is installed its own
serializer handles gordo_b200 classes unchanged (they implement the hooks); this module exists so
the package is usable, and testable, without gordo.

``redirect_gordo=True`` maps ``gordo.machine.model.*`` paths onto ``gordo_b200.machine.model.*`` so
an existing project YAML builds on the B200 path without edits.
"""
import importlib
from typing import Any

_REDIRECT = ("gordo.machine.model.", "gordo_b200.machine.model.")


def import_location(path: str, redirect_gordo: bool = False):
    if redirect_gordo and path.startswith(_REDIRECT[0]):
        path = _REDIRECT[1] + path[len(_REDIRECT[0]):]
    module, _, name = path.rpartition(".")
    if not module:
        raise ValueError(f"'{path}' is not a dotted import path")
    return getattr(importlib.import_module(module), name)


def _looks_like_definition(obj) -> bool:
    if isinstance(obj, str):
        return "." in obj and " " not in obj and obj[0].isalpha() and _importable(obj)
    return isinstance(obj, dict) and len(obj) == 1 and isinstance(next(iter(obj)), str) \
        and _looks_like_definition(next(iter(obj)))


def _importable(path: str) -> bool:
    try:
        import_location(path, redirect_gordo=True)
        return True
    except Exception:
        return False


def from_definition(definition: Any, redirect_gordo: bool = False):
    """Build the object graph described by ``definition`` (str path or {path: params})."""
    if isinstance(definition, str):
        return import_location(definition, redirect_gordo)()
    if not (isinstance(definition, dict) and len(definition) == 1):
        raise ValueError(f"not a model definition: {definition!r}")
    path, params = next(iter(definition.items()))
    cls = import_location(path, redirect_gordo)
    params = dict(params or {})
    if hasattr(cls, "from_definition"):
        return cls.from_definition(params)
    return cls(**_load_params(params, redirect_gordo))


def _load_params(params: dict, redirect_gordo: bool):
    out = {}
    for key, value in params.items():
        if key == "steps" and isinstance(value, list):
            steps = []
            for i, step in enumerate(value):
                obj = from_definition(step, redirect_gordo) if _looks_like_definition(step) else step
                steps.append(obj if isinstance(obj, tuple) else (f"step_{i}", obj))
            out[key] = steps
        elif _looks_like_definition(value):
            out[key] = from_definition(value, redirect_gordo)
        elif isinstance(value, dict):
            out[key] = _load_params(value, redirect_gordo)
        else:
            out[key] = value
    return out


def into_definition(obj) -> Any:
    """Inverse of from_definition for the objects this package builds."""
    path = f"{obj.__class__.__module__}.{obj.__class__.__name__}"
    # looked up on the class: the detectors are transparent into base_estimator (diff.py:78-86), so an instance
    # lookup would find a bare KerasAutoEncoder's hook and describe the detector with the estimator's parameters
    if hasattr(type(obj), "into_definition"):
        return {path: obj.into_definition()}
    params = obj.get_params(deep=False) if hasattr(obj, "get_params") else {}
    out = {}
    for k, v in params.items():
        if k == "steps":
            out[k] = [into_definition(s[1]) for s in v]
        elif hasattr(type(v), "get_params") or hasattr(type(v), "into_definition"):
            out[k] = into_definition(v)
        else:
            out[k] = v
    return {path: out}
