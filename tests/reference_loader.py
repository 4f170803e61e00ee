"""
Load the REAL reference modules of the builder / serializer seam from /root/reference with their absent
third-party dependencies stubbed (TensorFlow, Keras, gordo-core, xarray, simplejson, dataclasses-json) --
the same technique tests/golden/make_golden.py uses for diff.py.  Test infrastructure only; nothing here is
copied from the reference, its files are executed where they lie.  Only usable where /root/reference exists
(the build container; not the GPU box).
"""
import importlib.util
import json
import os
import pydoc
import sys
import types
from dataclasses import dataclass, field, asdict
from typing import Any, Dict, Optional

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "gordo"))


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
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


class StubDataset:
    """Stands in for gordo_core's dataset: hands back the frames it was built from."""
    registry: Dict[str, Any] = {}

    def __init__(self, key):
        self.key = key

    @classmethod
    def from_dict(cls, d):
        return cls(d["key"])

    def to_dict(self):
        return {"key": self.key}

    def get_data(self):
        return StubDataset.registry[self.key]

    def get_metadata(self):
        return {"stub": True}


@dataclass
class _CV:
    scores: Dict[str, Any] = field(default_factory=dict)
    cv_duration_sec: Optional[float] = None
    splits: Dict[str, Any] = field(default_factory=dict)


@dataclass
class _ModelBuild:
    model_offset: int = 0
    model_creation_date: Optional[str] = None
    model_builder_version: str = "0.0.0"
    cross_validation: _CV = field(default_factory=_CV)
    model_training_duration_sec: Optional[float] = None
    model_meta: Dict[str, Any] = field(default_factory=dict)


@dataclass
class _DatasetBuild:
    query_duration_sec: Optional[float] = None
    dataset_meta: Dict[str, Any] = field(default_factory=dict)


@dataclass
class _Build:
    model: _ModelBuild = field(default_factory=_ModelBuild)
    dataset: _DatasetBuild = field(default_factory=_DatasetBuild)


@dataclass
class _Metadata:
    user_defined: Dict[str, Any] = field(default_factory=dict)
    build_metadata: _Build = field(default_factory=_Build)

    def to_dict(self):
        return asdict(self)


class StubMachine:
    """The attributes ModelBuilder touches of gordo.machine.Machine (machine.py), nothing more."""

    def __init__(self, name, model, dataset, evaluation=None, metadata=None, project_name="p", runtime=None):
        self.name, self.model, self.project_name = name, model, project_name
        self.dataset = dataset if isinstance(dataset, StubDataset) else StubDataset.from_dict(dataset)
        self.evaluation = dict(evaluation or {"cv_mode": "full_build"})
        self.metadata = metadata if isinstance(metadata, _Metadata) else _Metadata(**(metadata or {})) \
            if not isinstance(metadata, dict) or "build_metadata" not in (metadata or {}) else _Metadata()
        self.runtime = runtime or {}

    @classmethod
    def from_dict(cls, d, back_compatibles=None, default_data_provider=None):
        return cls(d["name"], d["model"], d["dataset"], d.get("evaluation"), None, d.get("project_name", "p"), d.get("runtime"))

    def to_dict(self):
        return {"name": self.name, "model": self.model, "dataset": self.dataset.to_dict(), "evaluation": self.evaluation,
                "metadata": self.metadata.to_dict(), "project_name": self.project_name, "runtime": self.runtime}


def load():
    """Returns {'serializer', 'build_model', 'builder_utils'}: the reference's own modules."""
    import pandas as pd
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, **_):
            return pd.concat([self, other.to_frame().T])
        pd.DataFrame.append = _append

    class _Anything:
        def __init__(self, *a, **k):
            pass
    tf = _stub("tensorflow", random=types.SimpleNamespace(set_seed=lambda s: None))
    tfk = _stub("tensorflow.keras", Sequential=type("Sequential", (), {}))
    tf.keras = tfk
    _stub("keras"); _stub("keras.src"); _stub("keras.src.callbacks", Callback=type("Callback", (), {}))
    _stub("xarray", DataArray=_Anything, Dataset=_Anything)
    _stub("simplejson", **{k: getattr(json, k) for k in ("dump", "dumps", "load", "loads")})

    def import_location(location, import_path=None):
        obj = pydoc.locate(location)
        if obj is None:
            raise ImportError(f'Unable to import "{location}"')
        return obj
    _stub("gordo_core")
    _stub("gordo_core.import_utils", import_location=import_location, BackCompatibleLocations=dict)
    _stub("gordo_core.base", GordoBaseDataset=StubDataset)
    _stub("gordo_core.sensor_tag", SensorTag=type("SensorTag", (), {"__init__": lambda self, name: setattr(self, "name", name)}))

    g = os.path.join(REF, "gordo")
    _load("gordo", os.path.join(g, "__init__.py")).__path__ = [g]
    _pkg("gordo.machine", os.path.join(g, "machine"))
    sys.modules["gordo.machine"].Machine = StubMachine
    sys.modules["gordo.machine"].load_model_config = lambda metadata: metadata
    _stub("gordo.machine.metadata", BuildMetadata=_Build, ModelBuildMetadata=_ModelBuild,
          DatasetBuildMetadata=_DatasetBuild, CrossValidationMetaData=_CV, Metadata=_Metadata)
    _pkg("gordo.machine.model", os.path.join(g, "machine", "model"))
    _load("gordo.machine.model.base", os.path.join(g, "machine", "model", "base.py"))
    _load("gordo.machine.model.utils", os.path.join(g, "machine", "model", "utils.py"))
    _pkg("gordo.util", os.path.join(g, "util"))
    _load("gordo.util.disk_registry", os.path.join(g, "util", "disk_registry.py"))
    _pkg("gordo.workflow", os.path.join(g, "workflow"))
    _pkg("gordo.workflow.config_elements", os.path.join(g, "workflow", "config_elements"))
    # normalized_config.py:66-106 needs gordo-core / yaml machinery; ModelBuilder only reads the evaluation defaults
    nc = type("NormalizedConfig", (), {"DEFAULT_CONFIG_GLOBALS": {"evaluation": {
        "cv_mode": "full_build", "scoring_scaler": "sklearn.preprocessing.MinMaxScaler",
        "metrics": ["explained_variance_score", "r2_score", "mean_squared_error", "mean_absolute_error"]}}})
    _stub("gordo.workflow.config_elements.normalized_config", NormalizedConfig=nc)
    _pkg("gordo.serializer", os.path.join(g, "serializer"))
    _load("gordo.serializer.utils", os.path.join(g, "serializer", "utils.py"))
    fd = _load("gordo.serializer.from_definition", os.path.join(g, "serializer", "from_definition.py"))
    idf = _load("gordo.serializer.into_definition", os.path.join(g, "serializer", "into_definition.py"))
    ser = _load("gordo.serializer.serializer", os.path.join(g, "serializer", "serializer.py"))
    s = sys.modules["gordo.serializer"]
    for mod in (fd, idf, ser):
        for k, v in mod.__dict__.items():
            if callable(v) and not k.startswith("_"):
                setattr(s, k, v)
    sys.modules["gordo"].serializer = s
    _pkg("gordo.builder", os.path.join(g, "builder"))
    bm = _load("gordo.builder.build_model", os.path.join(g, "builder", "build_model.py"))
    bu = _load("gordo.builder.utils", os.path.join(g, "builder", "utils.py"))
    return {"serializer": s, "build_model": bm, "builder_utils": bu}


def unload():
    for k in [k for k in sys.modules if k == "gordo" or k.startswith("gordo.") or k.startswith("gordo_core")
              or k in ("tensorflow", "tensorflow.keras", "keras", "keras.src", "keras.src.callbacks", "xarray", "simplejson")]:
        del sys.modules[k]
