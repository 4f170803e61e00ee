"""
sklearn-compatible estimators with the surface of gordo/machine/model/models.py:

    KerasBaseEstimator (:36-357), KerasAutoEncoder (:360-398),
    KerasLSTMBaseEstimator (:463-698), KerasLSTMForecast (:701), KerasLSTMAutoEncoder (:707)

Same constructor (``kind`` + kwargs), ``from_definition`` / ``into_definition`` hooks, ``fit`` /
``predict`` / ``score`` / ``get_params`` / ``get_metadata``, error behaviour and pickling contract
-- but ``self.model`` is a topology + a float32 parameter vector and every fit / predict is a
launch of libgordo_b200.so on the current CUDA device (a fleet of ONE Machine; fleets of thousands
go through gordo_b200.builder / gordo_b200.fleet).  There is no CPU fallback.

Selectable purely from Machine YAML:
    gordo_b200.machine.model.models.KerasAutoEncoder:
        kind: feedforward_hourglass
"""
import importlib
import logging
import threading
from copy import copy, deepcopy
from importlib.util import find_spec
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np
import pandas as pd
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.exceptions import NotFittedError
from sklearn.metrics import explained_variance_score

from gordo_b200.machine.model.base import GordoBase
from gordo_b200.machine.model.factories import *  # noqa: F401,F403  (registers the factories)
from gordo_b200.machine.model.register import register_model_builder

logger = logging.getLogger(__name__)

# one fitted model object is shared by the server's gthread workers (gordo/server/utils.py:334-335):
# creating its device-side serving state is serialised, predicting is not
_SERVING_LOCK = threading.RLock()


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("gordo_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def _values(a):
    return a.values if hasattr(a, "values") and not isinstance(a, np.ndarray) else a


class _Model:
    """What ``self.model`` holds: the topology, the float32 parameters, the optimizer state."""

    def __init__(self, topology, params: np.ndarray):
        self.topology = topology
        self.params = np.ascontiguousarray(params, np.float32)
        self.adam_mv: Optional[np.ndarray] = None
        self.adam_t: int = 0

    def count_params(self) -> int:
        return int(self.params.size)


class History:
    """Stand-in for keras.callbacks.History (attributes read at models.py:339-357)."""

    def __init__(self, history=None, params=None, epoch=None):
        self.history = history or {}
        self.params = params or {}
        self.epoch = epoch or []


class EarlyStopping:
    """
    [3P] keras.callbacks.EarlyStopping (Keras 3.3.3 semantics) for the per-epoch loop of ``fit``:
    monitor / min_delta / patience / mode / baseline / restore_best_weights / start_from_epoch.
    The best weights are restored at the end of training whenever ``restore_best_weights`` is set.
    """

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False, start_from_epoch=0):
        self.monitor, self.min_delta, self.patience = monitor, abs(float(min_delta)), int(patience)
        self.baseline, self.restore_best_weights, self.start_from_epoch = baseline, bool(restore_best_weights), int(start_from_epoch)
        if mode not in ("auto", "min", "max"):
            mode = "auto"
        if mode == "auto":
            mode = "max" if (monitor.endswith("acc") or monitor.endswith("accuracy") or monitor.endswith("auc")) else "min"
        self.mode = mode
        self.wait, self.stopped_epoch, self.best_epoch, self.best_weights = 0, 0, 0, None
        self.best = np.inf if mode == "min" else -np.inf

    def _is_improvement(self, value, reference):
        return value < reference - self.min_delta if self.mode == "min" else value > reference + self.min_delta

    def on_epoch_end(self, epoch, logs, get_weights) -> bool:
        """Returns True when training must stop."""
        current = logs.get(self.monitor)
        if current is None or epoch < self.start_from_epoch:
            return False
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights, self.best_epoch = get_weights(), epoch
        self.wait += 1
        if self._is_improvement(current, self.best):
            self.best, self.best_epoch = current, epoch
            if self.restore_best_weights:
                self.best_weights = get_weights()
            if self.baseline is None or self._is_improvement(current, self.baseline):
                self.wait = 0
            return False
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            return True
        return False


def build_callbacks(definitions):
    """
    Callback definitions of a Machine YAML (gordo/serializer/from_definition.py:352-373) -> callback
    objects.  Supported: ``tensorflow.keras.callbacks.EarlyStopping`` (any ``...EarlyStopping`` path) or an
    object exposing the same attributes; anything else fails loudly.
    """
    out = []
    for cb in definitions or []:
        if isinstance(cb, EarlyStopping):
            out.append(cb)
        elif isinstance(cb, dict) and len(cb) == 1 and str(next(iter(cb))).endswith("EarlyStopping"):
            out.append(EarlyStopping(**(next(iter(cb.values())) or {})))
        elif isinstance(cb, str) and cb.endswith("EarlyStopping"):
            out.append(EarlyStopping())
        elif type(cb).__name__ == "EarlyStopping":
            out.append(EarlyStopping(**{k: getattr(cb, k) for k in ("monitor", "min_delta", "patience", "mode", "baseline",
                                                                    "restore_best_weights", "start_from_epoch") if hasattr(cb, k)}))
        else:
            raise NotImplementedError(f"Keras callback {cb!r} is not supported by gordo_b200 (EarlyStopping only)")
    return out


class KerasBaseEstimator(BaseEstimator, GordoBase):
    supported_fit_args = ["batch_size", "epochs", "verbose", "callbacks", "validation_split", "shuffle",
                          "class_weight", "initial_epoch", "steps_per_epoch", "validation_batch_size",
                          "max_queue_size", "workers", "use_multiprocessing"]
    # kwargs consumed by this implementation and never forwarded to a factory
    _b200_args = ("precision", "l1_batch_norm")

    def __init__(self, kind: Union[str, Callable], **kwargs) -> None:
        """
        kind: a registered factory name, a dotted path to a factory, or a factory callable
        (registered on the fly) -- models.py:53-94.  kwargs: factory arguments and/or fit
        arguments (epochs, batch_size, shuffle, validation_split ...), plus ``precision``
        ("f32" default | "bf16": tensor-core inference | "f16x3": fp32-grade tensor-core inference) and ``l1_batch_norm`` ("sum" default, as
        Keras 3.3.3 | "mean").
        """
        self.kind = self.load_kind(kind)
        self.kwargs: Dict[str, Any] = kwargs
        self._history = None
        self.model = None

    # ------------------------------------------------------------------ kind / definition codec
    @staticmethod
    def parse_module_path(module_path) -> Tuple[Optional[str], str]:
        parts = module_path.split(".")
        return (None, parts[0]) if len(parts) == 1 else (".".join(parts[:-1]), parts[-1])

    def load_kind(self, kind):
        if callable(kind):
            register_model_builder(type=self.__class__.__name__)(kind)
            return kind.__name__
        module_name, class_name = self.parse_module_path(kind)
        if module_name is None:
            if class_name not in register_model_builder.factories.get(self.__class__.__name__, {}):
                raise ValueError(f"kind: {kind} is not an available model for type: {class_name}!")
        else:
            has_error = True
            try:
                has_error = not find_spec(module_name)
            except ModuleNotFoundError:
                pass
            if has_error:
                raise ValueError(f"kind: {kind}, unable to find module: '{module_name}'")
        return kind

    @classmethod
    def extract_supported_fit_args(cls, kwargs):
        return {arg: kwargs[arg] for arg in cls.supported_fit_args if arg in kwargs}

    @classmethod
    def from_definition(cls, definition: dict):
        """Handler for gordo.serializer.from_definition (models.py:146-159)."""
        definition = copy(definition)
        kind = definition.pop("kind")
        return cls(kind, **definition)

    def into_definition(self) -> dict:
        """Handler for gordo.serializer.into_definition (models.py:161-171)."""
        definition = copy(self.kwargs)
        definition["kind"] = self.kind
        return definition

    @property
    def sk_params(self):
        return self.kwargs

    def get_params(self, deep=False, **params):
        out = {"kind": self.kind}
        out.update(self.kwargs)
        return out

    def __sklearn_is_fitted__(self):
        # sklearn.pipeline.Pipeline.predict/score call check_is_fitted on the last step
        return self.model is not None

    def set_params(self, **params):
        if "kind" in params:
            self.kind = self.load_kind(params.pop("kind"))
        self.kwargs.update(params)
        return self

    def __repr__(self):
        return f"{self.__class__.__name__}(kind={self.kind!r}, " + ", ".join(f"{k}={v!r}" for k, v in self.kwargs.items()) + ")"

    # ------------------------------------------------------------------ pickling (models.py:185-210)
    def __getstate__(self):
        # numpy + plain python only; the device-side serving cache never travels with the pickle
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_gb200_")}

    def __setstate__(self, state):
        self.__dict__ = state
        return self

    # ------------------------------------------------------------------ shapes
    @staticmethod
    def get_n_features_out(y) -> Union[int, tuple]:
        if len(y.shape) == 1:
            raise ValueError("Unsupported number of the output dataset dimensions %d" % len(y.shape))
        return y.shape[1] if len(y.shape) == 2 else y.shape[1:]

    @staticmethod
    def get_n_features(X) -> Union[int, tuple]:
        if len(X.shape) == 1:
            raise ValueError("Unsupported number of the output dataset dimensions %d" % len(X.shape))
        return X.shape[1] if len(X.shape) == 2 else X.shape[2]

    # ------------------------------------------------------------------ model construction
    def _factory(self):
        module_name, class_name = self.parse_module_path(self.kind)
        if module_name is None:
            return register_model_builder.factories[self.__class__.__name__][self.kind]
        module = importlib.import_module(module_name)
        if not hasattr(module, class_name):
            raise ValueError("kind: %s, unable to find class %s in module '%s'" % (self.kind, class_name, module_name))
        return getattr(module, class_name)

    def _topology(self):
        kw = {k: v for k, v in self.sk_params.items() if k not in self._b200_args}
        return self._factory()(**deepcopy(kw))

    def _fit_arg(self, name, default, overrides):
        if name in overrides:
            return overrides[name]
        return self.kwargs.get(name, default)

    @property
    def _precision(self) -> str:
        p = self.kwargs.get("precision", "f32")
        if p not in ("f32", "bf16", "f16x3"):
            raise ValueError("precision must be 'f32', 'bf16' or 'f16x3'")
        return p

    def get_metadata(self):
        """{"history": {metric: [per epoch], "params": {...}}} once fitted (models.py:339-357)."""
        if self._history is not None:
            history = self._history.history
            history["params"] = self._history.params
            return {"history": history}
        return {}


class KerasAutoEncoder(KerasBaseEstimator, TransformerMixin):
    """Feed-forward autoencoder estimator (models.py:360-398)."""

    def fit(self, X, y, **kwargs):
        torch = _torch()
        from gordo_b200.fleet import FFFleet
        if isinstance(y, np.ndarray) and y.ndim == 1:
            y = y.reshape(-1, 1)
        self.kwargs.update({"n_features": self.get_n_features(X), "n_features_out": self.get_n_features_out(y)})
        same = y is X
        X = np.asarray(_values(X)); y = X if same else np.asarray(_values(y))
        callbacks = build_callbacks(self._fit_arg("callbacks", None, kwargs))
        epochs = int(self._fit_arg("epochs", 1, kwargs))
        batch_size = int(self._fit_arg("batch_size", None, kwargs) or 32)
        shuffle = bool(self._fit_arg("shuffle", True, kwargs))
        vsplit = float(self._fit_arg("validation_split", 0.0, kwargs) or 0.0)
        l1_mean = self.kwargs.get("l1_batch_norm", "sum") == "mean"

        dev = torch.device("cuda", torch.cuda.current_device())
        seed = int(np.random.randint(0, 2 ** 31 - 1))            # np.random.seed(...) => reproducible builds
        gen = torch.Generator(device=dev); gen.manual_seed(seed)
        if self.model is None:
            topo = self._topology()
            if topo.n_in != X.shape[1] or topo.n_out != y.shape[1]:
                raise ValueError("factory topology does not match the data shape")
            self.model = _Model(topo, topo.glorot_init(1, gen, dev)[0].cpu().numpy())
        topo = self.model.topology
        n = len(X)
        n_train = n if not vsplit else int(np.floor(n * (1.0 - vsplit)))
        xd = torch.as_tensor(np.ascontiguousarray(X, np.float32), device=dev)
        yd = None if same or (X.shape == y.shape and np.array_equal(X, y)) else \
            torch.as_tensor(np.ascontiguousarray(y, np.float32), device=dev)
        fleet = FFFleet(topo, 1, dev)
        params = torch.as_tensor(self.model.params[None].copy(), device=dev)
        mv = None if self.model.adam_mv is None else torch.as_tensor(self.model.adam_mv[None].copy(), device=dev)
        t = torch.tensor([self.model.adam_t], dtype=torch.int64, device=dev)
        lo = torch.zeros(1, dtype=torch.int64, device=dev); hi = torch.full((1,), n_train, dtype=torch.int64, device=dev)
        hist = {"loss": [], "accuracy": []}
        if vsplit:
            hist["val_loss"] = []
        # Keras evaluates the validation split and runs callbacks after every epoch: one epoch per launch then
        per_epoch = bool(vsplit) or bool(callbacks)
        epochs_run = 0
        for e0 in range(0, epochs, 1 if per_epoch else epochs):
            ne = 1 if per_epoch else epochs
            pool = poff = None
            if shuffle and n_train > 0:
                pool = torch.stack([torch.randperm(n_train, generator=gen, device=dev) for _ in range(ne)]
                                   ).to(torch.int32).reshape(-1).contiguous()
                poff = torch.zeros(1, dtype=torch.int64, device=dev)
            hl, ha, mv, t = fleet.fit_jobs(xd, yd, lo, hi, params, epochs=ne, batch_size=batch_size,
                                           perm_pool=pool, perm_off=poff, l1_mean=l1_mean, adam_mv=mv, adam_t=t)
            hist["loss"] += [float(v) for v in hl[0].cpu()]
            hist["accuracy"] += [float(v) for v in ha[0].cpu()]
            epochs_run += ne
            if vsplit and n > n_train:
                # val_loss = MSE + activity-regulariser loss, batch-size weighted like Keras' evaluate()
                from gordo_b200.fleet import Schedule
                fleet.set_params(params)
                vs = Schedule(rows_lo=[n_train], rows_hi=[n], rows_total=n)
                vp = fleet.score(vs, xd, yd, precision="f32", columns=("total-anomaly-unscaled", "activity-l1"))
                mse = float(vp["total-anomaly-unscaled"][n_train:n].mean())
                act = vp["activity-l1"][n_train:n].double()
                nv = n - n_train
                if l1_mean:
                    reg = float(act.mean())
                else:           # per batch the activity loss is SUMMED over its rows: weight by the batch sizes
                    sizes = torch.full((-(-nv // batch_size),), batch_size, dtype=torch.float64, device=dev)
                    sizes[-1] = nv - batch_size * (len(sizes) - 1)
                    per_batch = torch.zeros_like(sizes).index_add_(0, torch.arange(nv, device=dev) // batch_size, act)
                    reg = float((per_batch * sizes).sum() / nv)
                hist["val_loss"].append(mse + reg)
            if callbacks:
                logs = {k: v[-1] for k, v in hist.items() if v}
                get_w = lambda: (params.clone(), mv.clone(), t.clone())
                if any(cb.on_epoch_end(epochs_run - 1, logs, get_w) for cb in callbacks):
                    break
        for cb in callbacks:
            if cb.restore_best_weights and cb.best_weights is not None:
                params = cb.best_weights[0].clone()
        self.model.params = params[0].cpu().numpy()
        self.model.adam_mv = mv[0].cpu().numpy()
        self.model.adam_t = int(t[0])
        steps = -(-n_train // batch_size)
        self._history = History(hist, {"verbose": 0, "epochs": epochs, "steps": steps}, list(range(epochs_run)))
        return self

    def predict(self, X, **kwargs) -> np.ndarray:
        """ŷ [len(X), n_features_out] float32 (models.py:289-300)."""
        torch = _torch()
        from gordo_b200.fleet import FFFleet, Schedule
        if self.model is None:
            raise NotFittedError(f"This {self.__class__.__name__} has not been fitted yet.")
        X = np.asarray(_values(X))
        topo = self.model.topology
        if X.ndim != 2 or X.shape[1] != topo.n_in:
            raise ValueError(f"X must be [n, {topo.n_in}], got {X.shape}")
        dev = torch.device("cuda", torch.cuda.current_device())
        # the device copy of the weights (and their bf16 operand image) is kept between calls
        key = (dev.index, id(self.model), hash(self.model.params.tobytes()))
        with _SERVING_LOCK:
            cached = self.__dict__.get("_gb200_serving")
            if cached is not None and cached[0] == key:
                fleet = cached[1]
            else:
                fleet = FFFleet(topo, 1, dev)
                fleet.set_params(torch.as_tensor(self.model.params[None], device=dev))
                self.__dict__["_gb200_serving"] = (key, fleet)
        xd = torch.as_tensor(np.ascontiguousarray(X, np.float32), device=dev)
        prec = fleet.auto_precision(self._precision)
        return fleet.predict(Schedule.single(len(X)), xd, precision=prec).cpu().numpy()

    def transform(self, X, **kwargs) -> np.ndarray:
        return self.predict(X, **kwargs)

    def score(self, X, y, sample_weight: Optional[np.ndarray] = None, **kwargs) -> float:
        """Explained variance of the reconstruction (models.py:365-398)."""
        if self.model is None:
            raise NotFittedError(f"This {self.__class__.__name__} has not been fitted yet.")
        return explained_variance_score(_values(y), self.predict(X, **kwargs))


class KerasRawModelRegressor(KerasAutoEncoder):
    """
    A regressor from a raw Keras config (models.py:401-460): ``kind = {"spec": {...Sequential: {layers: [...]}},
    "compile": {...}}``.  The reference hands the spec to Keras and accepts any graph; here the subset the B200 kernels
    run is accepted -- a ``Sequential`` of ``Dense`` layers (``units``, ``activation``, ``input_shape`` / ``input_dim``,
    optional ``activity_regularizer`` L1), ``loss: mse`` and an Adam optimizer -- and built as a feed-forward topology;
    anything else (other layers, SGD, kernel regularisers, ``use_bias: false``) raises NotImplementedError naming it.
    """

    _expected_keys = ("spec", "compile")

    def load_kind(self, kind):
        return kind

    def __repr__(self):
        from pprint import pformat
        return f"{self.__class__.__name__}(kind: {pformat(self.kind)})"

    @staticmethod
    def _single(defn, what):
        if isinstance(defn, str):
            return defn, {}
        if isinstance(defn, dict) and len(defn) == 1:
            k, v = next(iter(defn.items()))
            return k, dict(v or {})
        raise ValueError(f"{what}: expected a class path or a one-key mapping, got {defn!r}")

    def _topology(self):
        from gordo_b200.fleet import FFTopology
        from gordo_b200.machine.model.factories.utils import adam_from, loss_from
        if not all(k in self.kind for k in self._expected_keys):
            raise ValueError(f"Expected spec to have keys: {self._expected_keys}, but found {self.kind.keys()}")
        path, body = self._single(self.kind["spec"], "spec")
        if not path.endswith("Sequential"):
            raise NotImplementedError(f"KerasRawModelRegressor on B200 runs Sequential models of Dense layers, not {path}")
        widths, acts, l1 = [self.kwargs.get("n_features")], [], []
        for i, layer in enumerate(body.get("layers", [])):
            lpath, lkw = self._single(layer, f"layer {i}")
            if not lpath.endswith(".Dense"):
                raise NotImplementedError(f"layer {i} ({lpath}): only Dense layers run on the B200 feed-forward kernels")
            lkw = dict(lkw)
            shape = lkw.pop("input_shape", None)
            n_in = lkw.pop("input_dim", shape[-1] if shape else None)
            if i == 0 and n_in is not None:
                widths[0] = int(n_in)
            reg = lkw.pop("activity_regularizer", None)
            strength = 0.0
            if reg is not None:
                rpath, rkw = self._single(reg, f"layer {i} activity_regularizer")
                if set(rkw) - {"l1", "l"} or not (rpath.lower().endswith("l1") or rpath.endswith("L1L2")):
                    raise NotImplementedError(f"layer {i}: activity regulariser {reg!r} (only L1 is implemented)")
                strength = float(rkw.get("l1", rkw.get("l", 0.01)))
            units = lkw.pop("units", None)
            act = lkw.pop("activation", None) or "linear"
            if units is None:
                raise ValueError(f"layer {i}: Dense needs `units`")
            if lkw.pop("use_bias", True) is not True:
                raise NotImplementedError(f"layer {i}: use_bias=False")
            lkw.pop("name", None)
            if lkw:
                raise NotImplementedError(f"layer {i}: Dense arguments {sorted(lkw)} are not implemented on B200")
            widths.append(int(units)); acts.append(act); l1.append(strength)
        if len(widths) < 2:
            raise ValueError("the Sequential spec holds no layers")
        if widths[0] is None:
            raise ValueError("the first Dense layer needs input_shape / input_dim (or call fit first)")
        compile_kw = dict(self.kind["compile"] or {})
        loss_from(compile_kw, "mean_squared_error")
        opath, okw = self._single(compile_kw.get("optimizer", "adam"), "compile.optimizer")
        if opath.rpartition(".")[2].lower() != "adam":
            raise NotImplementedError(f"optimizer {opath}: the B200 training kernel implements Adam")
        return FFTopology(widths, acts, l1, adam_from("Adam", okw))


class KerasLSTMBaseEstimator(KerasBaseEstimator, TransformerMixin):
    """Many-to-one LSTM autoencoder / 1-step forecast (models.py:463-698)."""

    lookahead = 0

    def __init__(self, kind: Union[Callable, str], lookback_window: int = 1, batch_size: int = 32, **kwargs) -> None:
        self.lookback_window = lookback_window
        self.batch_size = batch_size
        kwargs["lookback_window"] = lookback_window
        kwargs["batch_size"] = batch_size
        super().__init__(kind, **kwargs)

    def get_params(self, deep=False, **params):
        out = super().get_params(deep)
        out["lookback_window"] = self.lookback_window
        out["batch_size"] = self.batch_size
        return out

    def get_metadata(self):
        metadata = super().get_metadata()
        metadata.update({"forecast_steps": self.lookahead})
        return metadata

    def _validate_and_fix_size_of_X(self, X):
        if X.ndim == 1:
            logger.info(f"Reshaping X from an array to an matrix of shape {(len(X), 1)}")
            X = X.reshape(len(X), 1)
        if self.lookback_window >= X.shape[0]:
            raise ValueError("For KerasLSTMForecast lookback_window must be < size of X")
        return X

    def fit(self, X, y, **kwargs):
        torch = _torch()
        from gordo_b200.lstm import LSTMFleet
        X = np.asarray(_values(X)); y = np.asarray(_values(y))
        X = self._validate_and_fix_size_of_X(X)
        if y.ndim == 1:
            y = y.reshape(len(y), 1)
        self.kwargs.update({"n_features": X.shape[1], "n_features_out": y.shape[1]})
        epochs = int({**self.kwargs, **kwargs}.get("epochs", 1))
        dev = torch.device("cuda", torch.cuda.current_device())
        gen = torch.Generator(device=dev); gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
        if self.model is None:
            topo = self._topology()
            self.model = _Model(topo, topo.init_params(1, gen, dev)[0].cpu().numpy())
        topo = self.model.topology
        fleet = LSTMFleet(topo, 1, self.lookahead, dev)
        xd = torch.as_tensor(np.ascontiguousarray(X, np.float32), device=dev)
        yd = torch.as_tensor(np.ascontiguousarray(y, np.float32), device=dev)
        params = torch.as_tensor(self.model.params[None].copy(), device=dev)
        hl, pl = fleet.fit_jobs(xd, yd, np.array([0]), np.array([len(X)]), params, epochs=epochs,
                                batch_size=int(self.batch_size))
        self.model.params = params[0].cpu().numpy()
        # gordo keeps the History of the PRIMER fit (models.py:285-286 captures it, :615 never refreshes it)
        self._history = History({"loss": [float(pl[0])]}, {"verbose": 0, "epochs": 1, "steps": 1}, [0])
        self.history_main_ = {"loss": [float(v) for v in hl[0].cpu()]}
        return self

    def predict(self, X, **kwargs) -> np.ndarray:
        """[n - lookback_window + 1 - lookahead, n_features_out] float32 (models.py:618-660)."""
        torch = _torch()
        from gordo_b200.fleet import Schedule
        from gordo_b200.lstm import LSTMFleet
        if self.model is None:
            raise NotFittedError(f"This {self.__class__.__name__} has not been fitted yet.")
        X = self._validate_and_fix_size_of_X(np.asarray(_values(X)))
        dev = torch.device("cuda", torch.cuda.current_device())
        # the device copy of the weights and the kernel scratch are kept between calls
        key = (dev.index, id(self.model), self.lookahead, hash(self.model.params.tobytes()))
        with _SERVING_LOCK:
            cached = self.__dict__.get("_gb200_serving")
            if cached is not None and cached[0] == key:
                fleet = cached[1]
            else:
                fleet = LSTMFleet(self.model.topology, 1, self.lookahead, dev)
                fleet.set_params(torch.as_tensor(self.model.params[None], device=dev))
                self.__dict__["_gb200_serving"] = (key, fleet)
        xd = torch.as_tensor(np.ascontiguousarray(X, np.float32), device=dev)
        prec = "bf16" if (self._precision == "bf16" and fleet.tc_eligible()) else "f32"       # "bf16": tcgen05 step kernel
        out, _ = fleet.predict(Schedule.single(len(X)), xd, precision=prec)
        return out.cpu().numpy()

    def transform(self, X, **kwargs):
        return self.predict(X, **kwargs)

    def score(self, X, y, sample_weight: Optional[np.ndarray] = None, **kwargs) -> float:
        if self.model is None:
            raise NotFittedError(f"This {self.__class__.__name__} has not been fitted yet.")
        out = self.predict(X, **kwargs)
        return explained_variance_score(np.asarray(_values(y))[-len(out):], out)


class KerasLSTMForecast(KerasLSTMBaseEstimator):
    lookahead = 1


class KerasLSTMAutoEncoder(KerasLSTMBaseEstimator):
    lookahead = 0
