"""
Fleet host layer: M independent Machines of one topology, their tensors held in PyTorch on one
B200, every arithmetic step a call into libgordo_b200.so (include/gordo_b200.h).

This is the batched twin of what the reference does one Machine at a time:
``for machine in machines: ModelBuilder(machine).build()`` (gordo/builder/local_build.py:69-70)
-> ``DiffBasedAnomalyDetector.cross_validate / fit / anomaly`` (gordo/machine/model/anomaly/diff.py).
PyTorch is used for device memory, streams and RNG only.
"""
import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("gordo_b200 needs contiguous CUDA tensors")


@dataclass
class FFTopology:
    """Feed-forward stack as the factories of factories/feedforward_autoencoder.py build it."""
    widths: List[int]
    acts: List[str]
    l1: List[float] = field(default_factory=list)
    adam: Dict[str, float] = field(default_factory=lambda: dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7))

    def __post_init__(self):
        if not self.l1:
            self.l1 = [0.0] * (len(self.widths) - 1)
        self.arch = N.make_ff_arch(self.widths, self.acts, self.l1)

    @property
    def n_params(self) -> int:
        return sum(self.widths[i] * self.widths[i + 1] + self.widths[i + 1] for i in range(len(self.widths) - 1))

    @property
    def n_in(self) -> int:
        return self.widths[0]

    @property
    def n_out(self) -> int:
        return self.widths[-1]

    def key(self):
        return (tuple(self.widths), tuple(self.acts), tuple(self.l1))

    def glorot_init(self, n_machines: int, generator: torch.Generator, device) -> torch.Tensor:
        """[M, P] float32: glorot-uniform kernels, zero biases ([3P] Keras Dense defaults)."""
        out = torch.zeros((n_machines, self.n_params), dtype=torch.float32, device=device)
        o = 0
        for i in range(len(self.widths) - 1):
            a, b = self.widths[i], self.widths[i + 1]
            lim = float(np.sqrt(6.0 / (a + b)))
            w = (torch.rand((n_machines, a * b), generator=generator, device=device, dtype=torch.float32) * 2 - 1) * lim
            out[:, o:o + a * b] = w
            o += a * b + b
        return out


class Schedule:
    """Row layout of a fleet (gb200_fleet handle): Machine m owns rows [row_off[m], row_off[m+1])."""

    def __init__(self, row_counts: Optional[Sequence[int]] = None, *, rows_lo=None, rows_hi=None,
                 rows_total: Optional[int] = None):
        self._h = C.c_void_p()
        i64p = C.POINTER(C.c_int64)
        if row_counts is not None:
            rc = np.asarray(row_counts, dtype=np.int64)
            self.row_off = np.concatenate([[0], np.cumsum(rc)]).astype(np.int64)
            self.rows_lo, self.rows_hi = self.row_off[:-1].copy(), self.row_off[1:].copy()
            self.n_machines = len(rc)
            self.rows_total = int(self.row_off[-1])
            N.check(N.lib().gb200_fleet_create(C.byref(self._h), self.n_machines,
                                               self.row_off.ctypes.data_as(i64p)), "gb200_fleet_create")
        else:
            # explicit sub-ranges of a [rows_total, T] matrix (virtual Machines, e.g. CV test folds)
            self.rows_lo = np.ascontiguousarray(rows_lo, dtype=np.int64)
            self.rows_hi = np.ascontiguousarray(rows_hi, dtype=np.int64)
            self.n_machines = len(self.rows_lo)
            self.rows_total = int(rows_total if rows_total is not None else self.rows_hi.max())
            N.check(N.lib().gb200_fleet_create_ranges(C.byref(self._h), self.n_machines,
                                                      self.rows_lo.ctypes.data_as(i64p),
                                                      self.rows_hi.ctypes.data_as(i64p)), "gb200_fleet_create_ranges")

    @property
    def handle(self):
        return self._h

    _single: "Dict[tuple, Schedule]" = {}
    _single_lock = threading.Lock()

    @classmethod
    def single(cls, n_rows: int) -> "Schedule":
        """
        One Machine of ``n_rows`` rows, cached per (device, row count): a request path scores the same
        few frame lengths again and again, and creating a schedule costs device allocations.
        """
        key = (torch.cuda.current_device(), int(n_rows))
        with cls._single_lock:                  # gordo.server shares one model between gthread workers
            sch = cls._single.get(key)
            if sch is None:
                if len(cls._single) >= 64:
                    cls._single.pop(next(iter(cls._single)))
                sch = cls._single[key] = cls([int(n_rows)])
        return sch

    def __del__(self):
        try:
            if self._h:
                N.lib().gb200_fleet_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


SCORE_COLUMNS = ("model-output", "tag-anomaly-scaled", "tag-anomaly-unscaled",
                 "total-anomaly-scaled", "total-anomaly-unscaled",
                 "anomaly-confidence", "total-anomaly-confidence")


class FFFleet:
    """
    M feed-forward autoencoder Machines of one topology on one GPU.

    State (all CUDA float32): ``params`` [M,P]; input scaler ``in_scale``/``in_min`` [M,T_in]
    (the Pipeline's MinMaxScaler); error scaler ``err_scale`` [M,T_out] (the detector's
    MinMaxScaler, diff.py:173); thresholds ``feat_thr`` [M,T_out], ``agg_thr`` [M] once
    cross-validated (diff.py:256-264).
    """

    def __init__(self, topo: FFTopology, n_machines: int, device="cuda:0"):
        N.lib()
        self.topo = topo
        self.M = n_machines
        self.device = torch.device(device)
        self.params: Optional[torch.Tensor] = None
        self.in_scale = self.in_min = self.err_scale = None
        self.feat_thr = self.agg_thr = None
        self._packed = None                 # bf16 operand image (kept under this name: chunk views alias it)
        self._packed_version = -1
        self._packed_x3 = None              # f16x3 operand image
        self._packed_x3_version = -1
        self._version = 0
        self._pack_lock = threading.Lock()

    def out_rows(self, n_rows: int) -> int:
        """Rows ``predict`` returns for ``n_rows`` input rows (a Dense stack: one per row; cf. LSTMFleet.out_rows)."""
        return int(n_rows)

    # ------------------------------------------------------------------ parameters
    def set_params(self, params: torch.Tensor):
        params = params.to(self.device, torch.float32).contiguous()
        if params.shape != (self.M, self.topo.n_params):
            raise ValueError(f"params must be [{self.M}, {self.topo.n_params}], got {tuple(params.shape)}")
        self.params = params
        self._version += 1

    def init_params(self, seed: int = 0):
        g = torch.Generator(device=self.device); g.manual_seed(int(seed))
        self.set_params(self.topo.glorot_init(self.M, g, self.device))

    def tc_eligible(self, precision: str = "bf16") -> bool:
        return N.lib().gb200_ff_packed_bytes_prec(C.byref(self.topo.arch), N.PREC_CODES[precision]) > 0

    def auto_precision(self, requested: str = "bf16") -> str:
        """
        The scorer to launch for a caller that asked for ``requested``:
        "bf16"  -> the bf16 tcgen05 path when the topology is eligible and wide enough to profit -- with every
                   layer at most 8 wide (5-tag Machines) the exact fp32 kernel is both faster and exact
                   (profiles/README.md r1g);
        "f16x3" -> the fp32-grade tcgen05 path (split fp16 operands) when eligible (tanh / sigmoid hidden layers);
        anything else, or not eligible -> the exact fp32 kernel.
        """
        if requested in ("bf16", "f16x3") and self.tc_eligible(requested) and max(self.topo.widths) > 8:
            return requested
        return "f32"

    def packed(self, precision: str = "bf16") -> torch.Tensor:
        """tcgen05 operand image of the current weights for ``precision`` (re-packed when they change)."""
        attr, vattr = ("_packed", "_packed_version") if precision == "bf16" else ("_packed_x3", "_packed_x3_version")
        with self._pack_lock:
            if getattr(self, attr) is None or getattr(self, vattr) != self._version:
                code = N.PREC_CODES[precision]
                nbytes = N.lib().gb200_ff_packed_bytes_prec(C.byref(self.topo.arch), code)
                if nbytes <= 0:
                    raise ValueError(f"topology is not eligible for the {precision} tensor-core path")
                packed = torch.empty((self.M, nbytes), dtype=torch.uint8, device=self.device)
                N.check(N.lib().gb200_ff_pack(C.byref(self.topo.arch), code, self.M, N.ptr(self.params),
                                              N.ptr(packed), _stream_ptr()), "gb200_ff_pack")
                if torch.cuda.current_stream() != torch.cuda.default_stream():
                    torch.cuda.current_stream().synchronize()      # another thread / stream may read it next
                setattr(self, attr, packed); setattr(self, vattr, self._version)
            return getattr(self, attr)

    # ------------------------------------------------------------------ scalers / thresholds
    @staticmethod
    def minmax_fit(x: torch.Tensor, rows_lo: torch.Tensor, rows_hi: torch.Tensor):
        """sklearn MinMaxScaler.fit over row ranges -> (scale [J,T], min_ [J,T])."""
        _require_cuda(x, rows_lo, rows_hi)
        J, T = rows_lo.numel(), x.shape[1]
        scale = torch.empty((J, T), dtype=torch.float32, device=x.device)
        mn = torch.empty_like(scale)
        N.check(N.lib().gb200_minmax_fit(J, N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(x), T, N.ptr(scale),
                                         N.ptr(mn), _stream_ptr()), "gb200_minmax_fit")
        return scale, mn

    @staticmethod
    def cv_scores(y: torch.Tensor, yhat: torch.Tensor, rows_lo: torch.Tensor, rows_hi: torch.Tensor,
                  scoring_scale: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        """
        The builder's cross-validation metrics (build_model.py:377-446) per row range and tag, from one
        device pass (gb200_cv_sums): explained-variance-score, r2-score, mean-squared-error,
        mean-absolute-error of MinMax-scaled y / yhat.  ``scoring_scale``: [J, T] scale_ of the
        scoring scaler fitted on the full y (None = unscaled).  Returns {metric: float64 [J, T]}.
        """
        _require_cuda(y, yhat, rows_lo, rows_hi)
        J, T = rows_lo.numel(), y.shape[1]
        sums = torch.empty((J, 5, T), dtype=torch.float64, device=y.device)
        N.check(N.lib().gb200_cv_sums(J, N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(y), N.ptr(yhat), T, N.ptr(sums),
                                      _stream_ptr()), "gb200_cv_sums")
        s = sums.cpu().numpy()
        n = (rows_hi - rows_lo).double().cpu().numpy()[:, None]
        sy, syy, se, see, sae = s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 4]
        var_y = syy / n - (sy / n) ** 2
        var_e = see / n - (se / n) ** 2
        sc = np.ones((J, T)) if scoring_scale is None else np.asarray(scoring_scale, np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            # sklearn semantics for a constant target: score 1 if the numerator is 0 too, else 0
            r2 = np.where(var_y > 0, 1.0 - (see / n) / var_y, np.where(see > 0, 0.0, 1.0))
            ev = np.where(var_y > 0, 1.0 - var_e / var_y, np.where(var_e > 0, 0.0, 1.0))
        return {"explained-variance-score": ev, "r2-score": r2,
                "mean-squared-error": (see / n) * sc ** 2, "mean-absolute-error": (sae / n) * np.abs(sc)}

    @staticmethod
    def score_outputs(model_out: torch.Tensor, y: torch.Tensor, out_row_off, y_row_off, *, err_scale=None,
                      feat_thr=None, agg_thr=None):
        """
        DiffBasedAnomalyDetector columns (diff.py:350-444) for PRECOMPUTED model output whose rows are
        offset against y (LSTM: the output is shorter than the input).  Machine m's output rows
        [out_row_off[m], out_row_off[m+1]) align with rows of y starting at y_row_off[m].
        """
        _require_cuda(model_out, y, err_scale, feat_thr, agg_thr)
        dev = model_out.device
        M = len(y_row_off)
        R, T = model_out.shape
        oo = torch.as_tensor(np.asarray(out_row_off, np.int64), device=dev)
        yo = torch.as_tensor(np.asarray(y_row_off, np.int64), device=dev)
        res = {"tag-anomaly-scaled": torch.empty((R, T), dtype=torch.float32, device=dev),
               "tag-anomaly-unscaled": torch.empty((R, T), dtype=torch.float32, device=dev),
               "total-anomaly-scaled": torch.empty((R,), dtype=torch.float32, device=dev),
               "total-anomaly-unscaled": torch.empty((R,), dtype=torch.float32, device=dev)}
        if feat_thr is not None:
            res["anomaly-confidence"] = torch.empty((R, T), dtype=torch.float32, device=dev)
        if agg_thr is not None:
            res["total-anomaly-confidence"] = torch.empty((R,), dtype=torch.float32, device=dev)
        N.check(N.lib().gb200_score_outputs(
            M, N.ptr(oo), N.ptr(yo), T, N.ptr(model_out), N.ptr(y), N.ptr(err_scale), N.ptr(feat_thr), N.ptr(agg_thr),
            N.ptr(res["tag-anomaly-scaled"]), N.ptr(res["tag-anomaly-unscaled"]), N.ptr(res["total-anomaly-scaled"]),
            N.ptr(res["total-anomaly-unscaled"]), N.ptr(res.get("anomaly-confidence")),
            N.ptr(res.get("total-anomaly-confidence")), _stream_ptr()), "gb200_score_outputs")
        return res

    @staticmethod
    def rolling_min_max(v: torch.Tensor, rows_lo: torch.Tensor, rows_hi: torch.Tensor, window: int = 6):
        """pandas rolling(window).min().max() per column over row ranges -> [J, C]."""
        v2 = v if v.dim() == 2 else v.unsqueeze(1)
        _require_cuda(v2, rows_lo, rows_hi)
        J, Cc = rows_lo.numel(), v2.shape[1]
        out = torch.empty((J, Cc), dtype=torch.float32, device=v.device)
        N.check(N.lib().gb200_rolling_min_max(J, N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(v2), Cc, int(window),
                                              N.ptr(out), _stream_ptr()), "gb200_rolling_min_max")
        return out

    SMOOTH_METHODS = {"smm": 0, "sma": 1, "ewma": 2}

    @staticmethod
    def smooth(v: torch.Tensor, rows_lo: torch.Tensor, rows_hi: torch.Tensor, method: str, window: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """
        pandas rolling(window).median() / .mean() / ewm(span=window).mean() of every column over row
        ranges (diff.py:302-308).  v: [R] or [R, C] float32 CUDA -> same shape.
        """
        if method not in FFFleet.SMOOTH_METHODS:
            raise ValueError(f"smoothing_method must be one of {sorted(FFFleet.SMOOTH_METHODS)}, got {method!r}")
        v2 = v if v.dim() == 2 else v.unsqueeze(1)
        _require_cuda(v2, rows_lo, rows_hi)
        if v2.dtype != torch.float32 or not v2.is_contiguous():
            raise ValueError("smooth: v must be contiguous float32")
        res = torch.full_like(v2, float("nan")) if out is None else (out if out.dim() == 2 else out.unsqueeze(1))
        N.check(N.lib().gb200_smooth(rows_lo.numel(), N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(v2), v2.shape[1],
                                     FFFleet.SMOOTH_METHODS[method], int(window), N.ptr(res), _stream_ptr()),
                "gb200_smooth")
        return res if v.dim() == 2 else res[:, 0]

    @staticmethod
    def quantile(v: torch.Tensor, rows_lo: torch.Tensor, rows_hi: torch.Tensor, q: float) -> torch.Tensor:
        """pandas DataFrame.quantile(q) (linear, NaN skipped) per column over row ranges -> [J, C] float64."""
        v2 = v if v.dim() == 2 else v.unsqueeze(1)
        _require_cuda(v2, rows_lo, rows_hi)
        if v2.dtype != torch.float32 or not v2.is_contiguous():
            raise ValueError("quantile: v must be contiguous float32")
        J, Cc = rows_lo.numel(), v2.shape[1]
        out = torch.empty((J, Cc), dtype=torch.float64, device=v.device)
        N.check(N.lib().gb200_quantile(J, N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(v2), Cc, float(q), N.ptr(out),
                                       _stream_ptr()), "gb200_quantile")
        return out

    # ------------------------------------------------------------------ inference + scoring
    def score(self, sched: Schedule, x: torch.Tensor, y: Optional[torch.Tensor] = None, *,
              precision: str = "bf16", columns: Sequence[str] = SCORE_COLUMNS,
              out: Optional[Dict[str, torch.Tensor]] = None, subset: Optional[slice] = None):
        """
        Fused predict + DiffBasedAnomalyDetector.anomaly columns for every row of every Machine
        (models.py:289-300 + diff.py:336-444).  x: [rows_total, T_in] float32 CUDA; y defaults
        to x.  Returns {column: tensor}; threshold-based columns only when thresholds are set.
        """
        if self.params is None:
            raise RuntimeError("fleet has no parameters (call set_params / init_params / fit)")
        if sched.n_machines != self.M:
            raise ValueError("schedule and fleet disagree on the number of Machines")
        _require_cuda(x, y)
        if x.dtype != torch.float32 or x.shape != (sched.rows_total, self.topo.n_in):
            raise ValueError(f"x must be float32 [{sched.rows_total}, {self.topo.n_in}]")
        R, To = sched.rows_total, self.topo.n_out
        want = set(columns)
        if y is None and self.topo.n_in != To:
            # the targets cannot alias the samples here.  A pure forward pass (estimator.predict) does not need them:
            # the kernel gets a zero target matrix and only the model output is read back
            if not want <= {"model-output", "activity-l1"}:
                raise ValueError("y is required when n_features_out != n_features (the error columns are |model-output - y|)")
            y = torch.zeros((R, To), dtype=torch.float32, device=x.device)
        if self.feat_thr is None:
            want.discard("anomaly-confidence")
        if self.agg_thr is None:
            want.discard("total-anomaly-confidence")
        res = dict(out) if out else {}

        def buf(name, shape):
            if name not in want:
                return None
            if name not in res:
                res[name] = torch.empty(shape, dtype=torch.float32, device=x.device)
            return res[name]

        mo = buf("model-output", (R, To)); ts = buf("tag-anomaly-scaled", (R, To))
        tu = buf("tag-anomaly-unscaled", (R, To)); tts = buf("total-anomaly-scaled", (R,))
        ttu = buf("total-anomaly-unscaled", (R,)); cf = buf("anomaly-confidence", (R, To))
        tcf = buf("total-anomaly-confidence", (R,))
        act = buf("activity-l1", (R,)) if precision == "f32" else None
        if precision in ("bf16", "f16x3"):
            prec, packed = N.PREC_CODES[precision], self.packed(precision)
        elif precision == "f32":
            prec, packed = N.PREC_F32, None
        else:
            raise ValueError("precision must be 'bf16', 'f16x3' or 'f32'")
        N.check(N.lib().gb200_ff_score(
            sched.handle, C.byref(self.topo.arch), prec, N.ptr(self.params), N.ptr(packed),
            N.ptr(self.in_scale), N.ptr(self.in_min), N.ptr(self.err_scale),
            N.ptr(self.feat_thr), N.ptr(self.agg_thr), N.ptr(x), N.ptr(y),
            N.ptr(mo), N.ptr(ts), N.ptr(tu), N.ptr(tts), N.ptr(ttu), N.ptr(cf), N.ptr(tcf), N.ptr(act),
            _stream_ptr()), "gb200_ff_score")
        return res

    def predict(self, sched: Schedule, x: torch.Tensor, precision: str = "f32") -> torch.Tensor:
        return self.score(sched, x, precision=precision, columns=("model-output",))["model-output"]

    # ------------------------------------------------------------------ training
    def fit_jobs(self, x: torch.Tensor, y: Optional[torch.Tensor], rows_lo: torch.Tensor, rows_hi: torch.Tensor,
                 params: torch.Tensor, *, in_scale=None, in_min=None, scale_slot=None,
                 epochs=1, batch_size=32, perm_pool=None, perm_off=None, l1_mean=False,
                 adam_mv=None, adam_t=None, want_history=True):
        """
        Keras ``Model.fit`` for J independent jobs in one launch (models.py:243-287).  ``params``
        [J,P] is updated in place.  Returns (hist_loss [J,epochs], hist_acc [J,epochs], adam_mv, adam_t).
        """
        _require_cuda(x, y, rows_lo, rows_hi, params, in_scale, in_min, scale_slot, perm_pool, perm_off, adam_mv, adam_t)
        J, P = params.shape
        if P != self.topo.n_params:
            raise ValueError("params width does not match the topology")
        if adam_mv is None:
            adam_mv = torch.zeros((J, 2 * P), dtype=torch.float32, device=x.device)
        if adam_t is None:
            adam_t = torch.zeros((J,), dtype=torch.int64, device=x.device)
        hl = torch.empty((J, epochs), dtype=torch.float32, device=x.device) if want_history else None
        ha = torch.empty((J, epochs), dtype=torch.float32, device=x.device) if want_history else None
        adam = N.Adam(**{k: float(v) for k, v in self.topo.adam.items()})
        N.check(N.lib().gb200_ff_fit(
            C.byref(self.topo.arch), C.byref(adam), J, N.ptr(rows_lo), N.ptr(rows_hi), N.ptr(scale_slot),
            N.ptr(in_scale), N.ptr(in_min), N.ptr(x), N.ptr(y), N.ptr(perm_off), N.ptr(perm_pool),
            int(epochs), int(batch_size), int(bool(l1_mean)), N.ptr(params), N.ptr(adam_mv), N.ptr(adam_t),
            N.ptr(hl), N.ptr(ha), _stream_ptr()), "gb200_ff_fit")
        return hl, ha, adam_mv, adam_t


def time_series_split_bounds(n: int, n_splits: int = 3):
    """sklearn TimeSeriesSplit(n_splits) as (train_end == test_start, test_end) pairs."""
    n_folds = n_splits + 1
    if n_folds > n:
        raise ValueError(f"Cannot have number of folds={n_folds} greater than the number of samples={n}.")
    test_size = n // n_folds
    return [(s, s + test_size) for s in range(n - n_splits * test_size, n, test_size)]
