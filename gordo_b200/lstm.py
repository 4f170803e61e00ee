"""
LSTM host layer: topology + fleet-batched predict / fit over libgordo_b200.so.

Replaces KerasLSTMBaseEstimator.fit / .predict (gordo/machine/model/models.py:557-660) and the
window generator of models.py:713-793: windows are never materialised; window k of a Machine is
rows [k, k+L) of its scaled sample matrix, target row k + L - 1 + lookahead.
"""
import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _native as N
from .fleet import Schedule, _stream_ptr, _require_cuda


# torch initialises its linalg backend lazily and not re-entrantly: the first torch.linalg.qr must not race
# between the builder's bucket threads
_QR_LOCK = threading.Lock()


@dataclass
class LSTMTopology:
    """Stacked-LSTM autoencoder as factories/lstm_autoencoder.py:70-103 builds it."""
    n_features: int
    n_features_out: int
    units: List[int]
    acts: List[str]
    out_func: str = "linear"
    lookback_window: int = 1
    adam: Dict[str, float] = field(default_factory=lambda: dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7))

    def arch(self, lookahead: int):
        return N.make_lstm_arch(self.n_features, self.n_features_out, self.units, self.acts, self.out_func,
                                self.lookback_window, lookahead)

    @property
    def n_params(self) -> int:
        n, n_in = 0, self.n_features
        for u in self.units:
            n += n_in * 4 * u + u * 4 * u + 4 * u
            n_in = u
        return n + n_in * self.n_features_out + self.n_features_out

    def key(self):
        return (self.n_features, self.n_features_out, tuple(self.units), tuple(self.acts), self.out_func,
                self.lookback_window)

    def init_params(self, n_machines: int, generator: torch.Generator, device) -> torch.Tensor:
        """
        [M, P] float32, [3P] Keras LSTM defaults: glorot-uniform W [in,4u] (fan_out = 4u), orthogonal U
        [u,4u], zero bias with the forget slice [u:2u] = 1 (unit_forget_bias); glorot Dense, zero bias.
        """
        out = torch.zeros((n_machines, self.n_params), dtype=torch.float32, device=device)
        o, n_in = 0, self.n_features
        for u in self.units:
            lim = float(np.sqrt(6.0 / (n_in + 4 * u)))
            out[:, o:o + n_in * 4 * u] = (torch.rand((n_machines, n_in * 4 * u), generator=generator, device=device) * 2 - 1) * lim
            o += n_in * 4 * u
            a = torch.randn((n_machines, 4 * u, u), generator=generator, device=device)
            with _QR_LOCK:
                q, r = torch.linalg.qr(a)                               # [M, 4u, u], orthonormal columns
            q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
            out[:, o:o + u * 4 * u] = q.transpose(-1, -2).reshape(n_machines, -1)      # U [u, 4u]
            o += u * 4 * u
            out[:, o + u:o + 2 * u] = 1.0
            o += 4 * u
            n_in = u
        lim = float(np.sqrt(6.0 / (n_in + self.n_features_out)))
        out[:, o:o + n_in * self.n_features_out] = (torch.rand((n_machines, n_in * self.n_features_out),
                                                               generator=generator, device=device) * 2 - 1) * lim
        return out


class LSTMFleet:
    """M LSTM autoencoder / forecast Machines of one topology on one GPU."""

    def __init__(self, topo: LSTMTopology, n_machines: int, lookahead: int = 0, device="cuda:0"):
        N.lib()
        self.topo, self.M, self.lookahead = topo, n_machines, int(lookahead)
        self.device = torch.device(device)
        self.arch = topo.arch(self.lookahead)
        self.params: Optional[torch.Tensor] = None
        self.in_scale = self.in_min = None
        self._scratch: Optional[torch.Tensor] = None

    def out_rows(self, n_rows: int) -> int:
        return max(int(n_rows) - self.topo.lookback_window + 1 - self.lookahead, 0)

    def set_params(self, params: torch.Tensor):
        params = params.to(self.device, torch.float32).contiguous()
        if params.shape != (self.M, self.topo.n_params):
            raise ValueError(f"params must be [{self.M}, {self.topo.n_params}], got {tuple(params.shape)}")
        self.params = params

    def _get_scratch(self, nbytes: int) -> torch.Tensor:
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=self.device)
        return self._scratch

    def tc_eligible(self) -> bool:
        return N.lib().gb200_lstm_scratch_bytes(C.byref(self.arch), 128, N.PREC_BF16_TC) > 0

    def predict(self, sched: Schedule, x: torch.Tensor, max_windows: int = 16384, precision: str = "f32"):
        """
        KerasLSTMBaseEstimator.predict for every Machine.  Returns (model_out [sum out_rows, T_out],
        out_row_off np.int64[M+1]).  precision "f32" (exact) or "bf16" (tcgen05 step kernel).
        """
        if self.params is None:
            raise RuntimeError("fleet has no parameters")
        _require_cuda(x)
        rows = sched.rows_hi - sched.rows_lo
        if (rows <= self.topo.lookback_window).any():
            raise ValueError("For KerasLSTMForecast lookback_window must be < size of X")
        out_off = np.concatenate([[0], np.cumsum([self.out_rows(r) for r in rows])]).astype(np.int64)
        out = torch.empty((int(out_off[-1]), self.topo.n_features_out), dtype=torch.float32, device=x.device)
        prec = {"f32": N.PREC_F32, "bf16": N.PREC_BF16_TC}[precision]
        nbytes = N.lib().gb200_lstm_scratch_bytes(C.byref(self.arch), int(max_windows), prec)
        if nbytes <= 0:
            raise ValueError("LSTM topology is not eligible for this precision")
        scratch = self._get_scratch(nbytes)
        d_off = torch.from_numpy(out_off).to(x.device)
        N.check(N.lib().gb200_lstm_predict(sched.handle, C.byref(self.arch), prec, N.ptr(self.params), N.ptr(self.in_scale),
                                           N.ptr(self.in_min), N.ptr(x), N.ptr(d_off), N.ptr(out), N.ptr(scratch),
                                           int(nbytes), _stream_ptr()), "gb200_lstm_predict")
        return out, out_off

    def fit_jobs(self, x, y, rows_lo: np.ndarray, rows_hi: np.ndarray, params: torch.Tensor, *, in_scale=None,
                 in_min=None, epochs=1, batch_size=32):
        """
        KerasLSTMBaseEstimator.fit for J jobs: primer step on the first window, then time-ordered
        batches (models.py:585-615).  ``params`` [J,P] updated in place.
        Returns (hist_loss [J,epochs], primer_loss [J]).
        """
        _require_cuda(x, y, params, in_scale, in_min)
        J = params.shape[0]
        lo = np.ascontiguousarray(rows_lo, np.int64); hi = np.ascontiguousarray(rows_hi, np.int64)
        if ((hi - lo) <= self.topo.lookback_window).any():
            raise ValueError("For KerasLSTMForecast lookback_window must be < size of X")
        nbytes = N.lib().gb200_lstm_fit_scratch_bytes(C.byref(self.arch), J, int(batch_size))
        scratch = self._get_scratch(nbytes)
        hl = torch.empty((J, epochs), dtype=torch.float32, device=x.device)
        pl = torch.empty((J,), dtype=torch.float32, device=x.device)
        adam = N.Adam(**{k: float(v) for k, v in self.topo.adam.items()})
        i64p = C.POINTER(C.c_int64)
        N.check(N.lib().gb200_lstm_fit(C.byref(self.arch), C.byref(adam), J, lo.ctypes.data_as(i64p),
                                       hi.ctypes.data_as(i64p), N.ptr(in_scale), N.ptr(in_min), N.ptr(x), N.ptr(y),
                                       int(epochs), int(batch_size), N.ptr(params), N.ptr(hl), N.ptr(pl),
                                       N.ptr(scratch), int(nbytes), _stream_ptr()), "gb200_lstm_fit")
        return hl, pl
