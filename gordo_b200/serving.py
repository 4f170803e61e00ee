"""
Fleet twin of the reference's anomaly request path.

gordo.server answers ``POST .../anomaly/prediction`` one Machine at a time:
``anomaly_df = g.model.anomaly(g.X, g.y, frequency=...)`` (gordo/server/blueprints/anomaly.py:50) with host
frames in and a host frame out.  ``FleetAnomalyServer.anomaly(X)`` is that call for M Machines of
one topology at once: HOST sample matrices in (pinned, or any array / list of frames), HOST anomaly
columns out -- every column of diff.py:310-458 materialised in pinned host memory, sliceable per
Machine and convertible to the reference's frame (``result.frame(m, index, frequency)``).

The call is bound by PCIe and host DRAM, not by the GPU (the fused scorer runs at ~4.6e9 rows/s, a
Gen5 x16 link moves 1e5 x 50 float32 rows in 0.4 ms).  So the work is split by what is cheapest to move:

  device : MinMax -> Dense stack -> |yhat - y| -> row totals (gb200_ff_score); sends back
           ``model-output`` and the three total columns, plus the rescaled matrices the plan keeps on the device;
  host   : the per-column rescalings the plan moves off the wire
           (``anomaly-confidence``, ``tag-anomaly-scaled``, ``tag-anomaly-unscaled`` -- in that order) are
           written by gb200_host_expand_columns straight into the response buffers while the next
           chunk is in flight.

``plan`` = how many of those three matrices the host derives (0..3; x.5 = alternate chunks use x and x+1, which
splits the load between the PCIe link and the host threads more finely).  One GPU alone is PCIe-bound and wants
about 3; eight ranks sharing two sockets are host-DRAM-bound and want 0 -- ``plan="auto"`` times the options on
the first call (all ranks do so together) and keeps the fastest.

Chunks of Machines flow through three streams (H2D, kernel, D2H) and three device buffer sets, so
both PCIe directions and the kernel overlap; a host worker thread expands chunk c while c+1 is on the wire.
"""
import ctypes as C
import queue
import threading
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .fleet import FFFleet, Schedule

MATRIX_COLUMNS = ("model-output", "tag-anomaly-scaled", "tag-anomaly-unscaled", "anomaly-confidence")
VECTOR_COLUMNS = ("total-anomaly-scaled", "total-anomaly-unscaled", "total-anomaly-confidence")
# order in which matrices move from the wire to the host threads as the plan grows
DERIVE_ORDER = ("anomaly-confidence", "tag-anomaly-scaled", "tag-anomaly-unscaled")


class FleetAnomalyResult:
    """Host columns of one ``FleetAnomalyServer.anomaly`` call (views of the server's pinned buffers)."""

    def __init__(self, columns: Dict[str, torch.Tensor], row_off: np.ndarray, x_host: torch.Tensor, tags=None):
        self.columns = columns
        self.row_off = row_off
        self.model_input = x_host
        self.tags = tags

    def machine(self, m: int) -> Dict[str, np.ndarray]:
        a, b = int(self.row_off[m]), int(self.row_off[m + 1])
        out = {"model-input": self.model_input[a:b].numpy()}
        out.update({k: v[a:b].numpy() for k, v in self.columns.items()})
        return out

    def groups(self, m: int, tags: Optional[Sequence] = None):
        """Machine m's columns as ``assemble_frame`` groups, in the reference's column order (diff.py:310-458)."""
        from gordo_b200.machine.model import utils as model_utils
        c = self.machine(m)
        T = c["model-input"].shape[1]
        tags = list(tags if tags is not None else (self.tags[m] if self.tags is not None else range(T)))
        names_in = model_utils._second_level(c["model-input"], tags)
        names_out = model_utils._second_level(c["model-output"], tags)
        tag_names = [model_utils._tag_name(t) for t in tags]
        groups = [("model-input", c["model-input"], names_in), ("model-output", c["model-output"], names_out),
                  ("tag-anomaly-scaled", c["tag-anomaly-scaled"], names_out),
                  ("total-anomaly-scaled", c["total-anomaly-scaled"], None),
                  ("tag-anomaly-unscaled", c["tag-anomaly-unscaled"], tag_names),
                  ("total-anomaly-unscaled", c["total-anomaly-unscaled"], None)]
        if "anomaly-confidence" in c:
            groups.append(("anomaly-confidence", c["anomaly-confidence"], names_out))
        if "total-anomaly-confidence" in c:
            groups.append(("total-anomaly-confidence", c["total-anomaly-confidence"], None))
        return groups

    def frame(self, m: int, index=None, frequency=None, tags: Optional[Sequence] = None):
        """The reference's anomaly frame for Machine m (model/utils.py:49-165 layout, diff.py:310-458 columns)."""
        from gordo_b200.machine.model import utils as model_utils
        return model_utils.assemble_frame(self.groups(m, tags), index, frequency)

    def parquet_bytes(self, m: int, index=None, frequency=None, tags: Optional[Sequence] = None, compression="snappy") -> bytes:
        """The ``?format=parquet`` response body of server/blueprints/anomaly.py:64-68 for Machine m, straight from the columns."""
        from gordo_b200.server.utils import columns_into_parquet_bytes
        return columns_into_parquet_bytes(self.groups(m, tags), index, frequency, compression)

    def to_dict(self, m: int, index=None, frequency=None, tags: Optional[Sequence] = None) -> dict:
        """The ``data`` member of the JSON response (server/blueprints/anomaly.py:70-72, server/utils.py:88-142)."""
        from gordo_b200.server.utils import columns_to_dict
        return columns_to_dict(self.groups(m, tags), index, frequency)


class FleetAnomalyServer:
    def __init__(self, fleet: FFFleet, row_counts: Sequence[int], *, precision: str = "bf16", n_chunks: int = 16,
                 plan="auto", n_threads: Optional[int] = None, tags=None):
        if fleet.params is None or fleet.err_scale is None:
            raise ValueError("FleetAnomalyServer needs a fitted fleet (params and error scaler)")
        rc = np.asarray(row_counts, np.int64)
        if len(rc) != fleet.M:
            raise ValueError("row_counts and fleet disagree on the number of Machines")
        self.fleet, self.precision, self.tags = fleet, fleet.auto_precision(precision), tags
        self.row_off = np.concatenate([[0], np.cumsum(rc)]).astype(np.int64)
        self.rows = int(self.row_off[-1])
        self.M = fleet.M
        dev = fleet.device
        T, To = fleet.topo.n_in, fleet.topo.n_out
        self.T, self.To = T, To
        self.has_feat, self.has_agg = fleet.feat_thr is not None, fleet.agg_thr is not None
        self.matrices = [c for c in MATRIX_COLUMNS if c != "anomaly-confidence" or self.has_feat]
        self.vectors = [c for c in VECTOR_COLUMNS if c != "total-anomaly-confidence" or self.has_agg]
        self.derivable = [c for c in DERIVE_ORDER if c in self.matrices]
        from .hostbind import effective_cpus
        n_eff = effective_cpus()                 # affinity mask capped by the container's CPU quota
        self.n_threads = int(n_threads) if n_threads else max(1, min(32, n_eff))
        # chunks of whole Machines
        n_chunks = max(1, min(int(n_chunks), self.M))
        b = np.linspace(0, self.M, n_chunks + 1).astype(int)
        self.chunks = [(int(a), int(c)) for a, c in zip(b[:-1], b[1:]) if c > a]
        max_rows = max(int(self.row_off[c] - self.row_off[a]) for a, c in self.chunks)
        self.n_slots = min(3, len(self.chunks))
        self.s_h2d, self.s_k, self.s_d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.dx = [torch.empty((max_rows, T), dtype=torch.float32, device=dev) for _ in range(self.n_slots)]
        self.dy = None                  # device targets, allocated on the first call that passes y != X
        self._max_rows = max_rows
        self.dout = [{**{c: torch.empty((max_rows, To), dtype=torch.float32, device=dev) for c in self.matrices},
                      **{c: torch.empty((max_rows,), dtype=torch.float32, device=dev) for c in self.vectors}}
                     for _ in range(self.n_slots)]
        self.host_out: Dict[str, torch.Tensor] = {
            **{c: torch.empty((self.rows, To), dtype=torch.float32, pin_memory=True) for c in self.matrices},
            **{c: torch.empty((self.rows,), dtype=torch.float32, pin_memory=True) for c in self.vectors}}
        # per-chunk fleet views + schedules, host copies of the per-column factors
        self.views = []
        for a, c in self.chunks:
            v = FFFleet(fleet.topo, c - a, dev)
            v.params = fleet.params[a:c]
            if self.precision == "bf16":
                v._packed = fleet.packed("bf16")[a:c]; v._packed_version = v._version
            elif self.precision == "f16x3":
                v._packed_x3 = fleet.packed("f16x3")[a:c]; v._packed_x3_version = v._version
            for name in ("in_scale", "in_min", "err_scale", "feat_thr", "agg_thr"):
                t = getattr(fleet, name)
                setattr(v, name, None if t is None else t[a:c])
            self.views.append((v, Schedule(np.diff(self.row_off[a:c + 1]))))
        self.h_err_scale = np.ascontiguousarray(fleet.err_scale.cpu().numpy(), np.float32)
        self.h_feat_thr = None if not self.has_feat else np.ascontiguousarray(fleet.feat_thr.cpu().numpy(), np.float32)
        self._lock = threading.Lock()             # one request at a time per server (buffers are reused)
        self._jobs: "queue.Queue" = queue.Queue()
        self._worker = threading.Thread(target=self._host_worker, daemon=True)
        self._worker.start()
        self._stage: Optional[torch.Tensor] = None
        self._stage_y: Optional[torch.Tensor] = None
        self.plan_timings: Dict[int, float] = {}
        if plan == "auto":
            self.plan: Optional[float] = None
        else:
            self.plan = self._clamp_plan(float(plan))

    # ------------------------------------------------------------------ construction from fitted models
    @classmethod
    def from_models(cls, models: Sequence, row_counts: Sequence[int], device=None, **kw) -> "FleetAnomalyServer":
        """
        From fitted ``DiffBasedAnomalyDetector[Pipeline[MinMaxScaler, KerasAutoEncoder]]`` objects of ONE
        topology (what FleetModelBuilder / serializer.load hand back).  Use ``group_by_topology`` first for a
        heterogeneous project.
        """
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        plans = [m._fused_plan() for m in models]
        if any(p is None for p in plans):
            raise ValueError("every model must be the standard fused composition (see DiffBasedAnomalyDetector._fused_plan)")
        from gordo_b200.machine.model.models import KerasLSTMBaseEstimator
        if any(isinstance(p[1], KerasLSTMBaseEstimator) for p in plans):
            raise ValueError("FleetAnomalyServer batches feed-forward Machines; LSTM Machines are served per Machine "
                             "(model.anomaly) or through LSTMFleet.predict + FFFleet.score_outputs")
        topo = plans[0][1].model.topology
        if any(p[1].model.topology.key() != topo.key() for p in plans):
            raise ValueError("models of one FleetAnomalyServer must share a topology")
        M = len(models)
        fleet = FFFleet(topo, M, dev)
        fleet.set_params(torch.as_tensor(np.stack([p[1].model.params for p in plans]), device=dev))
        f32 = lambda rows: torch.as_tensor(np.stack([np.asarray(r, np.float32) for r in rows]), device=dev)
        if all(p[0] is not None for p in plans):
            fleet.in_scale, fleet.in_min = f32([p[0].scale_ for p in plans]), f32([p[0].min_ for p in plans])
        elif any(p[0] is not None for p in plans):
            raise ValueError("either every model has an input MinMaxScaler or none")
        fleet.err_scale = f32([m.scaler.scale_ for m in models])
        feats = [m.__dict__.get("feature_thresholds_") for m in models]
        aggs = [m.__dict__.get("aggregate_threshold_") for m in models]
        if all(f is not None for f in feats):
            fleet.feat_thr = f32([np.asarray(f, np.float64) for f in feats])
        if all(a is not None for a in aggs):
            fleet.agg_thr = torch.as_tensor(np.asarray(aggs, np.float32), device=dev)
        prec = kw.pop("precision", plans[0][1]._precision)
        return cls(fleet, row_counts, precision=prec, **kw)

    @staticmethod
    def group_by_topology(models: Sequence) -> Dict[tuple, List[int]]:
        groups: Dict[tuple, List[int]] = {}
        for i, m in enumerate(models):
            p = m._fused_plan()
            key = None if p is None else (p[1].model.topology.key(), p[0] is None,
                                          m.__dict__.get("feature_thresholds_") is None)
            groups.setdefault(key, []).append(i)
        return groups

    # ------------------------------------------------------------------ host side
    def _host_worker(self):
        lib = N.lib()
        while True:
            job = self._jobs.get()
            if job is None:
                return
            ev, ci, derive, x_host, y_host, done = job
            try:
                ev.synchronize()
                if derive:
                    a, c = self.chunks[ci]
                    r0, r1 = int(self.row_off[a]), int(self.row_off[c])
                    ro = np.ascontiguousarray(self.row_off[a:c + 1])
                    hp = lambda name: C.c_void_p(self.host_out[name][r0:r1].data_ptr()) if name in derive else None
                    rc = lib.gb200_host_expand_columns(
                        c - a, ro.ctypes.data_as(C.c_void_p), self.To,
                        C.c_void_p(self.host_out["model-output"][r0:r1].data_ptr()), C.c_void_p(y_host[r0:r1].data_ptr()),
                        self.h_err_scale[a:c].ctypes.data_as(C.c_void_p),
                        None if self.h_feat_thr is None else self.h_feat_thr[a:c].ctypes.data_as(C.c_void_p),
                        hp("tag-anomaly-unscaled"), hp("tag-anomaly-scaled"), hp("anomaly-confidence"), self.n_threads)
                    if rc != 0:
                        raise RuntimeError("gb200_host_expand_columns: " + lib.gb200_last_error().decode())
                done.append(None)
            except Exception as e:       # surfaced by anomaly()
                done.append(e)
            finally:
                self._jobs.task_done()

    def _as_pinned(self, X, width: Optional[int] = None, which: str = "_stage") -> torch.Tensor:
        """The request's samples (or targets) as ONE pinned float32 [rows, width] matrix (no copy when it already is one)."""
        W = self.T if width is None else width
        if isinstance(X, torch.Tensor):
            if X.dtype == torch.float32 and X.is_pinned() and X.is_contiguous() and tuple(X.shape) == (self.rows, W):
                return X
            X = X.numpy()
        if getattr(self, which) is None:
            setattr(self, which, torch.empty((self.rows, W), dtype=torch.float32, pin_memory=True))
        st = getattr(self, which).numpy()
        if isinstance(X, (list, tuple)):
            if len(X) != self.M:
                raise ValueError(f"expected {self.M} Machines, got {len(X)}")
            for m, xm in enumerate(X):
                a, b = int(self.row_off[m]), int(self.row_off[m + 1])
                xv = np.asarray(getattr(xm, "values", xm))
                if xv.shape != (b - a, W):
                    raise ValueError(f"Machine {m}: expected {(b - a, W)}, got {xv.shape}")
                st[a:b] = xv
        else:
            xv = np.asarray(getattr(X, "values", X))
            if xv.shape != (self.rows, W):
                raise ValueError(f"expected {(self.rows, W)}, got {xv.shape}")
            st[:] = xv
        return getattr(self, which)

    # ------------------------------------------------------------------ the call
    def anomaly(self, X, y=None) -> FleetAnomalyResult:
        """
        X: pinned float32 [rows_total, T] (zero-copy), or an array / list of per-Machine arrays or frames
        (staged into pinned memory first).  y defaults to X (autoencoder).  Blocks until every host column
        is complete.
        """
        with self._lock:
            x_host = self._as_pinned(X)
            if y is None or y is X:
                if self.T != self.To:
                    raise ValueError("y may default to X only when n_features == n_features_out")
                y_host = x_host
            else:                                   # separate targets (model.anomaly(X, y): diff.py:336-344)
                y_host = self._as_pinned(y, self.To, "_stage_y")
                if self.dy is None:
                    self.dy = [torch.empty((self._max_rows, self.To), dtype=torch.float32, device=self.fleet.device)
                               for _ in range(self.n_slots)]
            if self.plan is None:
                self._calibrate(x_host, y_host)
            self._run(x_host, y_host, self.plan)
            return FleetAnomalyResult(self.host_out, self.row_off, x_host, self.tags)

    def _clamp_plan(self, p: float) -> float:
        p = max(0.0, min(float(p), float(len(self.derivable))))
        return round(p * 2) / 2.0                     # whole or half steps

    def _chunk_k(self, plan: float, ci: int) -> int:
        """Matrices chunk ``ci`` derives on the host under ``plan`` (x.5: odd chunks take one more)."""
        k = int(plan)
        if plan - k >= 0.5 and (ci & 1):
            k += 1
        return min(k, len(self.derivable))

    def _time_plan(self, x_host, y_host, p: float) -> float:
        self._run(x_host, y_host, p)                # warm
        t0 = time.perf_counter()
        self._run(x_host, y_host, p)
        dt = time.perf_counter() - t0
        self.plan_timings[p] = dt
        return dt

    def _calibrate(self, x_host, y_host):
        # whole plans first (most host derivation first: ties go to less PCIe), then the half steps next to the best
        best, best_t = 0.0, None
        for k in range(len(self.derivable), -1, -1):
            dt = self._time_plan(x_host, y_host, float(k))
            if best_t is None or dt < best_t * 0.97:
                best, best_t = float(k), dt
        if len(self.chunks) >= 4:
            for p in (best + 0.5, best - 0.5):
                if 0.0 <= p <= len(self.derivable):
                    dt = self._time_plan(x_host, y_host, p)
                    if dt < best_t * 0.97:
                        best, best_t = p, dt
        self.plan = best

    def _run(self, x_host, y_host, plan: float):
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event(); start.record(cur)
        for s in (self.s_h2d, self.s_k, self.s_d2h):
            s.wait_event(start)
        ev_k: List[Optional[torch.cuda.Event]] = [None] * len(self.chunks)
        ev_d: List[Optional[torch.cuda.Event]] = [None] * len(self.chunks)
        done: List = []
        for ci, ((a, c), (view, vs)) in enumerate(zip(self.chunks, self.views)):
            slot = ci % self.n_slots
            r0, r1 = int(self.row_off[a]), int(self.row_off[c])
            n = r1 - r0
            derive = tuple(self.derivable[:self._chunk_k(plan, ci)])
            dev_cols = tuple(cn for cn in self.matrices if cn not in derive) + tuple(self.vectors)
            dx = self.dx[slot][:n]
            with torch.cuda.stream(self.s_h2d):
                if ci >= self.n_slots:
                    self.s_h2d.wait_event(ev_k[ci - self.n_slots])        # the slot's previous kernel has read dx
                dx.copy_(x_host[r0:r1], non_blocking=True)
                dy = None
                if y_host is not x_host:
                    dy = self.dy[slot][:n]
                    dy.copy_(y_host[r0:r1], non_blocking=True)
                e_h = torch.cuda.Event(); e_h.record(self.s_h2d)
            out = {cname: self.dout[slot][cname][:n] for cname in dev_cols}
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(e_h)
                if ci >= self.n_slots:
                    self.s_k.wait_event(ev_d[ci - self.n_slots])          # the slot's previous results left the device
                view.score(vs, dx, dy, precision=self.precision, columns=dev_cols, out=out)
                ev_k[ci] = torch.cuda.Event(); ev_k[ci].record(self.s_k)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_k[ci])
                for cname in dev_cols:
                    self.host_out[cname][r0:r1].copy_(out[cname], non_blocking=True)
                ev_d[ci] = torch.cuda.Event(); ev_d[ci].record(self.s_d2h)
            self._jobs.put((ev_d[ci], ci, derive, x_host, y_host, done))
        self._jobs.join()                       # every chunk copied out and expanded
        for s in (self.s_h2d, self.s_k, self.s_d2h):
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
        errs = [e for e in done if e is not None]
        if errs:
            raise errs[0]

    # ------------------------------------------------------------------ accounting
    def bytes_per_call(self, plan: Optional[float] = None) -> Dict[str, int]:
        plan = self.plan if plan is None else plan
        plan = float(len(self.derivable)) if plan is None else float(plan)
        d2h = derived = 0
        for ci, (a, c) in enumerate(self.chunks):
            rows = int(self.row_off[c] - self.row_off[a])
            k = self._chunk_k(plan, ci)
            d2h += rows * ((len(self.matrices) - k) * self.To + len(self.vectors)) * 4
            derived += rows * k * self.To * 4
        h2d = self.rows * self.T * 4 + (self.rows * self.To * 4 if self.dy is not None else 0)
        return {"h2d": h2d, "d2h": d2h, "host_derived_bytes": derived}

    def kernel_launches_per_call(self) -> int:
        return len(self.chunks)

    def close(self):
        """Stop the host worker thread and wait for it (a thread still inside the library at interpreter shutdown aborts the process)."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        self._jobs.put(None)
        if self._worker.is_alive() and threading.current_thread() is not self._worker:
            self._worker.join(timeout=10)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
