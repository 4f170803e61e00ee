"""
Host-buffer scoring pipeline: pinned HOST samples in, pinned HOST anomaly columns out.

The reference-facing call path (server: gordo/server/blueprints/anomaly.py:50 -> model.anomaly)
hands over host arrays and expects host columns back; for a fleet that is PCIe-bound, so the
Machines are processed in chunks on two CUDA streams: chunk c+1's H2D copy and chunk c-1's D2H
copies overlap chunk c's fused kernel (PCIe is full duplex).
"""
from typing import Dict

import numpy as np
import torch

from .fleet import FFFleet, Schedule, SCORE_COLUMNS


class HostPipeline:
    def __init__(self, fleet: FFFleet, sched: Schedule, n_chunks: int = 8, precision: str = "bf16",
                 columns=SCORE_COLUMNS, machines: int = None):
        if getattr(sched, "row_off", None) is None:
            raise ValueError("HostPipeline needs a contiguous Schedule (row counts)")
        self.fleet, self.sched, self.precision = fleet, sched, precision
        self.columns = [c for c in columns
                        if not (c == "anomaly-confidence" and fleet.feat_thr is None)
                        and not (c == "total-anomaly-confidence" and fleet.agg_thr is None)]
        M = fleet.M if machines is None else max(1, min(int(machines), fleet.M))   # first M Machines only
        self.machines = M
        n_chunks = max(1, min(n_chunks, M))
        bounds = np.linspace(0, M, n_chunks + 1).astype(int)
        self.chunks = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        dev = fleet.device
        T, To = fleet.topo.n_in, fleet.topo.n_out
        max_rows = max(int(sched.row_off[b] - sched.row_off[a]) for a, b in self.chunks)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        self.dx = [torch.empty((max_rows, T), dtype=torch.float32, device=dev) for _ in range(2)]
        self.dout = [{c: torch.empty((max_rows, To) if c in ("model-output", "tag-anomaly-scaled",
                                                                "tag-anomaly-unscaled", "anomaly-confidence")
                                     else (max_rows,), dtype=torch.float32, device=dev)
                      for c in self.columns} for _ in range(2)]
        R = int(sched.row_off[M])
        self.rows = R
        self.host_out: Dict[str, torch.Tensor] = {
            c: torch.empty((R, To) if self.dout[0][c].dim() == 2 else (R,), dtype=torch.float32, pin_memory=True)
            for c in self.columns}
        self.views = []
        for a, b in self.chunks:
            v = FFFleet(fleet.topo, b - a, dev)
            v.params = fleet.params[a:b]
            if precision == "bf16":
                v._packed = fleet.packed()[a:b]; v._packed_version = v._version
            for name in ("in_scale", "in_min", "err_scale", "feat_thr", "agg_thr"):
                t = getattr(fleet, name)
                setattr(v, name, None if t is None else t[a:b])
            rc = np.diff(sched.row_off[a:b + 1])
            self.views.append((v, Schedule(rc)))
        self.h2d_bytes = R * T * 4
        self.d2h_bytes = sum(int(t.numel()) * 4 for t in self.host_out.values())

    def run(self, x_host: torch.Tensor) -> Dict[str, torch.Tensor]:
        """x_host: pinned [rows_total, T] float32.  Returns the pinned host columns (reused across calls)."""
        if not x_host.is_pinned():
            raise ValueError("x_host must be pinned host memory")
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event(); start.record(cur)
        for s in self.streams:
            s.wait_event(start)
        for i, ((a, b), (view, vs)) in enumerate(zip(self.chunks, self.views)):
            k = i & 1
            r0, r1 = int(self.sched.row_off[a]), int(self.sched.row_off[b])
            n = r1 - r0
            with torch.cuda.stream(self.streams[k]):
                dx = self.dx[k][:n]
                dx.copy_(x_host[r0:r1], non_blocking=True)
                out = {c: t[:n] for c, t in self.dout[k].items()}
                view.score(vs, dx, precision=self.precision, columns=self.columns, out=out)
                for c in self.columns:
                    self.host_out[c][r0:r1].copy_(out[c], non_blocking=True)
        for s in self.streams:
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
        return self.host_out
