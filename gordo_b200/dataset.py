"""
Upstream of X on the GPU: the arithmetic of ``dataset.get_data()`` between the data provider and the
matrix the builder trains on, for one Machine or a whole fleet in a handful of launches.

Reference call site: gordo/builder/build_model.py:208-213 (``GordoBaseDataset.from_dict(...).get_data()``).  The
implementation the reference calls is [3P] gordo-core 0.3.6 (requirements/full_requirements.txt:141, not vendored):
``gordo_core/time_series.py`` ``TimeSeriesDataset.join_timeseries`` / ``get_data`` and ``gordo_core/filters/rows.py``
``pandas_filter_rows`` / ``apply_buffer``.  The functions below keep those names, arguments and error behaviour:

    join_timeseries(series_iterable, resampling_startpoint, resampling_endpoint, resolution, ...) -> DataFrame
    pandas_filter_rows(df, filter_str, buffer_size=0)                                              -> DataFrame
    get_data(series, train_start_date, train_end_date, resolution="10T", row_filter="", ...)       -> DataFrame
    FleetTimeSeries(device).join(machines, ...) / .get_data(machines, ...)                         -> [JoinedMachine]

Kernels (gordo_b200/csrc/dataset.cu through the C-ABI, include/gordo_b200.h): gb200_resample (bin means / min / max /
sum / count / first / last of every raw series), gb200_interpolate (pandas ``interpolate(limit)`` / ``ffill(limit)``),
gb200_filter_rows (row predicates: dropna, thresholds, ``row_filter`` / ``known_filter_periods`` compiled from the
pandas-eval strings by :func:`compile_row_filter`, ``apply_buffer``) and gb200_compact_rows.  The host side only parses
strings, lays out descriptors and wraps results in DataFrames; without the CUDA library every call raises.

Not covered (raise NotImplementedError): ``filter_periods`` (rolling-median / isolation-forest period filters),
aggregation methods other than mean / min / max / sum / count / first / last, series of one Machine whose resampling
grids are out of phase (a resolution that does not divide the day with series starting on different days).
"""
import ast
import re
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import pandas as pd

from . import _native as N

AGG_CODES = {"mean": 0, "min": 1, "max": 2, "sum": 3, "count": 4, "first": 5, "last": 6}
INTERP_CODES = {"linear_interpolation": 1, "ffill": 2}
OPS = dict(CONST=0, COL=1, INDEX=2, NEG=3, ABS=4, NOT=5, ADD=6, SUB=7, MUL=8, DIV=9, POW=10, GT=11, GE=12, LT=13,
           LE=14, EQ=15, NE=16, AND=17, OR=18, ALL_FINITE=19, ALL_NOTNAN=20, ALL_BETWEEN=21)
MAX_OPS, MAX_CONSTS, MAX_STACK = 96, 48, 16
_ALIAS = {"T": "min", "H": "h", "S": "s", "L": "ms", "U": "us", "N": "ns"}
_UNIT_NS = {"ns": 1, "us": 1_000, "ms": 1_000_000, "s": 1_000_000_000}


class InsufficientDataError(ValueError):
    """gordo_core.exceptions.InsufficientDataError: fewer rows than ``n_samples_threshold`` survive."""


# --------------------------------------------------------------------------------------------- host-side parsing
def normalize_freq(freq: str) -> str:
    """pandas >= 2.2 spells the offsets gordo configs use ("10T", "8H", "30S") "10min", "8h", "30s"."""
    m = re.fullmatch(r"\s*(\d*\.?\d*)\s*([A-Za-z]+)\s*", str(freq))
    if not m:
        return freq
    n, unit = m.groups()
    return f"{n}{_ALIAS.get(unit, unit)}"


def _step_ns(resolution: str) -> int:
    step = int(pd.Timedelta(normalize_freq(resolution)).value)
    if step <= 0:
        raise ValueError(f"resolution {resolution!r} is not a positive fixed frequency")
    return step


def interpolation_limit_bins(interpolation_limit: Optional[str], resolution: str) -> Optional[int]:
    """``int(Timedelta(interpolation_limit) / Timedelta(resolution))`` bins; None = unlimited (join_timeseries)."""
    if not interpolation_limit:
        return None
    limit = int(pd.Timedelta(normalize_freq(interpolation_limit)).total_seconds()
                / pd.Timedelta(normalize_freq(resolution)).total_seconds())
    if limit <= 0:
        raise ValueError("Interpolation limit must be larger than given resolution")
    return limit


@dataclass(frozen=True)
class RowProgram:
    """Postfix program of gb200_filter_rows (include/gordo_b200.h GB200_OP_*)."""
    ops: Tuple[int, ...]
    args: Tuple[int, ...]
    consts: Tuple[float, ...]

    @staticmethod
    def all_notnan() -> "RowProgram":
        return RowProgram((OPS["ALL_NOTNAN"],), (0,), ())

    @staticmethod
    def all_between(low: float, high: float) -> "RowProgram":
        return RowProgram((OPS["ALL_BETWEEN"],), (0,), (float(low), float(high)))


def _rewrite(expr: str) -> Tuple[str, List[str]]:
    """`tag name` -> placeholder identifiers; `&` / `|` -> and / or with and/or precedence, as pandas' eval parser
    does (pandas/core/computation/expr.py `_replace_booleans`); quoted strings are left alone."""
    out, names, i, n = [], [], 0, len(expr)
    while i < n:
        ch = expr[i]
        if ch == "`":
            j = expr.index("`", i + 1)
            names.append(expr[i + 1:j]); out.append(f" __gb_col_{len(names) - 1}__ "); i = j + 1
        elif ch in "'\"":
            j = expr.index(ch, i + 1)
            out.append(expr[i:j + 1]); i = j + 1
        elif ch == "&":
            out.append(" and "); i += 2 if expr[i:i + 2] == "&&" else 1
        elif ch == "|":
            out.append(" or "); i += 2 if expr[i:i + 2] == "||" else 1
        else:
            out.append(ch); i += 1
    return "".join(out), names


def compile_row_filter(filter_str: Union[str, Sequence[str]], columns: Sequence[Any], ts_base_ns: int = 0,
                       index_tz=None) -> RowProgram:
    """
    ``row_filter`` / ``known_filter_periods`` (a string, or a list that is AND-ed: filters/rows.py
    ``_list_of_str_to_parenthesized_and``) -> postfix program.  The pandas ``DataFrame.eval`` subset gordo configs use:
    column names (bare or in backticks), ``index`` (compared with timestamp strings), numbers, ``+ - * / **``, unary
    ``-``, ``abs()``, comparisons (chained ones too), ``& | ~`` / ``and or not``, parentheses.
    """
    if not isinstance(filter_str, str):
        filter_str = " & ".join(f"({f})" for f in filter_str)
    src, quoted = _rewrite(filter_str)
    try:
        tree = ast.parse(src.strip(), mode="eval")
    except SyntaxError as exc:
        raise ValueError(f"row filter {filter_str!r} is not a valid expression: {exc}") from None
    col_index: Dict[Any, int] = {c: i for i, c in enumerate(columns)}
    ops: List[int] = []; args: List[int] = []; consts: List[float] = []

    def emit(op, arg=0):
        ops.append(OPS[op]); args.append(int(arg))

    def const(v: float):
        v = float(v)
        for i, c in enumerate(consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return emit("CONST", i)
        consts.append(v); emit("CONST", len(consts) - 1)

    def timestamp(s: str):
        t = pd.Timestamp(s)
        if t.tzinfo is None:
            t = t.tz_localize(index_tz or "UTC")
        const(float(t.value - int(ts_base_ns)))

    def column(name):
        if name not in col_index:
            raise ValueError(f"row filter {filter_str!r} names {name!r}, which is not a column of the data "
                             f"({list(columns)[:8]}...)")
        emit("COL", col_index[name])

    def visit(node, in_compare=False):
        if isinstance(node, ast.Expression):
            return visit(node.body)
        if isinstance(node, ast.BoolOp):
            visit(node.values[0])
            for v in node.values[1:]:
                visit(v); emit("AND" if isinstance(node.op, ast.And) else "OR")
            return
        if isinstance(node, ast.UnaryOp):
            visit(node.operand)
            if isinstance(node.op, (ast.Not, ast.Invert)):
                emit("NOT")
            elif isinstance(node.op, ast.USub):
                emit("NEG")
            elif not isinstance(node.op, ast.UAdd):
                raise NotImplementedError(f"unary operator in row filter {filter_str!r}")
            return
        if isinstance(node, ast.BinOp):
            table = {ast.Add: "ADD", ast.Sub: "SUB", ast.Mult: "MUL", ast.Div: "DIV", ast.Pow: "POW"}
            if type(node.op) not in table:
                raise NotImplementedError(f"operator {type(node.op).__name__} in row filter {filter_str!r}")
            visit(node.left); visit(node.right); emit(table[type(node.op)])
            return
        if isinstance(node, ast.Compare):
            table = {ast.Gt: "GT", ast.GtE: "GE", ast.Lt: "LT", ast.LtE: "LE", ast.Eq: "EQ", ast.NotEq: "NE"}
            terms = [node.left] + list(node.comparators)
            for k, op in enumerate(node.ops):
                if type(op) not in table:
                    raise NotImplementedError(f"comparison {type(op).__name__} in row filter {filter_str!r}")
                visit(terms[k], True); visit(terms[k + 1], True); emit(table[type(op)])
                if k:
                    emit("AND")
            return
        if isinstance(node, ast.Call):
            if isinstance(node.func, ast.Name) and node.func.id == "abs" and len(node.args) == 1 and not node.keywords:
                visit(node.args[0]); emit("ABS")
                return
            raise NotImplementedError(f"function call in row filter {filter_str!r} (only abs() is supported)")
        if isinstance(node, ast.Constant):
            if isinstance(node.value, bool):
                return const(1.0 if node.value else 0.0)
            if isinstance(node.value, (int, float)):
                return const(node.value)
            if isinstance(node.value, str) and in_compare:
                return timestamp(node.value)
            raise NotImplementedError(f"constant {node.value!r} in row filter {filter_str!r}")
        if isinstance(node, ast.Name):
            m = re.fullmatch(r"__gb_col_(\d+)__", node.id)
            if m:
                return column(quoted[int(m.group(1))])
            if node.id in col_index:
                return column(node.id)
            if node.id == "index":
                return emit("INDEX")
            return column(node.id)
        raise NotImplementedError(f"{type(node).__name__} in row filter {filter_str!r}")

    visit(tree)
    depth = peak = 0
    for op in ops:
        if op in (OPS["CONST"], OPS["COL"], OPS["INDEX"], OPS["ALL_FINITE"], OPS["ALL_NOTNAN"], OPS["ALL_BETWEEN"]):
            depth += 1
        elif op not in (OPS["NEG"], OPS["ABS"], OPS["NOT"]):
            depth -= 1
        peak = max(peak, depth)
    if len(ops) > MAX_OPS or len(consts) > MAX_CONSTS or peak > MAX_STACK:
        raise ValueError(f"row filter {filter_str!r} is too long for the device interpreter "
                         f"({len(ops)} operations / {len(consts)} constants / depth {peak}; "
                         f"limits {MAX_OPS} / {MAX_CONSTS} / {MAX_STACK})")
    return RowProgram(tuple(ops), tuple(args), tuple(consts))


# --------------------------------------------------------------------------------------------- device pipeline
def _torch():
    import torch
    return torch


def _stream_ptr():
    return _torch().cuda.current_stream().cuda_stream


def _grid_of(first_ns: int, last_ns: int, tz, freq: str, step: int, cache: dict) -> Tuple[int, int]:
    """(left edge of bin 0, number of bins) pandas gives a series spanning [first, last]: asked of pandas itself on a
    two-sample series, so origin / timezone rules are its own."""
    key = (first_ns, last_ns, str(tz), freq)
    if key not in cache:
        idx = pd.DatetimeIndex(np.array([first_ns, last_ns], dtype="datetime64[ns]"), tz="UTC")
        if tz is not None:
            idx = idx.tz_convert(tz)
        labels = pd.Series([np.nan, np.nan], index=idx).resample(freq, label="left").size().index
        a = labels.asi8
        if len(a) > 1 and a[-1] - a[0] != (len(a) - 1) * step:
            raise NotImplementedError("resampling bins of unequal length (a clock change inside the range)")
        cache[key] = (int(a[0]), int(len(a)))
    return cache[key]


@dataclass
class MachineSeries:
    """One Machine's raw inputs: its tag series (tz-aware DatetimeIndex, float values) and its train period."""
    series: Sequence[pd.Series]
    start: Any
    end: Any
    name: str = ""
    row_filter: Any = None                 # None: the fleet-wide argument of FleetTimeSeries.get_data applies
    known_filter_periods: Any = None


class JoinedMachine:
    """A Machine's joined grid on the device: ``values`` [rows, columns] float64, ``values_f32``, ``index_ns``."""

    def __init__(self, owner, group, job, columns, tz, name=""):
        self._owner, self._group, self._job = owner, group, job
        self.columns, self.tz, self.name = columns, tz, name

    def _range(self):
        g = self._group
        return int(g["lo_host"][self._job]), int(g["hi_host"][self._job])

    @property
    def values(self):
        a, b = self._range()
        return self._group["data"][a:b]

    @property
    def values_f32(self):
        a, b = self._range()
        return self._group["data_f32"][a:b]

    @property
    def index_ns(self):
        a, b = self._range()
        return self._group["ts"][a:b]

    def __len__(self):
        a, b = self._range()
        return b - a

    def frame(self) -> pd.DataFrame:
        idx = pd.DatetimeIndex(self.index_ns.cpu().numpy().astype("datetime64[ns]"), tz="UTC")
        if self.tz is not None:
            idx = idx.tz_convert(self.tz)
        cols = self.columns
        if cols and isinstance(cols[0], tuple):
            cols = pd.MultiIndex.from_tuples(cols, names=["tag", "aggregation_method"])
        return pd.DataFrame(self.values.cpu().numpy(), index=idx, columns=cols)


class FleetTimeSeries:
    """``join_timeseries`` / ``get_data`` for many Machines at once (Machines with the same number of columns share
    the filter / compaction launches)."""

    def __init__(self, device: str = "cuda:0", host_threads: Optional[int] = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("gordo_b200.dataset needs a CUDA device (there is no CPU path)")
        self.device = torch.device(device)
        N.lib()
        from .hostbind import effective_cpus
        self.host_threads = host_threads or max(1, min(16, effective_cpus()))
        self._staging: Dict[str, Any] = {}
        self._staging_events: Dict[str, Any] = {}

    # ------------------------------------------------------------------ raw samples -> device
    def _upload(self, name: str, parts: Sequence[np.ndarray], point_off: Sequence[int], dtype):
        """The series' arrays copied side by side into a pinned staging buffer (kept and reused between calls; filled
        by a few threads: numpy copies release the GIL) and from there to the device, slice by slice, asynchronously."""
        torch = _torch()
        n = int(point_off[-1])
        if name in self._staging_events:
            self._staging_events[name].synchronize()               # the previous call's copy has left the buffer
        stage = self._staging.get(name)
        if stage is None or stage.numel() < n or stage.dtype != dtype:
            stage = torch.empty((max(n, 1),), dtype=dtype, pin_memory=True)
            self._staging[name] = stage
        host = stage.numpy()
        dev_t = torch.empty((max(n, 1),), dtype=dtype, device=self.device)

        def fill(k0, k1):
            for k in range(k0, k1):
                host[point_off[k]:point_off[k + 1]] = parts[k]

        workers = max(1, min(self.host_threads, len(parts)))
        if workers == 1 or n < (1 << 22):
            fill(0, len(parts))
            dev_t[:n].copy_(stage[:n], non_blocking=True)
        else:
            # a few slices of series: while the DMA engine moves slice c, the threads fill slice c + 1
            from concurrent.futures import ThreadPoolExecutor
            n_slices = 4
            bounds = np.searchsorted(np.asarray(point_off), np.linspace(0, n, n_slices + 1)[1:-1]).tolist()
            slices = sorted(set([0] + [int(b) for b in bounds] + [len(parts)]))
            with ThreadPoolExecutor(max_workers=workers) as pool:
                for s0, s1 in zip(slices[:-1], slices[1:]):
                    cuts = np.linspace(s0, s1, workers + 1).astype(int)
                    list(pool.map(lambda ab: fill(*ab), zip(cuts[:-1], cuts[1:])))
                    a, b = int(point_off[s0]), int(point_off[s1])
                    if b > a:
                        dev_t[a:b].copy_(stage[a:b], non_blocking=True)
        self._staging_events[name] = torch.cuda.Event(); self._staging_events[name].record()
        return dev_t[:n] if n else dev_t[:0]

    # ------------------------------------------------------------------ resample + interpolate + inner join
    def join(self, machines: Sequence[MachineSeries], resolution: str, aggregation_methods: Union[str, List[str]] = "mean",
             interpolation_method: str = "linear_interpolation", interpolation_limit: Optional[str] = "8H") -> List[JoinedMachine]:
        torch = _torch()
        if interpolation_method not in INTERP_CODES:
            raise ValueError("Interpolation method should be either linear_interpolation or ffill")
        methods = [aggregation_methods] if isinstance(aggregation_methods, str) else list(aggregation_methods)
        for a in methods:
            if a not in AGG_CODES:
                raise NotImplementedError(f"aggregation method {a!r} (supported: {sorted(AGG_CODES)})")
        multi = not isinstance(aggregation_methods, str)
        freq, step = normalize_freq(resolution), _step_ns(resolution)
        limit = interpolation_limit_bins(interpolation_limit, resolution)
        nm = len(methods)
        cache: dict = {}
        t_src, v_src, point_off = [], [], [0]
        # nanoseconds per tick of the series' indexes (pandas >= 3 builds "us" indexes): when every series has the same
        # unit the ticks go up as they are and are scaled on the device, otherwise the host converts to nanoseconds
        units = {s.index.unit for mc in machines for s in mc.series if isinstance(s.index, pd.DatetimeIndex)}
        unit_ns = _UNIT_NS[units.pop()] if len(units) == 1 else 1
        desc = []                      # per series: machine, column, bin0, n_bins
        mgrid = []                     # per machine: g0, rows, n_columns, tz
        for mi, mc in enumerate(machines):
            start, end = pd.Timestamp(mc.start), pd.Timestamp(mc.end)
            if start.tzinfo is None or end.tzinfo is None:
                raise ValueError("resampling_startpoint / resampling_endpoint must be timezone aware")
            firsts, ends, tz = [], [], None
            for j, s in enumerate(mc.series):
                if len(s) == 0:
                    raise ValueError(f"series {s.name!r} has no samples")
                idx = s.index
                if not isinstance(idx, pd.DatetimeIndex) or idx.tz is None:
                    raise ValueError(f"series {s.name!r} needs a timezone-aware DatetimeIndex")
                tz = tz or idx.tz
                t = idx.asi8 if unit_ns != 1 or idx.unit == "ns" else idx.as_unit("ns").asi8
                v = s.to_numpy()
                if v.dtype != np.float64:
                    v = v.astype(np.float64)
                if not idx.is_monotonic_increasing:
                    order = np.argsort(t, kind="stable"); t, v = t[order], v[order]
                first, last = min(int(start.value), int(t[0]) * unit_ns), max(int(end.value), int(t[-1]) * unit_ns)
                b0, nb = _grid_of(first, last, idx.tz, freq, step, cache)
                t_src.append(t); v_src.append(v); point_off.append(point_off[-1] + len(t))
                desc.append((mi, j, b0, nb)); firsts.append(b0); ends.append(b0 + nb * step)
            g0, g1 = min(firsts), max(ends)
            if any((f - g0) % step for f in firsts):
                raise NotImplementedError("the resampling grids of this Machine's series are out of phase")
            mgrid.append((g0, (g1 - g0) // step, len(mc.series) * nm, tz))
        # groups: Machines with the same number of columns, rows back to back
        groups: Dict[int, dict] = {}
        flat_off = 0
        for mi, (g0, rows, C, tz) in enumerate(mgrid):
            g = groups.setdefault(C, dict(C=C, machines=[], lo=[], hi=[], rows=0))
            g["machines"].append(mi); g["lo"].append(g["rows"]); g["hi"].append(g["rows"] + rows); g["rows"] += rows
        for C, g in groups.items():
            g["base"] = flat_off; flat_off += g["rows"] * C
        where = {}
        for C, g in groups.items():
            for k, mi in enumerate(g["machines"]):
                where[mi] = (g, k)
        dev = self.device
        flat = torch.full((max(flat_off, 1),), float("nan"), dtype=torch.float64, device=dev)
        n_series = len(desc)
        bin0 = np.array([d[2] for d in desc], np.int64); nbins = np.array([d[3] for d in desc], np.int64)
        off = np.empty(n_series, np.int64); stride = np.empty(n_series, np.int64)
        for i, (mi, j, b0, nb) in enumerate(desc):
            g, k = where[mi]
            g0 = mgrid[mi][0]
            off[i] = g["base"] + (g["lo"][k] + (b0 - g0) // step) * g["C"] + j * nm
            stride[i] = g["C"]
        d_ts = self._upload("ts", t_src, point_off, torch.int64)
        if unit_ns and unit_ns != 1:
            d_ts.mul_(unit_ns)                                     # ticks -> nanoseconds, on the device
        d_val = self._upload("val", v_src, point_off, torch.float64)
        d_poff = torch.as_tensor(np.asarray(point_off, np.int64), device=dev)
        d_bin0 = torch.as_tensor(bin0, device=dev); d_nb = torch.as_tensor(nbins, device=dev)
        d_stride = torch.as_tensor(stride, device=dev)
        lib = N.lib()
        max_bins, total_bins = int(nbins.max()) if n_series else 0, int(nbins.sum())
        for a, method in enumerate(methods):
            d_off = torch.as_tensor(off + a, device=dev)
            for s0 in range(0, n_series, 65535):
                s1 = min(n_series, s0 + 65535)
                N.check(lib.gb200_resample(s1 - s0, N.ptr(d_poff[s0:]), N.ptr(d_ts), N.ptr(d_val), N.ptr(d_bin0[s0:]),
                                           N.ptr(d_nb[s0:]), N.ptr(d_off[s0:]), N.ptr(d_stride[s0:]), step, AGG_CODES[method],
                                           int(nbins[s0:s1].max()), int(point_off[s1] - point_off[s0]),
                                           int(nbins[s0:s1].sum()), N.ptr(flat), _stream_ptr()), "gb200_resample")
            N.check(lib.gb200_interpolate(n_series, N.ptr(d_nb), N.ptr(d_off), N.ptr(d_stride),
                                          INTERP_CODES[interpolation_method], -1 if limit is None else int(limit),
                                          N.ptr(flat), _stream_ptr()), "gb200_interpolate")
        out: List[Optional[JoinedMachine]] = [None] * len(machines)
        for C, g in groups.items():
            data = flat[g["base"]:g["base"] + g["rows"] * C].view(g["rows"], C)
            lo = torch.as_tensor(np.asarray(g["lo"], np.int64), device=dev)
            hi = torch.as_tensor(np.asarray(g["hi"], np.int64), device=dev)
            g0s = torch.as_tensor(np.array([mgrid[mi][0] for mi in g["machines"]], np.int64), device=dev)
            job = torch.repeat_interleave(torch.arange(len(g["machines"]), device=dev), hi - lo)
            ts = g0s[job] + (torch.arange(g["rows"], device=dev) - lo[job]) * step
            state = dict(C=C, data=data, ts=ts, lo=lo, hi=hi, lo_host=np.asarray(g["lo"]), hi_host=np.asarray(g["hi"]),
                         data_f32=None)
            self._apply(state, RowProgram.all_notnan(), 0, list(range(len(g["machines"]))))     # the dropna() of the join
            for k, mi in enumerate(g["machines"]):
                mc = machines[mi]
                cols = [(s.name, a) for s in mc.series for a in methods] if multi else [s.name for s in mc.series]
                out[mi] = JoinedMachine(self, state, k, cols, mgrid[mi][3], getattr(mc, "name", ""))
        return out

    # ------------------------------------------------------------------ one predicate stage over a group
    def _apply(self, state: dict, prog: RowProgram, buffer_size: int, jobs: Sequence[int], ts_base: int = 0):
        """keep-mask + compaction of the jobs in ``jobs`` (the others pass through unchanged)."""
        torch = _torch()
        dev = self.device
        data, ts, lo, hi = state["data"], state["ts"], state["lo"], state["hi"]
        R, C = data.shape
        if R == 0:                                   # nothing left to filter (e.g. series that never overlap)
            if state["data_f32"] is None:
                state["data_f32"] = torch.empty((0, C), dtype=torch.float32, device=dev)
            return
        keep = torch.ones((max(R, 1),), dtype=torch.uint8, device=dev)
        if len(jobs) and R:
            sel = torch.as_tensor(np.asarray(jobs, np.int64), device=dev)
            ops = (N.C.c_int32 * len(prog.ops))(*prog.ops); args = (N.C.c_int32 * len(prog.args))(*prog.args)
            consts = (N.C.c_double * max(len(prog.consts), 1))(*prog.consts)
            jlo, jhi = lo[sel].contiguous(), hi[sel].contiguous()       # named: both must be alive when the call reads them
            N.check(N.lib().gb200_filter_rows(len(jobs), N.ptr(jlo), N.ptr(jhi), N.ptr(data), C,
                                              N.ptr(ts), int(ts_base), ops, args, len(prog.ops), consts, len(prog.consts),
                                              int(buffer_size), N.ptr(keep), _stream_ptr()), "gb200_filter_rows")
        out = torch.empty_like(data); out32 = torch.empty(data.shape, dtype=torch.float32, device=dev)
        out_ts = torch.empty_like(ts); nlo = torch.empty_like(lo); nhi = torch.empty_like(hi)
        N.check(N.lib().gb200_compact_rows(lo.numel(), N.ptr(lo), N.ptr(hi), N.ptr(data), C, N.ptr(ts), N.ptr(keep),
                                           N.ptr(out), N.ptr(out32), N.ptr(out_ts), N.ptr(nlo), N.ptr(nhi), _stream_ptr()),
                "gb200_compact_rows")
        state["lo_host"], state["hi_host"] = nlo.cpu().numpy(), nhi.cpu().numpy()
        n = int(state["hi_host"][-1]) if len(state["hi_host"]) else 0
        state.update(data=out[:n], data_f32=out32[:n], ts=out_ts[:n], lo=nlo, hi=nhi)

    # ------------------------------------------------------------------ TimeSeriesDataset.get_data, minus the provider
    def get_data(self, machines: Sequence[MachineSeries], resolution: str = "10T", aggregation_methods="mean",
                 interpolation_method: str = "linear_interpolation", interpolation_limit: Optional[str] = "8H",
                 row_filter: Union[str, Sequence[str]] = "",
                 known_filter_periods=None, row_filter_buffer_size: int = 0, n_samples_threshold: int = 0,
                 low_threshold: Optional[float] = -1000, high_threshold: Optional[float] = 50000,
                 filter_periods=None) -> List[JoinedMachine]:
        """join -> row-count check -> known_filter_periods -> row_filter -> global thresholds (time_series.py get_data).
        ``row_filter`` / ``known_filter_periods`` apply to every Machine that does not carry its own (MachineSeries)."""
        if filter_periods:
            raise NotImplementedError("filter_periods (median / iforest period filters) are not part of this path")
        if low_threshold is not None and high_threshold is not None and low_threshold >= high_threshold:
            raise ValueError("Low threshold need to be larger than high threshold")
        joined = self.join(machines, resolution, aggregation_methods, interpolation_method, interpolation_limit)
        for jm in joined:
            if len(jm) <= n_samples_threshold:
                raise InsufficientDataError(
                    f"The length of the generated DataFrame ({len(jm)}) does not exceed the specified required threshold "
                    f"for number of rows ({n_samples_threshold}).")

        def per_machine(attr, fleet_value):
            return [getattr(mc, attr, None) if getattr(mc, attr, None) is not None else fleet_value for mc in machines]

        states = {id(jm._group): jm._group for jm in joined}
        position = {id(jm): i for i, jm in enumerate(joined)}
        for stage_values in (per_machine("known_filter_periods", known_filter_periods), per_machine("row_filter", row_filter)):
            for state in states.values():
                members = [jm for jm in joined if jm._group is state]
                base = int(state["ts"].min().item()) if state["ts"].numel() else 0
                by_prog: Dict[RowProgram, List[int]] = {}
                for jm in members:
                    f = stage_values[position[id(jm)]]
                    if not f:
                        continue
                    prog = compile_row_filter(f, jm.columns, base, jm.tz)
                    by_prog.setdefault(prog, []).append(jm._job)
                for prog, jobs in by_prog.items():
                    self._apply(state, prog, row_filter_buffer_size, jobs, base)
        if low_threshold is not None and high_threshold is not None:
            for state in states.values():
                self._apply(state, RowProgram.all_between(low_threshold, high_threshold), 0, list(range(state["lo"].numel())))
        return joined


# --------------------------------------------------------------------------------------------- gordo-core's names
def join_timeseries(series_iterable: Iterable[pd.Series], resampling_startpoint, resampling_endpoint, resolution: str,
                    aggregation_methods: Union[str, List[str]] = "mean", interpolation_method: str = "linear_interpolation",
                    interpolation_limit: Optional[str] = "8H", device: str = "cuda:0") -> pd.DataFrame:
    """``TimeSeriesDataset.join_timeseries`` (gordo_core/time_series.py) for one Machine, on the GPU."""
    fleet = FleetTimeSeries(device)
    jm = fleet.join([MachineSeries(list(series_iterable), resampling_startpoint, resampling_endpoint)], resolution,
                    aggregation_methods, interpolation_method, interpolation_limit)[0]
    return jm.frame()


def pandas_filter_rows(df: pd.DataFrame, filter_str: Union[str, Sequence[str]], buffer_size: int = 0,
                       device: str = "cuda:0") -> pd.DataFrame:
    """``pandas_filter_rows`` (gordo_core/filters/rows.py): the rows of ``df`` the expression keeps, with
    ``apply_buffer`` around the rejected ones; the mask is evaluated on the GPU."""
    torch = _torch()
    FleetTimeSeries(device)
    dev = torch.device(device)
    idx = df.index
    tz = getattr(idx, "tz", None)
    ts_host = idx.as_unit("ns").asi8 if isinstance(idx, pd.DatetimeIndex) else np.arange(len(df), dtype=np.int64)
    base = int(ts_host.min()) if len(df) else 0
    prog = compile_row_filter(filter_str, list(df.columns), base, tz)
    data = torch.as_tensor(np.array(df.to_numpy(np.float64), order="C", copy=True), device=dev)     # (a frame's array may be read-only)
    ts = torch.as_tensor(np.ascontiguousarray(ts_host), device=dev)
    lo = torch.zeros(1, dtype=torch.int64, device=dev); hi = torch.full((1,), len(df), dtype=torch.int64, device=dev)
    keep = torch.ones((max(len(df), 1),), dtype=torch.uint8, device=dev)
    if len(df):
        ops = (N.C.c_int32 * len(prog.ops))(*prog.ops); args = (N.C.c_int32 * len(prog.args))(*prog.args)
        consts = (N.C.c_double * max(len(prog.consts), 1))(*prog.consts)
        N.check(N.lib().gb200_filter_rows(1, N.ptr(lo), N.ptr(hi), N.ptr(data), data.shape[1], N.ptr(ts), base, ops, args,
                                          len(prog.ops), consts, len(prog.consts), int(buffer_size), N.ptr(keep), _stream_ptr()),
                "gb200_filter_rows")
    return df[keep[:len(df)].cpu().numpy().astype(bool)]


def get_data(series: Sequence[pd.Series], train_start_date, train_end_date, resolution: str = "10T", device: str = "cuda:0",
             **dataset_kwargs) -> pd.DataFrame:
    """``TimeSeriesDataset.get_data`` between ``load_series`` and the X / y column split, for one Machine."""
    fleet = FleetTimeSeries(device)
    return fleet.get_data([MachineSeries(list(series), train_start_date, train_end_date)], resolution, **dataset_kwargs)[0].frame()


class TimeSeriesDataset:
    """
    The dataset object of a Machine with ``get_data()`` on the GPU: constructor arguments, ``get_data() -> (X, y)`` and
    ``get_metadata()`` as gordo-core 0.3.6's ``gordo_core.time_series.TimeSeriesDataset`` (the object
    gordo/builder/build_model.py:208-213 builds with ``GordoBaseDataset.from_dict`` and calls).  ``data_provider`` is
    anything with ``load_series(train_start_date, train_end_date, tag_list) -> iterable of pandas Series`` (the
    gordo-core provider interface); fetching itself is out of scope.  Tags are plain names or objects with ``.name``.
    """

    def __init__(self, train_start_date, train_end_date, tag_list: Sequence[Any], target_tag_list: Optional[Sequence[Any]] = None,
                 data_provider: Any = None, resolution: Optional[str] = "10T", row_filter: Union[str, list] = "",
                 known_filter_periods: Optional[list] = None, aggregation_methods: Union[str, List[str]] = "mean",
                 row_filter_buffer_size: int = 0, n_samples_threshold: int = 0, low_threshold: Optional[float] = -1000,
                 high_threshold: Optional[float] = 50000, interpolation_method: str = "linear_interpolation",
                 interpolation_limit: Optional[str] = "8H", filter_periods: Optional[dict] = None, device: str = "cuda:0",
                 **kwargs):
        self.train_start_date, self.train_end_date = pd.Timestamp(train_start_date), pd.Timestamp(train_end_date)
        if self.train_start_date.tzinfo is None or self.train_end_date.tzinfo is None:
            raise ValueError(f"Timestamps ({train_start_date}, {train_end_date}) need to include timezone information")
        if self.train_start_date >= self.train_end_date:
            raise ValueError(f"train_end_date ({train_end_date}) must be after train_start_date ({train_start_date})")
        if resolution is None:
            raise NotImplementedError("resolution=None (join without resampling) is not part of this path")
        self.tag_list = list(tag_list)
        self.target_tag_list = list(target_tag_list) if target_tag_list else list(tag_list)
        self.data_provider, self.resolution = data_provider, resolution
        self.row_filter, self.known_filter_periods = row_filter, known_filter_periods or []
        self.aggregation_methods, self.row_filter_buffer_size = aggregation_methods, row_filter_buffer_size
        self.n_samples_threshold, self.low_threshold, self.high_threshold = n_samples_threshold, low_threshold, high_threshold
        self.interpolation_method, self.interpolation_limit = interpolation_method, interpolation_limit
        self.filter_periods, self.device = filter_periods, device
        self._metadata: Dict[str, Any] = {}

    @staticmethod
    def _name(tag) -> str:
        return getattr(tag, "name", tag)

    def get_data(self) -> Tuple[pd.DataFrame, Optional[pd.DataFrame]]:
        tags = list(dict.fromkeys(self._name(t) for t in self.tag_list + self.target_tag_list))
        wanted = [t for t in self.tag_list + self.target_tag_list]
        seen, load = set(), []
        for t in wanted:
            if self._name(t) not in seen:
                seen.add(self._name(t)); load.append(t)
        series = list(self.data_provider.load_series(self.train_start_date, self.train_end_date, load))
        data = get_data(series, self.train_start_date, self.train_end_date, self.resolution, device=self.device,
                        aggregation_methods=self.aggregation_methods, interpolation_method=self.interpolation_method,
                        interpolation_limit=self.interpolation_limit, row_filter=self.row_filter,
                        known_filter_periods=self.known_filter_periods, row_filter_buffer_size=self.row_filter_buffer_size,
                        n_samples_threshold=self.n_samples_threshold, low_threshold=self.low_threshold,
                        high_threshold=self.high_threshold, filter_periods=self.filter_periods)
        x_names = [self._name(t) for t in self.tag_list]
        y_names = [self._name(t) for t in self.target_tag_list]
        if not isinstance(self.aggregation_methods, str):
            pick = lambda names: [c for c in data.columns if c[0] in names]
            X, y = data[pick(x_names)], data[pick(y_names)]
        else:
            X, y = data[x_names], data[y_names]
        self._metadata = {"tag_loading_metadata": {"tags": tags}, "train_start_date_actual": X.index[0] if len(X) else None,
                          "train_end_date_actual": X.index[-1] if len(X) else None,
                          "summary_statistics": X.describe().to_dict(), "x_hist": {}, "row_count": len(X)}
        return X, y

    def get_metadata(self) -> Dict[str, Any]:
        return dict(self._metadata)
