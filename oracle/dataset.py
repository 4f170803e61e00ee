"""
TEST INFRASTRUCTURE (see oracle/__init__.py): pandas restatement of what `dataset.get_data()` does between the data
provider and the matrix the builder trains on -- SURVEY.md §8 f-4 "upstream".

Reference call site: gordo/builder/build_model.py:208-213 (`GordoBaseDataset.from_dict(...)`, `dataset.get_data()`),
gordo/machine/machine.py:191-195.  The implementation is third-party: **gordo-core 0.3.6** (pinned in
requirements/full_requirements.txt:141), `gordo_core/time_series.py` (`TimeSeriesDataset.get_data`,
`TimeSeriesDataset.join_timeseries`) and `gordo_core/filters/rows.py` (`pandas_filter_rows`, `apply_buffer`).
The package is NOT in /root/reference and cannot be installed here, so this file restates its published algorithm.

PARITY UNPINNED for this module: no gordo-core code or golden vector was available to check the restatement against;
what anchors it is pandas itself (every step below IS a pandas call, as in gordo-core) and the dataset defaults the
reference's own test pins (tests/gordo/workflow/test_config_elements.py:139-157: resolution "10T", aggregation "mean",
linear interpolation, row_filter "", buffer 0, n_samples_threshold 0).
"""
import re
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np
import pandas as pd

_ALIAS = {"T": "min", "H": "h", "S": "s", "L": "ms", "U": "us", "N": "ns"}


def normalize_freq(freq: str) -> str:
    """pandas >= 2.2 spells the offsets gordo configs use ("10T", "8H", "30S") "10min", "8h", "30s"."""
    m = re.fullmatch(r"\s*(\d*\.?\d*)\s*([A-Za-z]+)\s*", str(freq))
    if not m:
        return freq
    n, unit = m.groups()
    return f"{n}{_ALIAS.get(unit, unit)}"


def interpolation_limit_bins(interpolation_limit: Optional[str], resolution: str) -> Optional[int]:
    """join_timeseries: `limit = int(Timedelta(interpolation_limit) / Timedelta(resolution))`, None = unlimited."""
    if not interpolation_limit:
        return None
    limit = int(pd.Timedelta(normalize_freq(interpolation_limit)).total_seconds()
                / pd.Timedelta(normalize_freq(resolution)).total_seconds())
    if limit <= 0:
        raise ValueError("Interpolation limit must be larger than given resolution")
    return limit


def pad_series(series: pd.Series, start: pd.Timestamp, end: pd.Timestamp) -> pd.Series:
    """A NaN sample at the resampling start / end point when the series does not reach it, so that every series of a
    Machine resamples onto the same index."""
    tz = series.index[0].tzinfo
    start, end = start.astimezone(tz), end.astimezone(tz)
    if series.index[0] > start:
        series = pd.concat([pd.Series([np.nan], index=[start], name=series.name), series])
    if series.index[-1] < end:
        series = pd.concat([series, pd.Series([np.nan], index=[end], name=series.name)])
    return series


def join_timeseries(series_iterable: Iterable[pd.Series], resampling_startpoint, resampling_endpoint, resolution: str,
                    aggregation_methods: Union[str, List[str]] = "mean",
                    interpolation_method: str = "linear_interpolation",
                    interpolation_limit: Optional[str] = "8H") -> pd.DataFrame:
    """TimeSeriesDataset.join_timeseries: pad -> resample(resolution, label="left").agg(methods) -> interpolate with a
    limit -> inner join -> dropna."""
    if interpolation_method not in ("linear_interpolation", "ffill"):
        raise ValueError("Interpolation method should be either linear_interpolation or ffill")
    limit = interpolation_limit_bins(interpolation_limit, resolution)
    start, end = pd.Timestamp(resampling_startpoint), pd.Timestamp(resampling_endpoint)
    resampled = []
    for series in series_iterable:
        if len(series) == 0:
            raise ValueError(f"series {series.name!r} has no samples")
        padded = pad_series(series.astype(np.float64), start, end)
        rs = padded.resample(normalize_freq(resolution), label="left").agg(aggregation_methods)
        if isinstance(rs, pd.DataFrame):
            rs.columns = pd.MultiIndex.from_product([[series.name], rs.columns], names=["tag", "aggregation_method"])
        rs = rs.astype(np.float64)
        rs = rs.interpolate(limit=limit) if interpolation_method == "linear_interpolation" else rs.ffill(limit=limit)
        resampled.append(rs)
    joined = pd.concat(resampled, axis=1, join="inner")
    return joined.dropna()


def apply_buffer(mask: np.ndarray, buffer_size: int = 0) -> np.ndarray:
    """filters/rows.py apply_buffer: every False also clears the buffer_size entries on either side of it."""
    mask = np.asarray(mask, bool).copy()
    if buffer_size:
        for idx in np.where(~mask)[0]:
            mask[max(0, idx - buffer_size):min(len(mask), idx + buffer_size + 1)] = False
    return mask


def pandas_filter_rows(df: pd.DataFrame, filter_str: Union[str, Sequence[str]], buffer_size: int = 0) -> pd.DataFrame:
    """filters/rows.py pandas_filter_rows: `df.eval` of the expression (a list is AND-ed), buffered, applied."""
    if not isinstance(filter_str, str):
        filter_str = " & ".join(f"({f})" for f in filter_str)
    mask = apply_buffer(np.asarray(df.eval(filter_str), bool), buffer_size)
    return df[mask]


def get_data(series: Sequence[pd.Series], train_start_date, train_end_date, resolution: str = "10T",
             aggregation_methods="mean", interpolation_method="linear_interpolation", interpolation_limit="8H",
             row_filter: Union[str, Sequence[str]] = "", known_filter_periods: Optional[Sequence[str]] = None,
             row_filter_buffer_size: int = 0, n_samples_threshold: int = 0,
             low_threshold: Optional[float] = -1000, high_threshold: Optional[float] = 50000) -> pd.DataFrame:
    """TimeSeriesDataset.get_data between `load_series` and the X / y column split."""
    data = join_timeseries(series, train_start_date, train_end_date, resolution, aggregation_methods,
                           interpolation_method, interpolation_limit)
    if len(data) <= n_samples_threshold:
        raise ValueError(f"The length of the generated DataFrame ({len(data)}) does not exceed the specified "
                         f"required threshold for number of rows ({n_samples_threshold}).")
    if known_filter_periods:
        data = pandas_filter_rows(data, list(known_filter_periods), buffer_size=row_filter_buffer_size)
    if row_filter:
        data = pandas_filter_rows(data, row_filter, buffer_size=row_filter_buffer_size)
    if low_threshold is not None and high_threshold is not None:
        if low_threshold >= high_threshold:
            raise ValueError("Low threshold need to be larger than high threshold")
        data = data[((data > low_threshold) & (data < high_threshold)).all(axis=1)]
    return data
