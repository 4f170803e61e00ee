"""
Output-frame helpers with the contract of gordo/machine/model/utils.py:18-165:
``metric_wrapper`` and ``make_base_dataframe`` (MultiIndex columns ``start``, ``end``,
``model-input``, ``model-output``; alignment to the last ``len(model_output)`` rows).
The frame is assembled in one block instead of per-group ``DataFrame.join`` copies.
"""
import functools
import logging
from datetime import datetime, timedelta
from typing import List, Optional, Union

import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)


def _tag_name(tag) -> str:
    return str(getattr(tag, "name", tag))


def metric_wrapper(metric, scaler=None):
    """Metric on ``y_true[-len(y_pred):]`` vs ``y_pred``, optionally after ``scaler.transform``."""

    @functools.wraps(metric)
    def _wrapper(y_true, y_pred, *args, **kwargs):
        if scaler:
            y_true = scaler.transform(y_true)
            y_pred = scaler.transform(y_pred)
        return metric(y_true[-len(y_pred):], y_pred, *args, **kwargs)

    return _wrapper


def _second_level(values: np.ndarray, tags) -> List[str]:
    if values.shape[1] == len(tags):
        return [_tag_name(t) for t in tags]
    return [str(i) for i in range(values.shape[1])]


def time_columns(index, n_rows: int, frequency: Optional[timedelta]):
    """(normalized index, start strings, end strings) for the last ``n_rows`` rows."""
    idx = index[-n_rows:] if index is not None else pd.RangeIndex(n_rows)
    if isinstance(idx, pd.DatetimeIndex):
        start = [t.isoformat() for t in idx]
        end = [(t + frequency).isoformat() for t in idx] if frequency is not None else [None] * len(idx)
    else:
        start = [v.isoformat() if hasattr(v, "isoformat") else None for v in idx]
        end = [(v + frequency).isoformat() if isinstance(v, datetime) and frequency is not None else None
               for v in idx]
    return idx, start, end


def assemble_frame(groups, index, frequency: Optional[timedelta]) -> pd.DataFrame:
    """
    groups: ordered list of (top-level name, values [n] or [n, k], second-level names or None).
    Returns the MultiIndex-column frame with ``start`` / ``end`` first.
    """
    n = len(groups[0][1])
    idx, start, end = time_columns(index, n, frequency)
    cols = [("start", ""), ("end", "")]
    numeric = []
    for name, values, names in groups:
        v = np.asarray(values)
        if v.ndim == 1:
            cols.append((name, "")); numeric.append(v[:, None])
        else:
            cols.extend((name, s) for s in names); numeric.append(v)
    block = np.hstack([a.astype(np.float64, copy=False) for a in numeric]) if numeric else np.empty((n, 0))
    mi_full = _column_index(tuple(cols))
    times = np.empty((n, 2), dtype=object)
    times[:, 0] = start; times[:, 1] = end
    # two homogeneous blocks (object times, float64 values) side by side; the column index is cached
    df = pd.concat([pd.DataFrame(times, index=idx), pd.DataFrame(block, index=idx)], axis=1)
    df.columns = mi_full
    return df


@functools.lru_cache(maxsize=64)
def _column_index(cols: tuple) -> pd.MultiIndex:
    return pd.MultiIndex.from_tuples(list(cols))


def make_base_dataframe(tags, model_input: np.ndarray, model_output: np.ndarray,
                        target_tag_list: Optional[List] = None,
                        index: Optional[Union[np.ndarray, pd.Index]] = None,
                        frequency: Optional[timedelta] = None) -> pd.DataFrame:
    """
    ``model-input`` / ``model-output`` frame; the input is clipped to the last ``len(model_output)``
    rows (an LSTM outputs fewer rows than it reads), second-level labels are the tag names when
    the widths match and ``0..k-1`` otherwise.
    """
    target_tag_list = target_tag_list if target_tag_list is not None else tags
    model_output = np.asarray(getattr(model_output, "values", model_output))
    model_input = np.asarray(getattr(model_input, "values", model_input))[-len(model_output):, :]
    groups = [("model-input", model_input, _second_level(model_input, list(tags))),
              ("model-output", model_output, _second_level(model_output, list(target_tag_list)))]
    return assemble_frame(groups, index, frequency)
