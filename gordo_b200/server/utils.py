"""
Wire formats either side of the anomaly call, with the behaviour of
gordo/server/utils.py:

  dataframe_into_parquet_bytes / dataframe_from_parquet_bytes   (:47-85)   the ``?format=parquet`` response
  dataframe_to_dict / dataframe_from_dict                       (:88-195)  the nested-JSON response / request

plus the fleet fast paths that skip the DataFrame pivot: the anomaly columns already sit in host memory as one
float32 matrix per column group (serving.FleetAnomalyResult, DiffBasedAnomalyDetector._fused_columns), so

  columns_into_parquet_bytes(groups, index, frequency)   builds the Arrow table column by column and writes the
                                                         SAME parquet file the reference writes for the frame
                                                         (schema, pandas metadata, MultiIndex column names);
  columns_to_dict(groups, index, frequency)              builds the nested dict of dataframe_to_dict directly.

``groups`` is what ``model/utils.assemble_frame`` takes: [(top-level name, values [n] or [n, k], second-level
names or None)].  Host-side pandas / pyarrow code: nothing here touches the device.
"""
import io
from datetime import timedelta
from typing import Optional

import dateutil.parser
import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq

from gordo_b200.machine.model import utils as model_utils


def dataframe_into_parquet_bytes(df: pd.DataFrame, compression: str = "snappy") -> bytes:
    table = pa.Table.from_pandas(df)
    buf = pa.BufferOutputStream()
    pq.write_table(table, buf, compression=compression)
    return buf.getvalue().to_pybytes()


def dataframe_from_parquet_bytes(buf: bytes) -> pd.DataFrame:
    return pq.read_table(io.BytesIO(buf)).to_pandas()


def dataframe_to_dict(df: pd.DataFrame) -> dict:
    """MultiIndex-column frame -> {top: {sub: {index: value}}} (json.dumps-able); plain frames -> DataFrame.to_dict."""
    data = df.copy()
    if isinstance(data.index, pd.DatetimeIndex):
        data.index = data.index.astype(str)
    if isinstance(df.columns, pd.MultiIndex):
        return {col: (data[col].to_dict() if isinstance(data[col], pd.DataFrame) else pd.DataFrame(data[col]).to_dict())
                for col in data.columns.get_level_values(0)}
    return data.to_dict()


def dataframe_from_dict(data: dict) -> pd.DataFrame:
    """Inverse of dataframe_to_dict; the index is parsed as ISO datetimes, else as integers, and sorted."""
    if isinstance(data, dict) and any(isinstance(val, dict) for val in data.values()):
        try:
            keys = data.keys()
            df = pd.concat((pd.DataFrame.from_dict(data[key]) for key in keys), axis=1, keys=keys)
        except (ValueError, AttributeError):
            df = pd.DataFrame.from_dict(data)
    else:
        df = pd.DataFrame.from_dict(data)
    try:
        df.index = df.index.map(dateutil.parser.isoparse)
    except (TypeError, ValueError):
        df.index = df.index.map(int)
    df.sort_index(inplace=True)
    return df


# ----------------------------------------------------------------------------- fleet fast paths
def _template(groups, index, frequency):
    """The one-row frame of these groups: carries the column MultiIndex, dtypes and index type the reference would see."""
    first = [(name, np.asarray(v)[:1], names) for name, v, names in groups]
    idx0 = index[-len(np.asarray(groups[0][1])):][:1] if index is not None else None
    return model_utils.assemble_frame(first, idx0, frequency)


def columns_into_parquet_bytes(groups, index=None, frequency: Optional[timedelta] = None,
                               compression: str = "snappy") -> bytes:
    """
    ``dataframe_into_parquet_bytes(assemble_frame(groups, index, frequency))`` without building the frame: the
    schema (field names, pandas metadata) comes from the zero-row template, every value column is handed to Arrow
    as its own float64 array.  Reads back (dataframe_from_parquet_bytes) to the identical frame.
    """
    n = len(np.asarray(groups[0][1]))
    if n == 0:
        return dataframe_into_parquet_bytes(model_utils.assemble_frame(groups, index, frequency), compression)
    idx, start, end = model_utils.time_columns(index, n, frequency)
    template = pa.Table.from_pandas(_template(groups, index, frequency))
    schema = template.schema
    arrays = [pa.array(start, type=schema.field(0).type), pa.array(end, type=schema.field(1).type)]
    f = 2
    for name, values, names in groups:
        v = np.asarray(values)
        cols = [v] if v.ndim == 1 else [v[:, j] for j in range(v.shape[1])]
        for c in cols:
            arrays.append(pa.array(np.ascontiguousarray(c, dtype=np.float64), type=schema.field(f).type))
            f += 1
    # the index column(s) pandas metadata expects (a RangeIndex lives in the metadata only)
    while f < len(schema):
        field = schema.field(f)
        arrays.append(pa.array(idx, type=field.type) if not isinstance(idx, pd.RangeIndex) else pa.array(np.arange(n), type=field.type))
        f += 1
    if isinstance(idx, pd.RangeIndex):
        # the template's metadata describes RangeIndex(0, 1): patch stop to n
        import json
        meta = json.loads(schema.metadata[b"pandas"])
        for ic in meta.get("index_columns", []):
            if isinstance(ic, dict) and ic.get("kind") == "range":
                ic["stop"] = n
        schema = schema.with_metadata({**schema.metadata, b"pandas": json.dumps(meta).encode()})
    table = pa.Table.from_arrays(arrays, schema=schema)
    buf = pa.BufferOutputStream()
    pq.write_table(table, buf, compression=compression)
    return buf.getvalue().to_pybytes()


def columns_to_dict(groups, index=None, frequency: Optional[timedelta] = None) -> dict:
    """``dataframe_to_dict(assemble_frame(groups, index, frequency))`` built directly from the column groups."""
    n = len(np.asarray(groups[0][1]))
    idx, start, end = model_utils.time_columns(index, n, frequency)
    keys = list(idx.astype(str)) if isinstance(idx, pd.DatetimeIndex) else list(idx)
    # a top-level name whose second level is "" selects a Series named like the top level: {name: {name: {...}}}
    out = {"start": {"start": dict(zip(keys, start))}, "end": {"end": dict(zip(keys, end))}}
    for name, values, names in groups:
        v = np.asarray(values, dtype=np.float64)
        if v.ndim == 1:
            out[name] = {name: dict(zip(keys, v.tolist()))}
        else:
            out[name] = {sub: dict(zip(keys, v[:, j].tolist())) for j, sub in enumerate(names)}
    return out
