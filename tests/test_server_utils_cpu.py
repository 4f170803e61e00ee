"""
§8 f-4 wire formats (CPU): gordo_b200.server.utils against the reference's own gordo/server/utils.py (run from
/root/reference with flask / werkzeug / gordo-core stubbed, in a subprocess, where the reference exists), and the
fleet fast paths (column groups -> parquet bytes / nested dict without the DataFrame pivot) against the frame path.
"""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pandas as pd
import pytest

from gordo_b200.machine.model import utils as mu
from gordo_b200.server import utils as su

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _groups(n=40, T=3, seed=0):
    rng = np.random.default_rng(seed)
    tags = ["tag 1", "tag-2", "t3"][:T]
    return [("model-input", rng.random((n, T)), tags), ("model-output", rng.random((n, T)).astype(np.float32), tags),
            ("tag-anomaly-scaled", rng.random((n, T)), tags), ("total-anomaly-scaled", rng.random(n), None),
            ("tag-anomaly-unscaled", rng.random((n, T)), tags), ("total-anomaly-unscaled", rng.random(n), None),
            ("anomaly-confidence", rng.random((n, T)), tags), ("total-anomaly-confidence", rng.random(n), None)]


@pytest.mark.parametrize("with_time", [True, False])
def test_column_fast_paths_equal_the_frame_path(with_time):
    groups = _groups()
    index = pd.date_range("2020-03-01", periods=40, freq="10min", tz="UTC") if with_time else None
    freq = pd.Timedelta("10min") if with_time else None
    df = mu.assemble_frame(groups, index, freq)
    a = su.dataframe_from_parquet_bytes(su.dataframe_into_parquet_bytes(df))
    b = su.dataframe_from_parquet_bytes(su.columns_into_parquet_bytes(groups, index, freq))
    pd.testing.assert_frame_equal(a, b)
    pd.testing.assert_frame_equal(a, df, check_freq=False)
    assert list(b.columns) == list(df.columns) and isinstance(b.columns, pd.MultiIndex)
    d1, d2 = su.dataframe_to_dict(df), su.columns_to_dict(groups, index, freq)
    assert json.dumps(d1, sort_keys=True, default=str) == json.dumps(d2, sort_keys=True, default=str)
    back = su.dataframe_from_dict(json.loads(json.dumps(d2, default=str)))
    np.testing.assert_allclose(back["tag-anomaly-scaled"].to_numpy(), df["tag-anomaly-scaled"].to_numpy())
    assert len(back) == 40 and back.index.is_monotonic_increasing
    # a longer input than output (LSTM offset): the frame keeps the last len(output) index entries
    if with_time:
        long_index = pd.date_range("2020-03-01", periods=50, freq="10min", tz="UTC")
        c = su.dataframe_from_parquet_bytes(su.columns_into_parquet_bytes(groups, long_index, freq))
        assert c.index[0] == long_index[10] and len(c) == 40


def test_empty_and_plain_frames():
    groups = [(n, np.asarray(v)[:0], s) for n, v, s in _groups()]
    assert len(su.dataframe_from_parquet_bytes(su.columns_into_parquet_bytes(groups))) == 0
    plain = pd.DataFrame({"a": [1.0, 2.0], "b": [3.0, 4.0]})
    assert su.dataframe_to_dict(plain) == plain.to_dict()
    pd.testing.assert_frame_equal(su.dataframe_from_dict(plain.to_dict()), plain)


@pytest.mark.skipif(not os.path.isdir("/root/reference/gordo"), reason="/root/reference is not on this box")
def test_codecs_match_the_reference_functions():
    script = "import sys; sys.path.insert(0, %r)\n" % ROOT + textwrap.dedent("""
        import importlib.util, json, sys, types
        import numpy as np, pandas as pd
        def stub(name, **kw):
            m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m; return m
        stub("flask", request=None, g=None, jsonify=None, make_response=None, Response=object)
        stub("werkzeug"); stub("werkzeug.exceptions", NotFound=Exception, UnprocessableEntity=Exception, InternalServerError=Exception)
        g = stub("gordo"); g.__path__ = []; g.serializer = None
        stub("gordo.serializer")
        srv = stub("gordo.server"); srv.__path__ = ["/root/reference/gordo/server"]
        stub("gordo.server.properties", get_tags=None, get_target_tags=None)
        spec = importlib.util.spec_from_file_location("gordo.server.utils", "/root/reference/gordo/server/utils.py")
        ref = importlib.util.module_from_spec(spec); sys.modules["gordo.server.utils"] = ref; spec.loader.exec_module(ref)
        from gordo_b200.server import utils as su
        from gordo_b200.machine.model import utils as mu
        from tests.test_server_utils_cpu import _groups
        for index, freq in ((pd.date_range("2020-03-01", periods=40, freq="10min", tz="UTC"), pd.Timedelta("10min")), (None, None)):
            groups = _groups()
            df = mu.assemble_frame(groups, index, freq)
            # the reference's codec on the frame vs ours on the frame and on the raw column groups
            assert ref.dataframe_to_dict(df) == su.dataframe_to_dict(df) == su.columns_to_dict(groups, index, freq)
            want = ref.dataframe_from_parquet_bytes(ref.dataframe_into_parquet_bytes(df))
            pd.testing.assert_frame_equal(want, ref.dataframe_from_parquet_bytes(su.dataframe_into_parquet_bytes(df)))
            pd.testing.assert_frame_equal(want, ref.dataframe_from_parquet_bytes(su.columns_into_parquet_bytes(groups, index, freq)))
            pd.testing.assert_frame_equal(want, su.dataframe_from_parquet_bytes(ref.dataframe_into_parquet_bytes(df)))
            d = json.loads(json.dumps(ref.dataframe_to_dict(df), default=str))
            pd.testing.assert_frame_equal(ref.dataframe_from_dict(d), su.dataframe_from_dict(d))
        print("OK")
    """)
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + "\n" + r.stderr
