"""
CPU tests of the host side of gordo_b200.dataset (SURVEY.md §8 f-4 upstream; reference call site
gordo/builder/build_model.py:208-213 -> gordo-core 0.3.6 TimeSeriesDataset.get_data): the row-filter compiler
against pandas' own ``DataFrame.eval`` (a Python restatement of the device interpreter runs the compiled program),
the frequency / limit helpers, hand-checked facts about the oracle's ``join_timeseries``, and the no-GPU behaviour.
"""
import numpy as np
import pandas as pd
import pytest

from gordo_b200 import dataset as ds
from oracle import dataset as ods


def run_program(prog: ds.RowProgram, data: np.ndarray, ts_rel: np.ndarray) -> np.ndarray:
    """The stack machine of gb200_filter_rows (csrc/dataset.cu filter_rows_kernel), one row at a time."""
    O = ds.OPS
    out = np.zeros(len(data), bool)
    with np.errstate(all="ignore"):
        for r, row in enumerate(data):
            st = []
            for op, arg in zip(prog.ops, prog.args):
                if op == O["CONST"]: st.append(prog.consts[arg])
                elif op == O["COL"]: st.append(float(row[arg]))
                elif op == O["INDEX"]: st.append(float(ts_rel[r]))
                elif op == O["NEG"]: st[-1] = -st[-1]
                elif op == O["ABS"]: st[-1] = abs(st[-1])
                elif op == O["NOT"]: st[-1] = 0.0 if st[-1] != 0.0 else 1.0
                elif op == O["ALL_FINITE"]: st.append(float(np.isfinite(row).all()))
                elif op == O["ALL_NOTNAN"]: st.append(float((~np.isnan(row)).all()))
                elif op == O["ALL_BETWEEN"]: st.append(float(((row > prog.consts[arg]) & (row < prog.consts[arg + 1])).all()))
                else:
                    y = np.float64(st.pop()); x = np.float64(st[-1])
                    st[-1] = float({O["ADD"]: lambda: x + y, O["SUB"]: lambda: x - y, O["MUL"]: lambda: x * y,
                                    O["DIV"]: lambda: x / y, O["POW"]: lambda: x ** y, O["GT"]: lambda: x > y,
                                    O["GE"]: lambda: x >= y, O["LT"]: lambda: x < y, O["LE"]: lambda: x <= y,
                                    O["EQ"]: lambda: x == y, O["NE"]: lambda: x != y,
                                    O["AND"]: lambda: (x != 0) and (y != 0), O["OR"]: lambda: (x != 0) or (y != 0)}[op]())
            assert len(st) == 1
            out[r] = st[0] != 0.0
    return out


def _frame(n=400, seed=0):
    rng = np.random.default_rng(seed)
    idx = pd.date_range("2020-04-08 00:00:00+00:00", periods=n, freq="10min")
    df = pd.DataFrame(rng.normal(10, 20, (n, 4)), index=idx, columns=["TAG 1", "tag-2.PV", "plain", "GRA-TE  -23-0733.PV"])
    df.iloc[rng.integers(0, n, 15), 1] = np.nan
    df.iloc[rng.integers(0, n, 5), 2] = 0.0
    return df


EXPRESSIONS = [
    "`TAG 1` > 5",
    "`TAG 1` > 5 & `tag-2.PV` < 20",
    "`TAG 1` > 5 | plain <= 0",
    "(`TAG 1` > 0) & (`tag-2.PV` > -10) | ~(plain > 3)",
    "plain * 2 + 1 >= `TAG 1` - 3",
    "-plain < 4 and not `TAG 1` > 30",
    "abs(`GRA-TE  -23-0733.PV`) < 15",
    "0 < `TAG 1` < 25",
    "`TAG 1` / plain > 1",
    "`tag-2.PV` == `tag-2.PV`",
    "`tag-2.PV` != `tag-2.PV`",
    "plain ** 2 > 100",
    "~('2020-04-08 04:00:00+00:00' <= index <= '2020-04-08 10:00:00+00:00')",
    "index > '2020-04-09 00:00:00+00:00' & plain > 0",
    ["`TAG 1` > 0", "plain < 50", "~('2020-04-08 04:00:00+00:00' <= index <= '2020-04-08 10:00:00+00:00')"],
]


@pytest.mark.parametrize("expr", EXPRESSIONS, ids=[str(e)[:40] for e in EXPRESSIONS])
def test_compiled_row_filter_equals_pandas_eval(expr):
    df = _frame()
    base = int(df.index.as_unit("ns").asi8.min())
    prog = ds.compile_row_filter(expr, list(df.columns), base, df.index.tz)
    got = run_program(prog, df.to_numpy(np.float64), (df.index.as_unit("ns").asi8 - base).astype(np.float64))
    text = expr if isinstance(expr, str) else " & ".join(f"({e})" for e in expr)
    want = np.asarray(df.eval(text), bool)
    np.testing.assert_array_equal(got, want)
    assert 0 < want.sum() < len(df) or "!=" in text or "==" in text       # the case discriminates


def test_row_filter_errors_and_limits():
    cols = ["a", "b"]
    with pytest.raises(ValueError, match="not a column"):
        ds.compile_row_filter("c > 1", cols)
    with pytest.raises(ValueError, match="not a valid expression"):
        ds.compile_row_filter("a > > 1", cols)
    with pytest.raises(NotImplementedError):
        ds.compile_row_filter("sin(a) > 0", cols)
    with pytest.raises(ValueError, match="too long"):
        ds.compile_row_filter(" & ".join(f"(a > {i})" for i in range(60)), cols)
    p = ds.compile_row_filter("a > 1 & a > 1", cols)
    assert p.consts == (1.0,)                                     # constants are shared
    assert ds.RowProgram.all_between(-1000, 50000).consts == (-1000.0, 50000.0)


def test_frequency_and_limit_helpers():
    assert ds.normalize_freq("10T") == "10min" and ds.normalize_freq("8H") == "8h" and ds.normalize_freq("30S") == "30s"
    assert ds.normalize_freq("10min") == "10min" and ds.normalize_freq("1D") == "1D"
    assert ds.interpolation_limit_bins("8H", "10T") == 48 and ds.interpolation_limit_bins("48H", "10T") == 288
    assert ds.interpolation_limit_bins(None, "10T") is None
    with pytest.raises(ValueError, match="larger than given resolution"):
        ds.interpolation_limit_bins("1T", "10T")
    assert ods.interpolation_limit_bins("8H", "10T") == 48 and ods.normalize_freq("2T") == "2min"


def test_oracle_join_timeseries_hand_checked():
    """Facts worked out by hand: bins are labelled by their left edge and anchored at midnight; a NaN sample pads a
    series to the resampling start / end; interpolation is forward-only with a bin limit; rows with a NaN are dropped."""
    t = lambda s: pd.Timestamp(f"2020-01-01 {s}+00:00")
    a = pd.Series([1.0, 3.0, 10.0, 20.0], index=[t("00:01"), t("00:09"), t("00:41"), t("00:59")], name="a")
    b = pd.Series([5.0, 7.0], index=[t("00:12"), t("00:55")], name="b")
    got = ods.join_timeseries([a, b], t("00:00"), t("01:00"), "10T", interpolation_limit="20T")
    # a: bins 00:00 -> mean(1,3)=2, 00:10..00:30 empty, 00:40 -> 10, 00:50 -> 20, 01:00 pad (NaN -> carried 20)
    #    linear with limit 2: 00:10 -> 4, 00:20 -> 6, 00:30 stays NaN
    # b: 00:00 NaN (leading, never filled), 00:10 -> 5, 00:20 / 00:30 -> 5.5 / 6.0 (limit 2), 00:40 NaN, 00:50 -> 7, 01:00 -> 7
    want = pd.DataFrame({"a": [4.0, 6.0, 20.0, 20.0], "b": [5.0, 5.5, 7.0, 7.0]},
                        index=pd.DatetimeIndex([t("00:10"), t("00:20"), t("00:50"), t("01:00")]))
    pd.testing.assert_frame_equal(got, want, check_freq=False, check_names=False)
    kept = ods.pandas_filter_rows(want, "a > 5", buffer_size=1)
    assert list(kept.index) == [t("00:50"), t("01:00")]           # 00:10 rejected, its neighbour 00:20 with it
    assert list(ods.apply_buffer(np.array([1, 1, 0, 1, 1, 1], bool), 1)) == [True, False, False, False, True, True]


def test_dataset_needs_the_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = pd.Series([1.0], index=[pd.Timestamp("2020-01-01 00:00:00+00:00")], name="a")
    with pytest.raises(RuntimeError, match="CUDA"):
        ds.join_timeseries([s], pd.Timestamp("2020-01-01 00:00:00+00:00"), pd.Timestamp("2020-01-01 01:00:00+00:00"), "10T")
    with pytest.raises(RuntimeError, match="CUDA"):
        ds.pandas_filter_rows(pd.DataFrame({"a": [1.0]}), "a > 0")
