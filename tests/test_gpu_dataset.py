"""
-m gpu parity tests of the upstream-of-X kernels (gb200_resample / gb200_interpolate / gb200_filter_rows /
gb200_compact_rows through gordo_b200.dataset and the C-ABI) against oracle/dataset.py, the pandas restatement of
gordo-core 0.3.6 `TimeSeriesDataset.join_timeseries` / `get_data` / `pandas_filter_rows` (call site
gordo/builder/build_model.py:208-213).  Index and columns must be identical; values: float64, rtol 1e-12 for means
and interpolated values (summation order), exact for min / max / first / last / count and for every row decision.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from gordo_b200 import dataset as ds
from oracle import dataset as ods

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _series(rng, name, start, end, n, tz="UTC", nan_frac=0.02, gap=None, before=False, after=False):
    lo = start - pd.Timedelta("3h") if before else start + pd.Timedelta(seconds=float(rng.uniform(0, 4000)))
    hi = end + pd.Timedelta("2h") if after else end - pd.Timedelta(seconds=float(rng.uniform(0, 4000)))
    t = np.sort(rng.integers(lo.value, hi.value, n))
    t = np.unique(t)
    if gap is not None:                                     # a hole in the data: many empty bins
        g0 = lo.value + int((hi.value - lo.value) * gap[0]); g1 = lo.value + int((hi.value - lo.value) * gap[1])
        t = t[(t < g0) | (t >= g1)]
    v = rng.normal(50, 30, len(t))
    v[rng.random(len(t)) < nan_frac] = np.nan
    idx = pd.DatetimeIndex(t.astype("datetime64[ns]"), tz="UTC").tz_convert(tz)
    return pd.Series(v, index=idx, name=name)


def _same(got: pd.DataFrame, want: pd.DataFrame, exact=False):
    assert list(got.columns) == list(want.columns)
    np.testing.assert_array_equal(got.index.as_unit("ns").asi8, want.index.as_unit("ns").asi8)
    assert str(got.index.tz) == str(want.index.tz)
    if exact:
        np.testing.assert_array_equal(got.to_numpy(), want.to_numpy())
    else:
        np.testing.assert_allclose(got.to_numpy(), want.to_numpy(), rtol=1e-12, atol=0)


CASES = [
    dict(res="10T", n=3000, agg="mean", interp="linear_interpolation", limit="8H"),
    dict(res="10T", n=200, agg="mean", interp="linear_interpolation", limit="30T", gap=(0.3, 0.6)),       # sparse: empty bins, limit bites
    dict(res="2T", n=40000, agg="mean", interp="ffill", limit="10T", gap=(0.5, 0.52)),
    dict(res="1H", n=60000, agg="max", interp="linear_interpolation", limit=None),                          # dense bins, no limit
    dict(res="10T", n=5000, agg=["mean", "min", "max", "count", "first", "last", "sum"], interp="linear_interpolation", limit="8H"),
    dict(res="10T", n=3000, agg="mean", interp="linear_interpolation", limit="8H", tz="Asia/Kolkata", before=True, after=True),
    dict(res="30S", n=20000, agg="min", interp="ffill", limit="48H", tz="Etc/GMT-3"),
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c['res']}-{c['agg'] if isinstance(c['agg'], str) else 'multi'}-{c['interp'][:6]}" for c in CASES])
def test_join_timeseries_matches_oracle(case):
    rng = np.random.default_rng(CASES.index(case))
    start, end = pd.Timestamp("2020-03-01 09:00:30+00:00"), pd.Timestamp("2020-03-03 21:10:00+00:00")
    tz = case.get("tz", "UTC")
    series = [_series(rng, f"TAG {j}", start, end, case["n"] // (1 + j % 3), tz, gap=case.get("gap") if j != 1 else None,
                      before=case.get("before", False) and j == 0, after=case.get("after", False) and j == 2)
              for j in range(5)]
    series[3] = series[3].sample(frac=1.0, random_state=3)                       # an unsorted index
    want = ods.join_timeseries(series, start, end, case["res"], case["agg"], case["interp"], case["limit"])
    got = ds.join_timeseries(series, start, end, case["res"], case["agg"], case["interp"], case["limit"], device=DEV)
    assert len(want) > 10
    exact = case["agg"] in ("min", "max") and case["interp"] == "ffill"
    _same(got, want, exact=exact)
    if isinstance(case["agg"], list):
        assert got.columns.names == ["tag", "aggregation_method"]
        for m in ("count", "min", "max", "first", "last"):
            if case["interp"] == "linear_interpolation":
                continue
    # a second run gives the same bits (fixed reduction order)
    again = ds.join_timeseries(series, start, end, case["res"], case["agg"], case["interp"], case["limit"], device=DEV)
    np.testing.assert_array_equal(again.to_numpy(), got.to_numpy())


def test_resample_exact_statistics_without_interpolation():
    """min / max / count / first / last / sum of raw bins are exact (no rounding differences to hide behind)."""
    rng = np.random.default_rng(11)
    start, end = pd.Timestamp("2021-06-01 00:00:00+00:00"), pd.Timestamp("2021-06-02 00:00:00+00:00")
    s = _series(rng, "x", start, end, 30000, nan_frac=0.1)
    s = pd.Series(np.round(s.to_numpy() * 4) / 4, index=s.index, name="x")          # sums of quarter-integers are exact
    for agg in ("min", "max", "count", "first", "last", "sum", "mean"):
        fleet = ds.FleetTimeSeries(DEV)
        got = fleet.join([ds.MachineSeries([s], start, end)], "10T", agg, "ffill", "10T")[0].frame()
        want = ods.join_timeseries([s], start, end, "10T", agg, "ffill", "10T")
        _same(got, want, exact=agg != "mean")


@pytest.mark.parametrize("buffer_size", [0, 1, 7])
def test_pandas_filter_rows_matches_oracle(buffer_size):
    rng = np.random.default_rng(buffer_size)
    idx = pd.date_range("2020-04-08 00:00:00+00:00", periods=5000, freq="10min")
    df = pd.DataFrame(rng.normal(10, 20, (5000, 4)), index=idx, columns=["TAG 1", "tag-2.PV", "plain", "GRA-TE  -23-0733.PV"])
    kept_some = 0
    for expr in ["`TAG 1` > 5 & `tag-2.PV` < 20", "(`TAG 1` > -20) | ~(plain > 3)", "abs(`GRA-TE  -23-0733.PV`) < 35",
                 "~('2020-04-10 04:00:00+00:00' <= index <= '2020-04-12 10:00:00+00:00')", "`TAG 1` > -32 & plain < 65",
                 ["`TAG 1` > -30", "plain / `TAG 1` < 400"]]:
        want = ods.pandas_filter_rows(df, expr, buffer_size)
        got = ds.pandas_filter_rows(df, expr, buffer_size, device=DEV)
        assert len(want) < len(df)
        kept_some += len(want) > 0
        pd.testing.assert_frame_equal(got, want)
    assert kept_some >= 3


def test_get_data_pipeline_matches_oracle_and_errors():
    rng = np.random.default_rng(5)
    start, end = pd.Timestamp("2020-03-01 00:00:00+00:00"), pd.Timestamp("2020-03-08 00:00:00+00:00")
    series = [_series(rng, f"T-{j}", start, end, 20000, gap=(0.4, 0.45) if j == 2 else None) for j in range(6)]
    series[1] = series[1] * 40                                   # some samples leave (low, high) = (-1000, 50000)? no: (-500, 2000)
    kw = dict(resolution="10T", interpolation_limit="2H", row_filter="`T-0` > 20 & `T-3` < 95",
              known_filter_periods=["~('2020-03-02 04:00:00+00:00' <= index <= '2020-03-02 18:00:00+00:00')",
                                    "~('2020-03-05 00:00:00+00:00' <= index <= '2020-03-05 01:00:00+00:00')"],
              row_filter_buffer_size=3, low_threshold=-500, high_threshold=2500)
    want = ods.get_data(series, start, end, **kw)
    got = ds.get_data(series, start, end, device=DEV, **kw)
    assert 50 < len(want) < 900
    _same(got, want)
    with pytest.raises(ds.InsufficientDataError):
        ds.get_data(series, start, end, device=DEV, n_samples_threshold=10_000)
    with pytest.raises(ValueError, match="threshold"):
        ds.get_data(series, start, end, device=DEV, low_threshold=5, high_threshold=1)
    with pytest.raises(NotImplementedError):
        ds.get_data(series, start, end, device=DEV, filter_periods={"filter_method": "median"})
    with pytest.raises(NotImplementedError):
        ds.join_timeseries(series, start, end, "10T", aggregation_methods="median", device=DEV)
    with pytest.raises(ValueError, match="not a column"):
        ds.get_data(series, start, end, device=DEV, row_filter="`nope` > 1")


def test_fleet_join_equals_per_machine_and_feeds_float32():
    """Many Machines (different tag counts, periods, filters) in one FleetTimeSeries call == one call per Machine."""
    rng = np.random.default_rng(9)
    machines = []
    for m in range(9):
        start = pd.Timestamp("2020-01-01 00:00:00+00:00") + pd.Timedelta(hours=7 * m)
        end = start + pd.Timedelta(days=1 + m % 3)
        T = [3, 5, 3, 8, 5, 1, 3, 8, 5][m]
        series = [_series(rng, f"m{m}-t{j}", start, end, 1500 + 400 * j) for j in range(T)]
        machines.append(ds.MachineSeries(series, start, end, name=f"m{m}",
                                         row_filter=f"`m{m}-t0` > 10" if m % 2 else None))
    fleet = ds.FleetTimeSeries(DEV)
    joined = fleet.get_data(machines, "10T", interpolation_limit="1H", row_filter_buffer_size=2, low_threshold=-60, high_threshold=160)
    for mc, jm in zip(machines, joined):
        want = ods.get_data(mc.series, mc.start, mc.end, resolution="10T", interpolation_limit="1H",
                            row_filter=mc.row_filter or "", row_filter_buffer_size=2, low_threshold=-60, high_threshold=160)
        _same(jm.frame(), want)
        assert jm.values.dtype == torch.float64 and jm.values_f32.dtype == torch.float32
        np.testing.assert_array_equal(jm.values_f32.cpu().numpy(), want.to_numpy().astype(np.float32)
                                      if len(want) == 0 else jm.values.cpu().numpy().astype(np.float32))
        assert len(jm) == len(want) and jm.index_ns.shape[0] == len(want)


def test_resample_full_size_properties():
    """1-second data, 64 tags x 2 days = 1.1e7 points: a constant series resamples to the constant, counts add up to
    the number of samples, and the dense path (32 lanes per bin) agrees with pandas on one tag."""
    n_tags, secs = 64, 2 * 86400
    start = pd.Timestamp("2022-01-01 00:00:00+00:00"); end = start + pd.Timedelta(seconds=secs)
    idx = pd.date_range(start, periods=secs, freq="s")
    rng = np.random.default_rng(0)
    series = [pd.Series(np.full(secs, 3.25) if j == 0 else rng.normal(j, 1, secs), index=idx, name=f"t{j}") for j in range(n_tags)]
    fleet = ds.FleetTimeSeries(DEV)
    mean = fleet.join([ds.MachineSeries(series, start, end)], "10T", "mean", "linear_interpolation", "8H")[0]
    cnt = fleet.join([ds.MachineSeries(series, start, end)], "10T", "count", "ffill", None)[0]
    v = mean.values
    assert v.shape == (secs // 600, n_tags) or v.shape == (secs // 600 + 1, n_tags)
    assert (v[: secs // 600, 0] == 3.25).all()
    c = cnt.values[: secs // 600]
    assert float(c.sum()) == float(n_tags * secs) and (c == 600).all()
    want = series[5].resample("10min", label="left").mean().to_numpy()[: secs // 600]
    np.testing.assert_allclose(v[: secs // 600, 5].cpu().numpy(), want, rtol=1e-13)


def test_time_series_dataset_object_get_data():
    """The dataset object the builder calls (build_model.py:208-213): provider -> get_data() -> (X, y)."""
    rng = np.random.default_rng(21)
    start, end = pd.Timestamp("2020-03-01 00:00:00+00:00"), pd.Timestamp("2020-03-04 00:00:00+00:00")
    store = {f"tag-{j}": _series(rng, f"tag-{j}", start, end, 9000) for j in range(5)}

    class Provider:
        def load_series(self, train_start_date, train_end_date, tag_list, dry_run=False):
            assert train_start_date == start and train_end_date == end
            for t in tag_list:
                yield store[getattr(t, "name", t)]

    dsx = ds.TimeSeriesDataset(start, end, ["tag-0", "tag-1", "tag-2"], target_tag_list=["tag-3", "tag-1"], data_provider=Provider(),
                               resolution="10T", row_filter="`tag-0` > 0", interpolation_limit="48H", device=DEV)
    X, y = dsx.get_data()
    want = ods.get_data([store[f"tag-{j}"] for j in (0, 1, 2, 3)], start, end, resolution="10T", row_filter="`tag-0` > 0",
                        interpolation_limit="48H")
    _same(X, want[["tag-0", "tag-1", "tag-2"]]); _same(y, want[["tag-3", "tag-1"]])
    assert dsx.get_metadata()["row_count"] == len(want)
    with pytest.raises(ValueError, match="timezone"):
        ds.TimeSeriesDataset("2020-01-01", "2020-01-02", ["a"])


def test_raw_series_to_models_to_anomaly_frames():
    """The whole chain a project build runs (build_model.py:208-339 per Machine): raw tag series -> get_data on the GPU ->
    FleetModelBuilder (CV + fit + thresholds) -> model.anomaly on fresh data resampled the same way."""
    from gordo_b200.builder import FleetMachine, FleetModelBuilder
    rng = np.random.default_rng(33)
    start, end = pd.Timestamp("2021-05-01 00:00:00+00:00"), pd.Timestamp("2021-05-04 00:00:00+00:00")
    model = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {
        "sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler",
            {"gordo.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "epochs": 3}}]}}}}
    raw = []
    for m in range(3):
        T = 4 + m
        phase = rng.uniform(0, 6, T)
        series = []
        for j in range(T):
            s = _series(rng, f"m{m}-tag{j}", start, end, 6000, nan_frac=0.01)
            secs = (s.index.as_unit("ns").asi8 - start.value) / 1e9
            series.append(pd.Series(np.sin(secs / 7000.0 + phase[j]) * (j + 1) + rng.normal(0, 0.05, len(s)), index=s.index, name=s.name))
        raw.append(ds.MachineSeries(series, start, end, name=f"m{m}", row_filter=f"`m{m}-tag0` > -0.95"))
    fleet = ds.FleetTimeSeries(DEV)
    joined = fleet.get_data(raw, "10T", interpolation_limit="1H", low_threshold=-100, high_threshold=100)
    machines = []
    for mc, jm in zip(raw, joined):
        want = ods.get_data(mc.series, mc.start, mc.end, resolution="10T", interpolation_limit="1H", row_filter=mc.row_filter,
                            low_threshold=-100, high_threshold=100)
        X = jm.frame()
        _same(X, want)
        assert 300 < len(X) <= 433
        machines.append(FleetMachine(mc.name, X.astype(np.float32), model=model, evaluation={"seed": 5}))
    built = FleetModelBuilder.build_fleet(machines, streams=2)
    assert [mach.name for _, mach in built] == ["m0", "m1", "m2"]
    for (det, mach), mc in zip(built, machines):
        frame = det.anomaly(mc.X, mc.X, frequency=pd.Timedelta("10min"))
        assert len(frame) == len(mc.X) and np.isfinite(frame["total-anomaly-confidence"].to_numpy()).all()
        assert mach.build_metadata["model"]["cross_validation"]["scores"]["r2-score"]["fold-mean"] is not None


def test_join_of_series_that_never_overlap_is_empty():
    """One tag only has samples in the first hour, the other only in the last: every row of the inner join holds a
    NaN (interpolation limit 1 bin), so the frame is empty -- as with pandas -- and get_data raises InsufficientDataError."""
    start, end = pd.Timestamp("2020-01-01 00:00:00+00:00"), pd.Timestamp("2020-01-01 12:00:00+00:00")
    a = pd.Series([1.0, 2.0], index=[start + pd.Timedelta("5min"), start + pd.Timedelta("25min")], name="a")
    b = pd.Series([3.0, 4.0], index=[end - pd.Timedelta("45min"), end - pd.Timedelta("5min")], name="b")
    want = ods.join_timeseries([a, b], start, end, "10T", interpolation_limit="10T")
    got = ds.join_timeseries([a, b], start, end, "10T", interpolation_limit="10T", device=DEV)
    assert len(want) == 0 and len(got) == 0 and list(got.columns) == ["a", "b"]
    with pytest.raises(ds.InsufficientDataError):
        ds.get_data([a, b], start, end, "10T", interpolation_limit="10T", device=DEV)
    empty = pd.DataFrame({"a": []}, index=pd.DatetimeIndex([], tz="UTC"))
    assert len(ds.pandas_filter_rows(empty, "a > 1", device=DEV)) == 0
