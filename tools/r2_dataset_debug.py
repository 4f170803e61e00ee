import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from tests.test_gpu_dataset import _series, CASES
from oracle import dataset as ods
from gordo_b200 import dataset as ds, _native as N
case = CASES[0]
rng = np.random.default_rng(0)
start, end = pd.Timestamp("2020-03-01 09:00:30+00:00"), pd.Timestamp("2020-03-03 21:10:00+00:00")
series = [_series(rng, f"TAG {j}", start, end, case["n"] // (1 + j % 3), "UTC") for j in range(5)]
want = ods.join_timeseries(series, start, end, "10T")
fleet = ds.FleetTimeSeries("cuda:0")
orig_apply = fleet._apply
def spy(state, prog, buf, jobs, ts_base=0):
    d = state["data"]
    print("before stage: rows", d.shape, "rows with NaN", int(torch.isnan(d).any(1).sum()), "first 8 rows NaN per col", torch.isnan(d[:8]).int().tolist())
    orig_apply(state, prog, buf, jobs, ts_base)
    print("after stage: rows", state["data"].shape, "lo/hi", state["lo_host"], state["hi_host"])
fleet._apply = spy
jm = fleet.join([ds.MachineSeries(series, start, end)], "10T")[0]
got = jm.frame()
print("want", len(want), want.index[0], "got", len(got), got.index[0])
# direct mask check
d = torch.tensor([[1.0, float("nan")], [2.0, 3.0], [float("nan"), float("nan")], [4.0, 5.0]], dtype=torch.float64, device="cuda:0")
lo = torch.zeros(1, dtype=torch.int64, device="cuda:0"); hi = torch.full((1,), 4, dtype=torch.int64, device="cuda:0")
keep = torch.full((4,), 7, dtype=torch.uint8, device="cuda:0")
prog = ds.RowProgram.all_notnan()
ops = (N.C.c_int32 * 1)(*prog.ops); args = (N.C.c_int32 * 1)(*prog.args); consts = (N.C.c_double * 1)()
rc = N.lib().gb200_filter_rows(1, N.ptr(lo), N.ptr(hi), N.ptr(d), 2, None, 0, ops, args, 1, consts, 0, 0, N.ptr(keep), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("rc", rc, N.lib().gb200_last_error(), "keep", keep.tolist())
