"""
-m gpu parity tests of the fused scorer, through the C-ABI (gordo_b200.fleet -> libgordo_b200.so),
against the oracle on the same seeded inputs.  Tolerances:
  GB200_PREC_F32     : |yhat - oracle| <= 2e-5 abs on O(1) outputs (fp32 FMA order vs BLAS); the
                       derived columns inherit that error times their column factor (|err_scale|,
                       1/threshold) plus 2e-4 rel: float32 device arithmetic vs the reference's
                       float64 pandas arithmetic on the same float32 yhat.
  GB200_PREC_BF16_TC : vs a bf16-operand emulation of the same stack: mean |err| <= 6e-4 and max
                       <= 1e-2 (tanh.approx + accumulation order flip an occasional bf16 rounding
                       of a hidden activation; a layout / descriptor error is O(1)); vs the fp32
                       oracle <= 3e-2 abs (bf16 operands through 7 layers).
"""
import numpy as np
import pytest
import torch

from tests.gpu_util import make_fleet_case, oracle_score, fleet_from_case, ff_forward_bf16

pytestmark = pytest.mark.gpu

CASES = {
    # BASELINE.json configs[0]: 1 Machine, feedforward hourglass, 10 tags x 1000 steps
    "c1_1x10x1000": dict(seed=1, row_counts=[1000], T=10),
    "ragged_small_T7": dict(seed=2, row_counts=[1, 127, 128, 129, 0, 1000, 5], T=7),
    "c2_shape_3x50": dict(seed=3, row_counts=[4096, 1003, 2500], T=50),
    "c5_shape_5tags": dict(seed=4, row_counts=[300] * 9, T=5),
    "relu_no_thresholds": dict(seed=5, row_counts=[700, 64], T=20, func="relu", thresholds=False),
    "sigmoid_T33_E2": dict(seed=6, row_counts=[513], T=33, func="sigmoid", encoding_layers=2),
    "separate_y_T12_to_4": dict(seed=7, row_counts=[400, 333], T=12, T_out=4),
    "unaligned_T3": dict(seed=8, row_counts=[131, 77, 9], T=3),
}


def _compare(res, case, forward, tol_out, rtol, atol):
    off = 0
    for m, n in enumerate(case["row_counts"]):
        want = oracle_score(case, m, forward)
        for key, w in want.items():
            got = res[key][off:off + n].double().cpu().numpy()
            es = float(np.abs(case["err_scale"][m]).max())
            ift = 1.0 / float(case["feat_thr"][m].min()) if case["feat_thr"] is not None else 1.0
            if key == "model-output":
                np.testing.assert_allclose(got, w, rtol=0, atol=tol_out, err_msg=f"machine {m} {key}")
            elif key.startswith("total-"):
                # mean of squares: error ~ 2*|d|*|delta| -> relative to the value, plus a floor
                np.testing.assert_allclose(got, w, rtol=20 * rtol, atol=atol, err_msg=f"machine {m} {key}")
            else:
                factor = {"tag-anomaly-unscaled": 1.0, "tag-anomaly-scaled": es, "anomaly-confidence": ift}[key]
                np.testing.assert_allclose(got, w, rtol=rtol, atol=2 * tol_out * factor + atol,
                                           err_msg=f"machine {m} {key}")
        off += n


@pytest.mark.parametrize("name", list(CASES))
def test_ff_score_f32_matches_oracle(name):
    case = make_fleet_case(**CASES[name])
    fl, sched, X, Y = fleet_from_case(case)
    res = fl.score(sched, X, Y, precision="f32")
    torch.cuda.synchronize()
    _compare(res, case, None, 2e-5, 2e-4, 2e-6)
    if case["feat_thr"] is None:
        assert "anomaly-confidence" not in res and "total-anomaly-confidence" not in res


@pytest.mark.parametrize("name", list(CASES))
def test_ff_score_tc_matches_bf16_emulation_and_oracle(name):
    case = make_fleet_case(**CASES[name])
    fl, sched, X, Y = fleet_from_case(case)
    assert fl.tc_eligible()
    res = fl.score(sched, X, Y, precision="bf16")
    torch.cuda.synchronize()
    # (a) tight: against the oracle run with bf16-rounded operands -- catches any layout / descriptor error
    off = 0
    for m, n in enumerate(case["row_counts"]):
        want = oracle_score(case, m, ff_forward_bf16)["model-output"]
        got = res["model-output"][off:off + n].cpu().numpy()
        scale = max(1.0, float(np.abs(want).max(initial=0)))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-2 * scale, err_msg=f"machine {m}")
        if n:
            assert float(np.abs(got - want).mean()) <= 6e-4 * scale, f"machine {m}"
        off += n
    # (b) loose: against the fp32 oracle (the stated bf16 tolerance)
    off = 0
    for m, n in enumerate(case["row_counts"]):
        want = oracle_score(case, m)["model-output"]
        got = res["model-output"][off:off + n].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=3e-2 * max(1.0, float(np.abs(want).max(initial=0))))
        off += n
    # (c) the score columns must be exactly consistent with the kernel's own yhat
    mo = res["model-output"].double()
    y = (Y if Y is not None else X).double()
    d = (mo - y).abs()
    np.testing.assert_allclose(res["tag-anomaly-unscaled"].double().cpu().numpy(), d.cpu().numpy(), rtol=1e-6, atol=1e-7)
    es = torch.repeat_interleave(fl.err_scale.double().abs(), torch.tensor(case["row_counts"], device=mo.device), dim=0)
    np.testing.assert_allclose(res["tag-anomaly-scaled"].double().cpu().numpy(), (d * es).cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res["total-anomaly-unscaled"].double().cpu().numpy(), (d ** 2).mean(1).cpu().numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(res["total-anomaly-scaled"].double().cpu().numpy(), ((d * es) ** 2).mean(1).cpu().numpy(), rtol=1e-4, atol=1e-7)
    if fl.feat_thr is not None:
        ft = torch.repeat_interleave(fl.feat_thr.double(), torch.tensor(case["row_counts"], device=mo.device), dim=0)
        np.testing.assert_allclose(res["anomaly-confidence"].double().cpu().numpy(), (d / ft).cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_ff_score_wide_topology_weights_from_l2():
    # T=100 hourglass: 37 422 params -> the fp32 kernel streams weights through L1/L2
    case = make_fleet_case(seed=9, row_counts=[300, 200], T=100)
    fl, sched, X, Y = fleet_from_case(case)
    res = fl.score(sched, X, Y, precision="f32")
    _compare(res, case, None, 5e-5, 3e-4, 3e-6)


def test_predict_only_and_subranges():
    case = make_fleet_case(seed=10, row_counts=[1000, 600], T=10)
    fl, sched, X, Y = fleet_from_case(case)
    full = fl.predict(sched, X)
    from gordo_b200.fleet import Schedule
    # virtual Machines = sub-ranges (how CV test folds are scored): rows outside stay untouched
    sub = Schedule(rows_lo=[250, 1100], rows_hi=[500, 1600], rows_total=1600)
    out = {"model-output": torch.full((1600, 10), -7.0, device=X.device)}
    res = fl.score(sub, X, None, precision="f32", columns=("model-output",), out=out)
    got = res["model-output"]
    assert torch.equal(got[250:500], full[250:500]) and torch.equal(got[1100:1600], full[1100:1600])
    assert bool((got[:250] == -7).all()) and bool((got[500:1100] == -7).all())


def test_abi_rejects_bad_arguments():
    import ctypes as C
    from gordo_b200 import _native as N
    lib = N.lib()
    h = C.c_void_p()
    off = (C.c_int64 * 3)(0, 10, 5)
    assert lib.gb200_fleet_create(C.byref(h), 2, off) != 0 and b"non-decreasing" in lib.gb200_last_error()
    with pytest.raises(ValueError):
        N.make_ff_arch([4, 3, 4], ["tanh", "swish"])
    case = make_fleet_case(seed=11, row_counts=[10], T=4)
    fl, sched, X, Y = fleet_from_case(case)
    with pytest.raises(ValueError):
        fl.score(sched, X[:5].contiguous())
