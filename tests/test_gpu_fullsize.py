"""
-m gpu property tests at BASELINE.json's full sizes, where the oracle cannot follow: the fused
scorer on c2 (128 Machines x 50 tags x 100 000 rows) and a c5-shaped boundary stress (10 000 tiny
5-tag Machines).  Size-independent properties: determinism, column identities recomputed from the
kernel's own yhat, fp32 vs tensor-core agreement, Machine independence (a sub-fleet scored alone
gives bit-identical rows), untouched rows outside the schedule.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fleet(M, T, rows, seed, widths=None):
    from gordo_b200.fleet import FFFleet, FFTopology, Schedule
    from gordo_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
    topo = feedforward_hourglass(T)
    fl = FFFleet(topo, M, DEV)
    fl.init_params(seed)
    g = torch.Generator(device=DEV); g.manual_seed(seed + 1)
    X = torch.rand((M * rows, T), generator=g, device=DEV)
    sched = Schedule([rows] * M)
    lo = torch.arange(M, device=DEV, dtype=torch.int64) * rows
    fl.in_scale, fl.in_min = FFFleet.minmax_fit(X, lo, lo + rows)
    fl.err_scale = fl.in_scale.clone()
    fl.feat_thr = torch.rand((M, T), generator=g, device=DEV) * 0.4 + 0.1
    fl.agg_thr = torch.rand((M,), generator=g, device=DEV) * 0.09 + 0.01
    return fl, sched, X


def _check_identities(fl, res, X, rows, machines):
    T = X.shape[1]
    for m in machines:
        sl = slice(m * rows, (m + 1) * rows)
        mo = res["model-output"][sl].double(); y = X[sl].double()
        d = (mo - y).abs()
        es = fl.err_scale[m].double().abs(); ft = fl.feat_thr[m].double()
        assert torch.allclose(res["tag-anomaly-unscaled"][sl].double(), d, rtol=1e-6, atol=1e-7)
        assert torch.allclose(res["tag-anomaly-scaled"][sl].double(), d * es, rtol=1e-5, atol=1e-7)
        assert torch.allclose(res["total-anomaly-unscaled"][sl].double(), (d ** 2).mean(1), rtol=1e-4, atol=1e-8)
        assert torch.allclose(res["total-anomaly-scaled"][sl].double(), ((d * es) ** 2).mean(1), rtol=1e-4, atol=1e-8)
        assert torch.allclose(res["anomaly-confidence"][sl].double(), d / ft, rtol=1e-5, atol=1e-7)
        assert torch.allclose(res["total-anomaly-confidence"][sl].double(),
                              ((d * es) ** 2).mean(1) / fl.agg_thr[m].double(), rtol=1e-4, atol=1e-7)


def test_c2_full_size_properties():
    from gordo_b200.fleet import FFFleet, Schedule
    M, T, rows = 128, 50, 100_000
    fl, sched, X = _fleet(M, T, rows, seed=2)
    res = fl.score(sched, X, precision="bf16")
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(v).all()) for v in res.values())
    _check_identities(fl, res, X, rows, machines=[0, 1, 63, 127])
    # determinism: a second pass is bit-identical
    ref = {k: v[: 3 * rows].clone() for k, v in res.items()}
    chk = {k: v.clone() for k, v in res.items() if v.dim() == 1}
    res2 = fl.score(sched, X, precision="bf16", out=res)
    assert all(torch.equal(res2[k][: 3 * rows], ref[k]) for k in ref) and all(torch.equal(res2[k], chk[k]) for k in chk)
    # Machine independence: Machines 5..7 scored alone as their own fleet give identical rows
    sub = FFFleet(fl.topo, 3, DEV)
    sub.set_params(fl.params[5:8]); sub.in_scale = fl.in_scale[5:8].contiguous(); sub.in_min = fl.in_min[5:8].contiguous()
    sub.err_scale = fl.err_scale[5:8].contiguous(); sub.feat_thr = fl.feat_thr[5:8].contiguous(); sub.agg_thr = fl.agg_thr[5:8].contiguous()
    r3 = sub.score(Schedule([rows] * 3), X[5 * rows: 8 * rows].contiguous(), precision="bf16")
    for k in r3:
        assert torch.equal(r3[k], res2[k][5 * rows: 8 * rows]), k
    # tensor-core path vs the exact fp32 kernel on a slice of Machines (stated bf16 tolerance)
    few = FFFleet(fl.topo, 2, DEV)
    few.set_params(fl.params[:2]); few.in_scale = fl.in_scale[:2].contiguous(); few.in_min = fl.in_min[:2].contiguous()
    few.err_scale = fl.err_scale[:2].contiguous()
    f32 = few.predict(Schedule([rows] * 2), X[: 2 * rows].contiguous(), precision="f32")
    assert float((f32 - res2["model-output"][: 2 * rows]).abs().max()) < 3e-2


def test_c5_shape_many_tiny_machines():
    """10 000 Machines x 5 tags (rows reduced to 2 000 each): every CTA crosses Machine boundaries
    constantly -- weights re-staged per Machine, ragged last tiles (2000 = 15*128 + 80)."""
    from gordo_b200.fleet import Schedule
    M, T, rows = 10_000, 5, 2_000
    fl, sched, X = _fleet(M, T, rows, seed=3)
    out = {"model-output": torch.full((M * rows, T), float("nan"), device=DEV)}
    res = fl.score(sched, X, precision="bf16", out=out)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(res["model-output"]).all())          # every row of every Machine was written
    _check_identities(fl, res, X, rows, machines=[0, 1, 4999, 9998, 9999])
    f32 = fl.score(sched, X, precision="f32", columns=("model-output",))["model-output"]
    # bf16 tolerance is relative to the output scale (tiny 5-4-4-3 glorot nets reach |yhat| ~ 2-3)
    assert float((f32 - res["model-output"]).abs().max()) < 3e-2 * max(1.0, float(f32.abs().max()))
    # thresholds + scaler kernels at fleet scale agree with torch reductions
    lo = torch.arange(M, device=DEV, dtype=torch.int64) * rows
    xr = X.view(M, rows, T)
    assert torch.allclose(fl.in_scale, 1.0 / (xr.max(1).values - xr.min(1).values), rtol=1e-5)
    thr = fl.rolling_min_max(res["total-anomaly-scaled"], lo, lo + rows, 6)[:, 0]
    ts = res["total-anomaly-scaled"].view(M, rows)
    want = ts.unfold(1, 6, 1).min(-1).values.max(1).values
    assert torch.equal(thr, want)


def test_c4_full_size_lstm_properties():
    """
    c4: KerasLSTMAutoEncoder 200 tags, lookback 128, hourglass 167-133-100-100-133-167, 100 000 rows
    (99 873 windows) on the tcgen05 pipeline.  Windows are independent sequences, so: a series cut at
    another offset gives the same windows bit for bit (different tile alignment), one Machine split into
    two overlapping Machines gives the same rows, and the exact fp32 kernel agrees on a slice within the
    stated bf16 tolerance.
    """
    from gordo_b200.fleet import Schedule
    from gordo_b200.lstm import LSTMFleet
    from gordo_b200.machine.model.factories.lstm_autoencoder import lstm_hourglass
    T, L, n = 200, 128, 100_000
    topo = lstm_hourglass(T, lookback_window=L)
    g = torch.Generator(device=DEV); g.manual_seed(5)
    X = torch.rand((n, T), generator=g, device=DEV)
    one = LSTMFleet(topo, 1, 0, DEV)
    one.set_params(topo.init_params(1, g, DEV))
    assert one.tc_eligible()
    out, off = one.predict(Schedule([n]), X, precision="bf16")
    assert int(off[-1]) == n - L + 1 and out.shape == (n - L + 1, T) and bool(torch.isfinite(out).all())
    # (a) the same series entered 131 rows later: window k of X[131:] is window k + 131 of X
    a = 131
    out_a, _ = one.predict(Schedule([n - a]), X[a:].contiguous(), precision="bf16")
    assert torch.equal(out_a, out[a:])
    # (b) one Machine as two overlapping Machines of the same model
    n1 = 40_007
    two = LSTMFleet(topo, 2, 0, DEV)
    two.set_params(one.params.repeat(2, 1))
    Xs = torch.cat([X[:n1], X[n1 - L + 1:]])
    out2, off2 = two.predict(Schedule([n1, n - (n1 - L + 1)]), Xs, precision="bf16")
    assert int(off2[-1]) == n - L + 1 and torch.equal(out2, out)
    # (c) exact fp32 kernel on the first 300 windows
    ref, _ = one.predict(Schedule([300 + L - 1]), X[:300 + L - 1].contiguous(), precision="f32")
    scale = ref.abs().max().item()
    assert (out[:300] - ref).abs().max().item() <= 3e-2 * max(1.0, scale)


def test_c2_full_size_fit_properties():
    """
    Training at the c2 row count (8 jobs x 100 000 rows x 50 tags, batch 32, 1 epoch = 3 125 dependent Adam
    steps per job): a job trained inside the fleet launch equals the same job trained alone bit for bit
    (jobs share nothing, whatever CTA order they get), a rerun reproduces itself, the loss is finite and
    lower than at the start.
    """
    from gordo_b200.builder import segmented_randperm
    J, rows, T = 8, 100_000, 50
    fl, sched, X = _fleet(J, T, rows, seed=21)
    dev = torch.device(DEV)
    g = torch.Generator(device=DEV); g.manual_seed(9)
    lo = torch.arange(J, device=DEV, dtype=torch.int64) * rows
    hi = lo + rows - torch.arange(J, device=DEV, dtype=torch.int64) * 3_001       # jobs of different length
    n_job = (hi - lo).cpu().numpy()
    pool = segmented_randperm(n_job, g, dev)
    poff = torch.as_tensor(np.concatenate([[0], np.cumsum(n_job)[:-1]]).astype(np.int64), device=DEV)
    P0 = fl.params.clone()

    def run(sel):
        p = P0[sel].clone()
        hl, _, _, _ = fl.fit_jobs(X, None, lo[sel].contiguous(), hi[sel].contiguous(), p,
                                  in_scale=fl.in_scale[sel].contiguous(), in_min=fl.in_min[sel].contiguous(),
                                  epochs=1, batch_size=32, perm_pool=pool, perm_off=poff[sel].contiguous())
        torch.cuda.synchronize()
        return p, hl

    sel_all = torch.arange(J, device=DEV)
    p_all, h_all = run(sel_all)
    p_again, h_again = run(sel_all)
    assert torch.equal(p_all, p_again) and torch.equal(h_all, h_again)
    p_one, h_one = run(torch.tensor([5], device=DEV))
    assert torch.equal(p_one[0], p_all[5]) and torch.equal(h_one[0], h_all[5])
    assert bool(torch.isfinite(p_all).all()) and bool(torch.isfinite(h_all).all())
    # the epoch-mean loss is far below the loss of the untrained net on the same rows
    res0 = fl.score(sched, X, precision="f32", columns=("model-output",))
    xs = X * fl.in_scale.repeat_interleave(rows, 0) + fl.in_min.repeat_interleave(rows, 0)
    start_mse = ((res0["model-output"] - xs) ** 2).mean().item()
    assert h_all.max().item() < 0.5 * start_mse
