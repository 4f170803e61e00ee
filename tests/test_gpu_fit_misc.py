"""
-m gpu parity tests of MinMaxScaler.fit, the rolling-min-max thresholds and the training kernel,
through the C-ABI, against the oracle on the same seeded inputs (identical initial weights and
batch permutations are fed to both).
"""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import dense, factories
from oracle.anomaly import rolling_min_max
from oracle.scaler import MinMaxScaler

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _i64(a):
    return torch.tensor(np.asarray(a, np.int64), device=DEV)


def test_minmax_fit_matches_sklearn_semantics():
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(0)
    X = (rng.random((20000, 13)) * rng.uniform(0.1, 50, 13) + rng.uniform(-9, 9, 13)).astype(np.float32)
    X[:, 4] = 2.5                         # constant column -> scale 1 (sklearn _handle_zeros_in_scale)
    lo, hi = [0, 100, 9000, 19999, 5], [100, 9000, 20000, 20000, 5]
    scale, mn = FFFleet.minmax_fit(torch.from_numpy(X).to(DEV), _i64(lo), _i64(hi))
    for j, (a, b) in enumerate(zip(lo, hi)):
        if b > a:
            sc = MinMaxScaler().fit(X[a:b])
            np.testing.assert_allclose(scale[j].cpu().numpy(), sc.scale_.astype(np.float32), rtol=1e-6)
            np.testing.assert_allclose(mn[j].cpu().numpy(), sc.min_.astype(np.float32), rtol=1e-6, atol=1e-7)
    # wide matrix (more tags than threads per block)
    Xw = rng.random((300, 300)).astype(np.float32)
    scale, mn = FFFleet.minmax_fit(torch.from_numpy(Xw).to(DEV), _i64([0]), _i64([300]))
    sc = MinMaxScaler().fit(Xw)
    np.testing.assert_allclose(scale[0].cpu().numpy(), sc.scale_.astype(np.float32), rtol=1e-6)


@pytest.mark.parametrize("window", [6, 144])
def test_rolling_min_max_matches_pandas(window):
    from gordo_b200.fleet import FFFleet
    rng = np.random.default_rng(1)
    V = rng.random((30000, 7)).astype(np.float32)
    V[777, 3] = np.nan
    lo, hi = [0, 5000, 29990, 100], [5000, 29990, 30000, 100 + window - 1]
    out = FFFleet.rolling_min_max(torch.from_numpy(V).to(DEV), _i64(lo), _i64(hi), window).cpu().numpy()
    for j, (a, b) in enumerate(zip(lo, hi)):
        want = pd.DataFrame(V[a:b].astype(np.float64)).rolling(window).min().max().to_numpy()
        np.testing.assert_allclose(out[j], want.astype(np.float32), rtol=0, atol=0, equal_nan=True)
        np.testing.assert_allclose(rolling_min_max(V[a:b], window), want, equal_nan=True)
    v1 = FFFleet.rolling_min_max(torch.from_numpy(V[:, 0].copy()).to(DEV), _i64([0]), _i64([30000]), window)
    assert v1.shape == (1, 1)


@pytest.mark.parametrize("T,func,l1_mean,batch", [(10, "tanh", False, 32), (50, "tanh", False, 32),
                                                   (7, "relu", True, 16), (5, "sigmoid", False, 32),
                                                   (6, "tanh", False, 10), (9, "tanh", True, 33)])      # batches that are not multiples of 4
def test_ff_fit_matches_oracle(T, func, l1_mean, batch):
    from gordo_b200.fleet import FFFleet, FFTopology
    rng = np.random.default_rng(100 + T)
    spec = factories.feedforward_hourglass(T, func=func)
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    fl = FFFleet(topo, 1, DEV)
    n_jobs, epochs = 3, 2
    rows = [650, 321, 64]
    X = rng.random((sum(rows), T)).astype(np.float32) * 3 - 1
    lo = np.concatenate([[0], np.cumsum(rows)[:-1]]); hi = np.cumsum(rows)
    inits, perms, scalers = [], [], []
    for j in range(n_jobs):
        inits.append(dense.ff_flatten(dense.ff_init(spec, rng)))
        perms.append([rng.permutation(rows[j]) for _ in range(epochs)])
        scalers.append(MinMaxScaler().fit(X[lo[j]:hi[j]]))
    # ---- oracle
    want_params, want_hist = [], []
    for j in range(n_jobs):
        p = dense.ff_unflatten(inits[j], spec["widths"])
        xs = scalers[j].transform(X[lo[j]:hi[j]]).astype(np.float32)
        h, _ = dense.ff_fit(spec, p, xs, X[lo[j]:hi[j]], epochs=epochs, batch_size=batch, perms=perms[j],
                            l1_mode="mean" if l1_mean else "sum")
        want_params.append(dense.ff_flatten(p)); want_hist.append(h)
    # ---- GPU
    params = torch.from_numpy(np.stack(inits)).to(DEV)
    pool = np.concatenate([np.concatenate(p) for p in perms]).astype(np.int32)
    poff = np.concatenate([[0], np.cumsum([epochs * r for r in rows])[:-1]]).astype(np.int64)
    in_scale = torch.from_numpy(np.stack([s.scale_ for s in scalers]).astype(np.float32)).to(DEV)
    in_min = torch.from_numpy(np.stack([s.min_ for s in scalers]).astype(np.float32)).to(DEV)
    hl, ha, mv, t = fl.fit_jobs(torch.from_numpy(X).to(DEV), None, _i64(lo), _i64(hi), params,
                                in_scale=in_scale, in_min=in_min, epochs=epochs, batch_size=batch,
                                perm_pool=torch.from_numpy(pool).to(DEV), perm_off=_i64(poff), l1_mean=l1_mean)
    torch.cuda.synchronize()
    steps = [epochs * -(-r // batch) for r in rows]
    assert t.cpu().tolist() == steps
    for j in range(n_jobs):
        got = params[j].cpu().numpy()
        # fp32 summation-order differences accumulate over the Adam steps: a few 1e-5 after ~40 steps
        np.testing.assert_allclose(got, want_params[j], rtol=0, atol=3e-4, err_msg=f"job {j}")
        np.testing.assert_allclose(hl[j].cpu().numpy(), want_hist[j]["loss"], rtol=2e-4)
        np.testing.assert_allclose(ha[j].cpu().numpy(), want_hist[j]["accuracy"], atol=2.0 / rows[j])
    # training moved the weights and reduced the loss
    assert float((params.cpu() - torch.from_numpy(np.stack(inits))).abs().max()) > 1e-3
    assert float(hl[0, -1]) < float(hl[0, 0])


def test_ff_fit_continues_from_adam_state():
    """A second fit on the same model continues from the trained weights AND optimizer state (models.py:282)."""
    from gordo_b200.fleet import FFFleet, FFTopology
    rng = np.random.default_rng(5)
    spec = factories.feedforward_hourglass(6)
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    fl = FFFleet(topo, 1, DEV)
    X = rng.random((256, 6)).astype(np.float32)
    init = dense.ff_flatten(dense.ff_init(spec, rng))
    Xd = torch.from_numpy(X).to(DEV)
    lo, hi = _i64([0]), _i64([256])
    p2 = torch.from_numpy(init[None].copy()).to(DEV)
    _, _, mv, t = fl.fit_jobs(Xd, None, lo, hi, p2, epochs=1)
    fl.fit_jobs(Xd, None, lo, hi, p2, epochs=1, adam_mv=mv, adam_t=t)
    p1 = torch.from_numpy(init[None].copy()).to(DEV)
    fl.fit_jobs(Xd, None, lo, hi, p1, epochs=2)
    torch.cuda.synchronize()
    np.testing.assert_allclose(p2.cpu().numpy(), p1.cpu().numpy(), atol=1e-6)
    assert int(t[0]) == 16
