"""
-m gpu: the tensor-core GEMMs of the LSTM training step (tcgen05 kind::tf32, 3xTF32 split) against the
CUDA-core GEMM on the same fits.  GB200_LSTM_GEMM forces a variant for every launch ("simt", "tc" =
tensor cores with 16-byte staging where the operands allow it, "tcs" = tensor cores with scalar
staging); topologies are chosen so that each operand layout (K-major / MN-major image, aligned / ragged)
is exercised.  Tolerance: the 3xTF32 product drops the lo*lo term (< 2^-21 relative per product).
"""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fit(variant, n_features, units, lookback, rows, jobs=3, batch=32, seed=0):
    from gordo_b200.lstm import LSTMFleet, LSTMTopology
    topo = LSTMTopology(n_features=n_features, n_features_out=n_features, lookback_window=lookback,
                        units=list(units), acts=["tanh"] * len(units), out_func="tanh")
    g = torch.Generator(device=DEV); g.manual_seed(seed)
    fl = LSTMFleet(topo, jobs, 0, DEV)
    P = topo.init_params(jobs, g, DEV)
    X = torch.rand((jobs * rows, n_features), generator=g, device=DEV)
    lo = np.arange(jobs, dtype=np.int64) * rows
    old = os.environ.get("GB200_LSTM_GEMM")
    os.environ["GB200_LSTM_GEMM"] = variant
    try:
        hl, pl = fl.fit_jobs(X, X, lo, lo + rows, P, epochs=1, batch_size=batch)
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["GB200_LSTM_GEMM"]
        else:
            os.environ["GB200_LSTM_GEMM"] = old
    return P.cpu().numpy(), hl.cpu().numpy()


@pytest.mark.parametrize("variant", ["tc", "tcs"])
@pytest.mark.parametrize("n_features,units,lookback,rows", [
    (6, (5, 5), 5, 70),          # only dz.W^T of layer 1 is aligned (both images K-major)
    (6, (8,), 5, 70),            # only the U-gradient GEMM is 16-byte aligned (MN-major images)
    (8, (6,), 5, 70),            # W-gradient (MN-major) and input projection (A K-major, B MN-major)
    (8, (8, 12, 8), 6, 100),     # everything aligned, dz.W^T (both K-major) for the upper layers
    (7, (9, 5), 4, 61),          # nothing aligned: scalar staging in both variants, ragged tiles, partial last batch
    (200, (167, 100), 16, 60),   # c4-sized widths: several 128-wide tiles per GEMM, K tails
])
def test_tensor_core_gemm_matches_cuda_core_gemm(variant, n_features, units, lookback, rows):
    P_ref, h_ref = _fit("simt", n_features, units, lookback, rows)
    P_tc, h_tc = _fit(variant, n_features, units, lookback, rows)
    scale = np.abs(P_ref).max()
    assert np.isfinite(P_tc).all()
    np.testing.assert_allclose(P_tc, P_ref, rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(h_tc, h_ref, rtol=2e-5)


def _fit_env(env, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return _fit("simt", **kw) if "GB200_LSTM_GEMM" not in env else _fit(env["GB200_LSTM_GEMM"], **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("batch,lookback,rows,units", [
    (32, 5, 70, (8, 6)),
    (10, 3, 45, (5,)),           # batch not a multiple of 16: padded sequence groups in the recurrence kernels
    (48, 4, 120, (12, 7)),       # three 16-sequence groups, partial last batch
    (32, 1, 40, (6,)),           # lookback 1: no recurrent step at all
    (7, 6, 30, (40, 33)),        # wide layers on a tiny batch: clusters of 4 CTAs
])
def test_cluster_recurrence_matches_per_step_kernels(batch, lookback, rows, units):
    """lstm_rec_fwd/bwd (one launch per layer) vs the per-time-step launch path (GB200_LSTM_REC=0)."""
    kw = dict(n_features=6, units=units, lookback=lookback, rows=rows, jobs=2, batch=batch)
    P_ref, h_ref = _fit_env({"GB200_LSTM_REC": "0"}, **kw)
    P_rec, h_rec = _fit_env({"GB200_LSTM_REC": "1"}, **kw)
    assert np.isfinite(P_rec).all()
    scale = np.abs(P_ref).max()
    np.testing.assert_allclose(P_rec, P_ref, rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(h_rec, h_ref, rtol=2e-5)


def test_fit_many_jobs_of_different_length():
    """40 jobs whose window counts differ: finished jobs idle while the others keep stepping."""
    from gordo_b200.lstm import LSTMFleet, LSTMTopology
    from oracle import factories, lstm as olstm
    spec = factories.lstm_model(4, None, lookback_window=3, encoding_dim=(5,), encoding_func=("tanh",),
                                decoding_dim=(4,), decoding_func=("tanh",), out_func="tanh")
    topo = LSTMTopology(spec["n_features"], spec["n_features_out"], spec["units"], spec["acts"], spec["out_func"],
                        spec["lookback_window"])
    g = torch.Generator(device=DEV); g.manual_seed(3)
    J = 40
    rows = np.array([20 + (7 * j) % 50 for j in range(J)], np.int64)
    lo = np.concatenate([[0], np.cumsum(rows)[:-1]]); hi = lo + rows
    X = torch.rand((int(rows.sum()), 4), generator=g, device=DEV)
    fl = LSTMFleet(topo, J, 0, DEV)
    P0 = topo.init_params(J, g, DEV)
    P = P0.clone()
    fl.fit_jobs(X, X, lo, hi, P, epochs=1, batch_size=8)
    Xh = X.cpu().numpy()
    for j in (0, 13, 39):
        p = olstm.lstm_unflatten(P0[j].cpu().numpy(), spec)
        olstm.lstm_fit(spec, p, Xh[lo[j]:hi[j]], Xh[lo[j]:hi[j]], lookback_window=3, lookahead=0, batch_size=8)
        np.testing.assert_allclose(P[j].cpu().numpy(), olstm.lstm_flatten(p), atol=3e-4)
