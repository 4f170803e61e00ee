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
