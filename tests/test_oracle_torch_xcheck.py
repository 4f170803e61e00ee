"""
Cross-checks of the oracle against INDEPENDENT implementations that do run here: PyTorch's own
``nn.LSTM`` (cell arithmetic, BPTT through autograd), ``nn.Linear`` stacks and ``optim.Adam`` -- code
written by neither the reference nor this repository.  Keras itself is not installable offline
(oracle/__init__.py, "PARITY UNPINNED"); these tests pin the algorithms the oracle restates (LSTM cell
with gate order i, f, c, o; MSE gradients; Adam) to a second framework's implementation of them.
The one documented difference: Keras adds epsilon outside the bias correction, torch inside.
"""
import numpy as np
import pytest
import torch

from oracle import dense, factories, lstm

torch.set_num_threads(1)


def _torch_lstm_stack(spec, params):
    """nn.LSTM modules loaded with the oracle's weights (Keras [in,4u] kernels -> torch [4u,in])."""
    mods, n_in = [], spec["n_features"]
    for (W, U, b), u in zip(params["lstm"], spec["units"]):
        m = torch.nn.LSTM(n_in, u, batch_first=True).double()
        with torch.no_grad():
            m.weight_ih_l0.copy_(torch.tensor(W.T, dtype=torch.float64))
            m.weight_hh_l0.copy_(torch.tensor(U.T, dtype=torch.float64))
            m.bias_ih_l0.copy_(torch.tensor(b, dtype=torch.float64))
            m.bias_hh_l0.zero_()
        mods.append(m); n_in = u
    lin = torch.nn.Linear(n_in, spec["n_features_out"]).double()
    with torch.no_grad():
        lin.weight.copy_(torch.tensor(params["dense"][0].T, dtype=torch.float64))
        lin.bias.copy_(torch.tensor(params["dense"][1], dtype=torch.float64))
    return mods, lin


def test_lstm_forward_and_bptt_match_torch_nn_lstm():
    rng = np.random.default_rng(11)
    spec = factories.lstm_model(5, lookback_window=7, encoding_dim=(8, 4), encoding_func=("tanh", "tanh"),
                                decoding_dim=(4, 6), decoding_func=("tanh", "tanh"), out_func="linear")
    params = lstm.lstm_init(spec, rng)
    for layer in params["lstm"]:
        layer[2][:] = rng.normal(size=layer[2].shape).astype(np.float32) * 0.1        # non-trivial biases
    xw = rng.random((9, 7, 5)).astype(np.float32); yb = rng.random((9, 5)).astype(np.float32)
    loss, grads, yhat = lstm.lstm_loss_and_grads(spec, params, xw, yb)
    mods, lin = _torch_lstm_stack(spec, params)
    h = torch.tensor(xw, dtype=torch.float64)
    for m in mods:
        h, _ = m(h)
    out = lin(h[:, -1])
    tl = ((out - torch.tensor(yb, dtype=torch.float64)) ** 2).mean()
    tl.backward()
    np.testing.assert_allclose(yhat, out.detach().numpy(), rtol=1e-4, atol=2e-6)
    assert abs(float(tl) - float(loss)) < 1e-6
    for (gW, gU, gb), m in zip(grads["lstm"], mods):
        np.testing.assert_allclose(gW, m.weight_ih_l0.grad.numpy().T, rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(gU, m.weight_hh_l0.grad.numpy().T, rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(gb, m.bias_ih_l0.grad.numpy(), rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(grads["dense"][0], lin.weight.grad.numpy().T, rtol=2e-3, atol=2e-6)


def test_lstm_predict_windows_match_torch_nn_lstm():
    """Sliding windows + stacked LSTM + Dense over a whole series (models.py:713-793 windowing)."""
    rng = np.random.default_rng(12)
    spec = factories.lstm_hourglass(6, lookback_window=4)
    params = lstm.lstm_init(spec, rng)
    X = rng.random((40, 6)).astype(np.float32)
    want = lstm.lstm_predict(spec, params, X, 4, 0)
    mods, lin = _torch_lstm_stack(spec, params)
    wins = torch.tensor(np.stack([X[k:k + 4] for k in range(40 - 4 + 1)]), dtype=torch.float64)
    h = wins
    for m in mods:
        h, _ = m(h)
    got = lin(h[:, -1]).detach().numpy()
    assert want.shape == got.shape == (37, 6)
    np.testing.assert_allclose(want, got, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("func", ["tanh", "relu"])
def test_ff_training_trajectory_matches_torch_adam(func):
    """40 mini-batch steps of the oracle's fit vs torch.nn.Linear + torch.optim.Adam on the same batches."""
    rng = np.random.default_rng(13)
    spec = factories.feedforward_hourglass(7, func=func)
    params = dense.ff_init(spec, rng)
    X = rng.random((320, 7)).astype(np.float32)
    perms = [rng.permutation(320) for _ in range(4)]
    layers = []
    for (W, b) in dense.ff_unflatten(dense.ff_flatten(params), spec["widths"]):
        lin = torch.nn.Linear(W.shape[0], W.shape[1]).double()
        with torch.no_grad():
            lin.weight.copy_(torch.tensor(W.T, dtype=torch.float64)); lin.bias.copy_(torch.tensor(b, dtype=torch.float64))
        layers.append(lin)
    act = {"tanh": torch.tanh, "relu": torch.relu, "linear": lambda z: z}
    opt = torch.optim.Adam([p for l in layers for p in l.parameters()], lr=1e-3, betas=(0.9, 0.999), eps=1e-7)
    tlosses = []
    for e in range(4):
        tot = 0.0
        for s0 in range(0, 320, 32):
            xb = torch.tensor(X[perms[e][s0:s0 + 32]], dtype=torch.float64)
            h = xb; reg = 0.0
            for lin, a, c1 in zip(layers, spec["acts"], spec["l1"]):
                h = act[a](lin(h))
                if c1:
                    reg = reg + c1 * h.abs().sum()          # activity_regularizer=l1(c1), summed over the batch (Keras 3.3.3)
            loss = ((h - xb) ** 2).mean() + reg
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss) * len(xb)
        tlosses.append(tot / 320)
    hist, _ = dense.ff_fit(spec, params, X, X, epochs=4, batch_size=32, perms=perms)
    np.testing.assert_allclose(hist["loss"], tlosses, rtol=2e-4)
    for (W, b), lin in zip(params, layers):
        # float32 (oracle) vs float64 (torch) arithmetic and the epsilon placement (Keras: outside the bias
        # correction): every weight moves by up to 40 x 1e-3; the two trajectories stay within 1e-4
        # (relu: a unit sitting at z ~ 0 can be on in one arithmetic and off in the other for a sample)
        tol = 1e-4 if func == "tanh" else 1e-3
        np.testing.assert_allclose(W, lin.weight.detach().numpy().T, atol=tol)
        np.testing.assert_allclose(b, lin.bias.detach().numpy(), atol=tol)
        assert np.abs(W - lin.weight.detach().numpy().T).mean() < (3e-5 if func == "tanh" else 1e-4)
