"""
Cross-check of the oracle's hand-written backward passes and Keras-form Adam against torch
autograd on CPU (the Keras reference itself is not installable here -- see oracle/__init__.py).
"""
import numpy as np
import pytest
import torch

from oracle import dense, factories, lstm

ACT = {"linear": lambda z: z, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid,
       "elu": torch.nn.functional.elu, "softplus": torch.nn.functional.softplus}


@pytest.mark.parametrize("l1_mode", ["sum", "mean"])
@pytest.mark.parametrize("func", ["tanh", "relu", "sigmoid", "elu", "softplus"])
def test_ff_grads_match_autograd(func, l1_mode):
    rng = np.random.default_rng(1)
    spec = factories.feedforward_hourglass(9, func=func)
    params = dense.ff_init(spec, rng)
    for W, b in params:
        b += rng.normal(0, 0.1, b.shape).astype(np.float32)
    xb = rng.random((13, 9)).astype(np.float32); yb = rng.random((13, 9)).astype(np.float32)
    loss, mse, grads, yhat = dense.ff_loss_and_grads(spec, params, xb, yb, l1_mode)
    tp = [(torch.tensor(W, dtype=torch.float64, requires_grad=True),
           torch.tensor(b, dtype=torch.float64, requires_grad=True)) for W, b in params]
    h = torch.tensor(xb, dtype=torch.float64); reg = 0.0
    for li, ((W, b), a) in enumerate(zip(tp, spec["acts"])):
        h = ACT[a](h @ W + b)
        if spec["l1"][li]:
            r = spec["l1"][li] * h.abs().sum()
            reg = reg + (r if l1_mode == "sum" else r / len(xb))
    tl = ((h - torch.tensor(yb, dtype=torch.float64)) ** 2).mean() + reg
    tl.backward()
    assert abs(float(tl) - float(loss)) < 1e-5
    for (gW, gb), (W, b) in zip(grads, tp):
        np.testing.assert_allclose(gW, W.grad.numpy(), rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(gb, b.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_lstm_grads_match_autograd():
    rng = np.random.default_rng(2)
    spec = factories.lstm_model(4, lookback_window=5, encoding_dim=(6, 3), encoding_func=("tanh", "relu"),
                                decoding_dim=(3, 5), decoding_func=("tanh", "tanh"))
    params = lstm.lstm_init(spec, rng)
    xw = rng.random((7, 5, 4)).astype(np.float32); yb = rng.random((7, 4)).astype(np.float32)
    loss, grads, yhat = lstm.lstm_loss_and_grads(spec, params, xw, yb)
    tl_params = [[torch.tensor(t, dtype=torch.float64, requires_grad=True) for t in layer]
                 for layer in params["lstm"]]
    Wd = torch.tensor(params["dense"][0], dtype=torch.float64, requires_grad=True)
    bd = torch.tensor(params["dense"][1], dtype=torch.float64, requires_grad=True)
    xs = torch.tensor(xw, dtype=torch.float64)
    for (W, U, b), a in zip(tl_params, spec["acts"]):
        u = U.shape[0]; h = torch.zeros(7, u, dtype=torch.float64); c = torch.zeros_like(h); seq = []
        for t in range(5):
            z = xs[:, t] @ W + h @ U + b
            i, f, g, o = torch.sigmoid(z[:, :u]), torch.sigmoid(z[:, u:2*u]), ACT[a](z[:, 2*u:3*u]), torch.sigmoid(z[:, 3*u:])
            c = f * c + i * g; h = o * ACT[a](c); seq.append(h)
        xs = torch.stack(seq, 1)
    out = xs[:, -1] @ Wd + bd
    tl = ((out - torch.tensor(yb, dtype=torch.float64)) ** 2).mean()
    tl.backward()
    assert abs(float(tl) - float(loss)) < 1e-6
    np.testing.assert_allclose(yhat, out.detach().numpy(), rtol=1e-4, atol=1e-5)
    for (gW, gU, gb), (W, U, b) in zip(grads["lstm"], tl_params):
        np.testing.assert_allclose(gW, W.grad.numpy(), rtol=1e-3, atol=2e-6)
        np.testing.assert_allclose(gU, U.grad.numpy(), rtol=1e-3, atol=2e-6)
        np.testing.assert_allclose(gb, b.grad.numpy(), rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(grads["dense"][0], Wd.grad.numpy(), rtol=1e-3, atol=2e-6)


def test_keras_adam_form():
    # w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps): eps OUTSIDE the bias correction
    w = np.array([1.0, -2.0], np.float32); g = np.array([0.5, 0.25], np.float32)
    opt = dense.Adam([w.shape]); w0 = w.copy()
    m = np.zeros(2); v = np.zeros(2); ref = w0.astype(np.float64)
    for t in range(1, 6):
        opt.step([w], [g])
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        ref = ref - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-7)
    np.testing.assert_allclose(w, ref, rtol=1e-5)


def test_ff_fit_learns_and_history():
    rng = np.random.default_rng(3)
    spec = factories.feedforward_hourglass(6)
    params = dense.ff_init(spec, rng)
    Z = rng.random((600, 2)); X = np.tanh(Z @ rng.normal(size=(2, 6))).astype(np.float32)
    perms = [rng.permutation(600) for _ in range(5)]
    hist, _ = dense.ff_fit(spec, params, X, X, epochs=5, batch_size=32, perms=perms)
    assert len(hist["loss"]) == 5 and hist["loss"][-1] < hist["loss"][0]
    assert 0.0 <= hist["accuracy"][-1] <= 1.0
    hist2, _ = dense.ff_fit(spec, dense.ff_init(spec, np.random.default_rng(3)), X, X, epochs=1,
                            validation_split=0.25)
    assert "val_loss" in hist2


def test_orthogonal_init_is_orthogonal():
    U = lstm.orthogonal(np.random.default_rng(0), 5, 20)
    np.testing.assert_allclose(U @ U.T, np.eye(5), atol=1e-5)
