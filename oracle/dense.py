"""
Oracle: feed-forward autoencoder arithmetic (test infrastructure, see oracle/__init__.py).

Restates what gordo/machine/model/models.py:243-300 delegates to [3P] Keras 3.3.3 through
scikeras 0.13.0: Dense forward, ``Model.fit`` batching, MSE, the L1
activity regulariser of feedforward_autoencoder.py:78-81, Keras-form Adam, the ``loss`` /
``accuracy`` history, and ``Model.predict``.  float32 throughout, as Keras computes.

Deterministic by construction: initial weights and the per-epoch permutations are INPUTS, so
the CUDA path can be fed the identical ones.  PARITY UNPINNED against real Keras (it cannot be
installed here); the backward pass is cross-checked against torch autograd in tests.
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- activations
def act_fwd(name, z):
    if name == "linear":
        return z
    if name == "tanh":
        return np.tanh(z)
    if name == "relu":
        return np.maximum(z, F32(0))
    if name == "sigmoid":
        return (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    if name == "elu":
        return np.where(z > 0, z, np.expm1(np.minimum(z, F32(0)))).astype(F32)
    if name == "softplus":
        return np.logaddexp(z, F32(0)).astype(F32)
    raise ValueError(f"unknown activation {name}")


def act_bwd(name, z, h):
    """d act / d z, given pre-activation z and output h."""
    if name == "linear":
        return np.ones_like(z)
    if name == "tanh":
        return F32(1) - h * h
    if name == "relu":
        return (z > 0).astype(F32)
    if name == "sigmoid":
        return h * (F32(1) - h)
    if name == "elu":
        return np.where(z > 0, F32(1), h + F32(1)).astype(F32)
    if name == "softplus":
        return (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    raise ValueError(f"unknown activation {name}")


# ----------------------------------------------------------------------------- parameters
def glorot_uniform(rng, fan_in, fan_out):
    # [3P] keras GlorotUniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out))
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(F32)


def ff_init(spec, rng):
    """List of (W [in,out], b [out]) float32; glorot-uniform kernels, zero bias."""
    w = spec["widths"]
    return [(glorot_uniform(rng, w[i], w[i + 1]), np.zeros(w[i + 1], F32))
            for i in range(len(w) - 1)]


def ff_param_count(widths):
    return sum(widths[i] * widths[i + 1] + widths[i + 1] for i in range(len(widths) - 1))


def ff_flatten(params):
    """Flat float32 vector: per layer W (row-major [in,out]) then b."""
    return np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in params]).astype(F32)


def ff_unflatten(flat, widths):
    out, o = [], 0
    for i in range(len(widths) - 1):
        a, b = widths[i], widths[i + 1]
        W = np.asarray(flat[o:o + a * b], F32).reshape(a, b).copy(); o += a * b
        bias = np.asarray(flat[o:o + b], F32).copy(); o += b
        out.append((W, bias))
    return out


# ----------------------------------------------------------------------------- forward
def ff_forward(spec, params, X, keep=False):
    """ŷ = Dense stack applied to X [n, T] (float32).  keep=True also returns (zs, hs)."""
    h = np.asarray(X, F32)
    zs, hs = [], [h]
    for (W, b), a in zip(params, spec["acts"]):
        z = (h @ W + b).astype(F32)
        h = act_fwd(a, z).astype(F32)
        if keep:
            zs.append(z); hs.append(h)
    return (h, zs, hs) if keep else h


def ff_predict(spec, params, X, batch_size=32):
    """[3P] Keras ``Model.predict``: float32 cast, batches of 32, no shuffling."""
    X = np.asarray(X, F32)
    out = np.empty((len(X), spec["widths"][-1]), F32)
    for s in range(0, len(X), batch_size):
        out[s:s + batch_size] = ff_forward(spec, params, X[s:s + batch_size])
    return out


# ----------------------------------------------------------------------------- loss + grad
def ff_loss_and_grads(spec, params, xb, yb, l1_mode="sum"):
    """
    One batch: returns (loss, mse, grads, yhat).  loss = mean((ŷ-y)^2 over all elements)
    + sum over regularised layers of l1 * sum|h| (l1_mode "sum": Keras 3.3.3, the activity
    loss is added as-is; "mean": divided by the batch size, Keras 2 / later Keras 3).
    """
    B = xb.shape[0]
    yhat, zs, hs = ff_forward(spec, params, xb, keep=True)
    diff = (yhat - yb).astype(F32)
    loss_kind = spec.get("loss", "mean_squared_error")
    if loss_kind in ("mean_squared_error", "mse"):
        base = F32(np.mean(diff * diff, dtype=F32))
        delta_h = (F32(2.0) / F32(diff.size)) * diff
    elif loss_kind in ("mean_absolute_error", "mae"):
        base = F32(np.mean(np.abs(diff), dtype=F32))
        delta_h = np.sign(diff).astype(F32) / F32(diff.size)
    else:
        raise ValueError(f"unsupported loss {loss_kind}")
    reg_scale = F32(1.0) if l1_mode == "sum" else F32(1.0 / B)
    reg = F32(0)
    for li, c in enumerate(spec["l1"]):
        if c:
            reg += F32(c) * reg_scale * F32(np.sum(np.abs(hs[li + 1]), dtype=F32))
    grads = [None] * len(params)
    for li in range(len(params) - 1, -1, -1):
        c = spec["l1"][li]
        if c:
            delta_h = delta_h + F32(c) * reg_scale * np.sign(hs[li + 1]).astype(F32)
        dz = (delta_h * act_bwd(spec["acts"][li], zs[li], hs[li + 1])).astype(F32)
        gW = (hs[li].T @ dz).astype(F32)
        gb = dz.sum(axis=0, dtype=F32)
        grads[li] = (gW, gb)
        if li > 0:
            delta_h = (dz @ params[li][0].T).astype(F32)
    return F32(base + reg), base, grads, yhat


# ----------------------------------------------------------------------------- Adam (Keras form)
class Adam:
    """
    [3P] keras.optimizers.Adam: alpha = lr*sqrt(1-b2^t)/(1-b1^t);
    m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= alpha*m/(sqrt(v)+eps).  eps is added AFTER the
    bias correction is folded into alpha -- this is not torch.optim.Adam.
    """

    def __init__(self, shapes, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.t = 0
        self.m = [np.zeros(s, F32) for s in shapes]
        self.v = [np.zeros(s, F32) for s in shapes]

    def step(self, tensors, grads):
        self.t += 1
        t = self.t
        # float32 scalar arithmetic as Keras does (ops.power on float32 tensors)
        b1t = F32(np.power(F32(self.b1), F32(t)))
        b2t = F32(np.power(F32(self.b2), F32(t)))
        alpha = F32(self.lr) * F32(np.sqrt(F32(1) - b2t)) / (F32(1) - b1t)
        for w, g, m, v in zip(tensors, grads, self.m, self.v):
            m += (g - m) * F32(1 - self.b1)
            v += (g * g - v) * F32(1 - self.b2)
            w -= alpha * m / (np.sqrt(v) + F32(self.eps))


def _flat_tensors(params):
    out = []
    for W, b in params:
        out += [W, b]
    return out


# ----------------------------------------------------------------------------- fit
def ff_fit(spec, params, X, y, *, epochs=1, batch_size=32, perms=None, l1_mode="sum",
           validation_split=0.0, adam_state=None):
    """
    [3P] Keras ``Model.fit`` on arrays: float32 cast; validation_split
    holds out the LAST rows before shuffling; one permutation of the training rows per epoch
    (``perms[e]``; None = no shuffle); last partial batch kept; history ``loss`` = sample-
    weighted running mean of the batch loss (MSE + activity loss), ``accuracy`` = mean of
    argmax(ŷ) == argmax(y) (categorical accuracy; binary accuracy at 0.5 when T_out == 1).
    Updates ``params`` in place.  Returns (history dict, Adam state).
    """
    X = np.asarray(X, F32); y = np.asarray(y, F32)
    n = len(X)
    n_train = n if not validation_split else int(np.floor(n * (1.0 - validation_split)))
    Xv, yv = X[n_train:], y[n_train:]
    X, y = X[:n_train], y[:n_train]
    tensors = _flat_tensors(params)
    opt = adam_state or Adam([t.shape for t in tensors], **spec["adam"])
    hist = {"loss": [], "accuracy": []}
    if validation_split:
        hist["val_loss"] = []; hist["val_accuracy"] = []
    for e in range(epochs):
        order = np.arange(n_train) if perms is None else np.asarray(perms[e])
        assert len(order) == n_train
        lsum = 0.0; asum = 0.0
        for s in range(0, n_train, batch_size):
            idx = order[s:s + batch_size]
            xb, yb = X[idx], y[idx]
            loss, _, grads, yhat = ff_loss_and_grads(spec, params, xb, yb, l1_mode)
            lsum += float(loss) * len(idx)
            asum += float(_accuracy_sum(yb, yhat))
            flat_g = []
            for gW, gb in grads:
                flat_g += [gW, gb]
            opt.step(tensors, flat_g)
        hist["loss"].append(lsum / n_train)
        hist["accuracy"].append(asum / n_train)
        if validation_split:
            vl = 0.0; va = 0.0
            for s in range(0, len(Xv), batch_size):
                xb, yb = Xv[s:s + batch_size], yv[s:s + batch_size]
                loss, _, _, yhat = ff_loss_and_grads(spec, params, xb, yb, l1_mode)
                vl += float(loss) * len(xb); va += float(_accuracy_sum(yb, yhat))
            hist["val_loss"].append(vl / max(len(Xv), 1))
            hist["val_accuracy"].append(va / max(len(Xv), 1))
    return hist, opt


def _accuracy_sum(y, yhat):
    if y.shape[1] == 1:
        return np.sum((yhat[:, 0] > 0.5).astype(F32) == y[:, 0])
    return np.sum(np.argmax(y, axis=1) == np.argmax(yhat, axis=1))


def explained_variance_score(y_true, y_pred):
    """[3P] sklearn.metrics.explained_variance_score, uniform average (models.py:398)."""
    y_true = np.asarray(y_true, np.float64); y_pred = np.asarray(y_pred, np.float64)
    d = y_true - y_pred
    num = np.mean((d - d.mean(axis=0)) ** 2, axis=0)
    den = np.mean((y_true - y_true.mean(axis=0)) ** 2, axis=0)
    nz_num, nz_den = num != 0, den != 0
    score = np.ones(y_true.shape[1])
    ok = nz_num & nz_den
    score[ok] = 1 - num[ok] / den[ok]
    score[nz_num & ~nz_den] = 0.0
    return float(np.mean(score))
