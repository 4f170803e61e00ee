"""
Oracle: LSTM autoencoder / forecast arithmetic (test infrastructure, see oracle/__init__.py).

Restates gordo/machine/model/models.py:557-660 (primer step, time-ordered batches, 10 000-
window predict), models.py:713-793 (create_keras_timeseriesgenerator, pinned by the reference
golden vectors tests/gordo/machine/model/test_model.py:239-311) and the [3P] Keras 3.3.3 LSTM
layer the factories of lstm_autoencoder.py:77-102 build: gate order i, f, c, o;
z = x.W + h.U + b; i,f,o = sigmoid; c~ = act(z_c); c = f*c + i*c~; h = o*act(c); h0 = c0 = 0.
float32 throughout.  PARITY UNPINNED against real Keras (not installable here); BPTT is
cross-checked against torch autograd in tests.
"""
import numpy as np

from .dense import F32, act_fwd, act_bwd, glorot_uniform, Adam


# ----------------------------------------------------------------------------- windowing
def window_count(n_rows, lookback_window, lookahead):
    return n_rows - lookback_window + 1 - lookahead


def window_index(n_rows, lookback_window, lookahead):
    """
    models.py:713-793 in index form.  Window k (k = 0..n_win-1) is sample rows
    [k, k+L) of X and target row k + L - 1 + lookahead of y.  Returns (starts, target_rows).
    """
    if lookahead < 0:
        raise ValueError(f"Value of `lookahead` can not be negative, is {lookahead}")
    n_win = max(window_count(n_rows, lookback_window, lookahead), 0)
    starts = np.arange(n_win)
    return starts, starts + lookback_window - 1 + lookahead


def timeseries_batches(X, y, batch_size, lookback_window, lookahead):
    """Materialise the generator's batches (small cases only): list of (bx [b,L,T], by [b,T])."""
    starts, tgt = window_index(len(X), lookback_window, lookahead)
    out = []
    for s in range(0, len(starts), batch_size):
        ks = starts[s:s + batch_size]
        bx = np.stack([X[k:k + lookback_window] for k in ks]) if len(ks) else np.empty((0,))
        out.append((bx, y[tgt[s:s + batch_size]]))
    return out


# ----------------------------------------------------------------------------- parameters
def orthogonal(rng, rows, cols):
    # [3P] keras OrthogonalInitializer(gain=1): QR of a normal (max, min) matrix, sign-fixed
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return q[:rows, :cols].astype(F32)


def lstm_init(spec, rng):
    """{'lstm': [(W [in,4u], U [u,4u], b [4u])...], 'dense': (Wd [u_last,T_out], bd)}."""
    layers, n_in = [], spec["n_features"]
    for u in spec["units"]:
        W = glorot_uniform(rng, n_in, 4 * u)
        U = orthogonal(rng, u, 4 * u)
        b = np.zeros(4 * u, F32); b[u:2 * u] = 1.0          # unit_forget_bias
        layers.append((W, U, b)); n_in = u
    Wd = glorot_uniform(rng, n_in, spec["n_features_out"])
    return {"lstm": layers, "dense": (Wd, np.zeros(spec["n_features_out"], F32))}


def lstm_param_count(spec):
    n, n_in = 0, spec["n_features"]
    for u in spec["units"]:
        n += n_in * 4 * u + u * 4 * u + 4 * u; n_in = u
    return n + n_in * spec["n_features_out"] + spec["n_features_out"]


def lstm_flatten(params):
    parts = []
    for W, U, b in params["lstm"]:
        parts += [W.ravel(), U.ravel(), b.ravel()]
    parts += [params["dense"][0].ravel(), params["dense"][1].ravel()]
    return np.concatenate(parts).astype(F32)


def lstm_unflatten(flat, spec):
    o, layers, n_in = 0, [], spec["n_features"]
    def take(shape):
        nonlocal o
        n = int(np.prod(shape)); a = np.asarray(flat[o:o + n], F32).reshape(shape).copy(); o += n
        return a
    for u in spec["units"]:
        layers.append((take((n_in, 4 * u)), take((u, 4 * u)), take((4 * u,)))); n_in = u
    T = spec["n_features_out"]
    return {"lstm": layers, "dense": (take((n_in, T)), take((T,)))}


def _sigmoid(z):
    return (F32(1) / (F32(1) + np.exp(-z))).astype(F32)


# ----------------------------------------------------------------------------- forward
def _layer_forward(W, U, b, act, xs, keep):
    B, L, _ = xs.shape
    u = U.shape[0]
    h = np.zeros((B, u), F32); c = np.zeros((B, u), F32)
    hs = np.empty((B, L, u), F32)
    cache = [] if keep else None
    for t in range(L):
        z = (xs[:, t] @ W + h @ U + b).astype(F32)
        i = _sigmoid(z[:, :u]); f = _sigmoid(z[:, u:2 * u])
        g = act_fwd(act, z[:, 2 * u:3 * u]).astype(F32); o = _sigmoid(z[:, 3 * u:])
        c_prev, h_prev = c, h
        c = (f * c_prev + i * g).astype(F32)
        ac = act_fwd(act, c).astype(F32)
        h = (o * ac).astype(F32)
        hs[:, t] = h
        if keep:
            cache.append((i, f, g, o, c_prev, h_prev, c, ac, z[:, 2 * u:3 * u]))
    return hs, cache


def lstm_forward(spec, params, xw, keep=False):
    """xw [B, L, T] float32 windows -> ŷ [B, T_out].  keep=True also returns caches."""
    xs = np.asarray(xw, F32)
    caches, inputs = [], []
    for (W, U, b), act in zip(params["lstm"], spec["acts"]):
        inputs.append(xs)
        xs, cache = _layer_forward(W, U, b, act, xs, keep)
        caches.append(cache)
    h_last = xs[:, -1]                      # last decoder LSTM has return_sequences=False
    Wd, bd = params["dense"]
    zd = (h_last @ Wd + bd).astype(F32)
    yhat = act_fwd(spec["out_func"], zd).astype(F32)
    return (yhat, (caches, inputs, h_last, zd)) if keep else yhat


def lstm_predict(spec, params, X, lookback_window, lookahead, batch_size=10000):
    """models.py:618-660: windows of X in batches of 10 000 -> [n_win, T_out]."""
    X = np.asarray(X, F32)
    if X.ndim == 1:
        X = X.reshape(len(X), 1)
    if lookback_window >= X.shape[0]:
        raise ValueError("For KerasLSTMForecast lookback_window must be < size of X")
    starts, _ = window_index(len(X), lookback_window, lookahead)
    out = np.empty((len(starts), spec["n_features_out"]), F32)
    for s in range(0, len(starts), batch_size):
        ks = starts[s:s + batch_size]
        xw = np.stack([X[k:k + lookback_window] for k in ks])
        out[s:s + len(ks)] = lstm_forward(spec, params, xw)
    return out


# ----------------------------------------------------------------------------- loss + grad
def lstm_loss_and_grads(spec, params, xw, yb):
    B = xw.shape[0]
    yhat, (caches, inputs, h_last, zd) = lstm_forward(spec, params, xw, keep=True)
    diff = (yhat - yb).astype(F32)
    loss = F32(np.mean(diff * diff, dtype=F32))
    dy = (F32(2.0) / F32(diff.size)) * diff
    dzd = (dy * act_bwd(spec["out_func"], zd, yhat)).astype(F32)
    Wd, _ = params["dense"]
    g_dense = ((h_last.T @ dzd).astype(F32), dzd.sum(axis=0, dtype=F32))
    L = xw.shape[1]
    n_layers = len(params["lstm"])
    dH = np.zeros((B, L, Wd.shape[0]), F32)
    dH[:, -1] = (dzd @ Wd.T).astype(F32)
    g_lstm = [None] * n_layers
    for li in range(n_layers - 1, -1, -1):
        W, U, b = params["lstm"][li]
        act = spec["acts"][li]
        u = U.shape[0]
        xs = inputs[li]
        gW = np.zeros_like(W); gU = np.zeros_like(U); gb = np.zeros_like(b)
        dX = np.zeros_like(xs)
        dh_next = np.zeros((B, u), F32); dc_next = np.zeros((B, u), F32)
        for t in range(L - 1, -1, -1):
            i, f, g, o, c_prev, h_prev, c, ac, zg = caches[li][t]
            dh = dH[:, t] + dh_next
            do = dh * ac
            dc = dc_next + dh * o * act_bwd(act, c, ac)
            di = dc * g; dg = dc * i; df = dc * c_prev
            dc_next = (dc * f).astype(F32)
            dz = np.concatenate([di * i * (1 - i), df * f * (1 - f),
                                 dg * act_bwd(act, zg, g), do * o * (1 - o)], axis=1).astype(F32)
            gW += xs[:, t].T @ dz; gU += h_prev.T @ dz; gb += dz.sum(axis=0, dtype=F32)
            dX[:, t] = dz @ W.T
            dh_next = (dz @ U.T).astype(F32)
        g_lstm[li] = (gW.astype(F32), gU.astype(F32), gb.astype(F32))
        dH = dX
    return loss, {"lstm": g_lstm, "dense": g_dense}, yhat


def _tensors(p):
    out = []
    for W, U, b in p["lstm"]:
        out += [W, U, b]
    out += [p["dense"][0], p["dense"][1]]
    return out


# ----------------------------------------------------------------------------- fit
def lstm_fit(spec, params, X, y, *, lookback_window, lookahead, batch_size=32, epochs=1):
    """
    models.py:557-616.  (1) primer: ONE Adam step on the first window alone (:585-597);
    (2) ``model.fit(generator, shuffle=False)``: batches of ``batch_size`` windows in time
    order, the Adam state carried over from the primer.  Returns (history_primer,
    history_main, adam): gordo's ``get_metadata()`` reports the PRIMER's History object
    (``self._history`` is captured at models.py:285-286 and never refreshed by :615).
    """
    X = np.asarray(X, F32); y = np.asarray(y, F32)
    if X.ndim == 1:
        X = X.reshape(len(X), 1)
    if y.ndim == 1:
        y = y.reshape(len(y), 1)
    if lookback_window >= X.shape[0]:
        raise ValueError("For KerasLSTMForecast lookback_window must be < size of X")
    tensors = _tensors(params)
    opt = Adam([t.shape for t in tensors], **spec["adam"])
    n_primer = lookahead + lookback_window
    (px, py), = timeseries_batches(X[:n_primer], y[:n_primer], 1, lookback_window, lookahead)[:1]
    loss, grads, _ = lstm_loss_and_grads(spec, params, px, py)
    opt.step(tensors, _tensors(grads))
    hist_primer = {"loss": [float(loss)]}
    starts, tgt = window_index(len(X), lookback_window, lookahead)
    hist = {"loss": []}
    for _ in range(epochs):
        lsum = 0.0
        for s in range(0, len(starts), batch_size):
            ks = starts[s:s + batch_size]
            xw = np.stack([X[k:k + lookback_window] for k in ks])
            loss, grads, _ = lstm_loss_and_grads(spec, params, xw, y[tgt[s:s + len(ks)]])
            lsum += float(loss) * len(ks)
            opt.step(tensors, _tensors(grads))
        hist["loss"].append(lsum / max(len(starts), 1))
    return hist_primer, hist, opt
