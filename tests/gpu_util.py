"""Helpers shared by the -m gpu parity tests (they call the C-ABI through gordo_b200.fleet)."""
import numpy as np
import torch

from oracle import dense, factories
from oracle.scaler import MinMaxScaler


def bf16_round(a):
    """Round-to-nearest-even to bfloat16 and back, on the CPU (emulates the tensor-core operands)."""
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def ff_forward_bf16(spec, params, X):
    """Oracle forward with bf16-rounded weights and activations, fp32 accumulate (what the TC path computes)."""
    h = bf16_round(X)
    n = len(params)
    for li, ((W, b), a) in enumerate(zip(params, spec["acts"])):
        # the bias rides inside the GEMM as a bf16 row of B (A carries a ones column)
        z = (h.astype(np.float64) @ bf16_round(W).astype(np.float64) + bf16_round(b).astype(np.float64)).astype(np.float32)
        h = dense.act_fwd(a, z).astype(np.float32)
        if li < n - 1:
            h = bf16_round(h)
    return h


def make_fleet_case(seed, row_counts, T, *, func="tanh", encoding_layers=3, cf=0.5, T_out=None,
                    thresholds=True, scale=True):
    """Seeded synthetic fleet: per-Machine data, weights, scalers, thresholds (numpy, float32)."""
    rng = np.random.default_rng(seed)
    spec = factories.feedforward_hourglass(T, T_out, encoding_layers=encoding_layers,
                                           compression_factor=cf, func=func)
    M = len(row_counts)
    To = spec["widths"][-1]
    Xs, Ys, P, in_s, in_m, es, ft, at = [], [], [], [], [], [], [], []
    for m in range(M):
        n = row_counts[m]
        X = (rng.random((n, T), dtype=np.float32) * rng.uniform(0.5, 20, T).astype(np.float32)
             + rng.uniform(-5, 5, T).astype(np.float32)).astype(np.float32)
        Y = X if To == T else rng.random((n, To), dtype=np.float32)
        params = dense.ff_init(spec, rng)
        for W, b in params:
            b += rng.normal(0, 0.05, b.shape).astype(np.float32)
        sx = MinMaxScaler().fit(X if n > 0 else np.zeros((1, T)))
        sy = MinMaxScaler().fit(Y if n > 0 else np.zeros((1, To)))
        Xs.append(X); Ys.append(Y); P.append(dense.ff_flatten(params))
        in_s.append(sx.scale_.astype(np.float32) if scale else np.ones(T, np.float32))
        in_m.append(sx.min_.astype(np.float32) if scale else np.zeros(T, np.float32))
        es.append(sy.scale_.astype(np.float32))
        ft.append(rng.uniform(0.05, 0.5, To).astype(np.float32)); at.append(np.float32(rng.uniform(0.01, 0.1)))
    return dict(spec=spec, X=Xs, Y=Ys, params=np.stack(P), in_scale=np.stack(in_s), in_min=np.stack(in_m),
                err_scale=np.stack(es), feat_thr=np.stack(ft) if thresholds else None,
                agg_thr=np.asarray(at, np.float32) if thresholds else None, row_counts=list(row_counts))


def oracle_score(case, m, forward=None):
    """The reference arithmetic for Machine m (diff.py:336-444) on the oracle, float64 scoring."""
    spec = case["spec"]
    X, Y = case["X"][m], case["Y"][m]
    params = dense.ff_unflatten(case["params"][m], spec["widths"])
    xs = (X * case["in_scale"][m] + case["in_min"][m]).astype(np.float32)
    yhat = (forward or dense.ff_forward)(spec, params, xs) if len(X) else np.zeros((0, spec["widths"][-1]), np.float32)
    d = np.abs(yhat.astype(np.float64) - Y.astype(np.float64))
    s = d * np.abs(case["err_scale"][m].astype(np.float64))
    out = {"model-output": yhat, "tag-anomaly-unscaled": d, "tag-anomaly-scaled": s,
           "total-anomaly-unscaled": (d ** 2).mean(axis=1) if len(X) else np.zeros(0),
           "total-anomaly-scaled": (s ** 2).mean(axis=1) if len(X) else np.zeros(0)}
    if case["feat_thr"] is not None:
        out["anomaly-confidence"] = d / case["feat_thr"][m]
        out["total-anomaly-confidence"] = out["total-anomaly-scaled"] / case["agg_thr"][m]
    return out


def fleet_from_case(case, device="cuda:0"):
    from gordo_b200.fleet import FFFleet, FFTopology, Schedule
    spec = case["spec"]
    topo = FFTopology(spec["widths"], spec["acts"], spec["l1"])
    fl = FFFleet(topo, len(case["row_counts"]), device)
    dev = torch.device(device)
    fl.set_params(torch.from_numpy(case["params"]))
    fl.in_scale = torch.from_numpy(case["in_scale"]).to(dev)
    fl.in_min = torch.from_numpy(case["in_min"]).to(dev)
    fl.err_scale = torch.from_numpy(case["err_scale"]).to(dev)
    if case["feat_thr"] is not None:
        fl.feat_thr = torch.from_numpy(case["feat_thr"]).to(dev)
        fl.agg_thr = torch.from_numpy(case["agg_thr"]).to(dev)
    sched = Schedule(case["row_counts"])
    T, To = spec["widths"][0], spec["widths"][-1]
    X = torch.from_numpy(np.concatenate(case["X"]) if sum(case["row_counts"]) else np.zeros((0, T), np.float32)).to(dev)
    Y = None if To == T and all(a is b for a, b in zip(case["X"], case["Y"])) else \
        torch.from_numpy(np.concatenate(case["Y"])).to(dev)
    return fl, sched, X, Y
