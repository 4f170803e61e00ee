"""
FleetBuild / FleetModelBuilder -- the batched twin of gordo.builder.ModelBuilder._build
(gordo/builder/build_model.py:192-339) for a whole project:

    for machine in machines: ModelBuilder(machine).build()          # local_build.py:69-70

becomes, per topology bucket, a handful of launches over ALL Machines:
  MinMaxScaler.fit (every CV fold + final)  -> gb200_minmax_fit
  Keras fit of every fold + the final model -> gb200_ff_fit        (one CTA per fit)
  fold predictions + validation errors      -> gb200_ff_score      (test folds = virtual Machines)
  rolling(6).min().max() thresholds         -> gb200_rolling_min_max
and the result is materialised as the same fitted objects a per-Machine build produces
(DiffBasedAnomalyDetector[Pipeline[MinMaxScaler, KerasAutoEncoder]]), ready for serializer.dump /
gordo.server.  Machines whose model is not that standard composition are built one at a time
through the estimator API (still on the GPU).

Two classes:
  FleetBuild          the batched engine: FleetBuild(machines).build() -> [(model, metadata)]
  FleetModelBuilder   the reference's builder seam (gordo/builder/utils.py:8-17, cli/cli.py:147-150): a
                      ``ModelBuilder`` subclass whenever ``gordo`` is importable -- same constructor, ``build()``
                      returns ``(model, machine)`` -- plus ``build_fleet`` / ``local_build``, the batched twins of
                      ``for machine in machines: ModelBuilder(machine).build()`` (local_build.py:67-70), and
                      ``build_sharded`` for one process per GPU.
"""
import datetime
import threading
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import pandas as pd
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler

from gordo_b200 import serializer
from gordo_b200.fleet import FFFleet, FFTopology, Schedule, time_series_split_bounds
from gordo_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector
from gordo_b200.machine.model.models import KerasAutoEncoder, KerasLSTMBaseEstimator, History, _Model


@dataclass
class FleetMachine:
    """What ModelBuilder needs of a Machine once its dataset is fetched (build_model.py:208-219)."""
    name: str
    X: Any                                  # DataFrame / ndarray [n, T]
    y: Any = None                           # defaults to X (autoencoder)
    model: Optional[dict] = None            # model definition (Machine YAML `model:`), default hourglass AE
    evaluation: Dict[str, Any] = field(default_factory=dict)   # cv_mode, n_splits, seed

    def definition(self):
        return self.model or {
            "gordo_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {
                "base_estimator": {"sklearn.pipeline.Pipeline": {"steps": [
                    "sklearn.preprocessing.MinMaxScaler",
                    {"gordo_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass"}}]}}}}


def extract_model_metadata(model) -> Dict[str, Any]:
    """
    ``ModelBuilder._extract_metadata_from_model`` (build_model.py:518-570): the ``get_metadata()`` of every GordoBase
    reachable from ``model`` -- through the last step of a Pipeline and through estimator-valued attributes --
    merged into one dict (outer objects first, so an inner estimator's keys win, as in the reference).
    """
    from sklearn.base import BaseEstimator
    from gordo_b200.machine.model.base import GordoBase
    out: Dict[str, Any] = {}
    if isinstance(model, Pipeline):
        return extract_model_metadata(model.steps[-1][1])
    if isinstance(model, GordoBase):
        out.update(model.get_metadata())
    for key, val in list(getattr(model, "__dict__", {}).items()):
        if key == "regressor" or key.startswith("_gb200_"):
            continue
        if isinstance(val, Pipeline):
            out.update(extract_model_metadata(val.steps[-1][1]))
        elif isinstance(val, (GordoBase, BaseEstimator)):
            out.update(extract_model_metadata(val))
    return out


def _metrics_dict(y: pd.DataFrame, scaler="sklearn.preprocessing.MinMaxScaler"):
    """ModelBuilder.build_metrics_dict (build_model.py:377-446): ``{metric}-{tag}`` per target tag and ``{metric}``."""
    from sklearn import metrics as skm
    from gordo_b200.machine.model.utils import metric_wrapper
    if isinstance(scaler, (str, dict)):
        scaler = serializer.from_definition(scaler)
    if scaler is not None:
        scaler.fit(y)

    def per_tag(fn, j):
        def _score(y_true, y_pred):
            return fn(np.asarray(getattr(y_true, "values", y_true))[:, j], np.asarray(getattr(y_pred, "values", y_pred))[:, j])
        return _score

    out = {}
    for name in DEFAULT_METRICS:
        fn = getattr(skm, name)
        ms = name.replace("_", "-")
        for j, col in enumerate(y.columns):
            out[f"{ms}-{str(col).replace(' ', '-')}"] = skm.make_scorer(metric_wrapper(per_tag(fn, j), scaler=scaler))
        out[ms] = skm.make_scorer(metric_wrapper(fn, scaler=scaler))
    return out


def build_split_dict(X, split_obj=None, bounds=None) -> dict:
    """ModelBuilder.build_split_dict (build_model.py:347-375): per fold the first / last index label of the train and
    test rows and their counts.  ``split_obj``: any sklearn splitter; ``bounds``: the (train_end, test_end) pairs of a
    TimeSeriesSplit (train = rows [0, s), test = rows [s, e)), which is what the batched build already holds."""
    index = X.index if hasattr(X, "index") else pd.RangeIndex(len(X))
    out: Dict[str, Any] = {}
    if bounds is not None:
        folds = [((0, s - 1), (s, e - 1), s, e - s) for s, e in bounds]
    else:
        folds = [((tr[0], tr[-1]), (te[0], te[-1]), len(tr), len(te)) for tr, te in split_obj.split(X)]
    for i, ((a0, a1), (b0, b1), n_tr, n_te) in enumerate(folds):
        out.update({f"fold-{i + 1}-train-start": index[a0], f"fold-{i + 1}-train-end": index[a1],
                    f"fold-{i + 1}-test-start": index[b0], f"fold-{i + 1}-test-end": index[b1]})
        out[f"fold-{i + 1}-n-train"] = int(n_tr)
        out[f"fold-{i + 1}-n-test"] = int(n_te)
    return out


def _plain_minmax(sc) -> bool:
    return (type(sc) is MinMaxScaler and tuple(getattr(sc, "feature_range", (0, 1))) == (0, 1)
            and not getattr(sc, "clip", False))


def segmented_randperm(lengths: Sequence[int], generator, device):
    """Independent random permutations of range(len) for every segment, concatenated (int32)."""
    import torch
    lengths_t = torch.as_tensor(np.asarray(lengths, np.int64), device=device)
    total = int(lengths_t.sum())
    if total == 0:
        return torch.empty(0, dtype=torch.int32, device=device)
    seg = torch.repeat_interleave(torch.arange(len(lengths), device=device), lengths_t)
    keys = torch.rand(total, generator=generator, device=device)
    order = torch.argsort(keys)
    order = order[torch.argsort(seg[order], stable=True)]          # random within, grouped by segment
    starts = torch.cumsum(lengths_t, 0) - lengths_t
    return (order - starts[seg]).to(torch.int32).contiguous()


def job_seeds(seed: int, n: int) -> List[int]:
    """
    The torch seeds of the ``n`` successive fits of ONE Machine's build, exactly as the per-Machine path
    draws them: ``ModelBuilder.set_seed`` seeds numpy (build_model.py:341-345), then every ``fit`` (CV folds
    in order, then the final fit) draws ``np.random.randint(0, 2**31 - 1)`` (models.py ``fit``).
    """
    rs = np.random.RandomState(int(seed))
    return [int(rs.randint(0, 2 ** 31 - 1)) for _ in range(n)]


class FleetBuild:
    def __init__(self, machines: Sequence[FleetMachine], device: Optional[str] = None, cv_precision: str = "f32",
                 streams: int = 1):
        """
        streams: topology buckets built concurrently, each on its own CUDA stream from its own host thread
        (a heterogeneous project is many small buckets; one bucket alone cannot fill 148 SMs).
        """
        self.machines = list(machines)
        self.device = device
        self.cv_precision = cv_precision
        self.streams = max(1, int(streams))
        self.last_bucket_count = 0

    # ------------------------------------------------------------------ public
    def build(self):
        """Returns [(fitted model, metadata dict)] in the order of ``machines``."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("gordo_b200 needs a CUDA device (there is no CPU fallback)")
        dev = torch.device(self.device) if self.device else torch.device("cuda", torch.cuda.current_device())
        results: List[Any] = [None] * len(self.machines)
        buckets: Dict[Any, List[int]] = {}
        prototypes: Dict[int, Any] = {}
        for i, mc in enumerate(self.machines):
            model = serializer.from_definition(mc.definition())
            prototypes[i] = model
            key = self._bucket_key(model, mc)
            if key is None:
                results[i] = self._build_one(model, mc)
            else:
                buckets.setdefault(key, []).append(i)
        self.last_bucket_count = len(buckets)

        self.bucket_log = []                       # (kind, tags, Machines, wall seconds) per bucket, for the bench

        def run(item):
            key, idxs = item
            t0 = time.time()
            out = self._build_bucket([self.machines[i] for i in idxs], [prototypes[i] for i in idxs], dev)
            X0 = self.machines[idxs[0]].X
            self.bucket_log.append(("lstm" if key[0] == "lstm" else "ff", int(np.shape(getattr(X0, "values", X0))[1]),
                                    len(idxs), time.time() - t0))
            return idxs, out

        # largest buckets first so the small ones fill the tail
        order = sorted(buckets.items(), key=lambda kv: -sum(len(np.asarray(getattr(self.machines[i].X, "values", self.machines[i].X)))
                                                            for i in kv[1]))
        if self.streams <= 1 or len(order) <= 1:
            done = [run(it) for it in order]
        else:
            from concurrent.futures import ThreadPoolExecutor
            main = torch.cuda.current_stream(dev)
            ready = torch.cuda.Event(); ready.record(main)
            local = threading.local()

            def run_on_stream(item):
                if not hasattr(local, "stream"):
                    torch.cuda.set_device(dev)
                    local.stream = torch.cuda.Stream(device=dev)
                    local.stream.wait_event(ready)
                with torch.cuda.stream(local.stream):
                    out = run(item)
                    local.stream.synchronize()
                return out
            with ThreadPoolExecutor(max_workers=min(self.streams, len(order))) as pool:
                done = list(pool.map(run_on_stream, order))
        for idxs, built in done:
            for i, r in zip(idxs, built):
                results[i] = r
        return results

    # ------------------------------------------------------------------ bucketing
    @staticmethod
    def _standard_parts(model):
        if type(model) is DiffBasedKFCVAnomalyDetector:
            if model.smoothing_method not in ("smm", "sma", "ewma") or not model.window:
                return None
        elif type(model) is not DiffBasedAnomalyDetector or model.shuffle:
            return None
        # the batched path fits both scalers with plain (0, 1) MinMax semantics (gb200_minmax_fit): any
        # other feature_range, or clip=True, builds through the estimator API like the reference
        if not _plain_minmax(model.scaler):
            return None
        be = model.base_estimator
        if isinstance(be, Pipeline) and len(be.steps) == 2 and _plain_minmax(be.steps[0][1]) \
                and (type(be.steps[1][1]) is KerasAutoEncoder or isinstance(be.steps[1][1], KerasLSTMBaseEstimator)):
            return be.steps[1][1]
        return None

    def _bucket_key(self, model, mc: FleetMachine):
        est = self._standard_parts(model)
        if est is None or est.kwargs.get("callbacks") or est.kwargs.get("validation_split"):
            return None
        if mc.evaluation.get("cv_mode", "full_build") not in ("full_build", "build_only"):
            return None
        X = np.asarray(getattr(mc.X, "values", mc.X))
        y = X if mc.y is None else np.asarray(getattr(mc.y, "values", mc.y))
        est.kwargs.update({"n_features": X.shape[1], "n_features_out": y.shape[1]})
        topo = est._topology()
        kfcv = None
        if type(model) is DiffBasedKFCVAnomalyDetector:
            if isinstance(est, KerasLSTMBaseEstimator):
                return None                  # the K-fold variant is batched for the feed-forward family only
            kfcv = (model.smoothing_method, float(model.threshold_percentile), bool(model.shuffle))
        if isinstance(est, KerasLSTMBaseEstimator):
            return ("lstm", type(est).__name__, topo.key(), int(est.kwargs.get("epochs", 1)), int(est.batch_size),
                    mc.evaluation.get("cv_mode", "full_build"), int(mc.evaluation.get("n_splits", 3)),
                    tuple(sorted(topo.adam.items())), est.kwargs.get("precision", "f32"), model.window)
        fit = (int(est.kwargs.get("epochs", 1)), int(est.kwargs.get("batch_size") or 32),
               bool(est.kwargs.get("shuffle", True)), est.kwargs.get("l1_batch_norm", "sum"),
               mc.evaluation.get("cv_mode", "full_build"), int(mc.evaluation.get("n_splits", 3)),
               tuple(sorted(topo.adam.items())), model.window, kfcv)
        return (topo.key(), fit)

    # ------------------------------------------------------------------ one-at-a-time fallback
    def _build_one(self, model, mc: FleetMachine, metrics: bool = False):
        """
        The reference's own sequence for one Machine (build_model.py:239-339): cross_validate -> fit -> offset ->
        metadata, through the estimator API.  ``metrics``: also the builder's CV scoring metrics
        (build_model.py:245-289: 4 metrics x (tags + 1) scorers on MinMax-scaled y / yhat).
        """
        np.random.seed(int(mc.evaluation.get("seed", 0)))
        X = mc.X if isinstance(mc.X, pd.DataFrame) else pd.DataFrame(np.asarray(mc.X))
        y = X if mc.y is None else (mc.y if isinstance(mc.y, pd.DataFrame) else pd.DataFrame(np.asarray(mc.y)))
        meta: Dict[str, Any] = {"name": mc.name, "cross_validation": {"scores": {}, "splits": {}}}
        t0 = time.time()
        cv_mode = mc.evaluation.get("cv_mode", "full_build")
        if cv_mode != "build_only" and hasattr(model, "predict"):
            from sklearn.model_selection import TimeSeriesSplit, cross_validate
            cv = TimeSeriesSplit(n_splits=int(mc.evaluation.get("n_splits", 3)))
            kwargs = {}
            if metrics:
                kwargs["scoring"] = _metrics_dict(y, mc.evaluation.get("scoring_scaler", "sklearn.preprocessing.MinMaxScaler"))
            if hasattr(model, "cross_validate"):
                cvo = model.cross_validate(X=X, y=y, cv=cv, **kwargs)
            else:
                cvo = cross_validate(model, X=X, y=y, cv=cv, return_estimator=True, **kwargs)
            for name in kwargs.get("scoring", {}):
                v = np.asarray(cvo[f"test_{name}"], np.float64)
                d = {"fold-mean": float(v.mean()), "fold-std": float(v.std()), "fold-max": float(v.max()), "fold-min": float(v.min())}
                d.update({f"fold-{i + 1}": float(x) for i, x in enumerate(v)})
                meta["cross_validation"]["scores"][name] = d
            meta["cross_validation"]["splits"] = build_split_dict(X, cv)
            meta["cv_duration_sec"] = time.time() - t0
        if cv_mode != "cross_val_only":
            t1 = time.time()
            model.fit(X, y)
            meta["model_training_duration_sec"] = time.time() - t1
            meta["model_offset"] = len(X) - len(model.predict(X))
        meta["model"] = extract_model_metadata(model)
        return model, meta

    # ------------------------------------------------------------------ the batched build
    def _build_bucket(self, mcs: List[FleetMachine], protos: List[Any], dev):
        import torch
        est0 = self._standard_parts(protos[0])
        if isinstance(est0, KerasLSTMBaseEstimator):
            return self._build_bucket_lstm(mcs, protos, dev)
        topo: FFTopology = est0._topology()
        epochs = int(est0.kwargs.get("epochs", 1)); batch = int(est0.kwargs.get("batch_size") or 32)
        do_shuffle = bool(est0.kwargs.get("shuffle", True))
        l1_mean = est0.kwargs.get("l1_batch_norm", "sum") == "mean"
        cv_mode = mcs[0].evaluation.get("cv_mode", "full_build")
        k = int(mcs[0].evaluation.get("n_splits", 3)) if cv_mode == "full_build" else 0
        M, T, To, P = len(mcs), topo.n_in, topo.n_out, topo.n_params

        Xs = [np.asarray(getattr(m.X, "values", m.X)) for m in mcs]
        alias = all(m.y is None or m.y is m.X for m in mcs)
        rows = np.array([len(x) for x in Xs], np.int64)
        off = np.concatenate([[0], np.cumsum(rows)])
        xd = torch.as_tensor(np.ascontiguousarray(np.concatenate(Xs), np.float32), device=dev)
        yd = None if alias else torch.as_tensor(np.ascontiguousarray(
            np.concatenate([np.asarray(getattr(m.y, "values", m.y)) if m.y is not None else x
                            for m, x in zip(mcs, Xs)]), np.float32), device=dev)
        ysrc = xd if yd is None else yd

        # jobs: per Machine k CV folds (training prefix) + the final fit on all rows
        lo, hi, te_lo, te_hi = [], [], [], []
        for m in range(M):
            bounds = time_series_split_bounds(int(rows[m]), k) if k else []
            for (s, e) in bounds:
                lo.append(off[m]); hi.append(off[m] + s); te_lo.append(off[m] + s); te_hi.append(off[m] + e)
            lo.append(off[m]); hi.append(off[m] + rows[m])
        J = len(lo)
        per = k + 1
        lo_t = torch.as_tensor(np.asarray(lo, np.int64), device=dev)
        hi_t = torch.as_tensor(np.asarray(hi, np.int64), device=dev)
        t_start = time.time()
        fleet = FFFleet(topo, M, dev)
        in_scale, in_min = FFFleet.minmax_fit(xd, lo_t, hi_t)                 # Pipeline's MinMaxScaler per job
        err_scale, _ = FFFleet.minmax_fit(ysrc, lo_t, hi_t)                   # detector scaler per job (diff.py:173)
        # every fit job draws its initial weights and its per-epoch permutations from ITS OWN generator, seeded as
        # the per-Machine path seeds it (job_seeds): a Machine's model does not depend on which bucket it shares
        gen = torch.Generator(device=dev)
        n_job = (np.asarray(hi) - np.asarray(lo)).astype(np.int64)
        is_kfcv = type(protos[0]) is DiffBasedKFCVAnomalyDetector
        kf_shuffle = is_kfcv and protos[0].shuffle and not do_shuffle
        params = torch.empty((J, P), dtype=torch.float32, device=dev)
        perms = []
        for m, mc in enumerate(mcs):
            for i, sd in enumerate(job_seeds(int(mc.evaluation.get("seed", 0)), per)):
                j = m * per + i
                gen.manual_seed(sd)
                params[j] = topo.glorot_init(1, gen, dev)[0]
                if do_shuffle and n_job[j] > 0:
                    perms += [torch.randperm(int(n_job[j]), generator=gen, device=dev) for _ in range(epochs)]
        pool = poff = None
        if kf_shuffle:
            # diff.py:166-170: rows shuffled once with random_state=0, then fed in that order every epoch
            hp = []
            for nj in n_job:
                idx = np.arange(int(nj)); np.random.RandomState(0).shuffle(idx)
                hp.append(np.tile(idx, epochs))
            pool = torch.as_tensor(np.concatenate(hp).astype(np.int32), device=dev)
        elif do_shuffle and perms:
            pool = torch.cat(perms).to(torch.int32).contiguous()
        if pool is not None:
            poff = torch.as_tensor(np.concatenate([[0], np.cumsum(n_job * epochs)[:-1]]).astype(np.int64), device=dev)
        hl, ha, _, _ = fleet.fit_jobs(xd, yd, lo_t, hi_t, params, in_scale=in_scale, in_min=in_min, epochs=epochs,
                                      batch_size=batch, perm_pool=pool, perm_off=poff, l1_mean=l1_mean)
        torch.cuda.current_stream().synchronize()
        t_fit = time.time() - t_start

        feat_pf = agg_pf = None
        win = protos[0].window
        sfeat_h = sagg_h = None
        if k:
            # score every test fold with its fold model: virtual Machines over sub-ranges
            fold_jobs = np.array([m * per + i for m in range(M) for i in range(k)])
            vfleet = FFFleet(topo, M * k, dev)
            sel = torch.as_tensor(fold_jobs, device=dev)
            vfleet.set_params(params[sel]); vfleet.in_scale = in_scale[sel].contiguous()
            vfleet.in_min = in_min[sel].contiguous(); vfleet.err_scale = err_scale[sel].contiguous()
            vs = Schedule(rows_lo=te_lo, rows_hi=te_hi, rows_total=int(off[-1]))
            prec = vfleet.auto_precision(self.cv_precision)
            init = None
            if is_kfcv:
                # diff.py:580-612: predictions start as zeros and the validation error as NaN; rows no test
                # fold covers keep them (with the builder's TimeSeriesSplit that is the first quarter)
                R = int(off[-1])
                zeros = torch.zeros((R, To), dtype=torch.float32, device=dev)
                init = {"model-output": zeros,
                        "tag-anomaly-unscaled": FFFleet.score_outputs(zeros, ysrc, [0, R], [0])["tag-anomaly-unscaled"],
                        "total-anomaly-scaled": torch.full((R,), float("nan"), dtype=torch.float32, device=dev)}
            res = vfleet.score(vs, xd, yd, precision=prec,
                               columns=("model-output", "tag-anomaly-unscaled", "total-anomaly-scaled"), out=init)
            tl = torch.as_tensor(np.asarray(te_lo, np.int64), device=dev)
            th = torch.as_tensor(np.asarray(te_hi, np.int64), device=dev)
            feat_pf = FFFleet.rolling_min_max(res["tag-anomaly-unscaled"], tl, th, 6).reshape(M, k, To)
            agg_pf = FFFleet.rolling_min_max(res["total-anomaly-scaled"], tl, th, 6).reshape(M, k)
            if is_kfcv:
                # thresholds = percentile of the smoothed validation errors over ALL rows of the Machine (diff.py:614-635)
                ml = torch.as_tensor(off[:-1].astype(np.int64), device=dev); mh = torch.as_tensor(off[1:].astype(np.int64), device=dev)
                det0 = protos[0]
                kf_feat = FFFleet.quantile(FFFleet.smooth(res["tag-anomaly-unscaled"], ml, mh, det0.smoothing_method, det0.window),
                                           ml, mh, det0.threshold_percentile)
                kf_agg = FFFleet.quantile(FFFleet.smooth(res["total-anomaly-scaled"], ml, mh, det0.smoothing_method, det0.window),
                                          ml, mh, det0.threshold_percentile)[:, 0]
            elif win is not None:       # the "smooth" thresholds: same statistic over rolling(window) (diff.py:241-248)
                sfeat_h = FFFleet.rolling_min_max(res["tag-anomaly-unscaled"], tl, th, win).reshape(M, k, To).double().cpu().numpy()
                sagg_h = FFFleet.rolling_min_max(res["total-anomaly-scaled"], tl, th, win).reshape(M, k).double().cpu().numpy()
            # the builder's CV metrics (build_model.py:245-289): scoring scaler = MinMaxScaler fitted on the
            # full y of each Machine = the final job's error scaler
            full_scale = err_scale[torch.arange(M, device=dev) * per + k].double().cpu().numpy()
            cv_metrics = FFFleet.cv_scores(ysrc, res["model-output"], tl, th, np.repeat(full_scale, k, axis=0))
        final = torch.arange(M, device=dev) * per + k
        fleet.set_params(params[final]); fleet.in_scale = in_scale[final].contiguous()
        fleet.in_min = in_min[final].contiguous(); fleet.err_scale = err_scale[final].contiguous()
        if k and is_kfcv:
            fleet.feat_thr = kf_feat.float().contiguous(); fleet.agg_thr = kf_agg.float().contiguous()
        elif k:
            fleet.feat_thr = feat_pf[:, -1].contiguous(); fleet.agg_thr = agg_pf[:, -1].contiguous()
        torch.cuda.current_stream().synchronize()
        t_total = time.time() - t_start
        self.last_fleet = fleet
        self.last_schedule_rows = rows

        # ---- materialise the per-Machine objects a sequential build would have produced
        P_host = fleet.params.cpu().numpy(); hl_h = hl.cpu().numpy(); ha_h = ha.cpu().numpy()
        xs_lo = (1.0 / in_scale.double()).cpu().numpy()           # data_range
        in_scale_h = in_scale.double().cpu().numpy(); in_min_h = in_min.double().cpu().numpy()
        es_h = err_scale.double().cpu().numpy()
        kf_feat_h = kf_feat.cpu().numpy() if (k and is_kfcv) else None
        kf_agg_h = kf_agg.cpu().numpy() if (k and is_kfcv) else None
        feat_h = feat_pf.double().cpu().numpy() if k else None
        agg_h = agg_pf.double().cpu().numpy() if k else None
        out = []
        for m, (mc, model) in enumerate(zip(mcs, protos)):
            j = m * per + k
            est = self._standard_parts(model)
            est.model = _Model(topo, P_host[m])
            est._history = History({"loss": [float(v) for v in hl_h[j]], "accuracy": [float(v) for v in ha_h[j]]},
                                   {"verbose": 0, "epochs": epochs, "steps": int(-(-rows[m] // batch))},
                                   list(range(epochs)))
            Xm = Xs[m]
            ym = Xm if (mc.y is None or mc.y is mc.X) else np.asarray(getattr(mc.y, "values", mc.y))
            _set_minmax(model.base_estimator.steps[0][1], in_scale_h[j], in_min_h[j], Xm, mc.X)
            ymin = ym.min(axis=0) if len(ym) else np.zeros(To)
            _set_minmax(model.scaler, es_h[j], -ymin * es_h[j], ym, mc.y if mc.y is not None else mc.X)
            tags = list(mc.y.columns) if isinstance(mc.y, pd.DataFrame) else (
                list(mc.X.columns) if isinstance(mc.X, pd.DataFrame) and alias else list(range(To)))
            if k and is_kfcv:
                model.feature_thresholds_ = pd.Series(kf_feat_h[m], index=tags, name=model.threshold_percentile)
                model.aggregate_threshold_ = float(kf_agg_h[m])
            elif k:
                model.feature_thresholds_per_fold_ = pd.DataFrame(feat_h[m], index=[f"fold-{i}" for i in range(k)], columns=tags)
                model.aggregate_thresholds_per_fold_ = {f"fold-{i}": float(agg_h[m, i]) for i in range(k)}
                model.feature_thresholds_ = pd.Series(feat_h[m, -1], index=tags, name=f"fold-{k - 1}")
                model.aggregate_threshold_ = float(agg_h[m, -1])
                _set_smooth_thresholds(model, None if sfeat_h is None else sfeat_h[m], None if sagg_h is None else sagg_h[m], tags, k)
            scores = _cv_score_dict({kk: v[m * k:(m + 1) * k] for kk, v in cv_metrics.items()}, tags) if k else {}
            # build_model.py:448-471: offset = len(X) - len(predict(X)); the feed-forward scorer emits one row per input row
            meta = {"name": mc.name, "model_offset": int(rows[m]) - fleet.out_rows(int(rows[m])), "model": extract_model_metadata(model),
                    "model_training_duration_sec": t_fit, "cv_duration_sec": (t_total - t_fit) if k else None,
                    "cross_validation": {"scores": scores,
                                         "splits": build_split_dict(mc.X, bounds=time_series_split_bounds(int(rows[m]), k)) if k else {}},
                    "fleet": {"machines_in_launch": M, "fit_jobs": J, "fit_duration_sec": t_fit,
                              "build_duration_sec": t_total},
                    "cv_fold_history": {f"fold-{i}": {"loss": [float(v) for v in hl_h[m * per + i]]} for i in range(k)}}
            out.append((model, meta))
        return out


def _set_smooth_thresholds(model, sfeat, sagg, tags, k):
    """The ``smooth_*`` threshold attributes of diff.py:241-264 ([k, T] / [k] arrays, or None without a window)."""
    if sfeat is None:
        model.smooth_feature_thresholds_per_fold_ = pd.DataFrame()
        model.smooth_aggregate_thresholds_per_fold_ = {}
        model.smooth_aggregate_threshold_ = None
        model.smooth_feature_thresholds_ = None
        return
    model.smooth_feature_thresholds_per_fold_ = pd.DataFrame(sfeat, index=[f"fold-{i}" for i in range(k)], columns=tags)
    model.smooth_aggregate_thresholds_per_fold_ = {f"fold-{i}": float(sagg[i]) for i in range(k)}
    model.smooth_feature_thresholds_ = pd.Series(sfeat[-1], index=tags, name=f"fold-{k - 1}")
    model.smooth_aggregate_threshold_ = float(sagg[-1])


def _build_bucket_lstm_impl(self, mcs, protos, dev):
    """
    Batched build of a bucket of LSTM Machines: every CV fold + final fit is a job of ONE
    gb200_lstm_fit call (lock-step batches, grid.z = jobs), the fold predictions one
    gb200_lstm_predict over the test ranges as virtual Machines, scored by gb200_score_outputs.
    """
    import torch
    from gordo_b200.lstm import LSTMFleet
    est0 = self._standard_parts(protos[0])
    topo = est0._topology()
    lookahead, L = est0.lookahead, topo.lookback_window
    epochs = int(est0.kwargs.get("epochs", 1)); batch = int(est0.batch_size)
    cv_mode = mcs[0].evaluation.get("cv_mode", "full_build")
    k = int(mcs[0].evaluation.get("n_splits", 3)) if cv_mode == "full_build" else 0
    M, T, To, per = len(mcs), topo.n_features, topo.n_features_out, k + 1
    Xs = [np.asarray(getattr(m.X, "values", m.X)) for m in mcs]
    Ys = [X if (m.y is None or m.y is m.X) else np.asarray(getattr(m.y, "values", m.y)) for m, X in zip(mcs, Xs)]
    rows = np.array([len(x) for x in Xs], np.int64)
    off = np.concatenate([[0], np.cumsum(rows)])
    xd = torch.as_tensor(np.ascontiguousarray(np.concatenate(Xs), np.float32), device=dev)
    yd = torch.as_tensor(np.ascontiguousarray(np.concatenate(Ys), np.float32), device=dev)
    lo, hi, te_lo, te_hi = [], [], [], []
    for m in range(M):
        for (s_, e_) in (time_series_split_bounds(int(rows[m]), k) if k else []):
            lo.append(off[m]); hi.append(off[m] + s_); te_lo.append(off[m] + s_); te_hi.append(off[m] + e_)
        lo.append(off[m]); hi.append(off[m] + rows[m])
    lo = np.asarray(lo, np.int64); hi = np.asarray(hi, np.int64)
    J = len(lo)
    lo_t = torch.as_tensor(lo, device=dev); hi_t = torch.as_tensor(hi, device=dev)
    t0 = time.time()
    in_scale, in_min = FFFleet.minmax_fit(xd, lo_t, hi_t)
    err_scale, _ = FFFleet.minmax_fit(yd, lo_t, hi_t)
    gen = torch.Generator(device=dev)
    params = torch.empty((J, topo.n_params), dtype=torch.float32, device=dev)
    for m, mc in enumerate(mcs):
        for i, sd in enumerate(job_seeds(int(mc.evaluation.get("seed", 0)), per)):
            gen.manual_seed(sd)
            params[m * per + i] = topo.init_params(1, gen, dev)[0]
    trainer = LSTMFleet(topo, J, lookahead, dev)
    hl, pl = trainer.fit_jobs(xd, yd, lo, hi, params, in_scale=in_scale, in_min=in_min, epochs=epochs, batch_size=batch)
    feat_h = agg_h = None
    if k:
        fold_jobs = np.array([m * per + i for m in range(M) for i in range(k)])
        sel = torch.as_tensor(fold_jobs, device=dev)
        vf = LSTMFleet(topo, M * k, lookahead, dev)
        vf.set_params(params[sel]); vf.in_scale = in_scale[sel].contiguous(); vf.in_min = in_min[sel].contiguous()
        vs = Schedule(rows_lo=te_lo, rows_hi=te_hi, rows_total=int(off[-1]))
        prec = "bf16" if (est0.kwargs.get("precision", "f32") == "bf16" and vf.tc_eligible()) else "f32"
        out, out_off = vf.predict(vs, xd, precision=prec)
        n_out = np.diff(out_off)
        y_off = np.asarray(te_hi, np.int64) - n_out                     # align to the LAST len(output) rows
        res = FFFleet.score_outputs(out, yd, out_off, y_off, err_scale=err_scale[sel].contiguous())
        ol = torch.as_tensor(out_off[:-1].copy(), device=dev); oh = torch.as_tensor(out_off[1:].copy(), device=dev)
        feat_h = FFFleet.rolling_min_max(res["tag-anomaly-unscaled"], ol, oh, 6).reshape(M, k, To).double().cpu().numpy()
        agg_h = FFFleet.rolling_min_max(res["total-anomaly-scaled"], ol, oh, 6).reshape(M, k).double().cpu().numpy()
        win = protos[0].window
        sfeat_h = sagg_h = None
        if win is not None:
            sfeat_h = FFFleet.rolling_min_max(res["tag-anomaly-unscaled"], ol, oh, win).reshape(M, k, To).double().cpu().numpy()
            sagg_h = FFFleet.rolling_min_max(res["total-anomaly-scaled"], ol, oh, win).reshape(M, k).double().cpu().numpy()
    torch.cuda.current_stream().synchronize()
    t_total = time.time() - t0
    P_host = params.cpu().numpy(); hl_h = hl.cpu().numpy(); pl_h = pl.cpu().numpy()
    in_scale_h = in_scale.double().cpu().numpy(); in_min_h = in_min.double().cpu().numpy(); es_h = err_scale.double().cpu().numpy()
    out_models = []
    for m, (mc, model) in enumerate(zip(mcs, protos)):
        j = m * per + k
        est = self._standard_parts(model)
        est.kwargs.update({"n_features": T, "n_features_out": To})
        est.model = _Model(topo, P_host[j])
        est._history = History({"loss": [float(pl_h[j])]}, {"verbose": 0, "epochs": 1, "steps": 1}, [0])     # primer (models.py:285)
        est.history_main_ = {"loss": [float(v) for v in hl_h[j]]}
        _set_minmax(model.base_estimator.steps[0][1], in_scale_h[j], in_min_h[j], Xs[m], mc.X)
        ymin = Ys[m].min(axis=0)
        _set_minmax(model.scaler, es_h[j], -ymin * es_h[j], Ys[m], mc.y if mc.y is not None else mc.X)
        tags = list(mc.y.columns) if isinstance(mc.y, pd.DataFrame) else (
            list(mc.X.columns) if isinstance(mc.X, pd.DataFrame) and To == T else list(range(To)))
        if k:
            model.feature_thresholds_per_fold_ = pd.DataFrame(feat_h[m], index=[f"fold-{i}" for i in range(k)], columns=tags)
            model.aggregate_thresholds_per_fold_ = {f"fold-{i}": float(agg_h[m, i]) for i in range(k)}
            model.feature_thresholds_ = pd.Series(feat_h[m, -1], index=tags, name=f"fold-{k - 1}")
            model.aggregate_threshold_ = float(agg_h[m, -1])
            _set_smooth_thresholds(model, None if sfeat_h is None else sfeat_h[m], None if sagg_h is None else sagg_h[m], tags, k)
        # build_model.py:448-471: len(X) - len(predict(X)), from the fleet's own output-row count
        meta = {"name": mc.name, "model_offset": int(rows[m]) - trainer.out_rows(int(rows[m])), "model": extract_model_metadata(model),
                "model_training_duration_sec": t_total, "cv_duration_sec": None,
                "cross_validation": {"scores": {}, "splits": build_split_dict(mc.X, bounds=time_series_split_bounds(int(rows[m]), k)) if k else {}},
                "fleet": {"machines_in_launch": M, "fit_jobs": J, "build_duration_sec": t_total}}
        out_models.append((model, meta))
    return out_models


FleetBuild._build_bucket_lstm = _build_bucket_lstm_impl


def _cv_score_dict(metrics: Dict[str, np.ndarray], tags) -> Dict[str, Dict[str, float]]:
    """
    ``scores`` in the layout of build_model.py:274-289: per metric ``{metric}-{tag}`` and the uniform
    average over tags ``{metric}``, each {fold-mean, fold-std, fold-max, fold-min, fold-1..k}.
    """
    out: Dict[str, Dict[str, float]] = {}

    def entry(vals):
        vals = np.asarray(vals, np.float64)
        d = {"fold-mean": float(vals.mean()), "fold-std": float(vals.std()), "fold-max": float(vals.max()),
             "fold-min": float(vals.min())}
        d.update({f"fold-{i + 1}": float(v) for i, v in enumerate(vals)})
        return d

    for name, per_tag in metrics.items():               # per_tag: [k folds, T]
        for j, tag in enumerate(tags):
            out[f"{name}-{str(tag).replace(' ', '-')}"] = entry(per_tag[:, j])
        out[name] = entry(per_tag.mean(axis=1))
    return out


def _set_minmax(scaler: MinMaxScaler, scale, min_, data, frame):
    """Give a sklearn MinMaxScaler the fitted state computed on the GPU."""
    scale = np.asarray(scale, np.float64); min_ = np.asarray(min_, np.float64)
    data = np.asarray(data)
    scaler.scale_ = scale
    scaler.min_ = min_
    scaler.data_min_ = data.min(axis=0).astype(np.float64) if len(data) else np.zeros_like(scale)
    scaler.data_max_ = data.max(axis=0).astype(np.float64) if len(data) else np.zeros_like(scale)
    scaler.data_range_ = scaler.data_max_ - scaler.data_min_
    scaler.n_features_in_ = data.shape[1]
    scaler.n_samples_seen_ = len(data)
    if isinstance(frame, pd.DataFrame) and all(isinstance(c, str) for c in frame.columns):
        scaler.feature_names_in_ = np.asarray(frame.columns, dtype=object)


# ======================================================================================== the builder seam
try:                                        # gordo installed: be a real ModelBuilder (builder/utils.py:8-17)
    from gordo.builder.build_model import ModelBuilder as _ReferenceModelBuilder
except Exception:                           # not installed (this image): same public contract, standalone
    _ReferenceModelBuilder = None

DEFAULT_METRICS = ("explained_variance_score", "r2_score", "mean_squared_error", "mean_absolute_error")


def redirect_definition(definition, enable: bool = True):
    """``gordo.machine.model.*`` class paths of a model definition -> ``gordo_b200.machine.model.*`` (deep copy)."""
    if not enable:
        return definition
    src, dst = "gordo.machine.model.", "gordo_b200.machine.model."
    if isinstance(definition, str):
        return dst + definition[len(src):] if definition.startswith(src) else definition
    if isinstance(definition, dict):
        return {redirect_definition(k): redirect_definition(v) for k, v in definition.items()}
    if isinstance(definition, (list, tuple)):
        return type(definition)(redirect_definition(v) for v in definition)
    return definition


def build_metadata_dict(meta: Dict[str, Any], builder_version: str = "gordo_b200") -> Dict[str, Any]:
    """A FleetBuild metadata record in the layout of gordo.machine.metadata.BuildMetadata (metadata.py:17-56)."""
    cv = meta.get("cross_validation", {})
    return {"model": {"model_offset": int(meta.get("model_offset", 0)),
                      "model_creation_date": str(datetime.datetime.now(datetime.timezone.utc).astimezone()),
                      "model_builder_version": builder_version,
                      "model_training_duration_sec": meta.get("model_training_duration_sec"),
                      "cross_validation": {"cv_duration_sec": meta.get("cv_duration_sec"),
                                           "scores": cv.get("scores", {}), "splits": cv.get("splits", {})},
                      "model_meta": meta.get("model", {})},
            "dataset": {"query_duration_sec": meta.get("query_duration_sec"), "dataset_meta": meta.get("dataset_meta", {})}}


class _StandaloneModelBuilder:
    """
    ``ModelBuilder``'s public contract (build_model.py:49-190) without gordo: ``machine`` is a ``FleetMachine``
    (data already fetched) or any object with ``name``, ``model``, ``evaluation`` and ``dataset.get_data()``.
    """

    def __init__(self, machine, back_compatibles=None, default_data_provider=None):
        self.machine = machine
        self.back_compatibles = back_compatibles
        self.default_data_provider = default_data_provider
        self._cached_model_path = None

    @property
    def cached_model_path(self):
        return self._cached_model_path

    @cached_model_path.setter
    def cached_model_path(self, value):
        self._cached_model_path = value

    @property
    def gordo_version(self):
        return "gordo_b200"

    def set_seed(self, seed: int):
        import random
        np.random.seed(seed)
        random.seed(seed)

    @staticmethod
    def _determine_offset(model, X) -> int:
        X = getattr(X, "values", X)
        out = model.predict(X) if hasattr(model, "predict") else model.transform(X)
        return len(X) - len(out)

    def build(self, output_dir=None, model_register_dir=None, replace_cache=False):
        """(model, machine) like ModelBuilder.build (build_model.py:104-190); the disk registry is gordo's own."""
        if model_register_dir:
            raise NotImplementedError("the model register (gordo.util.disk_registry) needs gordo itself")
        model, machine = self._build()
        if output_dir and getattr(machine, "evaluation", {}).get("cv_mode") != "cross_val_only":
            import json
            import os
            import pickle
            os.makedirs(output_dir, exist_ok=True)
            with open(os.path.join(output_dir, "model.pkl"), "wb") as f:      # serializer/serializer.py:189-196
                pickle.dump(model, f)
            with open(os.path.join(output_dir, "metadata.json"), "w") as f:
                json.dump({"name": machine.name, "metadata": {"build_metadata": machine.build_metadata}}, f, default=str)
            self.cached_model_path = output_dir
        return model, machine


class FleetModelBuilder(_ReferenceModelBuilder or _StandaloneModelBuilder):
    """
    ``--model-builder-class gordo_b200.builder.FleetModelBuilder`` (cli/cli.py:81-86,147-150).

    * ``FleetModelBuilder(machine).build(...)`` -- one Machine, the reference's contract: ``(model, machine)``.
      With gordo installed this IS ``ModelBuilder.build`` (dataset fetch, CV, fit, offset, metadata, cache,
      ``serializer.dump`` all inherited); ``gordo.machine.model.*`` class paths of the Machine's model are
      redirected to their ``gordo_b200`` twins first (``redirect_gordo``), so an unmodified project builds on the GPU.
    * ``FleetModelBuilder.build_fleet(machines)`` -- the whole project at once (FleetBuild), one entry per Machine.
    * ``FleetModelBuilder.build_sharded(machines, rank, world)`` -- one process per GPU, Machines dealt by cost.
    """

    redirect_gordo = True

    def _build(self):
        if _ReferenceModelBuilder is not None:
            if self.redirect_gordo:
                self.machine.model = redirect_definition(self.machine.model)
            return super()._build()
        mc = _as_fleet_machine(self.machine)
        model = serializer.from_definition(redirect_definition(mc.definition(), self.redirect_gordo))
        self.set_seed(int(mc.evaluation.get("seed", 0)))
        built, meta = FleetBuild([mc])._build_one(model, mc, metrics=True)
        mc.build_metadata = build_metadata_dict(meta, self.gordo_version)
        return built, mc

    # ------------------------------------------------------------------ the fleet
    @classmethod
    def build_fleet(cls, machines: Sequence[Any], device: Optional[str] = None, streams: int = 8,
                    cv_precision: str = "f32"):
        """
        Batched twin of ``for machine in machines: ModelBuilder(machine).build()`` (local_build.py:67-70).
        Returns ``[(model, machine)]``; ``machine.build_metadata`` (FleetMachine) or
        ``machine.metadata.build_metadata`` (gordo Machine) carries the reference's BuildMetadata record, and the
        raw FleetBuild record stays available as ``machine.fleet_metadata``.
        """
        fms = [_as_fleet_machine(m) for m in machines]
        for fm in fms:
            fm.model = redirect_definition(fm.definition(), cls.redirect_gordo)
        built = FleetBuild(fms, device=device, cv_precision=cv_precision, streams=streams).build()
        out = []
        for src, fm, (model, meta) in zip(machines, fms, built):
            record = build_metadata_dict(meta)
            target = src if not isinstance(src, FleetMachine) else fm
            if hasattr(target, "metadata") and hasattr(target.metadata, "build_metadata") and _ReferenceModelBuilder is not None:
                from gordo.machine.metadata import BuildMetadata
                target.metadata.build_metadata = BuildMetadata.from_dict(record)
            else:
                target.build_metadata = record
            target.fleet_metadata = meta
            out.append((model, target))
        return out

    @classmethod
    def local_build(cls, config_str: str, **kw):
        """``gordo.builder.local_build`` (local_build.py:14-70) with the Machines built as one fleet; needs gordo."""
        if _ReferenceModelBuilder is None:
            raise ImportError("FleetModelBuilder.local_build parses a gordo project config: gordo is not installed")
        import io
        from gordo.workflow.config_elements.normalized_config import NormalizedConfig
        from gordo.workflow.workflow_generator.workflow_generator import get_dict_from_yaml
        normed = NormalizedConfig(get_dict_from_yaml(io.StringIO(config_str)), project_name="local-build")
        return cls.build_fleet(normed.machines, **kw)

    # ------------------------------------------------------------------ one process per GPU
    @staticmethod
    def shard_indices(machines: Sequence[Any], world_size: int) -> List[List[int]]:
        """Machines dealt to ranks by estimated cost, largest first (partition.lpt)."""
        from gordo_b200.partition import lpt, machine_cost
        costs = []
        for m in machines:
            fm = _as_fleet_machine(m)
            n, T = np.shape(getattr(fm.X, "values", fm.X))[:2]
            d = str(fm.definition())
            lstm = "LSTM" in d
            look = 1
            if lstm:
                import re
                hit = re.search(r"lookback_window'?:\s*(\d+)", d)
                look = int(hit.group(1)) if hit else 1
            w = [T, max(1, round(0.83 * T)), max(1, round(0.67 * T)), max(1, round(0.5 * T))]
            f = 2.0 * 2 * sum(a * b for a, b in zip(w[:-1], w[1:]))
            costs.append(machine_cost(int(n), f * (4.0 if lstm else 1.0), 1, look))
        return lpt(costs, world_size)

    @classmethod
    def build_sharded(cls, machines: Sequence[Any], rank: int, world_size: int, *, gather: bool = False,
                      build_fn=None, **kw) -> Dict[int, Any]:
        """
        This rank's share of the project: ``{machine index: (model, machine)}``.  No data-path collective --
        Machines are independent; with ``gather=True`` (torch.distributed initialised) rank 0 receives every
        rank's results (pickled host objects), which is the host-side concat of SURVEY.md §8e.
        """
        mine = cls.shard_indices(machines, world_size)[rank]
        build = build_fn or (lambda ms: cls.build_fleet(ms, **kw))
        results = dict(zip(mine, build([machines[i] for i in mine])))
        if gather:
            import torch.distributed as dist
            parts = [None] * world_size if rank == 0 else None
            dist.gather_object(results, parts, dst=0)
            if rank == 0:
                results = {k: v for part in parts for k, v in part.items()}
        return results


def _as_fleet_machine(machine) -> FleetMachine:
    if isinstance(machine, FleetMachine):
        return machine
    if hasattr(machine, "dataset") and not hasattr(machine, "X"):
        ds = machine.dataset
        if _ReferenceModelBuilder is not None and not hasattr(ds, "get_data"):
            from gordo_core.base import GordoBaseDataset
            ds = GordoBaseDataset.from_dict(machine.dataset.to_dict())
        t0 = time.time()
        X, y = ds.get_data()
        fm = FleetMachine(name=machine.name, X=X, y=y, model=machine.model, evaluation=dict(machine.evaluation or {}))
        fm.query_duration_sec = time.time() - t0
        return fm
    return FleetMachine(name=machine.name, X=machine.X, y=getattr(machine, "y", None), model=getattr(machine, "model", None),
                        evaluation=dict(getattr(machine, "evaluation", {}) or {}))
