#!/usr/bin/env python
"""
Golden values of the builder's cross-validation metrics from the REAL reference (gordo/builder/build_model.py:377-446
`ModelBuilder.build_metrics_dict` + `metrics_from_list`, gordo/machine/model/utils.py:18-46 `metric_wrapper`), executed
unmodified from /root/reference through tests/reference_loader.py (TensorFlow / gordo-core stubbed; none of the stubbed
packages takes part in this arithmetic).  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_metrics_golden.py      ->  tests/golden/builder_metrics_golden.json
"""
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from tests import reference_loader
    ref = reference_loader.load()
    MB = ref["build_model"].ModelBuilder
    rng = np.random.default_rng(2024)
    n, T, offset = 240, 4, 3
    cols = ["TAG 1", "tag-2.PV", "GRA-TE  -23-0733.PV", "plain"]
    y = pd.DataFrame(rng.normal(5, 3, (n, T)) * np.array([1.0, 10.0, 0.1, 100.0]), columns=cols)
    y_pred = (y.to_numpy() + rng.normal(0, 1, (n, T)) * np.array([0.5, 4.0, 0.05, 60.0]))[offset:]      # a model with an offset (LSTM)

    from sklearn.base import BaseEstimator, RegressorMixin

    class Dummy(RegressorMixin, BaseEstimator):
        def fit(self, X, y=None):
            return self

        def predict(self, X):
            return y_pred

    metrics_list = MB.metrics_from_list(None)
    scorers = MB.build_metrics_dict(metrics_list, y, scaler="sklearn.preprocessing.MinMaxScaler")
    values = {k: float(s(Dummy(), y, y)) for k, s in scorers.items()}
    # split metadata (build_model.py:347-375) for a time-indexed frame under the builder's TimeSeriesSplit(3) and a KFold
    from sklearn.model_selection import KFold, TimeSeriesSplit
    Xdt = pd.DataFrame(y.to_numpy(), columns=cols, index=pd.date_range("2020-01-01", periods=n, freq="10min", tz="UTC"))
    splits = {"tss3": {k: str(v) for k, v in MB.build_split_dict(Xdt, TimeSeriesSplit(n_splits=3)).items()},
              "kfold4": {k: str(v) for k, v in MB.build_split_dict(Xdt, KFold(n_splits=4)).items()},
              "tss3_rangeindex": {k: str(v) for k, v in MB.build_split_dict(y, TimeSeriesSplit(n_splits=3)).items()}}
    # model metadata extraction (build_model.py:516-570) over nested estimators
    sys.path.insert(0, HERE)
    from metadata_structures import build_structures
    GordoBase = sys.modules["gordo.machine.model.base"].GordoBase
    extracted = {name: MB._extract_metadata_from_model(model) for name, model in build_structures(GordoBase).items()}
    out = {"columns": cols, "offset": offset, "y": y.to_numpy().tolist(), "y_pred": y_pred.tolist(),
           "metrics": [m.__name__ for m in metrics_list], "values": values, "splits": splits, "extracted_metadata": extracted}
    path = os.path.join(HERE, "builder_metrics_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {len(values)} scorers, e.g.", {k: values[k] for k in list(values)[:3]})


if __name__ == "__main__":
    main()
