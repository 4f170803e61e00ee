"""Nested estimator structures for the metadata-extraction golden: built once on the reference's GordoBase
(make_metrics_golden.py) and once on the mirror's (tests/test_oracle_golden.py)."""
from sklearn.base import BaseEstimator
from sklearn.compose import TransformedTargetRegressor
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler


def build_structures(GordoBase):
    class Leaf(GordoBase, BaseEstimator):
        def __init__(self, tag="leaf", extra=None):
            self.tag, self.extra = tag, extra

        def fit(self, X, y=None):
            return self

        def predict(self, X):
            return X

        def score(self, X, y=None):
            return 0.0

        def get_params(self, deep=False):
            return {"tag": self.tag, "extra": self.extra}

        def get_metadata(self):
            return {"history": {"loss": [3.0, 2.0], "params": {"epochs": 2}}, f"seen-{self.tag}": True}

    class Detector(GordoBase, BaseEstimator):
        def __init__(self, base_estimator=None, scaler=None):
            self.base_estimator, self.scaler = base_estimator, scaler

        def fit(self, X, y=None):
            return self

        def predict(self, X):
            return X

        def score(self, X, y=None):
            return 0.0

        def get_params(self, deep=False):
            return {"base_estimator": self.base_estimator, "scaler": self.scaler}

        def get_metadata(self):
            return {"feature-thresholds": [0.1, 0.2], "aggregate-threshold": 0.5, "history": {"loss": [9.0]}}

    ttr = TransformedTargetRegressor(regressor=Pipeline([("s", MinMaxScaler()), ("m", Leaf("inner"))]), transformer=MinMaxScaler())
    ttr.regressor_ = Pipeline([("s", MinMaxScaler()), ("m", Leaf("fitted-clone"))])      # what fit() leaves; `regressor` is skipped
    return {
        "leaf": Leaf(),
        "pipeline_last_step": Pipeline([("s", MinMaxScaler()), ("m", Leaf("last"))]),
        "pipeline_gordo_not_last": Pipeline([("m", Leaf("first")), ("s", MinMaxScaler())]),
        "detector_over_pipeline": Detector(Pipeline([("s", MinMaxScaler()), ("m", Leaf("p"))]), MinMaxScaler()),
        "detector_over_leaf_with_nested": Detector(Leaf("outer", extra=Leaf("nested")), MinMaxScaler()),
        "detector_over_ttr": Detector(ttr, MinMaxScaler()),
    }
