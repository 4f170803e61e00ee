"""
Oracle: [3P] sklearn.preprocessing.MinMaxScaler(feature_range=(0, 1)) and
sklearn.model_selection.TimeSeriesSplit (test infrastructure, see oracle/__init__.py).

Call sites in the reference: examples/config.yaml:75-82 (the Pipeline's input scaler),
gordo/machine/model/anomaly/diff.py:25,173 (the detector's error scaler, fitted on y after
training), builder/build_model.py:257-262 (TimeSeriesSplit(n_splits=3)).
Pinned against scikit-learn itself in tests/test_oracle_golden.py.
"""
import numpy as np


class MinMaxScaler:
    def fit(self, X):
        X = np.asarray(X, np.float64)
        self.data_min_ = np.nanmin(X, axis=0)
        self.data_max_ = np.nanmax(X, axis=0)
        rng = self.data_max_ - self.data_min_
        # sklearn _handle_zeros_in_scale: a constant column gets scale 1
        rng = np.where(rng < 10 * np.finfo(np.float64).eps, 1.0, rng)
        self.scale_ = 1.0 / rng
        self.min_ = 0.0 - self.data_min_ * self.scale_
        return self

    def transform(self, X):
        return np.asarray(X, np.float64) * self.scale_ + self.min_

    def fit_transform(self, X):
        return self.fit(X).transform(X)


def time_series_split(n_samples, n_splits=3):
    """Yield (train_idx, test_idx) exactly as sklearn TimeSeriesSplit(n_splits) does."""
    n_folds = n_splits + 1
    if n_folds > n_samples:
        raise ValueError("Cannot have number of folds greater than the number of samples")
    test_size = n_samples // n_folds
    idx = np.arange(n_samples)
    for test_start in range(n_samples - n_splits * test_size, n_samples, test_size):
        yield idx[:test_start], idx[test_start:test_start + test_size]


def kfold_split(n_samples, n_splits=5, seed=0):
    """
    Yield (train_idx, test_idx) as sklearn ``KFold(n_splits, shuffle=True, random_state=seed)`` does
    (the default cv of DiffBasedKFCVAnomalyDetector.cross_validate, diff.py:568): one Mersenne
    shuffle of arange(n), cut into folds of size n//k (+1 for the first n%k), each side returned in
    ascending order.
    """
    idx = np.arange(n_samples)
    np.random.RandomState(seed).shuffle(idx)
    sizes = np.full(n_splits, n_samples // n_splits, dtype=int)
    sizes[: n_samples % n_splits] += 1
    start = 0
    for s in sizes:
        mask = np.zeros(n_samples, dtype=bool)
        mask[idx[start:start + s]] = True
        yield np.where(~mask)[0], np.where(mask)[0]
        start += s


def shuffle_rows(n_samples, seed=0):
    """Row order of ``sklearn.utils.shuffle(X, y, random_state=seed)`` (diff.py:168-170)."""
    idx = np.arange(n_samples)
    np.random.RandomState(seed).shuffle(idx)
    return idx
