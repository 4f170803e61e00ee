"""AnomalyDetectorBase -- gordo/machine/model/anomaly/base.py:11-23."""
import abc
from datetime import timedelta
from typing import Optional

import pandas as pd
from sklearn.base import BaseEstimator

from gordo_b200.machine.model.base import GordoBase


class AnomalyDetectorBase(BaseEstimator, GordoBase, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def anomaly(self, X: pd.DataFrame, y: pd.DataFrame, frequency: Optional[timedelta] = None) -> pd.DataFrame:
        """Take X, y and optionally frequency; return a dataframe containing anomaly score(s)."""
