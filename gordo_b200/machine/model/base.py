"""
GordoBase -- the estimator protocol of gordo/machine/model/base.py:10-35.  When gordo itself is
importable the reference's ABC is re-exported so isinstance checks in gordo.builder / gordo.server
hold; otherwise an identical ABC is defined here.
"""
import abc

try:                                            # pragma: no cover - gordo is absent in the build container
    from gordo.machine.model.base import GordoBase  # type: ignore
except Exception:

    class GordoBase(abc.ABC):
        @abc.abstractmethod
        def __init__(self, **kwargs):
            """Initialize the model"""

        @abc.abstractmethod
        def get_params(self, deep=False):
            """Return a dict containing all parameters used to initialized object"""

        @abc.abstractmethod
        def score(self, X, y, sample_weight=None):
            """Score the model; must implement the correct default scorer based on model type"""

        @abc.abstractmethod
        def get_metadata(self):
            """Get model specific metadata, if any"""
