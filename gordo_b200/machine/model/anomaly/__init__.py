from .diff import DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector

__all__ = ["DiffBasedAnomalyDetector", "DiffBasedKFCVAnomalyDetector"]
