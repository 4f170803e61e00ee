"""Response codecs of the reference's server for the anomaly path (gordo/server/utils.py:47-195)."""
