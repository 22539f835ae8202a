"""Drop-in import paths of the evaluation helpers on the hot path (reference: evaluate/metrics.py, evaluate/common.py)."""
