"""``evaluate.evaluate_interpolations`` (reference: evaluate/evaluate_interpolations.py:41-63)."""
from superresolution_aniso_mri_b200.model_selection import evaluate_interpolation_performance  # noqa: F401
