"""``lpips.perceptual`` (reference: lpips/perceptual.py:6-33).  Implementation: superresolution_aniso_mri_b200.lpips_b200."""
from superresolution_aniso_mri_b200.lpips_b200 import PerceptualLoss  # noqa: F401
