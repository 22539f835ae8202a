"""Drop-in import path ``lpips.perceptual.PerceptualLoss`` (reference: lpips/perceptual.py)."""
