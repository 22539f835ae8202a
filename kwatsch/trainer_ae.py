"""``kwatsch.trainer_ae`` (reference: kwatsch/trainer_ae.py:16-109)."""
from superresolution_aniso_mri_b200.trainers import AEBaseTrainer  # noqa: F401
