"""``kwatsch.cardiac.trainer_ae`` (reference: kwatsch/cardiac/trainer_ae.py:8-182)."""
from superresolution_aniso_mri_b200.trainers import AETrainerEndToEnd  # noqa: F401
