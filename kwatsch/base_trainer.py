"""``kwatsch.base_trainer`` (reference: kwatsch/base_trainer.py).  Implementation: superresolution_aniso_mri_b200.trainers."""
from superresolution_aniso_mri_b200.trainers import BaseTrainer  # noqa: F401
