"""``kwatsch.brain.trainer_ae`` (reference: kwatsch/brain/trainer_ae.py:47-282)."""
from superresolution_aniso_mri_b200.trainers import AETrainerBrain, AETrainerExtension1Brain  # noqa: F401
