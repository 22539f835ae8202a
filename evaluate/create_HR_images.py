"""``evaluate.create_HR_images`` (reference: evaluate/create_HR_images.py:72-78, 110-196, 239-424): HR-volume creation
and scoring for the autoencoder path, metrics on the device."""
from superresolution_aniso_mri_b200.evaluation import compute_mean_metrics, compute_metrics  # noqa: F401
from superresolution_aniso_mri_b200.model_selection import (  # noqa: F401
    check_data_generator, create_hr_images, save_metrics_to_file)
