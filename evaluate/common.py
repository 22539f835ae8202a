"""``evaluate.common`` (reference: evaluate/common.py:36-39, 134-235): the evaluation twin of the synthesis loop."""
from superresolution_aniso_mri_b200.synthesis import create_super_volume_eval as create_super_volume  # noqa: F401
from superresolution_aniso_mri_b200.synthesis import latent_space_interp  # noqa: F401


def determine_last_slice(orig_num_slices, downsample_steps):
    return ((orig_num_slices - 1) // downsample_steps) * downsample_steps
