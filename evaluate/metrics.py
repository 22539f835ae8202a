"""``evaluate.metrics`` (reference: evaluate/metrics.py:29-45, 111-194): SSIM / PSNR over slice sets, on the device.
VIF and the LPIPS-as-metric wrapper are "next" rows and not provided."""
from superresolution_aniso_mri_b200.evaluation import (  # noqa: F401
    compute_psnr_for_batch, compute_ssim_for_batch, original_slice_ids)


def determine_original_sliceids(reference, downsample_steps, conv_interpol=False):
    return original_slice_ids(reference.shape[0], downsample_steps, conv_interpol)
