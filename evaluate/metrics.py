"""``evaluate.metrics`` (reference: evaluate/metrics.py:29-45, 65-108, 111-194, 210-242): SSIM / PSNR / VIF / LPIPS over
slice sets, on the device."""
from superresolution_aniso_mri_b200.evaluation import (  # noqa: F401
    compute_lpips_for_batch, compute_psnr_for_batch, compute_ssim_for_batch, compute_vif_for_batch, original_slice_ids,
    vif_slices)


def determine_original_sliceids(reference, downsample_steps, conv_interpol=False):
    return original_slice_ids(reference.shape[0], downsample_steps, conv_interpol)
