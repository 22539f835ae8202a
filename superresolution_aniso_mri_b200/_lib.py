"""ctypes binding of libaesr_b200.so (C-ABI declared in include/aesr_b200.h).

The product path has no CPU fallback: if the shared library is missing, or the device is not a compute-capability
10.x GPU, every operator raises ``RuntimeError`` -- loudly, never a silent eager-PyTorch substitute.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# AESR_B200_LIB: another build of the same library (A/B timing of kernel variants on one box, tools/gpu_*.sh)
LIB_PATH = os.environ.get("AESR_B200_LIB") or os.path.join(_HERE, "lib", "libaesr_b200.so")

_lib = None
_initialised_devices = set()
P, I, F = c_void_p, c_int, c_float

_SIGNATURES = {
    "aesr_init": (I, [I]),
    "aesr_last_error": (c_char_p, []),
    "aesr_sm_count": (I, []),
    "aesr_set_tuning": (I, [I, I]),
    "aesr_launch_count": (c_int64, []),
    "aesr_pack_conv3x3_weight": (I, [P, P, I, I, I, I, P]),
    "aesr_pack_conv3x3_weight_batch": (I, [P, P, P, I, I, I, P]),
    "aesr_conv3x3_fwd": (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, F, I, I, I, I, I, P]),
    "aesr_pack_conv3x3_weight_up2fold": (I, [P, P, I, I, I, P]),
    "aesr_conv3x3_up2_head_fwd": (I, [P, P, P, P, P, P, I, I, I, I, I, F, I, I, P]),
    "aesr_head_gather": (I, [P, P, P, P, I, I, I, c_size_t, I, P]),
    "aesr_stem_fold": (I, [P, P, P, P, P, I, P]),
    "aesr_stem_fwd": (I, [P, P, P, I, I, I, I, F, I, P]),
    "aesr_e0_fwd": (I, [P, P, P, P, I, I, I, I, I, P]),
    "aesr_head_fwd": (I, [P, P, P, P, P, I, I, I, I, c_size_t, I, I, P]),
    "aesr_lerp_latents": (I, [P, P, P, P, P, P, P, I, I, I, I, P]),
    "aesr_place_slices": (I, [P, P, P, I, I, I, P]),
    "aesr_copy_rows_async": (I, [P, c_size_t, c_size_t, P, c_size_t, c_size_t, c_size_t, c_size_t, c_size_t, I, P]),
    "aesr_lerp_pairs": (I, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "aesr_lerp_pairs_act": (I, [P, P, P, P, P, P, P, I, I, I, I, F, I, P]),
    # training step
    "aesr_bn_finalize": (I, [P, F, F, I, P, P, P, P, F, F, P, P, P, P, I, P]),
    "aesr_bn_apply": (I, [P, P, P, P, I, I, I, I, I, I, I, P]),
    "aesr_bn_bwd": (I, [P, P, P, P, P, P, F, P, P, P, P, I, I, I, I, I, I, I, F, F, I, P]),
    "aesr_mse": (I, [P, P, c_size_t, P, P, F, P]),
    "aesr_head_bwd": (I, [P, P, P, P, P, P, P, P, I, I, I, I, F, I, P]),
    "aesr_e0_bwd": (I, [P, P, P, P, I, I, I, I, P]),
    "aesr_wgrad3x3": (I, [P, P, P, P, I, I, I, I, I, I, I, P]),
    "aesr_mix_bwd": (I, [P, P, P, P, P, I, c_size_t, P]),
    "aesr_adam_step": (I, [P, P, P, P, c_size_t, F, F, F, F, F, I, P]),
    "aesr_adam_step_dev": (I, [P, P, P, P, c_size_t, F, F, F, F, F, P, P, P]),
    "aesr_vgg_conv1_fwd": (I, [P, P, P, P, I, I, I, P, P, I, I, P]),
    "aesr_vgg_conv1_bwd": (I, [P, P, P, I, I, I, P, I, F, P]),
    "aesr_maxpool_bwd": (I, [P, P, P, P, I, I, I, I, I, P]),
    "aesr_lpips_head": (I, [P, P, P, P, P, P, I, I, I, I, P]),
    # evaluation / data path
    "aesr_ssim_psnr": (I, [P, P, I, I, I, I, ctypes.c_double, P, P, P, P]),
    "aesr_vif_workspace_bytes": (c_size_t, [I, I, I]),
    "aesr_vif_quantize_u8": (I, [P, P, c_size_t, P]),
    "aesr_vif_mscale": (I, [P, P, I, I, I, P, P, ctypes.c_double, P, c_size_t, P, P]),
    "aesr_percentile_workspace_bytes": (c_size_t, []),
    "aesr_percentile_normalize": (I, [P, P, c_size_t, ctypes.c_double, ctypes.c_double, P, c_size_t, P, P]),
    "aesr_pad_crop_gather": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "aesr_augment_gather": (I, [P, P, P, P, P, P, P, ctypes.c_uint, I, I, I, I, I, P]),
    "aesr_gauss1d_axis0": (I, [P, P, P, I, I, c_size_t, P]),
}


# diagnostic entry points of lib/libaesr_b200_probe.so (include/aesr_b200_probe.h; not in the product library)
_PROBE_SIGNATURES = {
    "aesr_probe_launch_gap": (I, [P, I, I, I, I, I]),
    "aesr_probe_halo_conv": (I, [P, P, P, I, I, I, I, I, I, I, I, P]),
    "aesr_probe_umma_rate": (I, [P, I, I, I, I, I, I, I, I, I, P]),
    "aesr_probe_umma_pattern": (I, [P, I, I, I, I, I, I, I, I, P]),
    "aesr_probe_sync": (I, [P, I, I, P]),
    "aesr_probe_tmem_ld": (I, [P, I, I, I, I, P, P]),
}
_probe_lib = None


def load_probe():
    """The diagnostic build (product entry points + aesr_probe_*), built on demand: tools only."""
    global _probe_lib
    if _probe_lib is None:
        from . import build
        lib = ctypes.CDLL(build.build_library(probes=True))
        for name, (res, args) in {**_SIGNATURES, **_PROBE_SIGNATURES}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _probe_lib = lib
    return _probe_lib


def exported_symbols():
    """Names of every entry point include/aesr_b200.h declares (checked by tests/test_cabi.py)."""
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "aesr_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU / eager fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().aesr_last_error()
        raise RuntimeError("aesr_b200 %s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def lib_for_device(device_index: int):
    """Library handle with ``aesr_init(device)`` done (checks sm_100a, resolves the TMA encoder)."""
    lib = load()
    if device_index not in _initialised_devices:
        check(lib.aesr_init(int(device_index)), "init")
        _initialised_devices.add(device_index)
    return lib


def launch_count() -> int:
    return int(load().aesr_launch_count())
