"""``evaluate.find_best_model`` (reference: evaluate/find_best_model.py:25-137): checkpoint sweep on validation volumes.

  python -m evaluate.find_best_model --exper_dir DIR --epoch_range 500 900 --eval_patch_size 128 --data VOLUMES.npy

The reference builds its validation generator from dataset folders through SimpleITK / nibabel readers that are outside
the path (and absent in this image); here ``--data`` names a ``.npy`` / ``.npz`` of validation volumes [V,Z,H,W] in [0,1].
"""
import argparse
import os

import numpy as np

from superresolution_aniso_mri_b200.model_selection import (  # noqa: F401
    find_best_val_model, get_transforms, load_model_scores, store_top_scores)


def volumes_generator(path):
    arr = np.load(os.path.expanduser(path))
    vols = arr[arr.files[0]] if hasattr(arr, "files") else arr
    if vols.ndim == 3:
        vols = vols[None]
    return {i: {"image": np.asarray(v, dtype=np.float32), "patient_id": "vol%04d" % i,
                "spacing": np.array([1.0, 1.0, 1.0])} for i, v in enumerate(vols)}


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='Find best SR model')
    parser.add_argument('--epoch_range', type=int, nargs=2, default=[200, 201])
    parser.add_argument('--exper_dir', type=str, required=True)
    parser.add_argument('--eval_patch_size', type=int, default=None)
    parser.add_argument('--eval_axis', type=int, default=0)
    parser.add_argument('--downsample_steps', type=int, default=None)
    parser.add_argument('--data', type=str, required=True, help=".npy / .npz with validation volumes [V,Z,H,W] in [0,1]")
    args = parser.parse_args()
    epochs = np.arange(args.epoch_range[0], args.epoch_range[1] + 1)
    existing = [e for e in epochs if os.path.exists(os.path.join(os.path.expanduser(args.exper_dir), "models",
                                                                 "%d.models" % e))]
    find_best_val_model(volumes_generator(args.data), args.exper_dir, existing, ps_evaluate=args.eval_patch_size,
                        downsample_steps=args.downsample_steps, eval_axis=args.eval_axis)
