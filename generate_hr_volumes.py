"""``generate_hr_volumes.py`` -- drop-in CLI + functions for HR volume generation (reference: generate_hr_volumes.py).

Same flags (--exper_dir --model_nbr --num_interpolations --data_input_dir --output_dir --save), same functions
(create_super_volume, latent_space_interp, normalize_img, sitk_to_torch); the arithmetic runs in the sm_100a kernels of
superresolution_aniso_mri_b200.  Volume IO uses SimpleITK when it is installed (as the reference does); where it is absent
``.nii`` / ``.nii.gz`` / ``.mha`` / ``.mhd`` files go through the small reader / writer of
``superresolution_aniso_mri_b200.volume_io`` (same spacing bookkeeping, :177-182), and ``.npy`` volumes [z,y,x] are accepted.
"""
import argparse
import os
from pathlib import Path

import numpy as np
import torch

from superresolution_aniso_mri_b200 import volume_io
from superresolution_aniso_mri_b200.synthesis import create_super_volume, latent_space_interp  # noqa: F401

try:                                    # IO only -- no arithmetic on the path
    import SimpleITK as sitk
except ImportError:                     # pragma: no cover
    sitk = None


def normalize_img(img: np.ndarray, perc=(1, 99)) -> np.ndarray:
    """generate_hr_volumes.py:130-133 (float64 percentiles, clip to [0,1])."""
    min_val, max_val = np.percentile(img, perc)
    return ((img.astype(img.dtype) - min_val) / (max_val - min_val)).clip(0, 1)


def array_to_torch(np_img: np.ndarray) -> torch.Tensor:
    """generate_hr_volumes.py:104-111 after the SimpleITK read: [z,y,x] -> [z,1,y,x] fp32, normalised if needed."""
    np_img = np_img.astype(np.float32)
    if np_img.max() > 1 or np_img.min() < 0:
        np_img = normalize_img(np_img)
    return torch.from_numpy(np_img).float().unsqueeze(dim=1)


def sitk_to_torch(input_image) -> torch.Tensor:
    return array_to_torch(sitk.GetArrayFromImage(input_image))


def numpy_to_sitk(image_resolved, input_image, new_spacing=None):
    """generate_hr_volumes.py:114-127."""
    if image_resolved.ndim == 4:
        image_resolved = sitk.JoinSeries([sitk.GetImageFromArray(image_resolved[v], False)
                                          for v in range(image_resolved.shape[0])])
    else:
        image_resolved = sitk.GetImageFromArray(image_resolved)
    image_resolved.SetOrigin(input_image.GetOrigin())
    image_resolved.SetDirection(input_image.GetDirection())
    image_resolved.SetSpacing(new_spacing if new_spacing is not None else input_image.GetSpacing())
    return image_resolved


def load_images(input_dir: Path, suffix='.nii*'):
    """generate_hr_volumes.py:136-148 (+ .npy volumes)."""
    file_list = sorted(input_dir.rglob("*" + suffix)) or sorted(input_dir.rglob("*.mha")) or \
        sorted(input_dir.rglob("*.mhd")) or sorted(input_dir.rglob("*.npy"))
    if len(file_list) == 0:
        raise FileNotFoundError("Error - no files found in {} with extensions nii, mha, mhd, npy".format(input_dir))
    images = []
    for fname in file_list:
        if fname.suffix == ".npy":
            images.append((fname, np.load(str(fname))))
        elif sitk is None:
            images.append((fname, volume_io.read_volume(fname)))
        else:
            images.append((fname, sitk.ReadImage(str(fname))))
    return images


def synthesize_image(trainer, image, num_interpolations):
    """Per-file body of the reference main() (generate_hr_volumes.py:159-183)."""
    alpha_range = np.linspace(0, 1, num_interpolations + 2, endpoint=True)[1:-1]
    if isinstance(image, (np.ndarray, volume_io.Volume)):
        arr = image if isinstance(image, np.ndarray) else image.array
        frames = [arr] if arr.ndim == 3 else [arr[f] for f in range(arr.shape[0])]
        vols = [create_super_volume(trainer, array_to_torch(f), alpha_range, use_original=True)["upsampled_image"]
                .numpy().squeeze() for f in frames]
        np_img_hr = vols[0] if arr.ndim == 3 else np.stack(vols)
        if isinstance(image, np.ndarray):
            return np_img_hr
        # generate_hr_volumes.py:177-180: z spacing / (ni + 1); a 4-D series keeps 1 as its last spacing
        sp = image.GetSpacing()[:3]
        new_spacing_z = (sp[-1] / (num_interpolations + 1),) if arr.ndim == 3 else (sp[-1] / (num_interpolations + 1), 1,)
        new_spacing = np.asarray(sp[:2] + new_spacing_z).astype(np.float64)
        return volume_io.Volume(array=np_img_hr, spacing=tuple(float(v) for v in new_spacing), origin=image.origin,
                                direction=image.direction, fmt=image.fmt, header=image.header, byteorder=image.byteorder)
    num_frames = 1 if len(image.GetSize()) == 3 else image.GetSize()[-1]
    vols = []
    for f_id in range(num_frames):
        img = image if num_frames == 1 else image[:, :, :, int(f_id)]
        res = create_super_volume(trainer, sitk_to_torch(img), alpha_range, use_original=True, labels=None)
        vols.append(res["upsampled_image"].detach().cpu().numpy().squeeze())
    new_spacing_z = img.GetSpacing()[-1] / (num_interpolations + 1)
    new_spacing_z = (new_spacing_z,) if num_frames == 1 else (new_spacing_z, 1,)
    new_spacing = np.asarray(img.GetSpacing()[:2] + new_spacing_z).astype(np.float64)
    np_img_hr = vols[0] if num_frames == 1 else np.stack(vols)
    return numpy_to_sitk(np_img_hr, image, new_spacing=new_spacing)


def _source_spacing(vol):
    """Spacing recorded in the NIfTI header a Volume still carries (the file it was read from); its own otherwise."""
    if vol.fmt == "nifti" and vol.header is not None:
        nd = vol.array.ndim
        pix = np.frombuffer(vol.header, np.dtype("f4").newbyteorder(vol.byteorder), 8, 76)
        return tuple(float(abs(v)) for v in pix[1:nd + 1])
    return vol.spacing


def main(args, trainer, input_images, output_dir):
    images_hr = []
    for (fname, img) in input_images:
        images_hr.append((output_dir / fname.name, synthesize_image(trainer, img, args.num_interpolations)))
    return images_hr


def save_images(images_hr):
    for (fname, img) in images_hr:
        if isinstance(img, np.ndarray):
            np.save(str(fname), img)
        elif isinstance(img, volume_io.Volume):
            # `img` carries the source header with the NEW spacing already in .spacing: rescale from the source's spacing
            src = volume_io.Volume(array=img.array, spacing=_source_spacing(img), origin=img.origin, direction=img.direction,
                                   fmt=img.fmt, header=img.header, byteorder=img.byteorder)
            volume_io.write_volume(fname, img.array, like=src, spacing=img.spacing)
        else:
            sitk.WriteImage(img, str(fname))
        print("Save image HR {}".format(str(fname)))


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='Generate HR volumes')
    parser.add_argument('--exper_dir', type=str, default=None)
    parser.add_argument('--model_nbr', type=int, default=None)
    parser.add_argument('--num_interpolations', type=int, default=6)
    parser.add_argument('--data_input_dir', type=str, default=None)
    parser.add_argument('--output_dir', type=str, default=None)
    parser.add_argument('--save', action='store_true')
    cli = parser.parse_args()
    if cli.output_dir is None:
        cli.output_dir = cli.exper_dir + os.sep + "ni0{}".format(cli.num_interpolations)
    out_dir = Path(cli.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    inputs = load_images(Path(cli.data_input_dir))
    print("INFO - Found {} files to process in {}".format(len(inputs), cli.data_input_dir))
    from kwatsch.get_trainer import get_trainer_dynamic
    the_trainer, _ = get_trainer_dynamic(src_path=cli.exper_dir, model_nbr=cli.model_nbr, model_nbr_sr=None,
                                         eval_mode=True)
    results = main(cli, the_trainer, inputs, out_dir)
    if cli.save:
        save_images(results)
