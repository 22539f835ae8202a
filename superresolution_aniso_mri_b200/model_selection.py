"""The model-selection loop around the synthesis path: for every checkpoint of an
experiment, synthesise the dropped slices of every validation volume and score them.

Drop-ins for ``evaluate/create_HR_images.py::create_hr_images`` (:239-424, the autoencoder branch: ``interpol_filter``
is None), ``evaluate/evaluate_interpolations.py::evaluate_interpolation_performance`` (:41-63) and
``evaluate/find_best_model.py::find_best_val_model / store_top_scores / load_model_scores / get_transforms``
(:25-131) -- same signatures, result dictionaries and ``model_perf_*.npz`` files.  With synthesis, SSIM / PSNR and VIF
all on the device (``synthesis.create_super_volume_eval``, ``evaluation.compute_metrics``) a volume never returns to the
host between the decoder and the metric kernels; only the per-slice scores do.

Out of scope here: conventional interpolation baselines (SimpleITK / cv2 filters), NIfTI
output through SimpleITK (absent in this image: ``save_volumes`` writes ``.npy`` + a spacing sidecar instead) and label
volumes.
"""
from __future__ import annotations

import glob
import os
import types
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import evaluation as E
from . import synthesis


def check_data_generator(data_generator):
    """evaluate/create_HR_images.py:72-78."""
    if not isinstance(data_generator, types.GeneratorType):
        if isinstance(data_generator, dict):
            data_generator = data_generator.values()
        else:
            raise ValueError("ERROR - create_hr_images - data_generator is not a generator nor a dict")
    return data_generator


def save_metrics_to_file(result_dict, fname):
    """evaluate/create_HR_images.py:181-196."""
    keys = [m + s for s in ("", "_synth", "_recon") for m in ("ssim", "psnr", "vif", "lpips")]
    np.savez(fname, **{k: np.array(result_dict[k]) for k in keys})
    print("INFO - Saved results to {}".format(fname))


def _save_volume(new_images, pat_id, output_dir, file_suffix, myargs, spacing, origin, direction):
    pat_output_dir = os.path.join(output_dir, pat_id)
    os.makedirs(pat_output_dir, exist_ok=True)
    stem = pat_id + "_{}".format(myargs['model'] if file_suffix is None else file_suffix)
    np.save(os.path.join(pat_output_dir, stem + ".npy"), np.asarray(new_images, dtype=np.float32))
    np.savez(os.path.join(pat_output_dir, stem + "_geometry.npz"), spacing=np.asarray(spacing, dtype=np.float64),
             origin=np.asarray(origin if origin is not None else []), direction=np.asarray(direction if direction is not None else []))
    print("INFO - saved {}".format(os.path.join(pat_output_dir, stem + ".npy")))


@torch.no_grad()
def create_hr_images(data_generator, myargs, trainer=None, num_interpolations=None, downsample_steps=None,
                     is_4d=False, transform=None, expand_factor=None, use_original_slice=False,
                     normalize=False, generate_inbetween_slices=False, file_suffix=None, patient_id=None,
                     interpol_filter=None, save_volumes=False, output_dir=None, verbose=False,
                     compute_percept_loss=False, base_out_dir=None, eval_axis=0, is_arvc_labels=False,
                     resample=False):
    """evaluate/create_HR_images.py:239-424 for the autoencoder (``interpol_filter=None``).  Returns the result dict of
    per-volume metric lists when ``generate_inbetween_slices`` (the validation protocol: drop slices, synthesise them,
    score against the originals), else ``(None, None, None)`` like the reference."""
    assert (num_interpolations is not None or downsample_steps is not None)
    if interpol_filter is not None:
        raise NotImplementedError("aesr_b200: conventional interpolation baselines (SimpleITK / cv2) are not part of the "
                                  "autoencoder path")
    if resample or is_arvc_labels:
        raise NotImplementedError("aesr_b200: in-plane resampling / label volumes are outside the hot path")
    if num_interpolations is not None and downsample_steps is not None:
        if generate_inbetween_slices and num_interpolations + 1 != downsample_steps:
            raise ValueError("ERROR - num_interpolations {} must be equal to "
                             "downsample_steps {} - 1".format(num_interpolations, downsample_steps))
    percept_loss = None
    if compute_percept_loss:
        percept_loss = getattr(trainer, "percept_criterion", None)
    if base_out_dir is None and generate_inbetween_slices:
        base_out_dir = "images_sr"
    else:
        base_out_dir = "images_sr_ip"
    alpha_range = np.linspace(0, 1, num_interpolations + 2, endpoint=True)[1:-1] if num_interpolations is not None else None
    if save_volumes:
        if output_dir is not None:
            output_dir = os.path.join(os.path.expanduser(output_dir), base_out_dir)
        else:
            assert 'output_dir' in myargs
            output_dir = os.path.join(myargs['output_dir'], base_out_dir)
        print("INFO - saving output to {}".format(output_dir))
        os.makedirs(output_dir, exist_ok=True)
    data_generator = check_data_generator(data_generator)
    model = synthesis._model_of(trainer)
    dev = next(model.parameters()).device
    lists = [[] for _ in range(12)]
    for test_batch in data_generator:
        origin, direction = test_batch.get('origin'), test_batch.get('direction')
        if transform is not None:
            test_batch = transform(test_batch)
        images = test_batch['image']
        images = torch.from_numpy(images) if isinstance(images, np.ndarray) else images
        image_hr = test_batch.get('image_hr')
        if image_hr is not None and transform is not None:
            image_hr = transform({'image': image_hr})['image']
        pat_id = test_batch['patient_id'] if isinstance(test_batch['patient_id'], str) else str(test_batch['patient_id'])
        if patient_id is not None and patient_id != pat_id:
            continue
        spacing = np.asarray(test_batch.get('spacing', (1.0, 1.0, 1.0)), dtype=np.float64).copy()
        new_spacing_z = spacing[0] if generate_inbetween_slices else spacing[0] / (num_interpolations + 1)
        res = synthesis.create_super_volume_eval(trainer, images, alpha_range=alpha_range, use_original=use_original_slice,
                                                 downsample_steps=downsample_steps, hierarchical=False,
                                                 generate_inbetween_slices=generate_inbetween_slices,
                                                 keep_on_device=True)
        new_images = res['upsampled_image']
        spacing[0] = new_spacing_z
        if save_volumes:
            _save_volume(new_images.cpu().numpy(), pat_id, output_dir, file_suffix, myargs, spacing, origin, direction)
        if generate_inbetween_slices:
            ref = images if image_hr is None else image_hr
            ref = torch.from_numpy(ref) if isinstance(ref, np.ndarray) else ref
            E.compute_metrics(ref.to(dev), new_images.to(dev), downsample_steps, *lists,
                              compute_percept_loss=compute_percept_loss, percept_loss=percept_loss,
                              normalize=normalize, eval_axis=eval_axis, device=dev)
            if verbose:
                print("SSIM / PSRN / VIF: {:.3f} / {:.3f} / {:.3f}".format(lists[0][-1], lists[1][-1], lists[2][-1]))
    if save_volumes:
        suffix = "" if file_suffix is None else file_suffix
        if getattr(trainer, 'model_file', None) is not None:
            model_nbr = trainer.model_file.split(os.sep)[-1].replace('.models', '')
            readme = os.path.join(output_dir, "README_{}_".format(model_nbr) + suffix + ".txt")
        else:
            readme = os.path.join(output_dir, "README_" + suffix + ".txt")
        Path(readme).touch()
    if not generate_inbetween_slices:
        return None, None, None
    names = [m + s for s in ("", "_synth", "_recon") for m in ("ssim", "psnr", "vif", "lpips")]
    result_dict = dict(zip(names, lists))
    for tag, label in (("", "Total"), ("_recon", "Reconstruction"), ("_synth", "Synthesis")):
        if tag and eval_axis != 0:
            continue
        m = E.compute_mean_metrics(result_dict["ssim" + tag], result_dict["psnr" + tag], result_dict["vif" + tag],
                                   result_dict["lpips" + tag])
        print("{} - SSIM / PSRN / VIF / LPIPS: {:.3f} ({:.2f}) / {:.2f} ({:.2f}) / {:.3f} ({:.2f}) / {:.3f} ({:.2f})"
              .format(label, *m))
    return result_dict


def evaluate_interpolation_performance(trainer, myargs, data_generator, transform=None, downsample_steps=None,
                                       file_suffix=None, patient_id=None, eval_axis=0):
    """evaluate/evaluate_interpolations.py:41-63."""
    is_4d = True if myargs['dataset'] in ["ACDC", "ARVC"] else False
    return create_hr_images(data_generator, myargs, trainer, num_interpolations=downsample_steps - 1,
                            downsample_steps=downsample_steps, use_original_slice=False, is_4d=is_4d,
                            transform=transform, normalize=False, generate_inbetween_slices=True,
                            patient_id=patient_id, file_suffix=file_suffix, save_volumes=False, eval_axis=eval_axis,
                            compute_percept_loss=False, verbose=False)


class _PatchTransform:
    """AdjustToPatchSize + CenterCrop on the device (datasets/shared_transforms.py:389-447, 297-363) applied to the
    'image' entry of a batch dict, the way ``find_best_model.get_transforms(ps, to_tensor=False)`` composes them."""

    def __init__(self, patch: Optional[int]):
        self.patch = patch

    def __call__(self, batch):
        if self.patch is None:
            return batch
        out = dict(batch)
        img = batch['image']
        img = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
        vol = img[:, None] if img.dim() == 3 else img                  # [Z,H,W] -> [Z,1,H,W]
        vol = E.center_crop(E.adjust_to_patch_size(vol, self.patch), self.patch)
        vol = vol[:, 0] if img.dim() == 3 else vol
        out['image'] = vol.cpu().numpy() if isinstance(batch['image'], np.ndarray) else vol
        return out


def get_transforms(transform_patch_size, to_tensor=True):
    """evaluate/find_best_model.py:25-34."""
    return _PatchTransform(transform_patch_size)


def store_top_scores(model_nbr, top_scores, ssim_results, psnr_results, vif_results):
    """evaluate/find_best_model.py:37-42."""
    top_scores[model_nbr] = np.array([np.mean(np.array(ssim_results)), np.mean(np.array(psnr_results)),
                                      np.mean(np.array(vif_results))])
    return top_scores


def find_best_val_model(data_generator, exper_src_dir, epoch_range=None, ps_evaluate=None, eval_axis=0,
                        downsample_steps=None, patient_id=None, limit_4d=False, func_get_trainer=None):
    """evaluate/find_best_model.py:45-109: score every ``<exper>/models/<epoch>.models`` of ``epoch_range`` on the
    validation volumes, write ``model_perf_*`` / ``model_perf_synth_*`` .npz next to the experiment, return the
    {epoch: [mean SSIM, mean PSNR, mean VIF]} dict sorted by epoch."""
    if func_get_trainer is None:
        from kwatsch.get_trainer import get_trainer_dynamic as func_get_trainer
    exper_src_dir = os.path.expanduser(exper_src_dir)
    search_mask = os.path.join(os.path.join(exper_src_dir, "models"), "*.models")
    model_list = sorted(glob.glob(search_mask))
    if epoch_range is not None:
        epoch_range = [str(e) for e in epoch_range]
        model_list = sorted(m for m in model_list if os.path.basename(m).replace(".models", "") in epoch_range)
    print("INFO - find-best-validation-model - testing {} networks using p-size {} "
          " - eval_axis={}".format(len(model_list), ps_evaluate, eval_axis))
    if len(model_list) == 0:
        raise ValueError("Error no models found with search mask {}".format(search_mask))
    if epoch_range is None:
        epoch_range = [os.path.basename(m).replace(".models", "") for m in model_list]
    if isinstance(data_generator, types.GeneratorType):
        if limit_4d:
            data_generator = {i: t for i, t in enumerate(data_generator) if t['frame_id'] in [4, 11, 15]}
        else:
            data_generator = {i: t for i, t in enumerate(data_generator)}
        print("Transformed data generator into dict with len {}".format(len(data_generator)))
    top_scores, top_scores_synth = {}, {}
    best = {k: (None, -1.0) for k in ("ssim", "psnr", "vif", "ssim_synth", "psnr_synth", "vif_synth")}
    for model_nbr in epoch_range:
        trainer, e_args = func_get_trainer(src_path=exper_src_dir, model_nbr=model_nbr, eval_mode=True)
        if downsample_steps is None:
            if "downsample_steps" not in e_args.keys():
                raise ValueError("ERROR - Downsample steps need to be specified")
            downsample_steps = e_args["downsample_steps"]
        transform = get_transforms(transform_patch_size=ps_evaluate, to_tensor=False)
        result_dict = evaluate_interpolation_performance(trainer, e_args, data_generator, transform=transform,
                                                         downsample_steps=downsample_steps, file_suffix=None,
                                                         patient_id=patient_id, eval_axis=eval_axis)
        top_scores = store_top_scores(model_nbr, top_scores, result_dict['ssim'], result_dict['psnr'], result_dict['vif'])
        top_scores_synth = store_top_scores(model_nbr, top_scores_synth, result_dict['ssim_synth'],
                                            result_dict['psnr_synth'], result_dict['vif_synth'])
        for j, name in enumerate(("ssim", "psnr", "vif")):
            if top_scores[model_nbr][j] > best[name][1]:
                best[name] = (int(model_nbr), float(top_scores[model_nbr][j]))
            if top_scores_synth[model_nbr][j] > best[name + "_synth"][1]:
                best[name + "_synth"] = (int(model_nbr), float(top_scores_synth[model_nbr][j]))
    print("Top metrics: Mean SSIM/PSNR/VIF M-{}: {:.4f} / M-{}: {:.4f} M-{}: {:.4f}".format(
        best["ssim"][0], best["ssim"][1], best["psnr"][0], best["psnr"][1], best["vif"][0], best["vif"][1]))
    np_fname = os.path.join(exper_src_dir, "model_perf_{}_to_{}_axis{}.npz".format(epoch_range[0], epoch_range[-1], eval_axis))
    np.savez(np_fname, **top_scores)
    print("Top synthesis: Mean SSIM/PSNR/VIF M-{}: {:.4f} / M-{}: {:.4f} M-{}: {:.4f}".format(
        best["ssim_synth"][0], best["ssim_synth"][1], best["psnr_synth"][0], best["psnr_synth"][1],
        best["vif_synth"][0], best["vif_synth"][1]))
    print("Saved result dict to {}".format(np_fname))
    np_fname = os.path.join(exper_src_dir, "model_perf_synth_{}_to_{}_axis{}.npz".format(epoch_range[0], epoch_range[-1],
                                                                                        eval_axis))
    np.savez(np_fname, **top_scores_synth)
    return dict(sorted(top_scores.items()))


def load_model_scores(exper_dir, file_suffix='.npz', synthesis=False):
    """evaluate/find_best_model.py:112-137."""
    load_dir = os.path.expanduser(exper_dir)
    file_prefix = "model_perf_synth*" if synthesis else "model_perf*"
    files_to_load = glob.glob(os.path.join(load_dir, file_prefix + file_suffix))
    if len(files_to_load) == 0:
        print("INFO - nothing to load from {}".format(load_dir))
        return None
    results = {}
    for fname in files_to_load:
        if not synthesis and 'synth' in fname:
            continue
        np_files = np.load(fname)
        results.update({epoch: np_files[epoch] for epoch in np_files.files})
    epochs, ssim, psnr, vif = [], [], [], []
    for epoch, metrics in results.items():
        epochs.append(int(epoch)), ssim.append(metrics[0]), psnr.append(metrics[1]), vif.append(metrics[2])
    return results, np.array(epochs), np.array(ssim), np.array(psnr), np.array(vif)
