"""Training-pair sampling of the reference's datasets, host side: which (from, to, in-between) slices form a sample and
with which interpolation coefficients -- the numpy ``RandomState`` draw order of ``__getitem__`` preserved, so that a
seeded stream yields the reference's sequence of triplets; the slices themselves are gathered on the device.

  ACDC   datasets/ACDC/data4d_simple.py:191-205 (``ACDCDataset4DPairs.__getitem__``), ``_get_slice_step`` :245-251,
         ``_get_inbetween_sliceid`` :253-263
  brains datasets/common_brains.py:241-260 (``BrainDataset.__getitem__``), ``_get_slice_step`` :272-278,
         ``_get_inbetween_sliceid`` :280-282, ``determine_interpol_coefficients`` :117-119
  both   datasets/common.py:34-43 (``get_random_adjacent_slice``)
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch


def get_random_adjacent_slice(slice_id: int, num_slices: int, rs=np.random.RandomState(1234), step: int = 1):
    """datasets/common.py:34-43, same signature."""
    num_slices -= 1
    if slice_id + step > num_slices:
        return slice_id - step
    elif slice_id == 0:
        return step
    elif slice_id - step < 0:
        return slice_id + step
    else:
        return rs.choice([slice_id - step, slice_id + step])


def sample_triplet(slice_id_1: int, num_slices: int, rs: np.random.RandomState, kind: str = "acdc",
                   slice_selection: str = "adjacent_plus", downsample_steps: int = 2) -> Dict[str, object]:
    """The index part of one ``__getitem__`` call (see module docstring for the reference lines).  Draw order:
    [step: rs.choice for 'mix'] -> partner slice [rs.choice when both neighbours exist] -> [brain: in-between slice,
    rs.choice over the open interval] -> direction rs.choice([0, 1])."""
    assert kind in ("acdc", "brain") and slice_selection in ("adjacent", "adjacent_plus", "mix")
    far = 2 if kind == "acdc" else downsample_steps
    if slice_selection == "adjacent":
        step = 1
    elif slice_selection == "adjacent_plus":
        step = far
    else:
        step = rs.choice([1, far])
    slice_id_2 = get_random_adjacent_slice(slice_id_1, num_slices, rs=rs, step=step)
    if kind == "acdc":
        if (slice_id_1 + slice_id_2) % 2 == 0:
            inbetween, is_inbetween = (slice_id_1 + slice_id_2) // 2, 1
        else:
            inbetween, is_inbetween = slice_id_1, 0
    else:
        inbetween = rs.choice(np.arange(min(slice_id_1, slice_id_2) + 1, max(slice_id_1, slice_id_2)))
        is_inbetween = 1
    if rs.choice([0, 1]) == 0:
        s_from, s_to = slice_id_1, slice_id_2
    else:
        s_from, s_to = slice_id_2, slice_id_1
    if kind == "acdc":
        alpha_from = alpha_to = np.float32(0.5)
    else:
        gap = s_to - s_from
        alpha_from = np.float32(1 - ((inbetween - s_from) * 1 / gap))
        alpha_to = np.float32(1 - ((s_to - inbetween) * 1 / gap))
    return {"slice_idx_from": int(s_from), "slice_idx_to": int(s_to), "inbetween_slice_id": int(inbetween),
            "is_inbetween": np.float32(is_inbetween), "alpha_from": alpha_from, "alpha_to": alpha_to}


def gather_triplets(volume: torch.Tensor, triplets: Sequence[Dict[str, object]]) -> Dict[str, torch.Tensor]:
    """Stack the (from, to, in-between) slices of a device-resident volume [Z,H,W] into the dataset's sample layout
    'image' [B,3,H,W] (np.vstack of the three slices, data4d_simple.py:210-212 / common_brains.py:254-256) plus
    'alpha_from' / 'alpha_to' [B,1] fp32 and 'is_inbetween' [B]; one index_select, no host copy of image data."""
    idx = torch.as_tensor([[t["slice_idx_from"], t["slice_idx_to"], t["inbetween_slice_id"]] for t in triplets],
                          dtype=torch.long, device=volume.device)
    b = idx.shape[0]
    img = volume.index_select(0, idx.reshape(-1)).reshape(b, 3, *volume.shape[1:])
    f32 = dict(dtype=torch.float32, device=volume.device)
    return {"image": img,
            "alpha_from": torch.tensor([[float(t["alpha_from"])] for t in triplets], **f32),
            "alpha_to": torch.tensor([[float(t["alpha_to"])] for t in triplets], **f32),
            "is_inbetween": torch.tensor([float(t["is_inbetween"]) for t in triplets], **f32)}
