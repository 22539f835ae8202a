"""Device-resident training batch source -- the data flow between the datasets and ``trainer.train()`` in the reference
(train_cardiac_aesr.py:173-180, train_brain_aesr.py): ``DataLoader(dataset, batch_size, sampler)`` -> per sample
``__getitem__`` (triplet indices + transform chain, datasets/ACDC/data4d_simple.py:191-212, datasets/common_brains.py:241-260)
-> ``prepare_batch_pairs`` -> ``trainer.train(batch_item)``.

Here the volumes stay in HBM.  Per batch: the sampler picks (volume, slice) items, ``sampling.sample_triplet`` takes the
reference's RandomState draws for each item on the host (a few integers), ``sampling.gather_triplets`` stacks the three
slices with one index_select per volume shape, ``evaluation.augment_batch`` applies the whole transform chain in ONE kernel
(draws in the reference's order), and ``evaluation.prepare_batch_pairs`` lays the batch out as the trainers expect --
``for batch_item in loader: trainer.train(batch_item)`` reads like the reference's loop, with no host copy of image data.

Scope: the sampling / transform semantics of ``ACDCDataset4DPairs`` and ``BrainDataset``; reading
the datasets' files from disk stays with the caller (``volume_io`` / the CLI).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import evaluation as E
from . import sampling as S


class DeviceTripletLoader:
    """``volumes``: list of [Z,H,W] arrays / tensors (already intensity-normalised, like the datasets hold them).
    ``kind`` 'acdc' | 'brain'; ``slice_selection`` / ``downsample_steps`` as the dataset constructors take them;
    ``width`` / ``aug_patch`` / ``center`` / ``intensity_first`` describe the transform chain (see
    ``evaluation.augment_batch``: ACDC = AdjustToPatchSize(aug) -> CenterCrop -> RandomCrop(width) -> RandomIntensity ->
    RandomRotation; brains = RandomCrop -> RandomRotation -> RandomIntensity).  One RandomState ``rs`` drives the sampler
    permutation, the triplet draws and the transform draws, like the ``rs`` the reference threads through its dataset,
    transforms and loader.  ``drop_last`` like the reference's ``num_it_per_epoch = len(dataset) // batch_size``."""

    def __init__(self, volumes: Sequence, batch_size: int, kind: str = "acdc", slice_selection: str = "adjacent_plus",
                 downsample_steps: int = 2, width: int = 128, aug_patch: Optional[int] = None, center: bool = False,
                 intensity_first: Optional[bool] = None, rs: Optional[np.random.RandomState] = None, device="cuda:0",
                 shuffle: bool = True, augment: bool = True):
        assert kind in ("acdc", "brain")
        self.kind, self.sel, self.ds = kind, slice_selection, int(downsample_steps)
        self.batch_size, self.width, self.aug_patch, self.center = int(batch_size), int(width), aug_patch, bool(center)
        self.intensity_first = (kind == "acdc") if intensity_first is None else bool(intensity_first)
        self.rs = rs if rs is not None else np.random.RandomState(1234)
        self.device, self.shuffle, self.augment = torch.device(device), shuffle, augment
        self.volumes: List[torch.Tensor] = [torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).float().to(self.device)
                                            for v in volumes]
        # one item per (volume, slice) that can anchor a triplet (the datasets' _idcs): brains need an in-between slice
        # strictly inside the pair, i.e. at least `step` >= 2 slices of room on one side (common_brains.py:215-232)
        self.items = []
        for vi, v in enumerate(self.volumes):
            Z = v.shape[0]
            for z in range(Z):
                if kind == "brain" and slice_selection != "mix" and \
                        min(z + self.ds, Z - 1) - z < 2 and z - max(z - self.ds, 0) < 2:
                    continue
                self.items.append((vi, z, Z))
        if not self.items:
            raise ValueError("DeviceTripletLoader: no (volume, slice) item can anchor a triplet")

    def __len__(self) -> int:
        return len(self.items) // self.batch_size

    def _batch(self, picks: Sequence[int]) -> Dict[str, torch.Tensor]:
        # host: per sample, the draws of one ``__getitem__`` -- triplet indices first, then the transform chain
        trips, draws, owner = [], [], []
        for i in picks:
            vi, z, Z = self.items[i]
            while True:
                try:
                    t = S.sample_triplet(z, Z, self.rs, kind=self.kind, slice_selection=self.sel, downsample_steps=self.ds)
                    break
                except ValueError:          # 'mix' drew an adjacent pair in a brain set: empty open interval, draw again
                    continue
            trips.append(t)
            owner.append(vi)
            if self.augment:
                h, w = self.volumes[vi].shape[1:]
                _, _, hh, ww = E.augment_window(h, w, self.aug_patch, self.center)
                draws.append(E.augment_draw(self.rs, hh, ww, self.width, self.intensity_first))
        # device: gather + transform, one pass per group of equally sized volumes, rows written in batch order
        rows: List[Optional[torch.Tensor]] = [None] * len(picks)
        metas: List[Optional[Dict[str, torch.Tensor]]] = [None] * len(picks)
        for shape in sorted({tuple(self.volumes[o].shape[1:]) for o in owner}):
            sel = [k for k, o in enumerate(owner) if tuple(self.volumes[o].shape[1:]) == shape]
            imgs, parts = [], []
            for vi in sorted({owner[k] for k in sel}):
                ks = [k for k in sel if owner[k] == vi]
                g = S.gather_triplets(self.volumes[vi], [trips[k] for k in ks])
                imgs.append((ks, g))
            order = [k for ks, _ in imgs for k in ks]
            img = torch.cat([g["image"] for _, g in imgs], dim=0)
            if self.augment:
                img = E.augment_batch(img, None, self.width, aug_patch=self.aug_patch, center=self.center,
                                      intensity_first=self.intensity_first, device=self.device,
                                      draws=[draws[k] for k in order])
            pos = 0
            for ks, g in imgs:
                for j, k in enumerate(ks):
                    rows[k] = img[pos + j]
                    metas[k] = {m: g[m][j] for m in ("alpha_from", "alpha_to", "is_inbetween")}
                pos += len(ks)
        batch = {"image": torch.stack(rows, dim=0)}
        for m in ("alpha_from", "alpha_to", "is_inbetween"):
            batch[m] = torch.stack([mm[m] for mm in metas], dim=0)
        return E.prepare_batch_pairs(batch, expand_type="repeat")

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        n = len(self.items)
        perm = self.rs.permutation(n) if self.shuffle else np.arange(n)
        for b in range(len(self)):
            yield self._batch(perm[b * self.batch_size:(b + 1) * self.batch_size])
