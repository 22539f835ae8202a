"""Volume synthesis: encode the low-resolution slices, interpolate latents at every alpha, decode the in-between slices.

Drop-in for ``generate_hr_volumes.create_super_volume / latent_space_interp`` (generate_hr_volumes.py:12-101) and the
evaluation twin ``evaluate/common.create_super_volume`` (evaluate/common.py:134-235), with the reference's result
(slice order, lerp weights, clamp) but the minimal work: every slice is encoded ONCE (the reference re-encodes both
neighbours for every alpha), all (pair, alpha) latents are blended and decoded as one batch, and each synthesized slice
is written by the decoder's last kernel straight to its position ``i*(A+1)+1+k`` of a preallocated volume (the
reference builds it with an O(Z^2) chain of torch.cat).  Eval-mode BatchNorm is per-sample, so batching differently
from the reference cannot change a value.
"""
from __future__ import annotations

import collections
import concurrent.futures
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


def interp_weights(alpha_range: Sequence[float]):
    """fp32 weights exactly as torch forms them from a float64 python scalar: w_hi = fp32(alpha) multiplies the LATER
    slice, w_lo = fp32(1 - alpha) (subtraction in float64) the earlier one (generate_hr_volumes.py:50,88)."""
    a = np.asarray(alpha_range, dtype=np.float64)
    return a.astype(np.float32), (1.0 - a).astype(np.float32)


def _model_of(trainer_or_model, use_sr_model: bool = True):
    m = trainer_or_model
    if hasattr(m, "_use_sr_model"):
        return m._use_sr_model(use_sr_model)
    return getattr(m, "model", m)


def pair_plan(num_slices: int, num_alphas: int, w_hi: np.ndarray, w_lo: np.ndarray, device, slice_offset: int = 0,
              out_offset: int = 0):
    """Index / weight tables for all (pair i, alpha k) problems of one volume, m = i*A + k:
    out[i*(A+1)+1+k] = dec(w_hi[k] * z[i+1] + w_lo[k] * z[i])."""
    Z, A = num_slices, num_alphas
    i = np.repeat(np.arange(Z - 1, dtype=np.int64), A)
    k = np.tile(np.arange(A, dtype=np.int64), Z - 1)
    ia = (i + 1 + slice_offset).astype(np.int32)
    ib = (i + slice_offset).astype(np.int32)
    out_idx = (i * (A + 1) + 1 + k + out_offset).astype(np.int32)
    return ia, ib, w_hi[k].astype(np.float32), w_lo[k].astype(np.float32), out_idx


_PLAN_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_PLAN_CACHE_MAX = 16


def _synthesis_plan(V: int, Z: int, alpha_range: Sequence[float], dev: torch.device) -> dict:
    """Device-resident index / weight tables of a (V volumes, Z slices, alphas) synthesis problem, cached per shape: a
    steady stream of equally shaped batches then issues no host->device copies and no index arithmetic kernels (the
    pageable-memory copies of these tables used to block the host behind the previous batch's kernels, which kept the
    copy / compute pipeline of HostPipeline from running ahead, profiles/README.md r01x)."""
    a64 = np.asarray(alpha_range, dtype=np.float64)
    key = (V, Z, a64.tobytes(), dev.type, dev.index)
    plan = _PLAN_CACHE.get(key)
    if plan is not None:
        _PLAN_CACHE.move_to_end(key)
        return plan
    A = len(a64)
    Zo = (Z - 1) * (A + 1) + 1
    w_hi, w_lo = interp_weights(alpha_range)
    n = np.arange(V * Z, dtype=np.int64)
    q = np.arange(V * max(Z - 1, 0), dtype=np.int64)
    v_of, i_of = (q // (Z - 1), q % (Z - 1)) if Z > 1 else (q, q)
    oi_np = (v_of * Zo + i_of * (A + 1) + 1)[:, None] + np.arange(A, dtype=np.int64)[None, :]

    def dev_i32(x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev)

    plan = {"idx": dev_i32(n), "oi_kept": dev_i32((n // Z) * Zo + (n % Z) * (A + 1)),
            "pa": dev_i32(v_of * Z + i_of + 1), "pb": dev_i32(v_of * Z + i_of), "oi": dev_i32(oi_np.reshape(-1)),
            "wa": torch.from_numpy(w_hi.astype(np.float32)).to(dev), "wb": torch.from_numpy(w_lo.astype(np.float32)).to(dev),
            "one": torch.ones(1, dtype=torch.float32, device=dev), "zero": torch.zeros(1, dtype=torch.float32, device=dev)}
    _PLAN_CACHE[key] = plan
    while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
        _PLAN_CACHE.popitem(last=False)
    return plan


@torch.no_grad()
def synthesize_volumes(model, volumes: torch.Tensor, alpha_range: Sequence[float], use_original: bool = True,
                       decode_chunk: int = 4096, encode_chunk: int = 2048, out: Optional[torch.Tensor] = None,
                       place_kept: bool = True) -> torch.Tensor:
    """Batched synthesis of V independent volumes.  volumes: [V,Z,H,W] fp32 (device) -> [V,(Z-1)(A+1)+1,H,W] fp32.

    ``decode_chunk`` / ``encode_chunk`` bound the slices per kernel launch.  Large chunks win: measured on B200
    (profiles/README.md, chunk sweep) a launch of the persistent conv kernel carries ~10 us of fixed cost (launch,
    TMEM allocation, filter-bank load, tail wave), which at 256 slices per launch was 16 % of the step -- more than
    keeping the inter-layer tensors L2-resident ever bought.  The chunks are additionally capped so that the largest
    inter-layer tensor of a launch stays below ~4 GiB.  ``place_kept=False`` (with ``use_original``) leaves the kept
    slices of ``out`` unwritten: the caller already holds them (HostPipeline forms clamp(input) on the host)."""
    assert volumes.dim() == 4 and volumes.is_cuda
    V, Z, H, W = volumes.shape
    cap = max(1, (4 << 30) // ((H + 2) * (W + 2) * 64))           # stem output: 32 channels x 2 bytes per pixel
    decode_chunk, encode_chunk = max(1, min(decode_chunk, cap)), max(1, min(encode_chunk, cap))
    A = len(alpha_range)
    dev = volumes.device
    w_hi, w_lo = interp_weights(alpha_range)
    Zo = (Z - 1) * (A + 1) + 1
    vol = volumes.float().contiguous()
    if out is None:
        out = torch.empty((V, Zo, H, W), dtype=torch.float32, device=dev)
    flat_in = vol.view(V * Z, 1, H, W)
    # Interpolation behind dec.0: dec.0 is linear, so it is applied ONCE per encoded slice (un-rounded fp32 output) and
    # the blend + bias + LeakyReLU happen in lerp_pairs_act -- (Z-1)*A/Z times less dec.0 work, half the lerp bytes, and
    # the fp32 NCHW latent is never materialised.
    fold = bool(getattr(model, "fused_inference", False) and getattr(model, "linear_fold", False)
                and getattr(model, "scales", 0) >= 1)
    # ---- encode every slice once
    z_parts = []
    for s in range(0, V * Z, encode_chunk):
        if fold:
            z_parts.append(model.decode_pre_eval(model.encode_eval(flat_in[s:s + encode_chunk], nhwc_only=True)))
        else:
            z_parts.append(model.encode_eval(flat_in[s:s + encode_chunk]))
    z = z_parts[0] if len(z_parts) == 1 else torch.cat(z_parts, dim=0)
    bias0 = model.dec[0].bias.detach() if fold else None

    def blend_decode(pa_, pb_, wa_, wb_, oi_):
        if fold:
            a = ops.lerp_pairs_act(z, pa_, pb_, wa_, wb_, bias0)
        else:
            a = ops.lerp_pairs(z, pa_, pb_, wa_, wb_)
        model.decode_nhwc_eval(a, out=out, out_image_stride=H * W, out_index=oi_, after_first=fold)

    plan = _synthesis_plan(V, Z, alpha_range, dev)
    # ---- kept slices: originals (clamped, generate_hr_volumes.py:44,67) or reconstructions
    idx, oi = plan["idx"], plan["oi_kept"]
    if use_original:
        if place_kept:
            ops.place_slices(flat_in, out, oi, clamp=True)
    else:       # decode(z_i) = the blend with weights (1, 0) of slice i with itself (1*z + 0*z = z exactly)
        one, zero = plan["one"], plan["zero"]
        for s in range(0, V * Z, decode_chunk):
            e = min(s + decode_chunk, V * Z)
            blend_decode(idx[s:e], idx[s:e], one, zero, oi[s:e].contiguous())
    if A == 0 or Z < 2:
        return out
    # ---- all (volume, pair, alpha) lerp+decode problems
    #      pair q = v*(Z-1)+i blends z[v*Z+i+1] (weight w_hi[k]) and z[v*Z+i] (w_lo[k]); problem m = q*A + k lands at
    #      out[v*Zo + i*(A+1) + 1 + k].  The two fp32 operands of a pair are read once for all A alphas.
    pa, pb, oi, wa, wb = plan["pa"], plan["pb"], plan["oi"], plan["wa"], plan["wb"]
    pairs_per_chunk = max(1, decode_chunk // A)
    for s in range(0, V * (Z - 1), pairs_per_chunk):
        e = min(s + pairs_per_chunk, V * (Z - 1))
        blend_decode(pa[s:e], pb[s:e], wa, wb, oi[s * A:e * A])
    return out


class HostPipeline:
    """Host-buffer entry point for batched synthesis: pinned host volumes [V,Z,H,W] in, HR volumes
    [V,(Z-1)(A+1)+1,H,W] back in pinned host memory.  The V volumes are cut into groups; group g+1's host->device copy
    and group g-1's device->host copy run on their own streams while group g computes (double-buffered staging).

    Only the SYNTHESIZED slices cross PCIe on the way back (``host_kept``, default): the kept slices of the HR volume are
    clamp(input, 0, 1) of slices the host already holds (generate_hr_volumes.py:44,58-67), so a host worker thread writes
    them into ``host_out`` while the GPU computes, and the download is one strided 2-D copy per volume (rows = slice
    pairs, A slices each).  The device->host copy bounds this path (268 MB per 64-volume step at ~55 GB/s = 4.9 ms against
    4.2 ms of compute, profiles/README.md); leaving out the Z of (Z-1)(A+1)+1 kept slices cuts it by 16 % for A = 6.

    ``run(..., wait=True)`` (default) returns stream-ordered: the caller's stream has waited for the last device->host
    copy and the host worker has finished.  A caller that feeds a sequence of batches passes ``wait=False`` and calls ``wait()`` (or ``synchronize()``)
    once at the end: the staging buffers and their events persist across calls, so the copies of batch i overlap the
    compute of batch i+1 and the only serial parts left are the first upload and the last download."""

    def __init__(self, model, V, Z, H, W, alpha_range, groups: int = 2, chunk: int = 4096,
                 host_kept: Optional[bool] = None, d2h_streams: int = 2, host_workers: Optional[int] = None):
        self.model, self.ar, self.chunk = model, list(alpha_range), chunk
        if host_kept is None:
            host_kept = True
        self.host_kept = bool(host_kept) and len(self.ar) >= 1 and Z >= 2
        # The host workers' clamp of the kept slices must stay well below the download time (0.5 ms per 64-volume step on
        # 16 threads, 10 ms on one).  torchrun exports OMP_NUM_THREADS=1 to every rank, so torch's intra-op pool cannot be
        # relied on (the 2- and 4-GPU runs of profiles/r02l / r02z were bound by a single-threaded clamp): the group's
        # volumes are cut into slabs for an OWN pool of workers -- torch releases the GIL inside clamp, the slabs run in
        # parallel whatever the intra-op setting.  Workers = the cores this process may use (its share of the box under
        # one-process-per-GPU affinity), at most 8.
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        self.host_workers = host_workers or max(1, min(8, avail if avail < (os.cpu_count() or 1) else avail // local_world))
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.host_workers) if self.host_kept else None
        self.jobs = []
        dev = next(model.parameters()).device
        self.dev = dev
        groups = max(1, min(groups, V))
        self.bounds = [(g * V // groups, (g + 1) * V // groups) for g in range(groups)]
        gmax = max(e - s for s, e in self.bounds)
        A = len(self.ar)
        Zo = (Z - 1) * (A + 1) + 1
        self.d_in = [torch.empty(gmax, Z, H, W, device=dev) for _ in range(2)]
        self.d_out = [torch.empty(gmax, Zo, H, W, device=dev) for _ in range(2)]
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # download streams: each group's volumes are split between them (two copy engines in flight: 5.03 vs 5.13-5.30 ms
        # per step, profiles/r02x_d2h_streams.txt; a third changes nothing)
        self.s_out_extra = [torch.cuda.Stream(dev) for _ in range(max(0, int(d2h_streams) - 1))] if self.host_kept else []
        self.in_free = [torch.cuda.Event() for _ in range(2)]     # compute finished reading d_in[b]
        self.out_free = [torch.cuda.Event() for _ in range(2)]    # copy-out finished reading d_out[b]
        # bytes crossing PCIe per run() (what bench.py reports): all inputs up, synthesized (or all) slices down
        self.h2d_bytes = V * Z * H * W * 4
        self.d2h_bytes = V * ((Z - 1) * A if self.host_kept else Zo) * H * W * 4
        self.trace = None          # set to a list to collect (tag, group, timing event) for tools/e2e_probe.py
        self.used = [False, False]
        self.turn = 0                                             # staging buffer of the next group (persists over calls)

    def run(self, host_in: torch.Tensor, host_out: torch.Tensor, wait: bool = True) -> None:
        main = torch.cuda.current_stream(self.dev)
        for (s, e) in self.bounds:
            b = self.turn
            self.turn ^= 1
            n = e - s
            with torch.cuda.stream(self.s_in):
                if self.used[b]:
                    self.s_in.wait_event(self.in_free[b])
                self.d_in[b][:n].copy_(host_in[s:e], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.s_in)
            main.wait_event(ready)
            if self.used[b]:
                main.wait_event(self.out_free[b])
            self._mark("compute_start", s, main)
            synthesize_volumes(self.model, self.d_in[b][:n], self.ar, use_original=True, out=self.d_out[b][:n],
                               decode_chunk=self.chunk, encode_chunk=self.chunk, place_kept=not self.host_kept)
            self.in_free[b].record(main)
            done = torch.cuda.Event()
            done.record(main)
            self._mark("compute_end", s, main)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                self._mark("d2h_start", s, self.s_out)
                if self.host_kept:
                    # volume v: rows i = 0..Z-2 of A consecutive slices starting at slice i*(A+1)+1, pitch (A+1) slices
                    Z, A = host_in.shape[1], len(self.ar)
                    sl = host_out.shape[2] * host_out.shape[3] * 4          # bytes per slice
                    vol = host_out.shape[1] * sl
                    streams = [self.s_out] + self.s_out_extra
                    parts = [(k * n // len(streams), (k + 1) * n // len(streams)) for k in range(len(streams))]
                    for st, (v0, v1) in zip(streams, parts):
                        if v1 <= v0:
                            continue
                        if st is not self.s_out:
                            st.wait_event(done)
                        ops.copy_rows_async(host_out, self.d_out[b], (s + v0) * vol + sl, v0 * vol + sl, outer=v1 - v0,
                                            dst_outer_stride=vol, src_outer_stride=vol, rows=Z - 1, dpitch=(A + 1) * sl,
                                            spitch=(A + 1) * sl, width=A * sl, stream=st)
                        if st is not self.s_out:
                            ev = torch.cuda.Event()
                            ev.record(st)
                            self.s_out.wait_event(ev)
                else:
                    host_out[s:e].copy_(self.d_out[b][:n], non_blocking=True)
                self.out_free[b].record(self.s_out)
                self._mark("d2h_end", s, self.s_out)
            if self.host_kept:
                step = len(self.ar) + 1
                k = min(self.host_workers, n)
                for j in range(k):
                    a0, a1 = s + j * n // k, s + (j + 1) * n // k
                    self.jobs.append(self.pool.submit(torch.clamp, host_in[a0:a1], 0.0, 1.0, out=host_out[a0:a1, ::step]))
            self.used[b] = True
        if wait:
            self.wait()

    def _mark(self, tag, group, stream) -> None:
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((tag, group, ev))

    def wait(self) -> None:
        """Make the caller's current stream wait for every device->host copy issued so far (stream-ordered return)."""
        main = torch.cuda.current_stream(self.dev)
        for b in range(2):
            if self.used[b]:
                main.wait_event(self.out_free[b])
        self._join_host()

    def _join_host(self) -> None:
        jobs, self.jobs = self.jobs, []
        for j in jobs:
            j.result()

    def synchronize(self) -> None:
        """Block the host until every result issued so far is in host memory."""
        self._join_host()
        for b in range(2):
            if self.used[b]:
                self.out_free[b].synchronize()


@torch.no_grad()
def latent_space_interp(alpha, trainer, img1, img2, device="cuda", with_labels=False, hierarchical=False) -> dict:
    """generate_hr_volumes.py:72-101 / kwatsch/img_interpolation.py:57-89 (same signature and return dict)."""
    if with_labels or hierarchical:
        raise NotImplementedError("aesr_b200: label / hierarchical latents belong to other model families")
    model = _model_of(trainer)
    img1 = img1.float().to(device)
    img2 = img2.float().to(device)
    n = img1.shape[0]
    z = model.encode_eval(torch.cat([img1, img2], dim=0))
    hi, lo = interp_weights([float(alpha)])
    ia = torch.arange(n, dtype=torch.int32, device=z.device)
    a = ops.lerp_latents(z, ia, ia + n, torch.full((n,), float(hi[0]), device=z.device),
                         torch.full((n,), float(lo[0]), device=z.device))
    img = model.decode_nhwc_eval(a)
    return {"inter_image": img.detach().cpu(), "inter_label": None}


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> host through a PINNED buffer from torch's caching host allocator (one cudaMemcpyAsync at PCIe speed; a
    pageable ``.cpu()`` of a 50 MB volume takes ~20 ms).  The returned CPU tensor owns its buffer."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host


@torch.no_grad()
def create_super_volume(trainer, images: torch.Tensor, alpha_range, use_original: bool = False, labels=None,
                        keep_on_device: bool = False) -> dict:
    """generate_hr_volumes.py:12-69: images [Z,1,H,W] or [Z,H,W] -> {'upsampled_image': [(Z-1)(A+1)+1,H,W] (CPU; on the
    model's device with ``keep_on_device``, for callers that score the volume with the device metrics)}."""
    if labels is not None:
        raise NotImplementedError("aesr_b200: label volumes belong to the multi-channel model family")
    model = _model_of(trainer)
    if images.dim() == 4:
        images = images[:, 0]
    dev = next(model.parameters()).device
    vol = images.float().to(dev).unsqueeze(0)
    # the final torch.clamp(0, 1) of the reference (:67) is applied inside the kernels that write `out`
    out = synthesize_volumes(model, vol, alpha_range, use_original=use_original)[0]
    return {"upsampled_image": out if keep_on_device else _to_host(out), "upsampled_labels": None}


@torch.no_grad()
def create_super_volume_eval(trainer, images: torch.Tensor, alpha_range=None, use_original: bool = False,
                             hierarchical: bool = False, downsample_steps: Optional[int] = None,
                             generate_inbetween_slices: bool = False, train_patch_size=None, feature_dict=None,
                             labels=None, keep_on_device: bool = False) -> dict:
    """evaluate/common.py:134-235: optional slice dropping images[::d] after trimming (Z-1) % d tail slices, tail
    re-appended untouched.  Only the kept slices and the tail cross PCIe on the way in; the volume is assembled on the
    device and comes back in one pinned copy."""
    if labels is not None or hierarchical:
        raise NotImplementedError("aesr_b200: labels / hierarchical latents are out of scope")
    if generate_inbetween_slices and downsample_steps is None:
        downsample_steps = int(len(alpha_range) + 1)
    tail, orig_num = None, images.shape[0]
    if downsample_steps is not None or generate_inbetween_slices:
        remain = (orig_num - 1) % downsample_steps
        if remain != 0:
            if generate_inbetween_slices:
                tail = images[-remain:]
            images = images[:-remain]
        images = images[::downsample_steps]
    if alpha_range is None:
        alpha_range = [0.25, 0.5, 0.75]
    res = create_super_volume(trainer, images, alpha_range, use_original=use_original, keep_on_device=True)
    new_volume = res["upsampled_image"]
    if tail is not None:
        tail = tail.float().to(new_volume.device)
        if tail.dim() == 4:
            tail = tail[:, 0]
        new_volume = torch.cat([new_volume, torch.clamp(tail, 0, 1.)])
    n_alpha = len(alpha_range)
    pred_alphas = torch.cat([torch.FloatTensor([a]).expand(images.shape[0] - 1) for a in alpha_range]) \
        if n_alpha else None
    return {"upsampled_image": new_volume if keep_on_device else _to_host(new_volume), "upsampled_labels": None,
            "pred_alphas": pred_alphas}
