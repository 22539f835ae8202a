"""Host-side sharding for one-process-per-GPU runs (torch.distributed; NCCL over NVLink on the GPUs, gloo in CPU tests).

Inference shards independent units with NO data-path collective: volumes are independent, eval-mode BatchNorm is
per-sample.  Training is batch-sharded data parallelism: sample i of the global batch owns rows
i and B+i of ``image`` (its 'from' and 'to' slices, datasets/ACDC/data4d_simple.py:360) and row i of ``slice_between``
-- they must stay on the same rank -- and the only exchange is the mean of the flat gradient buffer (plus, in parity
mode, the per-channel BatchNorm sums).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), BEFORE
    pinned host buffers are allocated: cudaHostAlloc places pages on the allocating thread's node, and a host-buffer
    pipeline whose staging memory sits behind the other socket's interconnect pays for it on every device->host copy
    when eight ranks copy at once.  Returns the core list (None when NVML / affinity control is unavailable; nothing
    is changed then).  Safe to call once per rank at start-up (bench.py, generate_hr_volumes.py under torchrun)."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) slab of n units for this rank (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_volumes(volumes: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    s, e = shard_range(volumes.shape[0], rank, world)
    return volumes[s:e]


def shard_slice_pairs(num_slices: int, rank: int, world: int) -> Tuple[int, int]:
    """One large volume: contiguous slab of adjacent-slice pairs per rank; the slab [s, e) of pairs needs slices
    s .. e (one halo slice re-encoded at each internal boundary)."""
    return shard_range(num_slices - 1, rank, world)


def shard_batch_pairs(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Split a prepared batch (image [2B,1,H,W] = all 'from' then all 'to', slice_between [B,1,H,W], optional
    alpha_from / alpha_to [B,1]) so that every sample's three slices land on the same rank."""
    B = batch["slice_between"].shape[0]
    if batch["image"].shape[0] != 2 * B:
        raise ValueError("image must hold 2B slices (all 'from', then all 'to') for B slice_between rows")
    s, e = shard_range(B, rank, world)
    out = {"image": torch.cat([batch["image"][s:e], batch["image"][B + s:B + e]], dim=0),
           "slice_between": batch["slice_between"][s:e]}
    for k in ("alpha_from", "alpha_to"):
        if k in batch:
            out[k] = batch[k][s:e]
    return out


def average_gradients_(flat_grad: torch.Tensor, group=None):
    """In-place mean of the flat gradient buffer over ranks.  NCCL: one all-reduce with op AVG (no extra kernel);
    gloo (CPU tests) has no AVG, so SUM then scale."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat_grad
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad /= dist.get_world_size(group)
    return flat_grad


def sync_bn_sums_(stats: torch.Tensor, count: int, group=None) -> int:
    """Parity mode: global-batch BatchNorm statistics = all-reduce(SUM) of the per-rank [sum, sum^2] and the count."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return count
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return count * dist.get_world_size(group)


def gather_volume_shards(local: torch.Tensor, total: int, group=None) -> List[torch.Tensor]:
    """Collect every rank's HR volumes in global volume order (a single writer, e.g. the CLI saving files)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    bufs = [torch.empty((e - s,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for s, e in sizes]
    dist.all_gather(bufs, local.contiguous(), group=group) if len({e - s for s, e in sizes}) == 1 else \
        _all_gather_ragged(bufs, local, group)
    return bufs


def _all_gather_ragged(bufs, local, group):
    rank = dist.get_rank(group)
    for r, buf in enumerate(bufs):
        if r == rank:
            buf.copy_(local)
        dist.broadcast(buf, src=r, group=group)
