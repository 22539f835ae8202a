"""Concurrent device->host copy ceiling of the box: every rank (torchrun, one process per GPU) copies a 268 MB fp32 HR-volume
batch from its GPU into pinned host memory, all ranks at once, and rank 0 prints GB/s per GPU and in aggregate.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_probe.py [--numa]

--numa binds each rank to its GPU's NUMA-local cores (parallel.bind_to_gpu_numa) before the pinned allocation.
One cudaMemcpyAsync per copy (torch .copy_(non_blocking=True)); device timing with CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import parallel as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--numa", action="store_true")
ap.add_argument("--mb", type=int, default=268)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cores = P.bind_to_gpu_numa(local) if a.numa else None
n = a.mb * 1000 * 1000 // 4
d = torch.rand(n, device=dev)
h = torch.empty(n).pin_memory()
up = torch.rand(n // 6).pin_memory()
dup = torch.empty(n // 6, device=dev)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


res = {}
for name, both in (("d2h", False), ("d2h_with_h2d", True)):
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2 = torch.cuda.Stream()
    e0.record()
    for _ in range(a.reps):
        h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                dup.copy_(up, non_blocking=True)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gbs = a.reps * n * 4 / (float(ms.item()) * 1e-3) / 1e9
    res[name] = {"gbs_per_gpu": gbs, "gbs_aggregate": gbs * world, "ms_per_copy": float(ms.item()) / a.reps}
if rank == 0:
    print(json.dumps({"n_gpus": world, "mb_per_copy": a.mb, "numa_bound": bool(a.numa), "cores_rank0": len(cores) if cores else None,
                      "host_cpus": os.cpu_count(), **res}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
