#!/usr/bin/env python
"""Benchmark of the hot path: batched HR-volume synthesis (encode -> interpolate latents -> decode).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port) on host cores

One "step" = one pass of the hot path over one batch of synthetic volumes: V ACDC-shaped volumes [10,128,128] per GPU,
num_interpolations = 6 (the generate_hr_volumes.py default) -> V*54 synthesized slices per GPU per step.  The metric
is BASELINE.json's "synthesized HR slices/sec"; `value` has the inputs resident in HBM, `e2e` goes through the public
host-buffer API (pinned host volumes in, HR volumes back to pinned host memory, copies inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "synthesized_hr_slices_per_sec"
Z, SIZE, NI = 10, 128, 6
# algorithmic conv FLOPs (2*MAC) of the reference's formulation, per image (ACDC scales=2 @128^2)
ENC_GMAC, DEC_GMAC = 0.7722, 0.3822
STEM_GMAC = (32 * 130 * 130 + 9 * 32 * 32 * 130 * 130) / 1e9      # enc.0 + enc.1 (folded into the CUDA-core stem kernel)


def flops_per_step(V: int) -> float:
    """Minimal-work count: every slice encoded once, every synthesized slice decoded once."""
    return 2e9 * (V * Z * ENC_GMAC + V * (Z - 1) * NI * DEC_GMAC)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock + clock-event (throttle) reasons while the timed region runs.  NVML in-process (one sample every few
    ms, so that a 100 ms timed region is covered by tens of samples); nvidia-smi subprocess as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                    (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown"))

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt, self.period = index, [], threading.Event(), period_s
        self.nvml, self.handle, self.max_mhz, self.how = None, None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle, self.how = pynvml, h, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        mhz = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        try:
            mw = float(n.nvmlDeviceGetPowerUsage(h))
        except Exception:
            mw = 0.0
        self.rows.append((mhz, mask, mw))

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    parts = [x.strip() for x in out.strip().split(",")]
                    if len(parts) >= 6:
                        mask = 0
                        for (bit, _n), v in zip(self.NVML_REASONS[:4], parts[2:6]):
                            if v.lower().startswith("active"):
                                mask |= bit
                        self.max_mhz = float(parts[1])
                        self.rows.append((float(parts[0]), mask, 0.0))
            except Exception:
                pass
            self._halt.wait(self.period if self.nvml is not None else 0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[1]
        reasons = [name for bit, name in self.NVML_REASONS if mask & bit]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(self.rows), "power_w_max": max(r[2] for r in self.rows) / 1e3, "how": self.how}


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_synthesis_rate(sample_volumes: int, repeats: int):
    """The reference's algorithm for the path (generate_hr_volumes.create_super_volume as written: both neighbours
    re-encoded for every alpha) restated in oracle/aesr_oracle.py, on all host threads torch will use."""
    from oracle import aesr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    args = O.default_args(SIZE, 32)
    state = O.init_state(args, seed=892372)
    ar = O.alpha_range_for(NI)
    vols = [O.synthetic_volume(Z, SIZE, seed=1 + i) for i in range(sample_volumes)]
    O.create_super_volume(state, args, vols[0], ar, use_original=True)        # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for v in vols:
            O.create_super_volume(state, args, v, ar, use_original=True)
        best = min(best, time.perf_counter() - t0)
    return sample_volumes * (Z - 1) * NI / best, torch.get_num_threads(), best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    t0 = time.perf_counter()
    for _ in range(a.warmup):
        cpu_synthesis_rate(1, 1)
    rates = []
    for _ in range(a.steps):
        r, cores, _ = cpu_synthesis_rate(sample, 1)
        rates.append(r)
    wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sample * (Z - 1) * NI / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ACDC-shaped volumes [10,128,128], num_interpolations=6, ae_combined scales=2 "
                                   "(width 128, latent_width 32, latent 128)", "sample": "%d volumes per step" % sample},
            "cpu_baseline": {"value": value, "unit": "slices/s", "cores": cores, "kind": "port",
                             "sample": "%d volumes (%d synthesized slices) per step, %d steps, oracle port of "
                                       "generate_hr_volumes.create_super_volume" % (sample, sample * 54, a.steps)},
            "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------- this repo's arm (GPU)
def run_ours(a):
    import torch.distributed as dist
    from oracle import aesr_oracle as O          # synthetic inputs / seeded weights only (never computes on this arm)
    from superresolution_aniso_mri_b200 import _lib, ops, synthesis
    from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 GPU; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from superresolution_aniso_mri_b200 import build
    if rank == 0:
        build.build_library()
    if world > 1:
        dist.barrier()
    V = a.volumes
    args = O.default_args(SIZE, 32)
    margs = dict(args)
    margs["device"] = str(dev)
    torch.manual_seed(892372)
    model = VanillaACAI(margs)
    model.load_state_dict(O.calibrated_state(args))      # seeded synthetic checkpoint with O(1) activations
    model.eval()
    ar = O.alpha_range_for(NI)
    g = torch.Generator().manual_seed(100 + rank)
    host_in = torch.rand(V, Z, SIZE, SIZE, generator=g).pin_memory()
    dev_in = host_in.to(dev)
    Zo = (Z - 1) * (NI + 1) + 1
    dev_out = torch.empty(V, Zo, SIZE, SIZE, device=dev)
    host_out = torch.empty(V, Zo, SIZE, SIZE).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)          # 256 MB > 126 MB L2

    def step_device():
        synthesis.synthesize_volumes(model, dev_in, ar, use_original=True, out=dev_out,
                                     decode_chunk=a.chunk, encode_chunk=a.chunk)

    pipe = synthesis.HostPipeline(model, V, Z, SIZE, SIZE, ar, groups=a.groups, chunk=a.chunk)

    def step_host():
        # a stream of batches through the public host-buffer API: every step uploads its volumes from pinned host
        # memory and downloads its HR volumes; the copies of step i overlap the compute of step i+1, and the timed
        # region ends only after the LAST step's results are in host memory (pipe.wait() below)
        pipe.run(host_in, host_out, wait=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps, flush_l2):
        """K steps, device timing with CUDA events on the launching stream, L2 flushed between steps."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s in range(steps):
            if flush_l2:
                flush.zero_()
            ev[s][0].record()
            step_fn()
            ev[s][1].record()
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    ms_total = timed(step_device, a.steps, True)
    launches = (_lib.launch_count() - l0)
    clocks = sampler.stop() if sampler else None
    slices_per_step = world * V * (Z - 1) * NI
    value = slices_per_step * a.steps / (ms_total * 1e-3)

    # ---- e2e: pinned host volumes -> HR volumes in pinned host memory, H2D/D2H inside the timed region
    for _ in range(max(a.warmup, 3)):
        step_host()
    pipe.wait()
    barrier()
    t0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(a.steps):
        step_host()
    pipe.wait()
    e_end.record()
    barrier()
    e2e_ms = torch.tensor([e_start.elapsed_time(e_end)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = slices_per_step * a.steps / (float(e2e_ms.item()) * 1e-3)
    _ = time.perf_counter() - t0

    # ---- roofline of the dominant kernel family (tcgen05 conv3x3), timed per launch with CUDA events, untimed pass
    roof = None
    if rank == 0:
        peaks = load_peaks()
        ops.TIMING = []
        step_device()
        torch.cuda.synchronize()
        conv_ms = sum(e0.elapsed_time(e1) for name, e0, e1, fl, _d in ops.TIMING if name == "conv3x3")
        conv_fl = sum(fl for name, e0, e1, fl, _d in ops.TIMING if name == "conv3x3")
        n_conv = sum(1 for t in ops.TIMING if t[0] == "conv3x3")
        all_ms = sum(e0.elapsed_time(e1) for name, e0, e1, fl, _d in ops.TIMING)
        kernel_ms = {}
        for name, e0, e1, fl, _d in ops.TIMING:
            kernel_ms[name] = kernel_ms.get(name, 0.0) + e0.elapsed_time(e1)
        ops.TIMING = None
        # ALGORITHMIC FLOPs (reference formulation, minimal-work count: each LR slice encoded once, each
        # synthesized slice decoded once) of the layers the conv kernel family computes = everything but the stem
        # (enc.0 + enc.1, a CUDA-core kernel).  The algebraic folds change the EXECUTED MMA work (dec.0 runs once per
        # LR slice behind the interpolation; folded upsample convs execute the same MACs): reported separately.
        alg_fl = 2e9 * (V * Z * (ENC_GMAC - STEM_GMAC) + V * (Z - 1) * NI * DEC_GMAC)
        # algorithmic bytes of the conv launches of a step (DESIGN.md section 3 table, 16-bit activations): per encoded
        # slice enc.3 .. enc.15 + dec.0 on the latent, per synthesized slice dec.2 .. dec.12+head
        KB = 1024.0
        enc_bytes = (1056.25 + 264.06) + (264.06 + 528.13) + (528.13 + 128) + (128 + 256) + (256 + 256) + (256 + 256)
        dec_bytes = (128 + 128) + (128 + 256) + (256 + 256) + (256 + 256)
        alg_bytes = KB * (V * Z * enc_bytes + V * (Z - 1) * NI * dec_bytes)
        ach = alg_fl / (conv_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        # ncu --set full capture of the same conv launches (tools/profile_round.sh): the newest committed summary
        caps = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_conv_full.json"))
        if caps:
            with open(os.path.join(ROOT, "profiles", caps[-1])) as f:
                traffic = json.load(f)["mean_dram_bytes_per_launch"]
            traffic_src = "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, mean per conv launch)" % caps[-1]
        # memory-bound kernels of the step against the measured HBM copy bandwidth: algorithmic bytes
        HW, HWs = SIZE * SIZE, (SIZE + 2) * (SIZE + 2)
        n_enc, n_syn = V * Z, V * (Z - 1) * NI
        mem_bytes = {"stem": n_enc * (4 * HW + 64 * HWs),
                     "lerp": V * (Z - 1) * 2 * 4 * 64 * (HW // 16) + n_syn * 2 * 64 * (HW // 16),
                     "head": n_syn * (16 * HW + 4 * HW)}
        hbm = []
        for name in ("stem", "lerp", "head"):
            if name in kernel_ms and kernel_ms[name] > 0:
                gbs = mem_bytes[name] / (kernel_ms[name] * 1e-3) / 1e9
                hbm.append({"kernel": {"stem": "stem_mma", "lerp": "lerp_pairs_act", "head": "head_gather"}[name],
                            "algorithmic_bytes": mem_bytes[name], "ms": kernel_ms[name], "achieved_gbs": gbs,
                            "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
        roof = {"bound": "tensor", "kernel": "conv3x3_halo_kernel (tcgen05, all %d launches of a step)" % n_conv,
                "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_conv, 1),
                "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" % peaks["source"],
                "conv_ms_per_step": conv_ms, "all_kernels_ms_per_step": all_ms,
                "conv_share_of_kernel_time": conv_ms / all_ms if all_ms else None,
                "algorithmic_gflop_per_step": alg_fl / 1e9, "executed_gflop_per_step": conv_fl / 1e9,
                "executed_tflops": conv_fl / (conv_ms * 1e-3) / 1e12, "kernel_ms": kernel_ms,
                "memory_bound_kernels": hbm, "hbm_peak_gbs": peaks["hbm_gbs"]}

    train = bench_train(a, dev, rank, world, barrier) if a.train else None

    if rank == 0:
        cpu_rate, cores, cpu_t = cpu_synthesis_rate(a.cpu_sample, 3)
        line = {"metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16" if ops.DEFAULT_DTYPE == torch.float16 else "bf16",
                "data": "synthetic",
                "config": {"workload": "batched HR volume generation: %d ACDC-shaped volumes [10,128,128] per GPU per "
                                       "step, num_interpolations=6 -> %d synthesized slices/GPU/step; ae_combined "
                                       "scales=2 (width 128, latent_width 32, latent 128), seeded synthetic checkpoint"
                                       % (V, V * 54),
                           "volumes_per_gpu": V, "chunk": a.chunk, "l2": "256 MB flush buffer written between steps",
                           "accumulate": "fp32 (TMEM)", "sharding": "volumes over ranks, no collective"},
                "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": pipe.h2d_bytes,
                        "d2h_bytes_per_step": pipe.d2h_bytes, "d2h_note": "synthesized slices only; the kept slices = "
                        "clamp(input) are written into the pinned output by a host thread inside the timed region"
                        if pipe.host_kept else "whole HR volumes", "ms_per_step": float(e2e_ms.item()) / a.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "cpu_baseline": {"value": cpu_rate, "unit": "slices/s", "cores": cores, "kind": "port",
                                 "sample": "%d volumes (%d synthesized slices), best of 3, %.2f s; oracle port of "
                                           "generate_hr_volumes.create_super_volume (re-encodes per alpha like the "
                                           "reference)" % (a.cpu_sample, a.cpu_sample * 54, cpu_t)}}
        if train is not None:
            line["train"] = train
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_GFLOP_PER_STEP = 573.0      # ACDC B=12 step, reference formulation (2*MAC, convs only)


def cpu_train_rate(steps: int = 2):
    """The reference's train step (AETrainerEndToEnd.train restated in the oracle, autograd + Adam) on all host cores."""
    from oracle import aesr_oracle as O
    from oracle.make_golden import acdc_batch
    torch.set_num_threads(os.cpu_count() or 1)
    args = O.default_args(SIZE, 32)
    st = O.init_state(args, seed=892372)
    vgg = O.init_vgg(3)
    d = np.load(os.path.join(ROOT, "superresolution_aniso_mri_b200", "data", "lpips_vgg_lin_v0_1.npz"))
    lins = [torch.from_numpy(d["lin%d" % i]) for i in range(5)]
    adam = O.AdamState(st, lr=1e-5)
    img, mid = acdc_batch(0)
    O.train_step(st, args, adam, img, mid, vgg, lins, ex_loss_weight=0.05)          # warm-up
    t0 = time.perf_counter()
    for s in range(steps):
        O.train_step(st, args, adam, img, mid, vgg, lins, ex_loss_weight=0.05)
    dt = (time.perf_counter() - t0) / steps
    return 12.0 / dt, torch.get_num_threads(), dt


def bench_train(a, dev, rank, world, barrier):
    """BASELINE config 2: ACDC training step, B=12 per GPU (image [24,1,128,128] + slice_between [12,1,128,128]),
    MSE + 0.05 * LPIPS-VGG (seeded random-init VGG16 + shipped lin heads), Adam lr 1e-5; weak scaling, gradients
    averaged over ranks with NCCL.  samples/s, sample = one (from, to, between) triplet."""
    import torch.distributed as dist
    from oracle import aesr_oracle as O
    from oracle.make_golden import acdc_batch
    from networks.net_config import NetworkConfig
    from kwatsch.get_trainer import get_trainer_dynamic
    from superresolution_aniso_mri_b200 import ops
    targs = dict(NetworkConfig("ae_combined", "ACDC").architecture)
    targs.update(dataset="ACDC", model="ae_combined", ae_class="VanillaACAI", width=SIZE, latent_width=32, latent=128,
                 depth=32, lr=1e-5, weight_decay=0.0, epochs=10, device=str(dev), gpu_ids=[dev.index or 0],
                 ex_loss_weight1=0.05, use_percept_loss=False, use_loss_annealing=False, get_masks=False,
                 epoch_threshold=0, log_tensorboard=False, batch_size=12,
                 _vgg_state=[t for pair in O.init_vgg(3) for t in pair])
    torch.manual_seed(892372)
    tr = get_trainer_dynamic(targs)
    host = [tuple(t.pin_memory() for t in acdc_batch(i + 8 * rank)) for i in range(8)]
    devb = [(i.to(dev), m.to(dev)) for i, m in host]
    wa = torch.full((12,), 0.5, device=dev)
    eng, lp = tr.engine, tr.percept_criterion
    steps = a.train_steps

    def step_dev(s):
        img, mid = devb[s % 8]
        eng.step(img, mid, wa, wa, lpips=lp, ex_loss_weight=0.05, lr=1e-5)

    def step_host(s):
        img, mid = host[s % 8]
        tr.train({"image": img, "slice_between": mid}, keep_predictions=False)

    out = {}
    for name, fn in (("device", step_dev), ("e2e", step_host)):
        for s in range(3):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = float(ms.item())
    if rank != 0:
        return None
    ops.TIMING = []
    eng.world = 1            # per-kernel timing pass on rank 0 alone: no collective (the other ranks have moved on)
    step_dev(0)
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1, fl, _d in ops.TIMING:
        agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
    ops.TIMING = None
    cpu_rate, cores, cpu_dt = cpu_train_rate(2)
    ms_step = out["device"] / steps
    ach = TRAIN_GFLOP_PER_STEP / ms_step          # GFLOP / ms = TFLOP/s
    peaks = load_peaks()
    return {"metric": "train_samples_per_sec", "value": world * 12 * steps / (out["device"] * 1e-3), "unit": "samples/s",
            "steps": steps, "ms_per_step": ms_step, "scaling": "weak", "dtype": "bf16",
            "config": {"workload": "ACDC training step: B=12/GPU, 128x128, latent 128, MSE + 0.05*LPIPS-VGG, Adam 1e-5; "
                                   "gradient all-reduce (NCCL, AVG) over ranks"},
            "e2e": {"value": world * 12 * steps / (out["e2e"] * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": 36 * SIZE * SIZE * 4, "d2h_bytes_per_step": 16 + 48,
                    "ms_per_step": out["e2e"] / steps},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                         "note": "whole step (573 algorithmic GFLOP) over step time", "kernel_ms": agg},
            "cpu_baseline": {"value": cpu_rate, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "2 steps of B=12 after 1 warm-up, %.2f s/step, oracle port of "
                                       "AETrainerEndToEnd.train (autograd + Adam)" % cpu_dt}}


JSON_OUT = sys.stdout


def _json_only_stdout():
    """The contract is ONE JSON line on stdout: keep a private handle to the real stdout for it and point file
    descriptor 1 at stderr, so that banners written by libraries (e.g. NCCL's version line under NCCL_DEBUG=VERSION) end
    up on stderr instead of in front of the JSON."""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volumes", type=int, default=64, help="volumes per GPU per step")
    ap.add_argument("--chunk", type=int, default=4096, help="max slices per kernel launch")
    ap.add_argument("--groups", type=int, default=2, help="e2e: volume groups pipelined over copy/compute streams")
    ap.add_argument("--cpu-sample", type=int, default=4, dest="cpu_sample")
    ap.add_argument("--no-train", action="store_false", dest="train", help="skip the training-step measurement")
    ap.add_argument("--train-steps", type=int, default=30, dest="train_steps")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
