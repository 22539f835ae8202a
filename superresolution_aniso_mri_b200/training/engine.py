"""Fused training step of the ae_combined autoencoder on the sm_100a kernels.

One call = what ``AETrainerEndToEnd.train`` / ``AETrainerExtension1Brain.train`` do between ``x = batch['image']`` and
``opt_ae.step()`` (kwatsch/cardiac/trainer_ae.py:10-36, kwatsch/brain/trainer_ae.py:92-118):

    z      = enc(x)                    [2B]   train-mode BN (batch statistics, running stats updated)
    out    = dec(z)                    [2B]
    z_mix  = wa * z[:B] + wb * z[B:]   [B]    (0.5/0.5 cardiac, per-sample alphas brain)
    s_mix  = dec(z_mix)                [B]    second decoder pass, its own BN statistics
    z_ref  = enc(slice_between)        [B]    logged latent MSE only -- but it does update BN running stats
    loss   = MSE(out, x) + w * mean_b LPIPS(slice_between, s_mix)
    backward (dgrad + wgrad for the AE, dgrad only through VGG), Adam.

Merged passes: convolutions do not care which BatchNorm pass an image belongs to, so enc(x) and enc(slice_between) run as
ONE 3B-image launch per layer, and so do dec(z) and dec(z_mix); only the BatchNorm statistics are kept per pass
(``split`` = 2B: images [0, 2B) and [2B, 3B) have their own batch statistics and update the running statistics in the
reference's order, include/aesr_b200.h "Merged batches").  The decoder backward runs on the merged batch as well (weight
gradients of both passes land in the same buffers, as autograd would sum them); the encoder backward runs on the 2B-image
prefix -- enc(slice_between) only feeds a logged scalar.

No autograd: forward saves exactly the tensors the hand-written backward needs.  Parameters, gradients and Adam
moments live in flat fp32 buffers (parameters are views), so the optimizer is ONE kernel and data-parallel training is
ONE (two, overlapped) NCCL all-reduce(s) of the flat gradient buffer.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import ops, ops_train as T
from ..networks.acai_vanilla import BatchNormHolder, ConvHolder, VanillaACAI

SLOPE = 0.01


class _ConvRec:
    __slots__ = ("conv", "x_in", "prev", "bn")

    def __init__(self, conv, x_in, prev, bn=None):
        self.conv, self.x_in, self.prev, self.bn = conv, x_in, prev, bn


class _BnRec:
    __slots__ = ("bn", "a", "mean", "invstd", "mode", "split")

    def __init__(self, bn, a, mean, invstd, mode, split=0):
        self.bn, self.a, self.mean, self.invstd, self.mode, self.split = bn, a, mean, invstd, mode, split


class TrainForward:
    """Train-mode (batch-statistics BatchNorm) forward passes; also what ``VanillaACAI.encode/decode`` run when the
    module is in ``.train()`` mode outside the fused step (e.g. ``BaseTrainer.encode(x, eval=False)``)."""

    def __init__(self, model: VanillaACAI, sync_bn: bool = False, act_dtype: Optional[torch.dtype] = None):
        self.model = model
        self.sync_bn = sync_bn
        self.dev = next(model.parameters()).device
        # Training activations are bf16 like the gradients: the tensor-core weight-gradient GEMM multiplies a gradient
        # tile by an activation tile, and tcgen05 kind::f16 needs both operands in ONE 16-bit format (bf16 x fp16 is an
        # illegal instruction -- measured); gradients need bf16's range (dL/dx ~ 1e-6..1e-9).
        self.dtype = act_dtype or T.GRAD_DTYPE
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.w_fwd, self.w_bwd = {}, {}
        self.flat_p, self._pack_plan = None, None


class TrainEngine(TrainForward):
    def __init__(self, model: VanillaACAI, optimizer: Optional[torch.optim.Adam] = None, sync_bn: bool = False,
                 act_dtype: Optional[torch.dtype] = None):
        super().__init__(model, sync_bn, act_dtype)
        self.opt = optimizer
        self.params = [p for p in model.parameters()]
        self.step_count = 0
        # CUDA-graph replay of the whole step (168 launches; the host needs 3.2 ms to enqueue what the device runs in
        # 3.9 ms, and the launch gaps between ~60 short conv kernels cost another 11 %, tools/train_graph_probe.py).
        # Single-process only (the NCCL all-reduces of the data-parallel step stay eager); AESR_TRAIN_GRAPH=0 disables.
        self.use_graph = os.environ.get("AESR_TRAIN_GRAPH", "1") != "0"
        # data-parallel: capture the NCCL all-reduces inside the step graph as well.  OPT-IN (AESR_TRAIN_GRAPH_DP=1): measured
        # on 2 x B200 (profiles/r03c_dp_graph_check.txt) the replayed step matches the eager one (DP CHECK OK) and runs at
        # 3.16 instead of 3.47 ms, but dist.destroy_process_group() blocked while the captured graphs were alive -- call
        # release_graphs() before tearing the process group down (that teardown order has not been re-run on a GPU yet).
        self.use_graph_dp = os.environ.get("AESR_TRAIN_GRAPH_DP", "1") != "0"
        self._graphs = {}
        self._graph_seen = {}
        # Weight-gradient GEMMs on a second stream: wgrad(layer k) and the data-gradient chain (dgrad k -> BN backward ->
        # dgrad k-1 ...) both start from the same output gradient and do not depend on each other; the wgrad launches use
        # 28-136 CTAs (atomics / MMA cost model), so they fill SMs the short dgrad launches leave idle.  Joined before the
        # gradient all-reduce / Adam.  AESR_TRAIN_WGRAD_STREAM=0 keeps everything on one stream.
        self.overlap_wgrad = os.environ.get("AESR_TRAIN_WGRAD_STREAM", "1") != "0"
        self._side = None
        self._side_keep = []
        self._flatten()

    # ------------------------------------------------------------------ flat parameter / gradient / moment buffers
    def _flatten(self):
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=self.dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.grad: Dict[int, torch.Tensor] = {}
        off = 0
        self.offsets = []
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                self.grad[id(p)] = self.flat_g[off:off + k].view(p.shape)
                self.offsets.append((off, k))
                off += k
        enc_params = sum(p.numel() for p in self.model.enc.parameters())
        self.enc_numel = enc_params              # encoder parameters come first (module order)
        self._bind_optimizer_state(adopt=True)

    def _enc_deep_offset(self) -> int:
        """Offset (in the flat buffers) of the first parameter of the encoder's last two convs (module order = flat order)."""
        convs = [m for m in self.model.enc if isinstance(m, ConvHolder)]
        first = convs[-2].weight
        for p, (off, _k) in zip(self.params, self.offsets):
            if p is first:
                return off
        raise AssertionError("encoder parameter not found in the flat buffer")

    def _bind_optimizer_state(self, adopt: bool):
        """Make opt.state[p]['exp_avg' / 'exp_avg_sq'] views of the flat moment buffers so that
        ``opt.state_dict()`` stays a stock Adam state_dict (checkpoint layout of kwatsch/base_trainer.py:353-356)."""
        if self.opt is None:
            return
        for p, (off, k) in zip(self.params, self.offsets):
            st = self.opt.state[p]
            if adopt and "exp_avg" in st:       # state loaded from a checkpoint: copy it into the flat buffers
                self.flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count = int(st["step"]) if "step" in st else self.step_count
            st["exp_avg"] = self.flat_m[off:off + k].view(p.shape)
            st["exp_avg_sq"] = self.flat_v[off:off + k].view(p.shape)
            st["step"] = torch.tensor(float(self.step_count))

    def reload_optimizer_state(self):
        """Call after ``opt.load_state_dict`` (BaseTrainer.load)."""
        self._bind_optimizer_state(adopt=True)

    # ------------------------------------------------------------------ per-step derived tensors
    def _pack_all(self):
        """16-bit [9][Cout][Cin] forward filters and [9][Cin][Cout] mirrored data-gradient filters of every 3x3 conv, re-formed
        from the current fp32 parameters.  TrainEngine (flat parameter buffer): ONE launch per dtype over a static job
        table; the forward-only helper packs layer by layer."""
        if getattr(self, "flat_p", None) is None:
            self.w_fwd, self.w_bwd = {}, {}
            for seq in (self.model.enc, self.model.dec):
                for m in seq:
                    if isinstance(m, ConvHolder) and m.kernel_size == 3 and m.out_channels > 1:
                        self.w_fwd[id(m)] = ops.pack_conv3x3_weight(m.weight, dtype=self.dtype)
                        self.w_bwd[id(m)] = ops.pack_conv3x3_weight(m.weight, transpose_flip=True, dtype=T.GRAD_DTYPE)
            return
        if getattr(self, "_pack_plan", None) is None:
            off_of = {id(p): off for p, (off, _k) in zip(self.params, self.offsets)}
            convs = [m for seq in (self.model.enc, self.model.dec) for m in seq
                     if isinstance(m, ConvHolder) and m.kernel_size == 3 and m.out_channels > 1]
            plan = {}
            for dtype, kinds in ({self.dtype: (0, 1)} if self.dtype == T.GRAD_DTYPE else
                                 {self.dtype: (0,), T.GRAD_DTYPE: (1,)}).items():
                jobs, views, pos = [], [], 0
                for m in convs:
                    k = 9 * m.out_channels * m.in_channels
                    for flip in kinds:
                        jobs.append([off_of[id(m.weight)], pos, m.out_channels, m.in_channels, flip, 0])
                        views.append((m, flip, pos, k))
                        pos += k
                buf = torch.empty(pos, dtype=dtype, device=self.dev)
                for m, flip, p0, k in views:
                    shape = (9, m.in_channels, m.out_channels) if flip else (9, m.out_channels, m.in_channels)
                    (self.w_bwd if flip else self.w_fwd)[id(m)] = buf[p0:p0 + k].view(shape)
                plan[dtype] = (buf, torch.tensor(jobs, dtype=torch.int64, device=self.dev),
                               max(9 * m.out_channels * m.in_channels for m in convs))
            self._pack_plan = plan
        for buf, jobs, max_elems in self._pack_plan.values():
            ops.pack_conv3x3_weight_batch(self.flat_p, buf, jobs, max_elems)

    def _bn_train(self, bn: BatchNormHolder, stats: torch.Tensor, count: int, update_running: bool = True,
                  count1: int = 0):
        """``count1`` > 0: two passes of a merged batch ([2][2C] sums) -> [2, C] outputs, two running-stat updates."""
        if self.sync_bn and self.world > 1:
            dist.all_reduce(stats)
            count *= self.world
            count1 *= self.world
        out = T.bn_finalize(stats, count, bn.weight, bn.bias, bn.running_mean if update_running else None,
                            bn.running_var if update_running else None, bn.momentum, bn.eps, count1=count1)
        if update_running:
            bn.num_batches_tracked += 2 if count1 > 0 else 1
        return out

    # ------------------------------------------------------------------ forward passes (train mode)
    def encode_train(self, x: torch.Tensor, save: bool = True, split: int = 0):
        """``split``: images >= split are a second module call (own BatchNorm batch statistics), 0 = one call."""
        enc, sc = self.model.enc, self.model.scales
        n = x.shape[0]
        n0 = split if 0 < split < n else n
        recs: List[_ConvRec] = []
        a = ops.e0(x, enc[0].weight.reshape(-1), enc[0].bias, dtype=self.dtype)
        prev = "e0"
        i = 1
        for _ in range(sc):
            c1, c2, bn = enc[i], enc[i + 2], enc[i + 4]
            recs.append(_ConvRec(c1, a, prev))
            a1 = ops.conv3x3(a, self.w_fwd[id(c1)], c1.bias, act=ops.ACT_LEAKY)
            stats = torch.zeros((2 if n0 < n else 1) * 2 * c2.out_channels, dtype=torch.float32, device=self.dev)
            a2 = ops.conv3x3(a1, self.w_fwd[id(c2)], c2.bias, act=ops.ACT_LEAKY, stats=stats, stats_split=n0)
            hw = a2.shape[1] * a2.shape[2]
            scale, shift, mean, invstd = self._bn_train(bn, stats, n0 * hw, count1=(n - n0) * hw)
            a = T.bn_apply(a2, scale, shift, T.BN_POOL, split=n0)
            recs.append(_ConvRec(c2, a1, "leaky", _BnRec(bn, a2, mean, invstd, T.BN_POOL, n0)))
            prev = "bn"
            i += 6
        c1, c2 = enc[i], enc[i + 2]
        recs.append(_ConvRec(c1, a, prev))
        a1 = ops.conv3x3(a, self.w_fwd[id(c1)], c1.bias, act=ops.ACT_LEAKY)
        recs.append(_ConvRec(c2, a1, "leaky"))
        z, z16 = ops.conv3x3(a1, self.w_fwd[id(c2)], c2.bias, act=ops.ACT_NONE, out_mode=ops.OUT_NCHW_F32,
                             want_out2=True)
        return z, z16, (recs if save else None)

    def decode_train(self, z16: torch.Tensor, save: bool = True, split: int = 0):
        dec, sc = self.model.dec, self.model.scales
        n = z16.shape[0]
        n0 = split if 0 < split < n else n
        recs: List[_ConvRec] = []
        a, prev = z16, "latent"
        i = 0
        for _ in range(sc):
            c1, c2, bn = dec[i], dec[i + 2], dec[i + 4]
            recs.append(_ConvRec(c1, a, prev))
            a1 = ops.conv3x3(a, self.w_fwd[id(c1)], c1.bias, act=ops.ACT_LEAKY)
            stats = torch.zeros((2 if n0 < n else 1) * 2 * c2.out_channels, dtype=torch.float32, device=self.dev)
            a2 = ops.conv3x3(a1, self.w_fwd[id(c2)], c2.bias, act=ops.ACT_LEAKY, stats=stats, stats_split=n0)
            hw = a2.shape[1] * a2.shape[2]
            scale, shift, mean, invstd = self._bn_train(bn, stats, n0 * hw, count1=(n - n0) * hw)
            a = T.bn_apply(a2, scale, shift, T.BN_UP, split=n0)
            recs.append(_ConvRec(c2, a1, "leaky", _BnRec(bn, a2, mean, invstd, T.BN_UP, n0)))
            prev = "bn"
            i += 6
        c1, head = dec[i], dec[i + 2]
        recs.append(_ConvRec(c1, a, prev))
        a1 = ops.conv3x3(a, self.w_fwd[id(c1)], c1.bias, act=ops.ACT_LEAKY)
        w9c, hb = self._head_params(head)
        out = ops.head(a1, w9c, hb, sigmoid=True)
        return out, ((recs, a1, out, head, w9c) if save else None)

    def _head_params(self, head):
        """[9,32] fp32 filter + bias of dec.14, formed HERE (not through the model's version-keyed eval cache): inside a
        captured step graph the tensors must be recomputed on replay and owned by the graph's pool."""
        with torch.no_grad():
            return (head.weight.detach()[0].permute(1, 2, 0).reshape(9, -1).contiguous().float(),
                    head.bias.detach().float().contiguous())

    # ------------------------------------------------------------------ backward passes
    def _on_side(self, fn, keep=()):
        """Run ``fn`` on the second stream once everything enqueued so far on the current stream is done."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            out = fn()
        self._side_keep.append((keep, out))     # keep operands / results alive (allocator reuse) until the join
        return out

    def _wgrad(self, g, x_in, dW, dbias):
        if not self.overlap_wgrad or ops.TIMING is not None:
            T.wgrad3x3(g, x_in, dW, dbias)
            return
        self._on_side(lambda: T.wgrad3x3(g, x_in, dW, dbias), keep=(g, x_in))

    def _join_side(self):
        if self._side is not None and self._side_keep:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.dev).wait_event(ev)
            self._side_keep.clear()

    def _backward_convs(self, recs: List[_ConvRec], g: torch.Tensor, x_img: Optional[torch.Tensor] = None,
                        bias_done: bool = False, after_launch=None):
        """g: bf16 gradient w.r.t. the LAST conv's pre-activation output.  Returns the gradient w.r.t. the stage input
        (latent for the decoder, nothing for the encoder whose first op is enc.0).  ``g`` may cover only the first
        images of the saved activations (encoder: enc(slice_between) has no gradient): saved tensors are sliced to it.
        Bias gradients: the kernel that PRODUCES a layer's output gradient also sums it over the pixels (conv epilogue
        ``stats_split=-1``, BN backward ``dbias_conv``, head backward ``dbias_in``), so the weight-gradient launch only
        falls back to its own column-sum pass when nobody did (``bias_done`` tells about the incoming ``g``).
        ``after_launch(k)`` is called once everything that writes the gradients of layers >= k has been enqueued."""
        n = g.shape[0]
        fuse_bn_bias = not (self.sync_bn and self.world > 1)
        for k in range(len(recs) - 1, -1, -1):
            r = recs[k]
            x_in = r.x_in[:n]
            self._wgrad(g, x_in, self.grad[id(r.conv.weight)], None if bias_done else self.grad[id(r.conv.bias)])
            wt = self.w_bwd[id(r.conv)]
            db_prev = self.grad[id(recs[k - 1].conv.bias)] if k > 0 else None
            if r.prev == "leaky":
                prev_rec = recs[k - 1]
                if prev_rec.bn is not None:
                    raise AssertionError("a conv fed by a BN'd tensor is tagged 'bn'")
                g = ops.conv3x3(g, wt, None, mul_src=x_in, mul_mode=ops.MUL_LEAKY_GRAD, slope=SLOPE, stats=db_prev,
                                stats_split=-1)
                bias_done = True
            elif r.prev == "bn":
                bnrec = recs[k - 1].bn
                dnext = ops.conv3x3(g, wt, None)
                g = T.bn_bwd(dnext, bnrec.a[:n], bnrec.mean, bnrec.invstd, bnrec.bn.weight,
                             self.grad[id(bnrec.bn.weight)], self.grad[id(bnrec.bn.bias)], bnrec.mode, SLOPE,
                             sync_world=self.world if self.sync_bn else 1, split=bnrec.split if bnrec.split < n else 0,
                             dbias_conv=db_prev if fuse_bn_bias else None)
                bias_done = fuse_bn_bias
            if after_launch is not None:
                after_launch(k)
            if r.prev == "e0":
                d_a0 = ops.conv3x3(g, wt, None)
                e0 = self.model.enc[0]
                T.e0_bwd(d_a0, x_img, self.grad[id(e0.weight)].view(-1), self.grad[id(e0.bias)])
                return None
            if r.prev == "latent":
                return ops.conv3x3(g, wt, None)
        return None

    def decode_backward(self, ctx, dout: torch.Tensor):
        recs, a1, out, head, w9c = ctx
        g = T.head_bwd(dout, out, a1, w9c, self.grad[id(head.weight)].view(-1), self.grad[id(head.bias)], SLOPE,
                       dbias_in=self.grad[id(recs[-1].conv.bias)])
        return self._backward_convs(recs, g, bias_done=True)

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, image: torch.Tensor, slice_between: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
             lpips=None, ex_loss_weight: float = 0.0, combined: bool = True, do_update: bool = True,
             lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
             keep: bool = False) -> dict:
        """One optimisation step.  With ``use_graph`` (default, single process) the second and later calls of a given
        (shapes, loss weight, lr, ...) configuration replay a CUDA graph of the step: the inputs are copied into the
        graph's static buffers, the Adam step count travels through device memory, and the returned tensors are the
        graph's static outputs (valid until the next step -- the trainers read them right away)."""
        if lr is None and do_update:
            lr = self.opt.param_groups[0]["lr"]
        if (self.use_graph and (self.world == 1 or self.use_graph_dp) and do_update and not keep and ops.TIMING is None
                and not torch.cuda.is_current_stream_capturing()):
            return self._step_graphed(image, slice_between, wa, wb, lpips, ex_loss_weight, combined, lr, betas, eps,
                                      weight_decay)
        return self._step_impl(image, slice_between, wa, wb, lpips, ex_loss_weight, combined, do_update, lr, betas, eps,
                               weight_decay, keep)

    MAX_GRAPHS = 2      # captured configurations kept (each owns a private pool with one step's activations)

    def _step_graphed(self, image, slice_between, wa, wb, lpips, ex_loss_weight, combined, lr, betas, eps, weight_decay):
        # lr and the LPIPS weight travel through device memory (like the Adam step count): one graph serves a per-iteration
        # lr scheduler (kwatsch/base_trainer.py:18-22) and the per-epoch loss-weight annealing (:456-459).
        key = (tuple(image.shape), tuple(slice_between.shape), tuple(wa.shape), bool(combined),
               tuple(float(b) for b in betas), float(eps), float(weight_decay), id(lpips))
        ent = self._graphs.get(key)
        B = image.shape[0] // 2
        if ent is None:
            if self._graph_seen.get(key, 0) < 1:
                # first step of this configuration runs eagerly (one-time host work: shared-memory attributes, caches)
                if len(self._graph_seen) > 64:
                    self._graph_seen.clear()
                self._graph_seen[key] = 1
                return self._step_impl(image, slice_between, wa, wb, lpips, ex_loss_weight, combined, True, lr, betas,
                                       eps, weight_decay, False)
            while len(self._graphs) >= self.MAX_GRAPHS:           # least recently used configuration goes (with its pool)
                torch.cuda.synchronize(self.dev)
                self._graphs.pop(next(iter(self._graphs)))
            ent = {"x": image.detach().float().contiguous().clone(), "sb": slice_between.detach().float().contiguous().clone(),
                   "wa": wa.detach().clone(), "wb": wb.detach().clone(),
                   "step": torch.zeros(1, dtype=torch.int32, device=self.dev),
                   "lr": torch.full((1,), float(lr), dtype=torch.float32, device=self.dev),
                   "upstream": torch.full((B,), ex_loss_weight / B, dtype=torch.float32, device=self.dev)}
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            # thread-local capture mode: the NCCL watchdog thread polls events while the step is being captured
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                ent["res"] = self._step_impl(ent["x"], ent["sb"], ent["wa"], ent["wb"], lpips, ex_loss_weight, combined,
                                             True, lr, betas, eps, weight_decay, False, step_dev=ent["step"],
                                             lr_dev=ent["lr"], upstream_dev=ent["upstream"])
            ent["graph"] = graph
            self._graphs[key] = ent
        else:
            self._graphs[key] = self._graphs.pop(key)            # mark as most recently used
        ent["x"].copy_(image, non_blocking=True)
        ent["sb"].copy_(slice_between, non_blocking=True)
        ent["wa"].copy_(wa, non_blocking=True)
        ent["wb"].copy_(wb, non_blocking=True)
        self.step_count += 1
        ent["step"].fill_(self.step_count)
        ent["lr"].fill_(float(lr))
        ent["upstream"].fill_(ex_loss_weight / B)
        ent["graph"].replay()
        self._after_update()
        ent["res"]["ex_loss_weight"] = ex_loss_weight
        return ent["res"]

    def release_graphs(self) -> None:
        """Drop every captured step graph (and its private memory pool)."""
        torch.cuda.synchronize(self.dev)
        self._graphs.clear()
        self._graph_seen.clear()

    def _after_update(self):
        if self.opt is not None:
            for p in self.params:                # one tensor per parameter, like torch.optim.Adam keeps them
                self.opt.state[p]["step"] = torch.tensor(float(self.step_count))
        self.model.invalidate_cache()            # weights / BN buffers changed under the eval-path caches

    def _step_impl(self, image: torch.Tensor, slice_between: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
                   lpips=None, ex_loss_weight: float = 0.0, combined: bool = True, do_update: bool = True,
                   lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                   keep: bool = False, step_dev: Optional[torch.Tensor] = None, lr_dev: Optional[torch.Tensor] = None,
                   upstream_dev: Optional[torch.Tensor] = None) -> dict:
        m = self.model
        m.invalidate_cache()            # nothing derived by an earlier eval-mode call may leak into this step (or its graph)
        B = image.shape[0] // 2
        x = image.detach().float().contiguous()
        sb = slice_between.detach().float().contiguous()
        self._pack_all()
        self.flat_g.zero_()
        scal = torch.zeros(4, dtype=torch.float32, device=self.dev)      # [mse_recon, latent_mse, spare, spare]
        idx = torch.arange(B, dtype=torch.int32, device=self.dev)
        val = None
        if combined:
            # (The LPIPS trunk of the reference images depends on the batch only; running it on the second stream under the
            # autoencoder's forward was measured SLOWER -- 2.44 vs 2.39 ms per step: two 12-image trunk passes cost more than
            # one 24-image pass because the 8x8 / 16x16 layers are launch-bound -- so both halves stay one batch.)
            # enc(x) and enc(slice_between) as one 3B-image pass, dec(z) and dec(z_mix) as another (see module docstring)
            z3, z16, enc_ctx = self.encode_train(torch.cat([x, sb], dim=0), split=2 * B)
            z, z_ref = z3[:2 * B], z3[2 * B:]                 # z_ref: logged only; its pass updated the BN running stats
            _, z_mix = ops.lerp_latents(z3, idx, idx + B, wa, wb, want_nchw=True, out=z16[2 * B:])
            out3, dec_ctx = self.decode_train(z16, split=2 * B)
            out, s_mix = out3[:2 * B], out3[2 * B:]
            dout3 = torch.empty_like(out3)
            T.mse(out, x, scal[0:1], grad_out=dout3[:2 * B])
            T.mse(z_mix, z_ref, scal[1:2])
            upstream = upstream_dev if upstream_dev is not None else \
                torch.full((B,), ex_loss_weight / B, dtype=torch.float32, device=self.dev)
            val, _ = lpips.value_and_grad(sb, s_mix, upstream, normalize=True, grad_out=dout3[2 * B:])
            # ---- backward
            g3 = self.decode_backward(dec_ctx, dout3)                     # [3B,h,w,latent] bf16
            g_z = T.mix_bwd(g3[:2 * B], g3[2 * B:], wa, wb)
        else:
            z, z16, enc_ctx = self.encode_train(x)
            out, dec_ctx = self.decode_train(z16)
            dout = T.mse(out, x, scal[0:1], want_grad=True)
            _, z_mix = ops.lerp_latents(z, idx, idx + B, wa, wb, want_nchw=True, dtype=self.dtype)
            s_mix = None
            # AEBaseTrainer: eval-mode encode (no_grad).  Layer-per-kernel pipeline: its derived tensors (packed filters,
            # folded BN affine) are formed on the device, so the pass can be captured in the step graph (the folded stem
            # carries its filter as kernel parameters = a host round trip); dropped again so they cannot outlive the step.
            fused, m.fused_inference = m.fused_inference, False
            try:
                z_ref = m.encode_eval(sb)
            finally:
                m.fused_inference = fused
            m.invalidate_cache()
            T.mse(z_mix, z_ref, scal[1:2])
            g_z = self.decode_backward(dec_ctx, dout)                     # [2B,h,w,latent] bf16
        # Data parallel: three gradient buckets, each all-reduced (NCCL, AVG) as soon as the launches that write it are
        # enqueued -- decoder after the decoder backward, the two deepest encoder convs (75 % of the encoder's parameters)
        # after their weight gradients, the shallow rest at the end (the only exposed one: ~0.3 MB).  The collectives are
        # enqueued from the second stream (behind the weight-gradient GEMMs, which run there), so the data-gradient chain on
        # the main stream never waits for them.
        handles = []

        def reduce_bucket(lo, hi):
            def run():
                return dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.AVG, async_op=True)
            if self.overlap_wgrad and ops.TIMING is None:
                handles.append(self._on_side(run))
            else:
                handles.append(run())

        deep_lo = self._enc_deep_offset()
        if self.world > 1:
            reduce_bucket(self.enc_numel, self.flat_g.numel())

        def enc_hook(k):
            if self.world > 1 and k == len(enc_ctx) - 2:                  # enc.13 and enc.15: gradients enqueued
                reduce_bucket(deep_lo, self.enc_numel)

        self._backward_convs(enc_ctx, g_z, x_img=x, after_launch=enc_hook)
        if self.world > 1:
            reduce_bucket(0, deep_lo)
        self._join_side()
        for h in handles:
            h.wait()

        if do_update and step_dev is not None:       # graph capture: step count / lr are read from device memory at replay time
            T.adam_step_dev(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr, betas[0], betas[1], eps, weight_decay,
                            step_dev, lr_dev)
        elif do_update:
            self.step_count += 1
            if lr is None:
                lr = self.opt.param_groups[0]["lr"]
            T.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr, betas[0], betas[1], eps, weight_decay,
                        self.step_count)
            self._after_update()
        if not do_update:
            m.invalidate_cache()        # BN running statistics changed under the eval-path caches
        res = {"scalars": scal, "lpips_per_image": val, "B": B, "ex_loss_weight": ex_loss_weight}
        if keep:
            res.update(reconstruction=out, s_between_mix=s_mix, z=z, z_mix=z_mix)
        return res

    @staticmethod
    def logged_losses(res: dict) -> dict:
        """One device->host read for all logged scalars (the reference does five .item() syncs per step)."""
        scal = res["scalars"].tolist()
        out = {"loss_ae_dist": scal[0], "loss_latent_1": scal[1]}
        if res["lpips_per_image"] is not None:
            extra = res["ex_loss_weight"] * float(res["lpips_per_image"].mean().item())
            out["loss_ae_dist_extra"] = extra
            out["loss_ae_extra"] = extra
            out["loss_ae"] = scal[0] + extra
        else:
            out["loss_ae"] = scal[0]
        return out


# The forward-only helper shares the pass implementations with the engine (they only touch TrainForward state).
for _name in ("_pack_all", "_bn_train", "encode_train", "decode_train", "_head_params"):
    setattr(TrainForward, _name, getattr(TrainEngine, _name))
