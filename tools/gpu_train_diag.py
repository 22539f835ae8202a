"""GPU bring-up of the training step: per-parameter gradient comparison against the CPU oracle (autograd), loss values,
BN running stats, a few optimisation steps, rough timing.  Prints everything."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200 import build, ops  # noqa: E402
from superresolution_aniso_mri_b200.lpips_b200 import PerceptualLoss  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402
from superresolution_aniso_mri_b200.training.engine import TrainEngine  # noqa: E402

build.build_library()
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())


def load_lins():
    d = np.load(os.path.join(ROOT, "superresolution_aniso_mri_b200", "data", "lpips_vgg_lin_v0_1.npz"))
    return [torch.from_numpy(d["lin%d" % i]) for i in range(5)]


def make(width, lw, state):
    args = O.default_args(width, lw)
    margs = dict(args); margs["device"] = "cuda:0"
    m = VanillaACAI(margs); m.load_state_dict(state); m.train()
    return args, m


def run(width, lw, B, state_kind, brain=False, steps=3, ex_w=0.05):
    print("==== width %d lw %d B %d state %s brain %s" % (width, lw, B, state_kind, brain))
    args = O.default_args(width, lw)
    st = O.init_state(args, seed=892372) if state_kind == "rnd" else O.calibrated_state(args)
    _, model = make(width, lw, st)
    vgg = O.init_vgg(3)
    lp = PerceptualLoss(vgg_state=[t for pair in vgg for t in pair], device="cuda:0")
    eng = TrainEngine(model, None)
    g = torch.Generator().manual_seed(11)
    from oracle.make_golden import acdc_batch
    img, sb = acdc_batch(0, B=B, size=width)
    if brain:
        af = torch.tensor([[0.25], [0.5], [0.75], [0.5]] * (B // 4 + 1))[:B]
        at = 1 - af
    else:
        af = at = None
    wa = (af[:, 0] if brain else torch.full((B,), 0.5)).to(dev)
    wb = (at[:, 0] if brain else torch.full((B,), 0.5)).to(dev)
    # ---- LPIPS alone
    with torch.no_grad():
        want = O.lpips_forward(vgg, load_lins(), img[:B], sb, normalize=True).flatten()
    got = lp(img[:B].to(dev), sb.to(dev), normalize=True).flatten().cpu()
    print("lpips fwd: ref %s got %s relerr %.3e" % (want[:3].tolist(), got[:3].tolist(), ((got - want).abs() / want.abs()).max().item()))
    syn = img[:B].clone().requires_grad_(True)
    val = O.lpips_forward(vgg, load_lins(), syn, sb, normalize=True).mean()
    gref, = torch.autograd.grad(val, syn)
    up = torch.full((B,), 1.0 / B, device=dev)
    _, gg = lp.value_and_grad(sb.to(dev), img[:B].to(dev), up)
    print("lpips grad: rel l2 err %.3e (|g| %.3e)" % ((gg.cpu() - gref).norm().item() / gref.norm().item(), gref.norm().item()))
    # ---- gradients of one step
    st_o = {k: v.clone() for k, v in st.items()}
    lg = O.train_step(st_o, args, None, img, sb, vgg, load_lins(), alpha_from=af, alpha_to=at, ex_loss_weight=ex_w,
                      return_grads=True)
    res = eng.step(img.to(dev), sb.to(dev), wa, wb, lpips=lp, ex_loss_weight=ex_w, do_update=False, keep=True)
    logs = eng.logged_losses(res)
    for k in ("loss_ae_dist", "loss_ae_dist_extra", "loss_latent_1", "loss_ae"):
        print("  %-20s oracle %.6e ours %.6e rel %.2e" % (k, lg[k], logs[k], abs(lg[k] - logs[k]) / max(abs(lg[k]), 1e-30)))
    print("  recon max err %.3e  s_mix max err %.3e" % ((res["reconstruction"].cpu() - lg["reconstruction"]).abs().max().item(),
                                                          (res["s_between_mix"].cpu() - lg["s_between_mix"]).abs().max().item()))
    worst = 0
    for p_name, p in model.named_parameters():
        gr = lg["grads"][p_name]
        go = eng.grad[id(p)].cpu()
        rel = (go - gr).norm().item() / max(gr.norm().item(), 1e-30)
        cos = torch.nn.functional.cosine_similarity(go.flatten(), gr.flatten(), dim=0).item()
        worst = max(worst, rel)
        print("  grad %-14s |ref| %.3e rel-l2 %.3e cos %.5f" % (p_name, gr.norm().item(), rel, cos))
    print("  worst rel-l2 grad err %.3e" % worst)
    sd = model.state_dict()
    for k in sd:
        if "running" in k or "num_batches" in k:
            e = (sd[k].float().cpu() - st_o[k].float()).abs().max().item()
            print("  %-26s max err %.3e (ref max %.3e)" % (k, e, st_o[k].float().abs().max().item()))
    # ---- a few optimisation steps
    st_o = {k: v.clone() for k, v in st.items()}
    adam = O.AdamState(st_o, lr=1e-3)
    _, model = make(width, lw, st)
    eng = TrainEngine(model, None)
    for s in range(steps):
        img, sb = acdc_batch(s, B=B, size=width)
        lg = O.train_step(st_o, args, adam, img, sb, vgg, load_lins(), alpha_from=af, alpha_to=at, ex_loss_weight=ex_w)
        res = eng.step(img.to(dev), sb.to(dev), wa, wb, lpips=lp, ex_loss_weight=ex_w, lr=1e-3)
        logs = eng.logged_losses(res)
        print("  step %d loss_ae oracle %.6f ours %.6f | extra %.6f %.6f" % (s, lg["loss_ae"], logs["loss_ae"], lg["loss_ae_dist_extra"], logs["loss_ae_dist_extra"]))
    sd = model.state_dict()
    rel = max((sd[k].float().cpu() - st_o[k].float()).norm().item() / max(st_o[k].float().norm().item(), 1e-30) for k in sd if sd[k].dtype.is_floating_point)
    print("  max rel param diff after %d steps (lr 1e-3): %.3e" % (steps, rel))


for cfg in ((64, 16, 4, "rnd", False), (64, 16, 4, "cal", True)):
    try:
        run(*cfg)
    except Exception:
        traceback.print_exc()

# ---- timing of the BASELINE config 2 step
try:
    args = O.default_args(128, 32)
    _, model = make(128, 32, O.init_state(args, seed=892372))
    lp = PerceptualLoss(vgg_state=[t for pair in O.init_vgg(3) for t in pair], device="cuda:0")
    eng = TrainEngine(model, None)
    from oracle.make_golden import acdc_batch
    img, sb = acdc_batch(0)
    img, sb = img.to(dev), sb.to(dev)
    wa = torch.full((12,), 0.5, device=dev)
    for _ in range(3):
        eng.step(img, sb, wa, wa, lpips=lp, ex_loss_weight=0.05, lr=1e-5)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(10):
        eng.step(img, sb, wa, wa, lpips=lp, ex_loss_weight=0.05, lr=1e-5)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 10
    print("ACDC B=12 step: %.2f ms -> %.0f samples/s" % (dt * 1e3, 12 / dt))
    ops.TIMING = []
    eng.step(img, sb, wa, wa, lpips=lp, ex_loss_weight=0.05, lr=1e-5)
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1, fl in ops.TIMING:
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += e0.elapsed_time(e1)
    ops.TIMING = None
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("   %-14s n=%3d %.3f ms" % (k, n, ms))
except Exception:
    traceback.print_exc()
