"""GPU tier: the fused training step (forward + hand-written backward + Adam on the sm_100a kernels) against the CPU
oracle (autograd) and the reference's own 200-step loss curve (tests/golden/train_acdc_200.npz, produced by
``AETrainerEndToEnd.train`` of the unmodified reference on CPU).

Tolerance (BASELINE.json north_star): training loss curves within 1 % over 200 steps.  Gradients are compared per
parameter tensor by relative L2 error / cosine (activations fp16, gradient tensors bf16, fp32 accumulation)."""
import os

import numpy as np
import pytest
import torch

from oracle import aesr_oracle as O
from oracle.make_golden import acdc_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lins():
    d = np.load(os.path.join(ROOT, "superresolution_aniso_mri_b200", "data", "lpips_vgg_lin_v0_1.npz"))
    return [torch.from_numpy(d["lin%d" % i]) for i in range(5)]


def vgg_flat(seed=3):
    return [t for pair in O.init_vgg(seed) for t in pair]


def trainer_args(width=128, latent_width=32, dataset="ACDC", **over):
    from networks.net_config import NetworkConfig
    a = dict(NetworkConfig("ae_combined", dataset).architecture)
    a.update(dataset=dataset, model="ae_combined", ae_class="VanillaACAI", width=width, latent_width=latent_width,
             latent=128, depth=32, lr=1e-5, weight_decay=0.0, epochs=10, device="cuda:0", gpu_ids=[0],
             ex_loss_weight1=0.05, use_percept_loss=False, use_loss_annealing=False, get_masks=False,
             epoch_threshold=0, log_tensorboard=False, batch_size=12, _vgg_state=vgg_flat())
    a.update(over)
    return a


def make_trainer(args, seed=892372):
    from kwatsch.get_trainer import get_trainer_dynamic
    torch.manual_seed(seed)
    return get_trainer_dynamic(args)


def test_lpips_forward_and_gradient(cuda_lib, golden):
    from superresolution_aniso_mri_b200.lpips_b200 import PerceptualLoss
    dev = torch.device("cuda:0")
    g = golden("lpips_pins.npz")
    lp = PerceptualLoss(vgg_state=vgg_flat(), device=dev)
    gen = torch.Generator().manual_seed(21)
    a = torch.rand(3, 1, 64, 64, generator=gen)
    b = (a + 0.1 * torch.randn(3, 1, 64, 64, generator=gen)).clamp(0, 1)
    got = lp(a.to(dev), b.to(dev), normalize=True).cpu()
    assert got.shape == (3, 1, 1, 1)
    np.testing.assert_allclose(got.numpy(), g["lpips"], rtol=2e-3)           # reference PerceptualLoss output
    syn = b.clone().requires_grad_(True)
    val = O.lpips_forward(O.init_vgg(3), lins(), syn, a, normalize=True).sum()
    gref, = torch.autograd.grad(val, syn)
    _, gg = lp.value_and_grad(a.to(dev), b.to(dev), torch.ones(3, device=dev))
    rel = (gg.cpu() - gref).norm().item() / gref.norm().item()
    assert rel < 0.08, rel


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("cin,cout,n,h,w", [(32, 32, 2, 130, 130), (64, 32, 2, 65, 65), (32, 64, 1, 64, 64),
                                            (128, 64, 2, 32, 32), (128, 128, 1, 32, 32), (64, 64, 1, 5, 3)])
def test_wgrad3x3_vs_torch(cuda_lib, cin, cout, n, h, w, algo):
    """Weight gradient: tcgen05 path (pixels as GEMM K, MN-major operands from NHWC) and CUDA-core path vs torch."""
    import torch.nn.functional as F
    from superresolution_aniso_mri_b200 import ops_train as T
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(cin + cout + h)
    g = (torch.randn(n, h, w, cout, generator=gen) * 0.1).to(torch.bfloat16)
    x = torch.randn(n, h, w, cin, generator=gen).to(torch.bfloat16)
    wz = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    (F.conv2d(x.float().permute(0, 3, 1, 2), wz, padding=1) * g.float().permute(0, 3, 1, 2)).sum().backward()
    dW = torch.zeros(cout, cin, 3, 3, device=dev)
    db = torch.zeros(cout, device=dev)
    T.wgrad3x3(g.to(dev), x.to(dev), dW, db, algo=algo)
    T.wgrad3x3(g.to(dev), x.to(dev), dW, db, algo=algo)                      # accumulates
    scale = max(1.0, wz.grad.abs().max().item())
    assert (dW.cpu() - 2 * wz.grad).abs().max().item() < 2e-3 * scale
    assert (db.cpu() - 2 * g.float().sum(dim=(0, 1, 2))).abs().max().item() < 2e-3 * scale
    # relative L2 against fp32 autograd on the same bf16-rounded operands (fp32 accumulation either way)
    assert (dW.cpu() - 2 * wz.grad).norm().item() <= 2e-3 * (2 * wz.grad).norm().item()


def _trained_state():
    from collections import OrderedDict
    g = np.load(os.path.join(ROOT, "tests", "golden", "trained_ckpt.npz"), allow_pickle=False)
    return OrderedDict((k[len("state__"):], torch.from_numpy(g[k].copy())) for k in g.files if k.startswith("state__"))


@pytest.mark.parametrize("kind,brain", [("rnd", False), ("cal", True), ("trained", False)])
def test_step_gradients_match_oracle_autograd(cuda_lib, kind, brain):
    from superresolution_aniso_mri_b200.lpips_b200 import PerceptualLoss
    from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI
    from superresolution_aniso_mri_b200.training.engine import TrainEngine
    dev = torch.device("cuda:0")
    args = O.default_args(64, 16)
    st = O.init_state(args, seed=892372) if kind == "rnd" else O.calibrated_state(args) if kind == "cal" else _trained_state()
    margs = dict(args)
    margs["device"] = "cuda:0"
    model = VanillaACAI(margs)
    model.load_state_dict(st)
    model.train()
    lp = PerceptualLoss(vgg_state=vgg_flat(), device=dev)
    eng = TrainEngine(model, None)
    B = 4
    if kind == "trained":            # the checkpoint the reference trained, on the kind of images it was trained on
        v = O.mri_phantom(3 * B, 64, seed=6001)
        img, sb = torch.cat([v[0::3], v[2::3]], dim=0), v[1::3].clone()
    else:
        img, sb = acdc_batch(0, B=B, size=64)
    af = torch.tensor([[0.25], [0.5], [0.75], [0.5]]) if brain else None
    at = (1 - af) if brain else None
    wa = (af[:, 0] if brain else torch.full((B,), 0.5)).to(dev)
    wb = (at[:, 0] if brain else torch.full((B,), 0.5)).to(dev)
    st_o = {k: v.clone() for k, v in st.items()}
    lg = O.train_step(st_o, args, None, img, sb, O.init_vgg(3), lins(), alpha_from=af, alpha_to=at,
                      ex_loss_weight=0.05, return_grads=True)
    res = eng.step(img.to(dev), sb.to(dev), wa, wb, lpips=lp, ex_loss_weight=0.05, do_update=False, keep=True)
    logs = eng.logged_losses(res)
    # 'cal' (O(1)-activation stress checkpoint, bf16 training): the logged-only latent MSE sits at ~2e-3 relative and
    # moves in the 4th digit from run to run (fp32 atomics order of the BN statistic sums) -- measured 2.03e-3.
    lim = 2e-3 if kind == "rnd" else 5e-3
    for k in ("loss_ae_dist", "loss_ae_dist_extra", "loss_latent_1", "loss_ae"):
        assert abs(lg[k] - logs[k]) <= lim * abs(lg[k]) + 1e-9, (k, lg[k], logs[k])
    rows = []
    for name, p in model.named_parameters():
        gr, go = lg["grads"][name], eng.grad[id(p)].cpu()
        rows.append((name, (go - gr).norm().item() / max(gr.norm().item(), 1e-30),
                     torch.nn.functional.cosine_similarity(go.flatten(), gr.flatten(), dim=0).item(), gr.norm().item()))
    print("gradient parity, checkpoint %r (bf16 activations and gradients vs fp32 autograd):" % kind)
    for name, rel, cos, nrm in rows:
        print("  %-16s rel-L2 %.3e  cos %.5f  |g| %.3e" % (name, rel, cos, nrm))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "grad_parity_%s.txt" % kind), "w") as f:
            f.write("".join("%-16s rel-L2 %.3e  cos %.5f  |g| %.3e\n" % r for r in rows))
    for name, p in model.named_parameters():
        gr, go = lg["grads"][name], eng.grad[id(p)].cpu()
        rel = (go - gr).norm().item() / max(gr.norm().item(), 1e-30)
        cos = torch.nn.functional.cosine_similarity(go.flatten(), gr.flatten(), dim=0).item()
        # training runs in bf16 (activations AND gradients, fp32 accumulate).  Spec checkpoint (random init): tight.
        # 'cal' is the O(1)-activation stress checkpoint on which the bf16 forward itself already deviates by up to
        # 0.3 on [0,1] images; the deepest path (enc.0) is the worst tensor.
        # The sums behind these gradients use fp32 atomics (BN statistics, weight gradients): the worst tensor of the
        # stress checkpoint moves in the 3rd digit between runs (measured rel 0.47, cos 0.9285 .. 0.94).
        # 'trained' = the checkpoint the reference itself trained (tests/golden/trained_ckpt.npz), on the kind of images it was
        # trained on: measured rel-L2 1e-3 .. 8e-3 per tensor, cos >= 0.99997 (profiles/r05c_grad_parity_*.txt) -> 2e-2.
        # 'cal' is the synthetic stress checkpoint (He-gain weights, |gamma| up to 3): its error grows monotonically from the
        # head (6e-3 at dec.14) to the input (0.48 at enc.0) -- every layer's bf16-rounded activation flips some LeakyReLU /
        # ReLU masks of the backward pass, and the stress gains amplify that through 13 layers; it bounds the chain, it is
        # not what a trained model sees.
        lim_rel, lim_cos = (0.2, 0.98) if kind == "rnd" else (0.6, 0.90) if kind == "cal" else (2e-2, 0.9999)
        assert rel < lim_rel and cos > lim_cos, (name, rel, cos)
    sd = model.state_dict()
    for k in sd:                                                   # BN running statistics + counters (App. B item 7)
        if "running" in k:
            tol = dict(rtol=2e-3, atol=2e-4) if kind == "rnd" else dict(rtol=1e-2, atol=3e-3)   # bf16 activations
            assert torch.allclose(sd[k].cpu(), st_o[k], **tol), k
        if "num_batches" in k:
            assert int(sd[k]) == int(st_o[k]) == int(st[k]) + 2


def test_acdc_200_step_loss_curve_within_1_percent(cuda_lib, golden):
    """BASELINE config 2: B=12, 128x128, MSE + 0.05 LPIPS, Adam lr 1e-5, fixed cycle of 8 synthetic batches."""
    g = golden("train_acdc_200.npz")
    tr = make_trainer(trainer_args())
    assert type(tr).__name__ == "AETrainerEndToEnd"
    for s in range(200):
        img, mid = acdc_batch(s % 8)
        tr.train({"image": img, "slice_between": mid}, keep_predictions=False)
    for key in ("loss_ae", "loss_ae_dist", "loss_ae_dist_extra"):
        ours, ref = np.array(tr.losses[key]), g[key]
        rel = np.abs(ours - ref) / np.abs(ref)
        print(key, "max rel dev %.4f at step %d, mean %.4f" % (rel.max(), rel.argmax(), rel.mean()))
        assert rel.max() < 0.01, (key, rel.max(), int(rel.argmax()))
    assert tr.iters == 201 and int(tr.model.enc[5].num_batches_tracked) == 400


def test_trainer_interface_checkpoint_roundtrip_and_validate(cuda_lib, tmp_path):
    args = trainer_args(width=64, latent_width=16, dataset="dHCP", ex_loss_weight1=0.001, output_dir=str(tmp_path),
                        dir_models=str(tmp_path))
    tr = make_trainer(args)
    assert type(tr).__name__ == "AETrainerExtension1Brain"
    B = 4
    img, mid = acdc_batch(1, B=B, size=64)
    batch = {"image": img, "slice_between": mid, "alpha_from": torch.tensor([[0.25], [0.5], [0.75], [0.5]]),
             "alpha_to": torch.tensor([[0.75], [0.5], [0.25], [0.5]])}
    for _ in range(3):
        tr.train(batch, keep_predictions=True)
    assert set(tr.losses) >= {"loss_ae", "loss_ae_dist", "loss_ae_extra", "loss_ae_dist_extra", "loss_latent_1"}
    assert tr.train_predictions["reconstruction"].shape == (2 * B, 1, 64, 64)
    assert tr.train_predictions["slice_inbetween_mix"].shape == (B, 1, 64, 64)
    fname = os.path.join(str(tmp_path), "3.models")
    tr.save_models(fname, 3)
    ck = torch.load(fname, map_location="cpu")
    assert set(ck) == {"model_dict_ae", "optimizer_dict_ae", "epoch"}
    assert set(ck["optimizer_dict_ae"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}       # stock Adam state
    tr2 = make_trainer(dict(args), seed=1)
    tr2.load(fname)
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v, tr2.model.state_dict()[k]), k
    assert tr2.engine.step_count == tr.engine.step_count == 3                  # Adam state restored into the flat buffers
    assert torch.equal(tr2.engine.flat_m, tr.engine.flat_m) and torch.equal(tr2.engine.flat_v, tr.engine.flat_v)
    tr.train(batch, keep_predictions=False)
    tr2.train(batch, keep_predictions=False)
    # two trainers, same state, same batch: the BN statistics / gradient sums are fp32 atomics, so the forward loss repeats
    # only to ~1e-4 relative (measured spread 0 .. 1.1e-4 over the round's runs, gpurun_out/pytest_r01t_mma.log)
    assert abs(tr.losses["loss_ae"][-1] - tr2.losses["loss_ae"][-1]) < 5e-4 * abs(tr.losses["loss_ae"][-1]) + 1e-9
    # fp32 atomics make the weight-gradient sums order-dependent: allow a fraction of one lr-sized Adam step (1e-5)
    assert torch.allclose(tr.model.enc[1].weight, tr2.model.enc[1].weight, rtol=0, atol=4e-6)
    # eval-mode API used by the synthesis loops + validation bookkeeping
    z = tr.encode(img, eval=True)
    out = tr.decode(z, eval=True)
    assert z.shape == (2 * B, 128, 16, 16) and out.shape == (2 * B, 1, 64, 64)
    val = tr.validate(batch)
    assert "loss_ae" in val and len(tr.losses_test["loss_ae_dist_extra"]) == 1
    assert tr.model.training


@pytest.mark.parametrize("cfg", ["oasis", "dhcp"])
def test_brain_config_train_steps_match_oracle(cuda_lib, cfg):
    """BASELINE configs 3 / 4: OASIS (width 64, latent_width 16, B=16, alpha .5 / .5, weight 0.001) and dHCP (width 256,
    latent_width 64, B=4 of the 8 here to bound the CPU oracle, per-sample alphas in {.25,.5,.75}): three optimisation
    steps of AETrainerExtension1Brain against the oracle's autograd + Adam, logged losses within 1 %."""
    width, lw, B = (64, 16, 16) if cfg == "oasis" else (256, 64, 4)
    args = trainer_args(width=width, latent_width=lw, dataset="OASIS" if cfg == "oasis" else "dHCP", ex_loss_weight1=0.001,
                        batch_size=B)
    tr = make_trainer(args)
    assert type(tr).__name__ == "AETrainerExtension1Brain"
    oargs = O.default_args(width, lw)
    st = O.init_state(oargs, seed=892372)
    adam = O.AdamState(st, lr=1e-5)
    vgg = O.init_vgg(3)
    rs = np.random.RandomState(3)
    for step in range(3):
        img, mid = acdc_batch(step, B=B, size=width)
        if cfg == "oasis":
            af = torch.full((B, 1), 0.5)
        else:
            af = torch.from_numpy(rs.choice([0.25, 0.5, 0.75], size=(B, 1)).astype(np.float32))
        at = 1 - af
        lg = O.train_step(st, oargs, adam, img, mid, vgg, lins(), alpha_from=af, alpha_to=at, ex_loss_weight=0.001)
        tr.train({"image": img, "slice_between": mid, "alpha_from": af, "alpha_to": at}, keep_predictions=False)
        for k in ("loss_ae", "loss_ae_dist", "loss_ae_dist_extra"):
            ours, ref = tr.losses[k][-1], lg[k]
            assert abs(ours - ref) <= 0.01 * abs(ref) + 1e-9, (cfg, step, k, ours, ref)


def test_graphed_step_matches_eager_step(cuda_lib):
    """CUDA-graph replay of the training step (default) against the eager step: same losses and parameters after 6
    steps on a cycle of 3 batches, up to the run-to-run spread of the fp32-atomic sums (the Adam bias corrections are
    formed on the device in the graph, on the host otherwise); the stock optimizer state keeps counting."""
    runs = {}
    for graph in (True, False):
        tr = make_trainer(trainer_args(lr=1e-4))
        tr.engine.use_graph = graph
        for s in range(6):
            img, mid = acdc_batch(s % 3)
            tr.train({"image": img, "slice_between": mid}, keep_predictions=False)
        assert tr.engine.step_count == 6 and (len(tr.engine._graphs) == 1) == graph
        assert float(tr.opt_ae.state[tr.engine.params[0]]["step"]) == 6.0
        runs[graph] = (list(tr.losses["loss_ae"]), tr.engine.flat_p.clone())
    la, lb = runs[True][0], runs[False][0]
    assert len(la) == len(lb) == 6
    assert max(abs(a - b) / abs(b) for a, b in zip(la, lb)) < 2e-3, (la, lb)
    # early Adam steps move every weight by ~lr * sign(grad): where the gradient is noise around zero (fp32-atomic order)
    # two runs can drift apart by up to 2 * 6 * lr; almost everywhere they agree to ~1e-6 (measured: max 7.6e-4, typical 1e-6)
    diff = (runs[True][1] - runs[False][1]).abs()
    assert diff.max().item() < 1.3e-3 and diff.mean().item() < 2e-5


def test_plain_ae_trainer_steps_match_oracle(cuda_lib, golden):
    """``--model ae`` (kwatsch/trainer_ae.py:71-109): MSE-only step, logged latent loss from an EVAL-mode encode of
    slice_between (base_trainer.py:203-205) which leaves the module in eval mode when train() returns (:251-252; train()
    switches back at its start).  Four steps against the oracle (autograd + Adam) and the reference's own logged losses
    (tests/golden/train_small.npz 'plain': 32x32, B=4, seeded uniform batches)."""
    from networks.net_config import NetworkConfig
    g = golden("train_small.npz")
    args = trainer_args(width=32, latent_width=8, batch_size=4)
    args.update(NetworkConfig("ae", "ACDC").architecture)
    args.update(width=32, latent_width=8, latent=128, depth=32, model="ae")
    tr = make_trainer(args)
    assert type(tr).__name__ == "AEBaseTrainer" and tr.percept_criterion is None and not tr.combined
    oargs = O.default_args(32, 8)
    st = O.init_state(oargs, seed=892372)
    adam = O.AdamState(st, lr=args["lr"])
    gen = torch.Generator().manual_seed(11)
    for step in range(4):
        img = torch.rand(8, 1, 32, 32, generator=gen)
        sb = torch.rand(4, 1, 32, 32, generator=gen)
        lg = O.train_step(st, oargs, adam, img, sb, combined=False)
        tr.train({"image": img, "slice_between": sb}, keep_predictions=(step == 3))
        assert not tr.model.training                      # the quirk: eval mode after train()
        for k in ("loss_ae", "loss_latent_1"):
            ours, ref = tr.losses[k][-1], lg[k]
            assert abs(ours - ref) <= 0.01 * abs(ref) + 1e-9, (step, k, ours, ref)
            assert abs(ours - float(g["plain_" + k][step])) <= 0.01 * abs(float(g["plain_" + k][step])) + 1e-9
    assert "loss_ae_extra" not in tr.losses
    assert tr.train_predictions["slice_inbetween_mix"].shape == (4, 1, 32, 32)
    assert int(tr.model.enc[5].num_batches_tracked) == 4   # one train-mode encoder pass per step (the eval encode adds none)
    sd = tr.model.state_dict()
    for k, v in sd.items():
        if "running" in k:
            assert torch.allclose(v.cpu(), st[k], rtol=5e-3, atol=5e-4), k


def test_loss_annealing_weight_and_scheduler_share_one_graph(cuda_lib):
    """use_loss_annealing=True (kwatsch/cardiac/trainer_ae.py:80, table base_trainer.py:456-459): the LPIPS weight of epoch e
    is loss_weights[e]; with use_lr_scheduler the learning rate changes EVERY iteration (CosineAnnealingLR).  Both travel
    through device memory, so one captured graph serves all values -- checked against the oracle with the same weights / lrs."""
    import math
    args = trainer_args(width=64, latent_width=16, batch_size=4, use_loss_annealing=True, epochs=4, lr=1e-4,
                        use_lr_scheduler=True, lr_iter_max=8)
    tr = make_trainer(args)
    x = np.linspace(-5, 5, 4)
    want_w = (1.0 / (1.0 + np.exp(-x)) * 0.05)[::-1]
    np.testing.assert_allclose(tr.loss_weights, want_w, rtol=1e-12)
    oargs = O.default_args(64, 16)
    st = O.init_state(oargs, seed=892372)
    adam = O.AdamState(st, lr=1e-4)
    vgg = O.init_vgg(3)
    it = 0
    for epoch in range(3):
        for s in range(2):
            img, mid = acdc_batch(2 * epoch + s, B=4, size=64)
            adam.lr = 1e-4 * (1 + math.cos(math.pi * it / 8)) / 2          # CosineAnnealingLR(T_max=8, eta_min=0) closed form
            assert abs(tr.opt_ae.param_groups[0]["lr"] - adam.lr) < 1e-12
            lg = O.train_step(st, oargs, adam, img, mid, vgg, lins(), ex_loss_weight=float(want_w[epoch]))
            tr.train({"image": img, "slice_between": mid}, keep_predictions=False)
            for k in ("loss_ae", "loss_ae_dist", "loss_ae_dist_extra"):
                ours, ref = tr.losses[k][-1], lg[k]
                assert abs(ours - ref) <= 0.01 * abs(ref) + 1e-9, (epoch, s, k, ours, ref)
            it += 1
        tr.epoch += 1
    assert len(tr.engine._graphs) == 1 and tr.engine.step_count == 6
    # Adam moves every weight by ~lr per step, so the mean distance travelled measures the sum of the learning rates that
    # were actually applied: schedule 1, .96, .85, .69, .5, .31 (x 1e-4) = 4.3e-4 against 6e-4 for an lr stuck in a graph
    p0 = O.init_state(oargs, seed=892372)["dec.0.weight"]
    ours = (tr.model.dec[0].weight.detach().cpu() - p0).abs().mean().item()
    ref = (st["dec.0.weight"] - p0).abs().mean().item()
    assert ref > 1e-4 and abs(ours - ref) < 0.12 * ref, (ours, ref)


def test_dhcp_batch8_train_step(cuda_lib):
    """BASELINE config 4 at its real batch size (width 256, latent_width 64, B=8): two steps run, losses finite and equal
    to the oracle's within 1 % on the first step (one oracle step at this size takes a few seconds of CPU)."""
    args = trainer_args(width=256, latent_width=64, dataset="dHCP", ex_loss_weight1=0.001, batch_size=8)
    tr = make_trainer(args)
    B = 8
    rs = np.random.RandomState(5)
    oargs = O.default_args(256, 64)
    st = O.init_state(oargs, seed=892372)
    adam = O.AdamState(st, lr=1e-5)
    for step in range(2):
        img, mid = acdc_batch(step, B=B, size=256)
        af = torch.from_numpy(rs.choice([0.25, 0.5, 0.75], size=(B, 1)).astype(np.float32))
        tr.train({"image": img, "slice_between": mid, "alpha_from": af, "alpha_to": 1 - af}, keep_predictions=False)
        if step == 0:
            lg = O.train_step(st, oargs, adam, img, mid, O.init_vgg(3), lins(), alpha_from=af, alpha_to=1 - af,
                              ex_loss_weight=0.001)
            for k in ("loss_ae", "loss_ae_dist", "loss_ae_dist_extra"):
                assert abs(tr.losses[k][-1] - lg[k]) <= 0.01 * abs(lg[k]) + 1e-9, (k, tr.losses[k][-1], lg[k])
    assert all(np.isfinite(v).all() for v in tr.losses.values())


def test_merged_batch_batchnorm_equals_two_passes(cuda_lib):
    """The merged-batch BatchNorm entry points (conv statistics with ``stats_split``, bn_finalize with two passes, bn_apply /
    bn_bwd with ``split``) against the same kernels run once per pass."""
    from superresolution_aniso_mri_b200 import ops, ops_train as T
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(8)
    n, n0, h, w, c = 5, 3, 18, 22, 64
    x = torch.randn(n, h, w, c, generator=g).to(dt).to(dev)
    wt = ops.pack_conv3x3_weight((torch.randn(c, c, 3, 3, generator=g) / 24).to(dev), dtype=dt)
    b = (torch.randn(c, generator=g) * 0.1).to(dev)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(dev), (torch.randn(c, generator=g) * 0.1).to(dev)
    stats = torch.zeros(4 * c, device=dev)
    a = ops.conv3x3(x, wt, b, act=ops.ACT_LEAKY, stats=stats, stats_split=n0)
    parts, st_parts = [], []
    for sl in (slice(0, n0), slice(n0, n)):
        sp = torch.zeros(2 * c, device=dev)
        parts.append(ops.conv3x3(x[sl].contiguous(), wt, b, act=ops.ACT_LEAKY, stats=sp))
        st_parts.append(sp)
    assert torch.equal(a, torch.cat(parts))
    assert torch.allclose(stats, torch.cat(st_parts), rtol=1e-5, atol=1e-3)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    rm2, rv2 = rm.clone(), rv.clone()
    sc, sh, mean, inv = T.bn_finalize(stats, n0 * h * w, gamma, beta, rm, rv, 0.1, 1e-5, count1=(n - n0) * h * w)
    outs = [T.bn_finalize(st_parts[k], cnt * h * w, gamma, beta, rm2, rv2, 0.1, 1e-5) for k, cnt in ((0, n0), (1, n - n0))]
    for k in range(2):
        for got, want in zip((sc, sh, mean, inv), outs[k]):
            assert torch.allclose(got[k * c:(k + 1) * c], want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rm, rm2, rtol=1e-5, atol=1e-7) and torch.allclose(rv, rv2, rtol=1e-5, atol=1e-7)
    for mode in (T.BN_POOL, T.BN_UP):
        y = T.bn_apply(a, sc, sh, mode, split=n0)
        want = torch.cat([T.bn_apply(parts[k], outs[k][0], outs[k][1], mode) for k in range(2)])
        # scale / shift of the merged finalize agree with the per-pass ones to fp32 rounding: at most one bf16 ulp apart
        assert ((y.float() - want.float()).abs() <= 2.0 ** -7 * want.float().abs() + 1e-6).all()
        dn = (torch.randn(y.shape, generator=g) * 1e-3).to(torch.bfloat16).to(dev)
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        dg2, db2 = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        gg = T.bn_bwd(dn, a, mean, inv, gamma, dg, db, mode, split=n0)
        want_g = torch.cat([T.bn_bwd(dn[sl].contiguous(), parts[k], outs[k][2], outs[k][3], gamma, dg2, db2, mode)
                            for k, sl in enumerate((slice(0, n0), slice(n0, n)))])
        assert (gg.float() - want_g.float()).abs().max().item() <= 2e-2 * want_g.float().abs().max().item()
        assert torch.allclose(dg, dg2, rtol=1e-3, atol=1e-6) and torch.allclose(db, db2, rtol=1e-3, atol=1e-6)


def test_device_triplet_loader_feeds_trainer_and_follows_reference_draw_order(cuda_lib):
    """sample_triplet -> gather_triplets -> augment_batch -> prepare_batch_pairs as ONE batch source that
    ``trainer.train()`` consumes.  The batches equal a host replay of the reference's per-sample ``__getitem__`` order
    (triplet draws, then the transform draws) through the oracle's pinned ``sample_triplet`` / ``augment_sample``."""
    from superresolution_aniso_mri_b200.data_loader import DeviceTripletLoader
    dev = torch.device("cuda:0")
    vols = [O.mri_phantom(12, 150, seed=70 + i)[:, 0, :, :141].numpy() for i in range(3)]
    for kind, kw in (("acdc", dict(width=128, aug_patch=160, center=True)), ("brain", dict(width=64, downsample_steps=4))):
        rs = np.random.RandomState(99)
        loader = DeviceTripletLoader(vols, batch_size=4, kind=kind, rs=rs, device=dev, **kw)
        batches = [b for _, b in zip(range(3), loader)]
        rs2 = np.random.RandomState(99)
        perm = rs2.permutation(len(loader.items))
        for bi, b in enumerate(batches):
            want_img, want_af = [], []
            for i in perm[bi * 4:(bi + 1) * 4]:
                vi, z, Z = loader.items[i]
                t = O.sample_triplet(z, Z, rs2, kind=kind, slice_selection="adjacent_plus", downsample_steps=kw.get("downsample_steps", 2))
                img3 = np.stack([vols[vi][t["slice_idx_from"]], vols[vi][t["slice_idx_to"]], vols[vi][t["inbetween_slice_id"]]])
                out, _ = O.augment_sample(img3, rs2, width=kw["width"], aug_patch=kw.get("aug_patch"), center=kw.get("center", False),
                                          intensity_first=(kind == "acdc"))
                want_img.append(out)
                want_af.append(float(t["alpha_from"]))
            want = torch.from_numpy(np.stack(want_img))
            got = torch.cat([b["image"][:4], b["image"][4:], b["slice_between"]], dim=1).cpu()      # back to [B,3,H,W]
            assert got.shape == want.shape == (4, 3, kw["width"], kw["width"])
            assert (got - want).abs().max().item() <= 4 * 2.0 ** -23 * max(1.0, want.abs().max().item())
            assert torch.allclose(b["alpha_from"].cpu().reshape(-1), torch.tensor(want_af))
            assert b["image"].is_cuda and b["image"].shape == (8, 1, kw["width"], kw["width"])
    # end to end: the loader drives the ACDC trainer for a few iterations
    tr = make_trainer(trainer_args(width=64, latent_width=16, batch_size=4))
    loader = DeviceTripletLoader(vols, batch_size=4, kind="acdc", width=64, aug_patch=96, center=True,
                                 rs=np.random.RandomState(5), device=dev)
    n = 0
    for batch_item in loader:
        tr.train(batch_item, keep_predictions=False)
        n += 1
        if n == 4:
            break
    assert n == 4 and len(tr.losses["loss_ae"]) == 4 and all(np.isfinite(tr.losses["loss_ae"]))


def test_encode_decode_refuse_autograd_inputs(cuda_lib):
    """networks/acai_vanilla.py:130-138 is differentiable through autograd; this module is not -- it must say so."""
    from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI
    args = dict(O.default_args(64, 16))
    args["device"] = "cuda:0"
    m = VanillaACAI(args).eval()
    x = torch.rand(2, 1, 64, 64, device="cuda:0", requires_grad=True)
    with pytest.raises(RuntimeError, match="builds no autograd graph"):
        m.encode(x)
    with torch.no_grad():
        z = m.encode(x)
    assert not z.requires_grad
    with pytest.raises(RuntimeError, match="builds no autograd graph"):
        m(x)
    z2 = z.clone().requires_grad_(True)
    with pytest.raises(RuntimeError, match="builds no autograd graph"):
        m.decode(z2)
    assert m.decode(z).shape == (2, 1, 64, 64)              # plain tensors with grad mode on: fine
