"""GPU bring-up diagnostics (run on the B200 box): per-case error report for the conv kernel, the halo-descriptor
probe, and an end-to-end inference parity + rough timing.  Prints everything; never stops at the first failure."""
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superresolution_aniso_mri_b200 import _lib, ops, build  # noqa: E402

build.build_library()
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
lib = _lib.lib_for_device(0)
print("SMs", lib.aesr_sm_count())


def ref_conv(x_nhwc, w, b, act, scale=None, shift=None, mode=ops.OUT_SAME, dtype=torch.bfloat16):
    x = x_nhwc.float().permute(0, 3, 1, 2).cpu()
    wb = w.to(dtype).float().cpu()
    y = F.conv2d(x, wb, None if b is None else b.cpu(), padding=1)
    if act == ops.ACT_LEAKY:
        y = F.leaky_relu(y, 0.01)
    elif act == ops.ACT_RELU:
        y = F.relu(y)
    if scale is not None:
        y = y * scale.cpu()[None, :, None, None] + shift.cpu()[None, :, None, None]
    if mode == ops.OUT_AVGPOOL2:
        y = F.avg_pool2d(y, 2)
    elif mode == ops.OUT_UP2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return y


def run_case(cin, cout, n, h, w, act, mode, affine=False, seed=0, dtype=torch.bfloat16, algo=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g).to(dtype).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)).to(dev)
    b = torch.randn(cout, generator=g).to(dev) * 0.1
    sc = (torch.rand(cout, generator=g) + 0.5).to(dev) if affine else None
    sh = torch.randn(cout, generator=g).to(dev) * 0.1 if affine else None
    wp = ops.pack_conv3x3_weight(wt, dtype=dtype)
    try:
        res = ops.conv3x3(x, wp, b, act=act, scale=sc, shift=sh, out_mode=mode, algo=algo)
    except RuntimeError as e:
        if algo == 1 and "too large" in str(e):
            return True
        raise
    torch.cuda.synchronize()
    want = ref_conv(x, wt, b, act, sc, sh, mode if mode != ops.OUT_SAME_MAXPOOL2 else ops.OUT_SAME, dtype=dtype)
    if mode == ops.OUT_NCHW_F32:
        got = res.cpu()
    elif mode == ops.OUT_SAME_MAXPOOL2:
        got = res[0].float().permute(0, 3, 1, 2).cpu()
        got2 = res[1].float().permute(0, 3, 1, 2).cpu()
        e2 = (got2 - F.max_pool2d(want, 2)).abs().max().item()
        print("      maxpool out2 err %.4f" % e2)
    else:
        got = res.float().permute(0, 3, 1, 2).cpu()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    ok = err < 0.03 * max(ref, 1.0)
    print("conv %s algo%d %3d->%3d n%d %3dx%3d act%d mode%d aff%d : max err %.4f (ref max %.2f) %s"
          % (str(dtype)[6:], algo, cin, cout, n, h, w, act, mode, affine, err, ref, "OK" if ok else "FAIL"))
    return ok


cases = [
    (64, 64, 1, 16, 8, 0, 0, False), (64, 64, 2, 32, 32, 1, 0, False), (32, 32, 1, 16, 8, 0, 0, False),
    (32, 32, 2, 130, 130, 1, 0, False), (32, 32, 2, 130, 130, 1, 1, True), (32, 64, 2, 65, 65, 1, 0, False),
    (64, 64, 2, 65, 65, 1, 1, True), (64, 128, 2, 32, 32, 1, 0, False), (128, 128, 2, 32, 32, 0, 3, False),
    (128, 64, 3, 32, 32, 1, 0, False), (64, 64, 3, 32, 32, 1, 2, True), (64, 32, 3, 64, 64, 1, 0, False),
    (32, 32, 3, 64, 64, 1, 2, True), (32, 32, 3, 128, 128, 1, 0, False), (128, 256, 1, 16, 16, 2, 0, False),
    (256, 256, 1, 16, 16, 2, 4, False), (256, 512, 1, 8, 8, 2, 0, False), (512, 512, 1, 8, 8, 2, 0, False),
    (64, 64, 1, 55, 55, 1, 1, True), (64, 64, 300, 32, 32, 1, 0, False),
]
all_ok = True
for dtype, algo in ((torch.bfloat16, 1), (torch.float16, 1), (torch.float16, 2)):
    for c in cases:
        try:
            all_ok &= run_case(*c, dtype=dtype, algo=algo)
        except Exception:
            all_ok = False
            traceback.print_exc()
print("CONV ALL OK" if all_ok else "CONV FAILURES")

# ------------------------------------------------------------------ halo descriptor probe
try:
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 40, 40, 64, generator=g).to(torch.bfloat16).to(dev)
    wt = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
    wp = ops.pack_conv3x3_weight(wt, dtype=torch.bfloat16)
    want_full = F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), wt.to(torch.bfloat16).float().cpu(), padding=1)
    for (x0, y0) in ((8, 16), (0, 0), (32, 32)):
        want = torch.zeros(128, 64)
        for r in range(128):
            yy, xx = y0 + r // 8, x0 + r % 8
            if yy < 40 and xx < 40:
                want[r] = want_full[0, :, yy, xx]
        for pitch in (10,):
            for variant in (0,):
                out = torch.zeros(128, 64, device=dev)
                _lib.check(_lib.load_probe().aesr_probe_halo_conv(x.data_ptr(), wp.data_ptr(), out.data_ptr(), 1, 40, 40, x0, y0, 0,
                                                    pitch, variant, torch.cuda.current_stream().cuda_stream), "probe")
                torch.cuda.synchronize()
                valid = torch.tensor([(y0 + r // 8 < 40) and (x0 + r % 8 < 40) for r in range(128)])
                err = (out.cpu() - want)[valid].abs().max().item()
                print("halo probe tile(%d,%d) pitch %d variant %d: max err %.4f %s"
                      % (x0, y0, pitch, variant, err, "OK" if err < 0.02 else "WRONG"))
except Exception:
    traceback.print_exc()

# ------------------------------------------------------------------ end-to-end inference parity + timing
try:
    from oracle import aesr_oracle as O
    from superresolution_aniso_mri_b200 import synthesis
    from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI
    args = O.default_args(128, 32)
    st = O.calibrated_state(args)
    margs = dict(args); margs["device"] = "cuda:0"
    model = VanillaACAI(margs); model.load_state_dict(st); model.eval()
    vol = O.smooth_phantom(10, 128, seed=2)
    with torch.no_grad():
        z_ref = O.encode(st, args, vol)
        rec_ref = O.decode(st, args, z_ref)
    z = model.encode(vol.to(dev))
    rec = model.decode(z)
    print("encode: max err %.4e (z std %.3f, max %.2f)" % ((z.cpu() - z_ref).abs().max().item(), z_ref.std().item(), z_ref.abs().max().item()))
    print("decode(encode): max err %.4e" % (rec.cpu() - rec_ref).abs().max().item())
    rec2 = model.decode(z_ref.to(dev))
    print("decode(z_ref): max err %.4e" % (rec2.cpu() - rec_ref).abs().max().item())
    ar = O.alpha_range_for(6)
    want = O.create_super_volume(st, args, vol, ar, use_original=True)
    got = synthesis.create_super_volume(model, vol, ar, use_original=True)["upsampled_image"]
    d = (got - want).abs()
    print("create_super_volume ni=6: shape %s max err %.4e mean err %.3e" % (tuple(got.shape), d.max().item(), d.mean().item()))
    # rough timing: 64 volumes
    V = 64
    vols = torch.rand(V, 10, 128, 128, device=dev)
    for chunk in (64, 128, 256, 1024):
        for _ in range(2):
            synthesis.synthesize_volumes(model, vols, ar, decode_chunk=chunk, encode_chunk=chunk)
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 3
        for _ in range(reps):
            synthesis.synthesize_volumes(model, vols, ar, decode_chunk=chunk, encode_chunk=chunk)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / reps
        print("synthesize %d volumes chunk %d: %.2f ms -> %.0f synthesized slices/s" % (V, chunk, dt * 1e3, V * 54 / dt))
except Exception:
    traceback.print_exc()

# ------------------------------------------------------------------ per-layer timing (decoder/encoder shapes, big batch)
try:
    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    layers = [("E1 32->32@130", 32, 32, 130, 0, 160), ("E2 32->32@130 pool", 32, 32, 130, 1, 160),
              ("E3 32->64@65", 32, 64, 65, 0, 160), ("E4 64->64@65 pool", 64, 64, 65, 1, 160),
              ("E5 64->128@32", 64, 128, 32, 0, 160), ("E6 128->128@32 nchw", 128, 128, 32, 3, 160),
              ("D0 128->64@32", 128, 64, 32, 0, 864), ("D1 64->64@32 up", 64, 64, 32, 2, 864),
              ("D2 64->32@64", 64, 32, 64, 0, 864), ("D3 32->32@64 up", 32, 32, 64, 2, 864),
              ("D4 32->32@128", 32, 32, 128, 0, 864)]
    dt = torch.float16
    for name, cin, cout, hw, mode, n in layers:
        x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
        wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, dtype=dt)
        b = torch.zeros(cout, device=dev)
        out = torch.empty(ops.conv_out_shape(n, hw, hw, cout, mode), dtype=torch.float32 if mode == 3 else dt, device=dev)
        res = []
        for algo in (1, 2):
            ms = timeit(lambda: ops.conv3x3(x, wp, b, act=1, out_mode=mode, out=out, algo=algo))
            fl = 2.0 * n * hw * hw * 9 * cin * cout
            by = x.numel() * 2 + out.numel() * out.element_size()
            res.append("algo%d %.3f ms %.0f TF/s %.0f GB/s" % (algo, ms, fl / ms / 1e9, by / ms / 1e6))
        print("%-22s n=%d : %s" % (name, n, " | ".join(res)))
except Exception:
    traceback.print_exc()
