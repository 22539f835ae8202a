"""GPU tier: variants of the tcgen05 conv kernels against each other and against torch fp32 on 16-bit-rounded operands.
* conv3x3_fold_kernel (32 -> 32 layers, horizontal taps folded into N = 96; csrc/conv3x3_fold.cuh; opt-in) against torch fp32 and the
  tap-by-tap halo kernel.  Reference layers: enc.3, dec.8, dec.10 (/root/reference/networks/acai_vanilla.py:55,92,94).
* TMA-store epilogues (staged cp.async.bulk.tensor stores) against per-thread global stores: bit-identical.
Tolerance of the torch comparisons: a few output ulps of the 16-bit storage type, as for the other conv kernels (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_lib):
    return torch.device("cuda:0")


def _case(n, h, w, seed, dtype, dev, affine):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, 32, generator=g).to(dtype).to(dev)
    wt = (torch.randn(32, 32, 3, 3, generator=g) / np.sqrt(32 * 9)).to(dev)
    b = (torch.randn(32, generator=g) * 0.1).to(dev)
    sc = (torch.rand(32, generator=g) + 0.5).to(dev) if affine else None
    sh = (torch.randn(32, generator=g) * 0.1).to(dev) if affine else None
    return x, wt, b, sc, sh


def _ref(x, wt, b, act, sc, sh, mode, dtype):
    y = F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), wt.to(dtype).float().cpu(), b.cpu(), padding=1)
    if act == 1:
        y = F.leaky_relu(y, 0.01)
    if sc is not None:
        y = y * sc.cpu()[None, :, None, None] + sh.cpu()[None, :, None, None]
    if mode == 1:
        y = F.avg_pool2d(y, 2)
    return y


# n, h, w, act, mode (0 same, 1 avg-pool), affine.  14-column output tiles: widths around multiples of 14, one-pixel images,
# odd pooled extents (130 -> 65, 65 -> 32), more super-tiles than SMs (160 x 24 x 24), tall images (T = 2 super-tiles).
FOLD_CASES = [(1, 16, 8, 0, 0, False), (2, 130, 130, 1, 1, True), (2, 64, 64, 1, 0, True), (2, 128, 128, 1, 0, False),
              (1, 3, 5, 1, 1, False), (3, 37, 29, 1, 0, True), (1, 1, 1, 1, 0, False), (2, 15, 14, 0, 0, False),
              (2, 14, 28, 1, 1, True), (160, 24, 24, 1, 0, True), (2, 65, 65, 1, 1, False), (1, 9, 15, 1, 0, False),
              (5, 220, 220, 1, 0, True), (1, 2, 2, 1, 1, False)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("tile", [(0, 0), (1, 4), (2, 2)])
@pytest.mark.parametrize("case", FOLD_CASES)
def test_fold_conv_vs_torch_fp32_and_halo_kernel(dev, case, tile, dtype):
    from superresolution_aniso_mri_b200 import ops
    n, h, w, act, mode, affine = case
    if dtype == torch.bfloat16 and (tile != (0, 0) or n > 3):
        pytest.skip("bf16 covered on the automatic tile shape")
    x, wt, b, sc, sh = _case(n, h, w, h * 131 + w, dtype, dev, affine)
    wp = ops.pack_conv3x3_weight(wt, dtype=dtype)
    try:
        ops.set_tuning(ops.TUNE_FOLD, 1)
        ops.set_tuning(ops.TUNE_CONV_T, tile[0])
        ops.set_tuning(ops.TUNE_CONV_NBUF, tile[1])
        got = ops.conv3x3(x, wp, b, act=act, scale=sc, shift=sh, out_mode=mode)            # ALGO_AUTO + AESR_FOLD: fold kernel
    finally:
        ops.set_tuning(ops.TUNE_FOLD, 0)
        ops.set_tuning(ops.TUNE_CONV_T, 0)
        ops.set_tuning(ops.TUNE_CONV_NBUF, 0)
    halo = ops.conv3x3(x, wp, b, act=act, scale=sc, shift=sh, out_mode=mode, algo=1)       # tap-by-tap halo kernel
    want = _ref(x, wt, b, act, sc, sh, mode, dtype)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    tol = 4 * ulp * max(1.0, want.abs().max().item()) + 2e-4 * np.sqrt(32 * 9)
    g = got.float().permute(0, 3, 1, 2).cpu()
    assert g.shape == want.shape
    assert (g - want).abs().max().item() <= tol
    # same operands, fp32 accumulation in a different order: at most one 16-bit ulp apart from the tap-by-tap kernel
    assert (got.float() - halo.float()).abs().max().item() <= 2 * ulp * max(1.0, want.abs().max().item())


def test_fold_conv_is_opt_in(dev, cuda_lib):
    """The automatic dispatch stays on the tap-by-tap halo kernel (bit-identical to algo = HALO); AESR_FOLD selects the fold."""
    from superresolution_aniso_mri_b200 import ops
    x, wt, b, _, _ = _case(4, 64, 64, 5, torch.float16, dev, False)
    wp = ops.pack_conv3x3_weight(wt, dtype=torch.float16)
    c = ops.conv3x3(x, wp, b, act=1)
    try:
        ops.set_tuning(ops.TUNE_FOLD, 1)
        a = ops.conv3x3(x, wp, b, act=1)
    finally:
        ops.set_tuning(ops.TUNE_FOLD, 0)
    h = ops.conv3x3(x, wp, b, act=1, algo=1)
    assert torch.equal(c, h)
    assert (a.float() - h.float()).abs().max().item() <= 2 * 2.0 ** -10 * max(1.0, h.float().abs().max().item())


def test_fold_conv_training_epilogues(dev):
    """act'(mul_src) multiplier, per-channel sums (bias gradient) and BatchNorm statistics with a pass split."""
    from superresolution_aniso_mri_b200 import ops
    dt = torch.float16
    ops.set_tuning(ops.TUNE_FOLD, 1)
    try:
        _fold_training_epilogues(dev, ops, dt)
    finally:
        ops.set_tuning(ops.TUNE_FOLD, 0)


def _fold_training_epilogues(dev, ops, dt):
    x, wt, b, _, _ = _case(4, 33, 20, 11, dt, dev, False)
    g = torch.Generator().manual_seed(12)
    src = torch.randn(4, 33, 20, 32, generator=g).to(dt).to(dev)
    wp = ops.pack_conv3x3_weight(wt, dtype=dt)
    y = F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), wt.to(dt).float().cpu(), b.cpu(), padding=1)
    # data-gradient launch: multiplier only
    got = ops.conv3x3(x, wp, b, act=0, mul_src=src, mul_mode=1)
    ym = y * torch.where(src.float().permute(0, 3, 1, 2).cpu() > 0, 1.0, 0.01)
    assert (got.float().permute(0, 3, 1, 2).cpu() - ym).abs().max().item() < 1e-2
    # data gradient + bias-gradient sums (stats_split < 0: sums only)
    s1 = torch.zeros(32, device=dev)
    got = ops.conv3x3(x, wp, b, act=0, mul_src=src, mul_mode=1, stats=s1, stats_split=-1)
    assert (got.float().permute(0, 3, 1, 2).cpu() - ym).abs().max().item() < 1e-2
    assert torch.allclose(s1.cpu(), ym.sum(dim=(0, 2, 3)), rtol=2e-3, atol=5e-2)
    # forward conv in front of a train-mode BatchNorm: sums / sums of squares of two passes of a merged batch
    s2 = torch.zeros(128, device=dev)
    got = ops.conv3x3(x, wp, b, act=1, stats=s2, stats_split=3)
    ya = F.leaky_relu(y, 0.01)
    assert (got.float().permute(0, 3, 1, 2).cpu() - ya).abs().max().item() < 1e-2
    for k, sl in enumerate((slice(0, 3), slice(3, 4))):
        assert torch.allclose(s2[64 * k:64 * k + 32].cpu(), ya[sl].sum(dim=(0, 2, 3)), rtol=2e-3, atol=5e-2)
        assert torch.allclose(s2[64 * k + 32:64 * k + 64].cpu(), (ya[sl] ** 2).sum(dim=(0, 2, 3)), rtol=2e-3, atol=5e-2)


# ------------------------------------------------------------------------------------------------ TMA-store epilogues
# cin, cout (GEMM N), n, h, w, mode (0 same, 5 depth-to-space), affine: staged cp.async.bulk.tensor stores (clipped at the image
# bounds by the hardware) against per-thread global stores (aesr_set_tuning key 10) -- the same values, so bit-identical.
TMA_STORE_CASES = [(64, 64, 3, 32, 32, 0, True), (64, 64, 2, 65, 33, 0, True), (32, 64, 2, 65, 65, 0, False), (64, 128, 2, 32, 32, 0, False),
                   (128, 128, 2, 17, 9, 0, False), (128, 64, 1, 1, 1, 0, True), (64, 128, 3, 20, 12, 5, False), (64, 128, 1, 1, 1, 5, False),
                   (32, 128, 2, 55, 55, 5, False), (64, 64, 149, 32, 32, 0, True), (64, 128, 40, 37, 29, 5, False)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", TMA_STORE_CASES)
def test_tma_store_epilogue_equals_per_thread_stores(dev, case, dtype):
    from superresolution_aniso_mri_b200 import ops
    cin, cout, n, h, w, mode, affine = case
    if dtype == torch.bfloat16 and n > 3:
        pytest.skip("bf16 covered on the small cases")
    g = torch.Generator().manual_seed(cin + cout + h * 7 + w)
    x = torch.randn(n, h, w, cin, generator=g).to(dtype).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)).to(dev)
    cb = cout // 4 if mode == 5 else cout
    b = (torch.randn(cb, generator=g) * 0.1).to(dev)
    sc = (torch.rand(cb, generator=g) + 0.5).to(dev) if affine else None
    sh = (torch.randn(cb, generator=g) * 0.1).to(dev) if affine else None
    wp = ops.pack_conv3x3_weight(wt, dtype=dtype)
    shape = ops.conv_out_shape(n, h, w, cout, mode)
    got = torch.full(shape, 7.0, dtype=dtype, device=dev)
    ops.conv3x3(x, wp, b, act=1, scale=sc, shift=sh, out_mode=mode, out=got, algo=1)
    want = torch.full(shape, 7.0, dtype=dtype, device=dev)
    try:
        ops.set_tuning(10, 1)
        ops.conv3x3(x, wp, b, act=1, scale=sc, shift=sh, out_mode=mode, out=want, algo=1)
    finally:
        ops.set_tuning(10, 0)
    assert torch.equal(got, want)
    assert not torch.any(got == 7.0) or torch.equal(got == 7.0, want == 7.0)       # every element written (7.0 is not a plausible value)


def test_tma_store_head_patches_equal_per_thread_stores(dev):
    """dec.12 + head: the tile's 128 patches through one 8 KB TMA store against 128 threads x 4 STG.128."""
    from superresolution_aniso_mri_b200 import ops
    dt = torch.float16
    for n, h, w in ((2, 64, 64), (3, 20, 12), (1, 1, 1), (1, 17, 9), (40, 33, 31)):
        g = torch.Generator().manual_seed(n + h + w)
        x = torch.randn(n, h, w, 32, generator=g).to(dt).to(dev)
        wp = ops.pack_conv3x3_weight_up2fold((torch.randn(32, 32, 3, 3, generator=g) / 17).to(dev), dtype=dt)
        b = (torch.randn(32, generator=g) * 0.1).to(dev)
        w9c = (torch.randn(9, 32, generator=g) / 17).contiguous()
        got = ops.conv3x3_up2_head(x, wp, b, w9c)
        try:
            ops.set_tuning(10, 1)
            want = ops.conv3x3_up2_head(x, wp, b, w9c)
        finally:
            ops.set_tuning(10, 0)
        assert torch.equal(got, want)
