// extern "C" entry points of the training-step kernels (included by api.cu; declared in include/aesr_b200.h).
#pragma once

namespace {
inline int grid_for(size_t total, int block, int per_sm = 8) {
    size_t g = (total + block - 1) / block;
    const size_t cap = static_cast<size_t>(g_sm_count) * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}
}  // namespace

extern "C" {

int aesr_bn_finalize(const float* stats, float count, float count1, int passes, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                     float* mean_out, float* invstd_out, int C, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!stats || !gamma || !beta || !scale || !shift || !mean_out || !invstd_out || C <= 0 || count <= 0 || passes < 1 ||
        passes > 2 || (passes == 2 && count1 <= 0))
        return fail(AESR_ERR_INVALID, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, count, count1, passes, gamma, beta, running_mean, running_var, momentum, eps, scale, shift, mean_out,
        invstd_out, C);
    return check_launch("bn_finalize");
}

int aesr_bn_apply(const void* a, const float* scale, const float* shift, void* out, int N, int H, int W, int C, int mode,
                  int dtype, int split, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!a || !scale || !shift || !out || C % 8 != 0 || mode < 0 || mode > 2) return fail(AESR_ERR_INVALID, "bn_apply: bad arguments");
    if (split <= 0 || split > N) split = N;
    const int Ho = mode == BN_POOL ? H / 2 : mode == BN_UP ? 2 * H : H, Wo = mode == BN_POOL ? W / 2 : mode == BN_UP ? 2 * W : W;
    const size_t total = static_cast<size_t>(N) * Ho * Wo * (C / 8);
    if (total == 0) return AESR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        bn_apply_kernel<true><<<grid_for(total, 256, 16), 256, 0, s>>>(static_cast<const uint16_t*>(a), scale, shift, static_cast<uint16_t*>(out), N, H, W, C, mode, split);
    else
        bn_apply_kernel<false><<<grid_for(total, 256, 16), 256, 0, s>>>(static_cast<const uint16_t*>(a), scale, shift, static_cast<uint16_t*>(out), N, H, W, C, mode, split);
    return check_launch("bn_apply");
}

int aesr_bn_bwd(const void* dnext, const void* a, const float* mean, const float* invstd, const float* gamma,
                float* sums, float slope, void* g_out, float* dgamma, float* dbeta, float* dbias_conv, int N, int H, int W,
                int C, int mode, int dtype, int phase, float count, float count1, int split, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!dnext || !a || !mean || !invstd || !gamma || !sums || !g_out || !dgamma || !dbeta || C % 32 != 0 || mode < 0 || mode > 2)
        return fail(AESR_ERR_INVALID, "bn_bwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (split <= 0 || split > N) split = N;
    const int passes = split < N ? 2 : 1;                            // merged batch: images [0, split) and [split, N)
    const size_t npix = static_cast<size_t>(N) * H * W;
    const bool do_reduce = phase != 2, do_apply = phase != 1;
    if (do_reduce) CUDA_TRY(cudaMemsetAsync(sums, 0, static_cast<size_t>(passes) * BN_SUMS * C * sizeof(float), s));
    if (C > 512) return fail(AESR_ERR_INVALID, "bn_bwd: C=%d > 512", C);
    const int groups = C / 8;
    const int rows = 256 / groups;                                   // pixel rows per block (C = 32: 64 ... C = 512: 4)
    const size_t npix_pass = passes == 2 ? static_cast<size_t>(split > N - split ? split : N - split) * H * W : npix;
    int gx = static_cast<int>((npix_pass + static_cast<size_t>(rows) * 4 - 1) / (static_cast<size_t>(rows) * 4));
    const int gx_cap = g_sm_count * 8 / passes;
    if (gx > gx_cap) gx = gx_cap;
    if (gx < 1) gx = 1;
    const dim3 rgrid(gx, passes);
    const int block = rows * groups;
    const size_t red_smem = static_cast<size_t>(rows) * BN_SUMS * C * sizeof(float);      // <= 32 KB
    const size_t total = npix * groups;
    if (count <= 0.f) count = static_cast<float>(static_cast<size_t>(split) * H * W);       // global-batch counts in SyncBN mode
    if (count1 <= 0.f) count1 = static_cast<float>(static_cast<size_t>(N - split) * H * W);
    // phase 0: the apply kernel banks dgamma / dbeta from the (local) sums; phase 1 banks them right after the local
    // reduce; phase 2 (sums all-reduced by the caller) must not bank them again.
    // dbias_conv: C global atomics per block at the end (same-address REDs: ~2k per address, a few us, fire-and-forget); a grid
    // capped at 4 blocks per SM to save them made the kernel 50 % slower (34 vs 23 us per launch, profiles/r06_train_launches.md)
    const int apply_grid = grid_for(total, 256, tune(7) == 9999 ? 4 : 16);
    float* dg_apply = phase == 0 ? dgamma : nullptr;
    float* db_apply = phase == 0 ? dbeta : nullptr;
    if (dtype == AESR_DT_FP16) {
        if (do_reduce) bn_bwd_reduce_kernel<true><<<rgrid, block, red_smem, s>>>(static_cast<const uint16_t*>(dnext), static_cast<const uint16_t*>(a), mean, invstd, sums, N, H, W, C, mode, split, slope);
        if (do_apply) bn_bwd_apply_kernel<true><<<apply_grid, 256, 0, s>>>(static_cast<const uint16_t*>(dnext), static_cast<const uint16_t*>(a), mean, invstd, gamma, sums, count, count1, split, slope, static_cast<uint16_t*>(g_out), dg_apply, db_apply, dbias_conv, N, H, W, C, mode);
    } else {
        if (do_reduce) bn_bwd_reduce_kernel<false><<<rgrid, block, red_smem, s>>>(static_cast<const uint16_t*>(dnext), static_cast<const uint16_t*>(a), mean, invstd, sums, N, H, W, C, mode, split, slope);
        if (do_apply) bn_bwd_apply_kernel<false><<<apply_grid, 256, 0, s>>>(static_cast<const uint16_t*>(dnext), static_cast<const uint16_t*>(a), mean, invstd, gamma, sums, count, count1, split, slope, static_cast<uint16_t*>(g_out), dg_apply, db_apply, dbias_conv, N, H, W, C, mode);
    }
    if (phase == 1) bn_bwd_accum_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, dgamma, dbeta, C, passes);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("bn_bwd");
}

int aesr_mse(const float* a, const float* b, size_t n, float* loss_acc, float* d, float grad_scale, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!a || !b || !loss_acc || n == 0) return fail(AESR_ERR_INVALID, "mse: bad arguments");
    mse_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, 1.f / static_cast<float>(n), loss_acc, d, grad_scale);
    return check_launch("mse");
}

int aesr_head_bwd(const float* dout, const float* out, const void* a_in, const float* w9c, void* g_in, float* dw9c,
                  float* dbias, float* dbias_in, int N, int H, int W, int C, float slope, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!dout || !out || !a_in || !w9c || !g_in || !dw9c || !dbias || C != 32) return fail(AESR_ERR_INVALID, "head_bwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint16_t* a16 = static_cast<const uint16_t*>(a_in);
    const long long tiles = static_cast<long long>(N) * ((H + HB_ROWS - 1) / HB_ROWS) * ((W + HB_PIX - 1) / HB_PIX);
    long long blocks = tiles;
    const long long cap = 3LL * g_sm_count;                        // three resident blocks per SM; 321 global atomics each at the end
    if (blocks > cap) blocks = cap;
    if (dtype == AESR_DT_FP16)
        head_bwd_fused_kernel<true><<<static_cast<int>(blocks), 256, 0, s>>>(dout, out, a16, w9c, static_cast<uint16_t*>(g_in), dw9c, dbias, dbias_in, N, H, W, slope);
    else
        head_bwd_fused_kernel<false><<<static_cast<int>(blocks), 256, 0, s>>>(dout, out, a16, w9c, static_cast<uint16_t*>(g_in), dw9c, dbias, dbias_in, N, H, W, slope);
    return check_launch("head_bwd");
}

int aesr_e0_bwd(const void* g, const float* x, float* dw, float* db, int N, int H, int W, int C, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!g || !x || !dw || !db || C != 32) return fail(AESR_ERR_INVALID, "e0_bwd: bad arguments (C must be 32)");
    const size_t total = static_cast<size_t>(N) * (H + 2) * (W + 2);
    int gx = static_cast<int>((total + 64 * 8 - 1) / (64 * 8));            // >= 8 pixels per thread
    if (gx > g_sm_count * 8) gx = g_sm_count * 8;
    if (gx < 1) gx = 1;
    e0_bwd_kernel<<<gx, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint16_t*>(g), x, dw, db, N, H, W, C);
    return check_launch("e0_bwd");
}

int aesr_wgrad3x3(const void* g, const void* x, float* dW, float* dbias, int N, int H, int W, int Cin, int Cout, int dtype,
                  int algo, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!g || !x || !dW || Cin % 32 != 0 || Cout % 32 != 0) return fail(AESR_ERR_INVALID, "wgrad3x3: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t npix = static_cast<size_t>(N) * H * W;
    const bool tc_ok = Cin <= 128 && Cout <= 128 && (Cin == 32 || Cin % 64 == 0) && (Cout == 32 || Cout % 64 == 0);
    if (algo == 1 && !tc_ok) return fail(AESR_ERR_INVALID, "wgrad3x3: tensor-core path needs channels in {32,64,128}");
    if (algo != 2 && tc_ok) {
        WgradParams p{};
        p.N = N; p.H = H; p.W = W; p.Cg = Cout; p.Cx = Cin;
        p.tiles_x = (W + 7) / 8; p.tiles_y = (H + 15) / 16; p.num_tiles = N * p.tiles_x * p.tiles_y;
        p.taps_per_group = Cin <= 32 ? 9 : Cin <= 64 ? 5 : 3;
        p.num_groups = (9 + p.taps_per_group - 1) / p.taps_per_group;
        int per = g_sm_count / p.num_groups;
        if (tune(7) > 0) {
            if (per > tune(7) / p.num_groups) per = tune(7) / p.num_groups;
        } else {
            // Every CTA ends with Cout x Cin x taps fp32 atomics on dW, so the launch costs about
            //   tiles / C * t_tile + C * t_atomics   per tap group;
            // measured (profiles/r02q_wgrad_cta_sweep.txt): ~24.5 ns per 1000 atomics chip-wide, a tile's MMAs at the
            // tensor-pipe rate of its N (0.7 us for a folded Cin = 32 tile).  C = sqrt(tiles * t_tile / t_atomics).
            const double cyc = Cin <= 32 ? 56.0 * 3 * 8 : (Cin <= 64 ? 48.0 : 64.0) * p.taps_per_group * 8;
            const double t_tile = cyc / 1900.0;                                                    // us
            const double t_atom = 24.5e-6 * static_cast<double>(Cout) * Cin * p.taps_per_group;   // us per CTA
            const int best = static_cast<int>(sqrt(static_cast<double>(p.num_tiles) * t_tile / t_atom) + 0.5);
            if (per > best) per = best;
        }
        if (per > p.num_tiles) per = p.num_tiles;
        if (per < 1) per = 1;
        p.ctas_per_group = per;
        p.x_fp16 = (dtype == AESR_DT_FP16);
        p.fold_dx = tune(6) == 0;
        p.stages = WG_MAX_STAGES;
        while (p.stages > 2 && wg_smem_bytes(Cout, Cin, p.stages) > g_max_smem_optin) --p.stages;
        p.dW = dW;
        CUtensorMap tg, tx;
        rc = make_act_tmap(&tg, g, N, H, W, Cout, Cout < 64 ? 32 : 64, 8, 16);
        if (rc != AESR_OK) return rc;
        rc = make_act_tmap(&tx, x, N, H, W, Cin, Cin < 64 ? 32 : 64, 10, 18);
        if (rc != AESR_OK) return rc;
        static int configured = 0;
        rc = set_max_smem(wgrad3x3_tc_kernel, &configured);
        if (rc != AESR_OK) return rc;
        wgrad3x3_tc_kernel<<<per * p.num_groups, WG_THREADS, wg_smem_bytes(Cout, Cin, p.stages), s>>>(tg, tx, p);
        rc = check_launch("wgrad3x3_tc");
        if (rc != AESR_OK) return rc;
        if (dbias) {
            const int groups = Cout / 8, rows = 256 / groups;
            int gx = static_cast<int>((npix + static_cast<size_t>(rows) * 4 - 1) / (static_cast<size_t>(rows) * 4));
            if (gx > g_sm_count * 8) gx = g_sm_count * 8;
            if (gx < 1) gx = 1;
            colsum_bf16_kernel<<<gx, rows * groups, static_cast<size_t>(rows) * Cout * sizeof(float), s>>>(
                static_cast<const uint16_t*>(g), dbias, npix, Cout);
            return check_launch("colsum_bf16");
        }
        return AESR_OK;
    }
    constexpr int CO_PER = 4;                       // 8 warps x 4 = 32 output channels per block
    const int gy = Cin / 32, gz = Cout / 32;
    int gx = g_sm_count * 4 / (gy * gz);
    if (gx < 1) gx = 1;
    if (static_cast<size_t>(gx) > npix) gx = static_cast<int>(npix);
    dim3 grid(gx, gy, gz);
    const uint16_t* g16 = static_cast<const uint16_t*>(g);
    const uint16_t* x16 = static_cast<const uint16_t*>(x);
    if (dtype == AESR_DT_FP16) wgrad3x3_kernel<true, CO_PER><<<grid, 256, 0, s>>>(g16, x16, dW, dbias, N, H, W, Cin, Cout);
    else wgrad3x3_kernel<false, CO_PER><<<grid, 256, 0, s>>>(g16, x16, dW, dbias, N, H, W, Cin, Cout);
    return check_launch("wgrad3x3");
}

int aesr_mix_bwd(const void* g_dec, const void* g_mix, const float* wa, const float* wb, void* g_z, int B,
                 size_t per_image, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!g_dec || !g_mix || !wa || !wb || !g_z || B <= 0) return fail(AESR_ERR_INVALID, "mix_bwd: bad arguments");
    mix_bwd_kernel<<<grid_for(2 * B * per_image, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(g_dec), static_cast<const uint16_t*>(g_mix), wa, wb, static_cast<uint16_t*>(g_z), B, per_image);
    return check_launch("mix_bwd");
}

int aesr_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!p || !g || !m || !v || n == 0 || step < 1) return fail(AESR_ERR_INVALID, "adam_step: bad arguments");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
    adam_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), nullptr, nullptr);
    return check_launch("adam_step");
}

int aesr_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, const int* step_dev, const float* lr_dev, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!p || !g || !m || !v || n == 0 || !step_dev) return fail(AESR_ERR_INVALID, "adam_step_dev: bad arguments");
    adam_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, step_dev, lr_dev);
    return check_launch("adam_step_dev");
}

int aesr_vgg_conv1_fwd(const float* img, const float* w, const float* b, void* out, int N, int H, int W,
                       const float* shift3, const float* scale3, int normalize, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!img || !w || !b || !out || !shift3 || !scale3) return fail(AESR_ERR_INVALID, "vgg_conv1_fwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (H >= 2 && W >= 2) {          // four pixels per thread
        const size_t totq = static_cast<size_t>(N) * H * ((W + 3) / 4) * 8;
        if (dtype == AESR_DT_FP16)
            vgg_conv1_fwd_quad_kernel<true><<<grid_for(totq, 256, 3), 256, 0, s>>>(img, w, b, static_cast<uint16_t*>(out), N, H, W, shift3[0], shift3[1], shift3[2], scale3[0], scale3[1], scale3[2], normalize);
        else
            vgg_conv1_fwd_quad_kernel<false><<<grid_for(totq, 256, 3), 256, 0, s>>>(img, w, b, static_cast<uint16_t*>(out), N, H, W, shift3[0], shift3[1], shift3[2], scale3[0], scale3[1], scale3[2], normalize);
        return check_launch("vgg_conv1_fwd_quad");
    }
    const size_t total = static_cast<size_t>(N) * H * W * 8;
    if (dtype == AESR_DT_FP16)
        vgg_conv1_fwd_kernel<true><<<grid_for(total, 256, 8), 256, 0, s>>>(img, w, b, static_cast<uint16_t*>(out), N, H, W, shift3[0], shift3[1], shift3[2], scale3[0], scale3[1], scale3[2], normalize);
    else
        vgg_conv1_fwd_kernel<false><<<grid_for(total, 256, 8), 256, 0, s>>>(img, w, b, static_cast<uint16_t*>(out), N, H, W, shift3[0], shift3[1], shift3[2], scale3[0], scale3[1], scale3[2], normalize);
    return check_launch("vgg_conv1_fwd");
}

int aesr_vgg_conv1_bwd(const void* g, const float* w, float* dimg, int N, int H, int W, const float* scale3,
                       int normalize, float out_scale, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!g || !w || !dimg || !scale3) return fail(AESR_ERR_INVALID, "vgg_conv1_bwd: bad arguments");
    const size_t totq = static_cast<size_t>(N) * H * ((W + 3) / 4);
    vgg_conv1_bwd_quad_kernel<<<grid_for(totq * 8, 256, 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(g), w, dimg, N, H, W, scale3[0], scale3[1], scale3[2], normalize, out_scale);
    return check_launch("vgg_conv1_bwd");
}

int aesr_maxpool_bwd(const void* a, const void* d_pooled, const void* g_tap, void* g_out, int N, int H, int W, int C,
                     int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!a || !g_out || (!d_pooled && !g_tap) || C % 8 != 0) return fail(AESR_ERR_INVALID, "maxpool_bwd: bad arguments (C % 8 == 0)");
    const size_t total = static_cast<size_t>(N) * H * W * (C / 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        maxpool_bwd_kernel<true><<<grid_for(total, 256, 16), 256, 0, s>>>(static_cast<const uint16_t*>(a), static_cast<const uint16_t*>(d_pooled), static_cast<const uint16_t*>(g_tap), static_cast<uint16_t*>(g_out), N, H, W, C);
    else
        maxpool_bwd_kernel<false><<<grid_for(total, 256, 16), 256, 0, s>>>(static_cast<const uint16_t*>(a), static_cast<const uint16_t*>(d_pooled), static_cast<const uint16_t*>(g_tap), static_cast<uint16_t*>(g_out), N, H, W, C);
    return check_launch("maxpool_bwd");
}

int aesr_lpips_head(const void* o0, const void* o1, const float* lin, float* val, const float* upstream, void* g1, int N,
                    int HW, int C, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!o0 || !o1 || !lin || (!val && !g1) || (g1 && !upstream)) return fail(AESR_ERR_INVALID, "lpips_head: bad arguments");
    if (C % 8 != 0 || C > 512 || ((C / 8) & (C / 8 - 1)) != 0)
        return fail(AESR_ERR_INVALID, "lpips_head: C=%d must be 8 * 2^k <= 512 (the VGG16 taps: 64, 128, 256, 512)", C);
    const size_t total = static_cast<size_t>(N) * HW;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint16_t* a = static_cast<const uint16_t*>(o0);
    const uint16_t* b = static_cast<const uint16_t*>(o1);
    const int lanes_per_px = (C / 8) < 32 ? (C / 8) : 32;
    const int grid = grid_for(total * lanes_per_px, 256, 8);
    if (g1) {       // backward (+ the forward value in the same pass when asked for)
        if (C <= 256) {
            if (dtype == AESR_DT_FP16) lpips_head_bwd_kernel<true, 1><<<grid, 256, 0, s>>>(a, b, lin, upstream, static_cast<uint16_t*>(g1), val, N, HW, C);
            else lpips_head_bwd_kernel<false, 1><<<grid, 256, 0, s>>>(a, b, lin, upstream, static_cast<uint16_t*>(g1), val, N, HW, C);
        } else {
            if (dtype == AESR_DT_FP16) lpips_head_bwd_kernel<true, 2><<<grid, 256, 0, s>>>(a, b, lin, upstream, static_cast<uint16_t*>(g1), val, N, HW, C);
            else lpips_head_bwd_kernel<false, 2><<<grid, 256, 0, s>>>(a, b, lin, upstream, static_cast<uint16_t*>(g1), val, N, HW, C);
        }
        return check_launch("lpips_head_bwd");
    }
    if (dtype == AESR_DT_FP16) lpips_head_fwd_kernel<true><<<grid, 256, 0, s>>>(a, b, lin, val, N, HW, C);
    else lpips_head_fwd_kernel<false><<<grid, 256, 0, s>>>(a, b, lin, val, N, HW, C);
    return check_launch("lpips_head_fwd");
}

// ------------------------------------------------------------------------------------------------------------------
// evaluation / data-path utilities
// ------------------------------------------------------------------------------------------------------------------
int aesr_ssim_psnr(const float* a, const float* b, int Z, int H, int W, int win, double data_range, double* ssim_sum,
                   double* sqerr_sum, unsigned int* min_key, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!a || !b || !ssim_sum || !sqerr_sum || !min_key || Z <= 0 || Z > 65535 || H < win || W < win)
        return fail(AESR_ERR_INVALID, "ssim_psnr: bad arguments (Z<=65535, H,W >= win)");
    if (win != 7 && win != 5) return fail(AESR_ERR_INVALID, "ssim_psnr: win must be 7 (default) or 5 (evaluate/metrics.py:151)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(ssim_sum, 0, Z * sizeof(double), s));
    CUDA_TRY(cudaMemsetAsync(sqerr_sum, 0, Z * sizeof(double), s));
    CUDA_TRY(cudaMemsetAsync(min_key, 0xFF, Z * sizeof(unsigned int), s));
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    dim3 grid((W + 15) / 16, (H + 15) / 16, Z);
    if (win == 7) ssim_psnr_kernel<7><<<grid, 256, 0, s>>>(a, b, H, W, c1, c2, ssim_sum, sqerr_sum, min_key);
    else ssim_psnr_kernel<5><<<grid, 256, 0, s>>>(a, b, H, W, c1, c2, ssim_sum, sqerr_sum, min_key);
    return check_launch("ssim_psnr");
}

size_t aesr_vif_workspace_bytes(int Z, int H, int W) {
    if (Z <= 0 || H <= 0 || W <= 0) return 0;
    return 12 * static_cast<size_t>(Z) * H * W + 256;       // 12 uint8 planes (see aesr_vif_mscale)
}

int aesr_vif_quantize_u8(const float* x, void* out_u8, size_t n, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!x || !out_u8 || n == 0) return fail(AESR_ERR_INVALID, "vif_quantize_u8: bad arguments");
    size_t grid = (n + 255) / 256;
    if (grid > static_cast<size_t>(g_sm_count) * 16) grid = static_cast<size_t>(g_sm_count) * 16;
    vif_quantize_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<uint8_t*>(out_u8), n);
    return check_launch("vif_quantize");
}

int aesr_vif_mscale(const void* ref_u8, const void* dist_u8, int Z, int H, int W, const double* weights,
                    const int* radii_host, double sigma_nsq, void* workspace, size_t workspace_bytes, double* num_den,
                    void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!ref_u8 || !dist_u8 || !weights || !radii_host || !workspace || !num_den || Z <= 0 || Z > 65535 || H <= 0 || W <= 0)
        return fail(AESR_ERR_INVALID, "vif_mscale: bad arguments (Z <= 65535)");
    if (workspace_bytes < aesr_vif_workspace_bytes(Z, H, W)) return fail(AESR_ERR_WORKSPACE, "vif_mscale: workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t plane = static_cast<size_t>(Z) * H * W;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint8_t *cur_r = ws, *cur_d = ws + plane, *tmp = ws + 2 * plane, *full_r = ws + 3 * plane, *full_d = ws + 4 * plane,
            *mu1 = ws + 5 * plane, *mu2 = ws + 6 * plane, *prr = ws + 7 * plane, *pdd = ws + 8 * plane,
            *prd = ws + 9 * plane, *g3 = ws + 10 * plane;      // g3: the three filtered products reuse [10, 12) + tmp
    CUDA_TRY(cudaMemsetAsync(num_den, 0, 2 * Z * sizeof(double), s));
    CUDA_TRY(cudaMemcpyAsync(cur_r, ref_u8, plane, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(cur_d, dist_u8, plane, cudaMemcpyDeviceToDevice, s));
    auto grid_for = [&](size_t n) {
        size_t g = (n + 255) / 256;
        const size_t cap = static_cast<size_t>(g_sm_count) * 16;
        return static_cast<int>(g < cap ? (g ? g : 1) : cap);
    };
    // gaussian_filter of `planes` images of h x w: axis 0 then axis 1, uint8 in between (scipy.ndimage.gaussian_filter)
    auto gauss = [&](const uint8_t* in, uint8_t* scratch, uint8_t* out, int planes, int h, int w, const double* wt, int lw) {
        const size_t n = static_cast<size_t>(planes) * h * w;
        vif_gauss1d_u8_kernel<<<grid_for(n), 256, 0, s>>>(in, scratch, planes, h, w, 0, wt, lw);
        vif_gauss1d_u8_kernel<<<grid_for(n), 256, 0, s>>>(scratch, out, planes, h, w, 1, wt, lw);
    };
    int h = H, w = W;
    const double* wt = weights;
    for (int scale = 1; scale <= 4; ++scale) {
        const int lw = radii_host[scale - 1];
        if (lw < 0 || lw > 64) return fail(AESR_ERR_INVALID, "vif_mscale: filter radius %d", lw);
        if (scale > 1) {
            gauss(cur_r, tmp, full_r, Z, h, w, wt, lw);
            gauss(cur_d, tmp, full_d, Z, h, w, wt, lw);
            const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
            const size_t n2 = static_cast<size_t>(Z) * h2 * w2;
            vif_subsample_kernel<<<grid_for(n2), 256, 0, s>>>(full_r, cur_r, Z, h, w);
            vif_subsample_kernel<<<grid_for(n2), 256, 0, s>>>(full_d, cur_d, Z, h, w);
            h = h2;
            w = w2;
        }
        const size_t n = static_cast<size_t>(Z) * h * w;
        gauss(cur_r, tmp, mu1, Z, h, w, wt, lw);
        gauss(cur_d, tmp, mu2, Z, h, w, wt, lw);
        vif_products_kernel<<<grid_for(n), 256, 0, s>>>(cur_r, cur_d, prr, pdd, prd, n);
        gauss(prr, tmp, full_r, Z, h, w, wt, lw);           // full_r / full_d / g3 are free here: filtered products
        gauss(pdd, tmp, full_d, Z, h, w, wt, lw);
        gauss(prd, tmp, g3, Z, h, w, wt, lw);
        vif_accumulate_kernel<<<Z, 256, 0, s>>>(mu1, mu2, full_r, full_d, g3, h * w, sigma_nsq, num_den);
        wt += 2 * lw + 1;
    }
    return check_launch("vif_mscale");
}

size_t aesr_percentile_workspace_bytes(void) { return 256 + 4 * 2048 * sizeof(unsigned int) + 64; }

int aesr_percentile_normalize(const float* x, float* out, size_t n, double q_lo, double q_hi, void* workspace,
                              size_t workspace_bytes, double* lo_hi_out, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!x || !workspace || n == 0 || q_lo < 0 || q_hi > 100 || q_lo > q_hi) return fail(AESR_ERR_INVALID, "percentile_normalize: bad arguments");
    if (workspace_bytes < aesr_percentile_workspace_bytes()) return fail(AESR_ERR_WORKSPACE, "percentile_normalize: workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    SelectState* st = reinterpret_cast<SelectState*>(ws);
    unsigned int* hist = reinterpret_cast<unsigned int*>(ws + 256);
    double* lo_hi = reinterpret_cast<double*>(ws + 256 + 4 * 2048 * sizeof(unsigned int));
    // numpy 'linear': virtual index q/100 * (n-1), neighbours floor / floor+1 (clamped)
    SelectState h{};
    const double vi[2] = {q_lo / 100.0 * static_cast<double>(n - 1), q_hi / 100.0 * static_cast<double>(n - 1)};
    double frac[2];
    for (int j = 0; j < 2; ++j) {
        const double fl = floor(vi[j]);
        frac[j] = vi[j] - fl;
        const uint64_t k = static_cast<uint64_t>(fl);
        h.rank[2 * j] = k;
        h.rank[2 * j + 1] = (k + 1 < n) ? k + 1 : n - 1;
    }
    CUDA_TRY(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(hist, 0, 4 * 2048 * sizeof(unsigned int), s));
    const int grid = grid_for(n, 256, 8);
    select_hist_kernel<21, 11><<<grid, 256, 0, s>>>(x, n, st, 4, hist);
    select_narrow_kernel<21, 11><<<1, 32, 0, s>>>(st, 4, hist, 0);
    select_hist_kernel<10, 11><<<grid, 256, 0, s>>>(x, n, st, 4, hist);
    select_narrow_kernel<10, 11><<<1, 32, 0, s>>>(st, 4, hist, 0);
    select_hist_kernel<0, 10><<<grid, 256, 0, s>>>(x, n, st, 4, hist);
    select_narrow_kernel<0, 10><<<1, 32, 0, s>>>(st, 4, hist, 1);
    percentile_finish_kernel<<<1, 1, 0, s>>>(st, frac[0], frac[1], lo_hi);
    g_launches.fetch_add(6, std::memory_order_relaxed);
    if (lo_hi_out) CUDA_TRY(cudaMemcpyAsync(lo_hi_out, lo_hi, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (out) normalize_apply_kernel<<<grid, 256, 0, s>>>(x, out, n, lo_hi);
    return check_launch("percentile_normalize");
}

int aesr_pad_crop_gather(const float* in, float* out, const int* top, const int* left, int B, int C, int Hin, int Win,
                         int Hout, int Wout, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!in || !out || !top || !left || B <= 0 || B > 65535 || C <= 0 || C > 65535) return fail(AESR_ERR_INVALID, "pad_crop_gather: bad arguments");
    int gx = (Hout * Wout + 255) / 256;
    if (gx > 64) gx = 64;
    pad_crop_gather_kernel<<<dim3(gx, C, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, top, left, C, Hin, Win, Hout, Wout);
    return check_launch("pad_crop_gather");
}

int aesr_augment_gather(const float* in, float* out, const int* top, const int* left, const int* rot_k, const float* gain,
                        const float* cutoff, unsigned chan_mask, int B, int C, int Hin, int Win, int P, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!in || !out || !top || !left || B <= 0 || B > 65535 || C <= 0 || C > 65535 || P <= 0 || Hin <= 0 || Win <= 0)
        return fail(AESR_ERR_INVALID, "augment_gather: bad arguments");
    if ((gain == nullptr) != (cutoff == nullptr)) return fail(AESR_ERR_INVALID, "augment_gather: gain and cutoff go together");
    int gx = (P * P + 255) / 256;
    if (gx > 64) gx = 64;
    augment_gather_kernel<<<dim3(gx, C, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, top, left, rot_k, gain, cutoff,
                                                                                       chan_mask, C, Hin, Win, P);
    return check_launch("augment_gather");
}

int aesr_gauss1d_axis0(const float* in, float* out, const double* taps, int lw, int Z, size_t HW, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!in || !out || !taps || lw < 0 || Z <= 0 || Z > 65535 || HW == 0 || in == out)
        return fail(AESR_ERR_INVALID, "gauss1d_axis0: bad arguments (out of place, Z <= 65535)");
    size_t gx = (HW + 255) / 256;
    if (gx > 4096) gx = 4096;
    gauss1d_axis0_kernel<<<dim3(static_cast<unsigned>(gx), Z), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, taps, lw, Z, HW);
    return check_launch("gauss1d_axis0");
}

}  // extern "C"
