/*
 * aesr_b200 -- C-ABI of the B200-native (sm_100a) kernels behind the ae_combined slice-synthesis hot path.
 *
 * The reference (qurAI-amsterdam/SuperResolution_aniso_MRI) is pure Python/PyTorch: it has no FFI of its own, every
 * "operator" below replaces a torch.nn / ATen / cuDNN library call made from the reference file:line cited beside it.
 * The Python host layer (superresolution_aniso_mri_b200/_lib.py, ctypes) is the only caller; INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (torch allocations);
 *     no ownership transfer, no hidden allocation, workspace (if any) is passed in.
 *   - stream-ordered: work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises.
 *   - return value: 0 on success, negative error code otherwise; aesr_last_error() gives the message.
 *   - thread-compatible: no global mutable state besides the once-initialised driver entry point.
 *   - activations are NHWC 16-bit internally (dtype: AESR_DT_BF16 or AESR_DT_FP16, fp32 accumulation);
 *     the public tensors (images, latents) are NCHW fp32.
 *   - no CPU fallback: aesr_init fails with AESR_ERR_ARCH unless the device is compute capability 10.x.
 */
#ifndef AESR_B200_H
#define AESR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AESR_OK 0
#define AESR_ERR_INVALID (-1)
#define AESR_ERR_CUDA (-2)
#define AESR_ERR_ARCH (-3)
#define AESR_ERR_WORKSPACE (-4)

/* 16-bit storage format of internal activations / packed filters */
#define AESR_DT_BF16 0
#define AESR_DT_FP16 1

/* activation fused into the conv epilogue */
#define AESR_ACT_NONE 0
#define AESR_ACT_LEAKY 1 /* nn.LeakyReLU(slope)   networks/acai_vanilla.py:17,55-56 */
#define AESR_ACT_RELU 2  /* nn.ReLU               lpips/pretrained_networks.py:107-116 (VGG16 features) */

/* output stage fused into the conv epilogue */
#define AESR_OUT_SAME 0          /* NHWC 16-bit [N,H,W,Cout] */
#define AESR_OUT_AVGPOOL2 1      /* nn.AvgPool2d(2) (floor)            networks/acai_vanilla.py:59  */
#define AESR_OUT_UP2 2           /* nn.Upsample(scale_factor=2) nearest networks/acai_vanilla.py:92  */
#define AESR_OUT_NCHW_F32 3      /* public latent: NCHW fp32 (+ optional NHWC 16-bit copy in out2) */
#define AESR_OUT_SAME_MAXPOOL2 4 /* full-res tap + nn.MaxPool2d(2) in out2  (VGG16 slices) */
#define AESR_OUT_SHUFFLE2 5      /* nn.Upsample(2) + the NEXT nn.Conv2d folded into one low-res conv with 4*C phase
                                    channels (weights from aesr_pack_conv3x3_weight_up2fold) + depth-to-space store:
                                    networks/acai_vanilla.py:92 followed by :87 / :96.  out NHWC 16-bit [N,2H,2W,Cout/4] */
#define AESR_OUT_SAME_F32 7      /* NHWC fp32 [N,H,W,Cout], accumulators not rounded: the decoder's first conv
                                    (networks/acai_vanilla.py:87) applied to the latents BEFORE the interpolation
                                    (generate_hr_volumes.py:88; a conv is linear) -- see aesr_lerp_pairs_act */

#define AESR_MUL_NONE 0
#define AESR_MUL_LEAKY_GRAD 1 /* dgrad epilogue: multiply by LeakyReLU'(mul_src) */
#define AESR_MUL_RELU_GRAD 2  /* dgrad epilogue: multiply by ReLU'(mul_src) */

/* conv kernel selection (aesr_conv3x3_fwd `algo`): AUTO picks HALO whenever the filter bank fits in shared memory */
#define AESR_ALGO_AUTO 0
#define AESR_ALGO_HALO 1   /* one TMA halo load per tile, nine row-shifted UMMA descriptors, resident filters */
#define AESR_ALGO_STREAM 2 /* one (A, B) TMA pair per filter tap */

/* Select device, verify compute capability 10.x, resolve cuTensorMapEncodeTiled.  Idempotent. */
int aesr_init(int device);
const char* aesr_last_error(void);
/* Profiling / tuning knobs (no reference counterpart; initial values come from the environment variables AESR_*):
 * key 0 = stage-isolation mask of the conv kernel (AESR_CONV_DEBUG), 1 = forced M-tiles per super-tile, 2 = forced TMEM
 * buffers, 3 = activation-ring stage cap, 4 = decoder head on warp-level mma.sync (AESR_HEAD_MMA), 5 = encoder stem
 * variant (1 CUDA cores, 2 one tf32 term; AESR_STEM_CUDA_CORES), 6 = weight gradient without the dx fold
 * (AESR_WGRAD_NO_FOLD), 7 = cap on the CTAs of a weight-gradient launch (AESR_WGRAD_CTAS), 8 = no split-K (AESR_NO_SPLITK),
 * 9 = 32 -> 32 layers on the horizontal-tap-fold kernel (AESR_FOLD; opt-in experiment, slower), 10 = per-thread global stores
 * instead of the staged TMA stores of the conv epilogues (AESR_NO_TMA_STORE, A/B measurements), 11 = head_gather with
 * thread-staged windows instead of TMA loads (AESR_HEAD_GATHER_NO_TMA).  value 0 = automatic. */
int aesr_set_tuning(int key, int value);
int aesr_sm_count(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t aesr_launch_count(void);

/* fp32 [Cout,Cin,3,3] (nn.Conv2d.weight layout) -> 16-bit [9][Cout][Cin] (forward) or, with transpose_flip != 0,
 * [9 (taps mirrored)][Cin][Cout] (the data-gradient conv).  Replaces cuDNN's internal filter transforms. */
int aesr_pack_conv3x3_weight(const float* w, void* packed, int Cout, int Cin, int transpose_flip, int dtype,
                             void* stream);

/* aesr_pack_conv3x3_weight for many filters in one launch: job j = 6 int64 in DEVICE memory {offset of the fp32 filter in
 * src_base (floats), offset of the packed filter in dst_base (16-bit elements), Cout, Cin, transpose_flip, 0};
 * max_elems = the largest 9*Cout*Cin.  The training step re-packs all filters after every optimizer update. */
int aesr_pack_conv3x3_weight_batch(const float* src_base, void* dst_base, const long long* jobs_dev, int n_jobs,
                                   int max_elems, int dtype, void* stream);

/* 3x3 / pad 1 / stride 1 convolution, implicit GEMM on tcgen05 tensor cores with TMA operand loads.
 * Replaces nn.Conv2d(k,k,3,padding=1) + activation (+ eval BatchNorm2d affine) (+ AvgPool2d / Upsample) of
 * networks/acai_vanilla.py:55-59,68-70,87-92,96 and the VGG16 convs of lpips/pretrained_networks.py:107-116.
 *   x        NHWC 16-bit [N,H,W,Cin],  Cin  in {32,64,128,256,512}
 *   w_packed 16-bit [9][Cout][Cin],    Cout multiple of 32
 *   y = act(conv(x) + bias) [* act'(mul_src)] ; stats += {sum y, sum y^2} per channel ; y = y*scale + shift ; out stage
 *   bias/scale/shift/out2/mul_src/stats may be NULL.  stats: fp32 [2*Cout], or [2][2*Cout] with 0 < stats_split < N:
 *   images >= stats_split (the second pass of a merged batch, see "Merged batches" below) accumulate into the second
 *   block.  stats_split < 0: stats is [Cout] and only the sums are accumulated -- a data-gradient launch whose output is the
 *   gradient of the previous layer's pre-activation delivers that layer's bias gradient on the way. */
int aesr_conv3x3_fwd(const void* x, const void* w_packed, const float* bias, const float* scale, const float* shift,
                     void* out, void* out2, const void* mul_src, float* stats, int N, int H, int W, int Cin, int Cout,
                     int act, float slope, int out_mode, int mul_mode, int dtype, int algo, int stats_split, void* stream);

/* Folded filter bank for AESR_OUT_SHUFFLE2: fp32 [Cout,Cin,3,3] -> 16-bit [9 low-res taps][4 phases][Cout][Cin],
 * phase 2a+b = hi-res pixel (2y+a, 2x+b); each entry is the sum of the original taps that read the same low-res
 * pixel of the nearest-upsampled input (exact in real arithmetic; sums formed in fp32, rounded once). */
int aesr_pack_conv3x3_weight_up2fold(const float* w, void* packed, int Cout, int Cin, int dtype, void* stream);

/* Decoder tail, first half: nn.Upsample(2) -> nn.Conv2d(32,32,3,padding=1) + LeakyReLU -> nn.Conv2d(32,1,3,padding=1)
 * (networks/acai_vanilla.py:92,96-98) in one tensor-core kernel.  x NHWC 16-bit [N,H,W,Cin] is the LOW-res input of the
 * upsample, w_folded the [9][128][Cin] bank of aesr_pack_conv3x3_weight_up2fold, head_w9c_host fp32 [9][32] the head
 * filter in HOST memory (the one host pointer of this ABI: it is passed to the kernel by value so that the epilogue reads
 * it from the constant bank; copied before the call returns).
 * head_w16 (device, may be NULL) = the same filter as 16-bit [16 taps (rows 9..15 zero)][32 channels]: when given, the
 * head conv itself runs on the tensor cores (activations written back to tensor memory as a 16-bit A operand, one
 * 128x16x32 GEMM per phase); when NULL it runs on the CUDA cores in fp32 from head_w9c_host.
 * The 32-channel hi-res activation never leaves the SM; per low-res pixel the kernel writes the 4x4 patch (origin
 * (2y-1, 2x-1)) of head-conv partial sums of its 2x2 hi-res block: partial fp32 [N,H,W,16]. */
int aesr_conv3x3_up2_head_fwd(const void* x, const void* w_folded, const float* bias, const float* head_w9c_host,
                              const void* head_w16, float* partial, int N, int H, int W, int Cin, int act, float slope,
                              int dtype, int algo, void* stream);

/* Decoder tail, second half: out(Y,X) = sigmoid(bias + the (up to) four overlapping patch entries), clamp(0,1)
 * (networks/acai_vanilla.py:98, generate_hr_volumes.py:67).  partial fp32 [N,h,w,16]; image n ([2h,2w] fp32) is written at
 * out + (out_index ? out_index[n] : n) * out_image_stride. */
int aesr_head_gather(const float* partial, const float* bias, float* out, const int* out_index, int N, int h, int w,
                     size_t out_image_stride, int apply_sigmoid, void* stream);

/* Encoder stem: enc.0 nn.Conv2d(1,32,1,padding=1) and enc.1 nn.Conv2d(32,32,3,padding=1) + LeakyReLU
 * (networks/acai_vanilla.py:51,55-56) composed into one 3x3 conv on the single input channel (both are linear).
 * aesr_stem_fold: weff[tap][co] = sum_ci w1[co][ci][tap]*w0[ci], beff likewise with b0 (once per parameter version).
 * aesr_stem_fwd: weff_beff_b1_host = fp32 [9*32 weff | 9*32 beff | 32 b1] in HOST memory (kernel parameters: the filter is
 * read from the constant bank); x fp32 [N,1,H,W] -> NHWC 16-bit [N,H+2,W+2,32]; border taps that fall off the (H+2)x(W+2) grid drop
 * their bias term, ring pixels see x = 0, exactly as the two zero paddings of the reference do. */
int aesr_stem_fold(const float* w0, const float* b0, const float* w1, float* weff, float* beff, int C, void* stream);
int aesr_stem_fwd(const float* x, const float* weff_beff_b1_host, void* out, int N, int H, int W, int C, float slope,
                  int dtype, void* stream);

/* enc.0: nn.Conv2d(1, C, 1, padding=1) (networks/acai_vanilla.py:51).  x fp32 [N,1,H,W] -> NHWC 16-bit [N,H+2,W+2,C]. */
int aesr_e0_fwd(const float* x, const float* w, const float* b, void* out, int N, int H, int W, int C, int dtype,
                void* stream);

/* dec.14/15: nn.Conv2d(C, 1, 3, padding=1) + nn.Sigmoid (networks/acai_vanilla.py:98) + clamp(0,1)
 * (generate_hr_volumes.py:67).  in NHWC 16-bit [N,H,W,C] (C = 32), w9c fp32 [9][C]; image n is written at
 * out + (out_index ? out_index[n] : n) * out_image_stride (floats), so synthesized slices land directly at their
 * position i*(A+1)+1+k inside the HR volume (generate_hr_volumes.py:58-60) without a concat pass. */
int aesr_head_fwd(const void* in, const float* w9c, const float* bias, float* out, const int* out_index, int N, int H, int W,
                  int C, size_t out_image_stride, int apply_sigmoid, int dtype, void* stream);

/* Latent interpolation (generate_hr_volumes.py:88, kwatsch/cardiac/trainer_ae.py:173,
 * kwatsch/brain/trainer_ae.py:265-266) fused with the NCHW fp32 -> NHWC 16-bit layout change:
 *   out[m] = wa[m] * z[ia[m]] + wb[m] * z[ib[m]]  (three separately rounded fp32 ops; ib[m] < 0: plain copy)
 *   z fp32 [*,C,HW]; ia/ib int32 [M]; wa/wb fp32 [M]; out_nhwc 16-bit [M,HW,C]; out_nchw fp32 [M,C,HW] or NULL. */
int aesr_lerp_latents(const float* z, const int* ia, const int* ib, const float* wa, const float* wb, void* out_nhwc,
                      float* out_nchw, int M, int C, int HW, int dtype, void* stream);

/* All K alpha steps of P slice pairs (generate_hr_volumes.py:49-53 loops alphas around the whole encode/lerp/decode;
 * here the two fp32 latents of a pair are read once): out[p*K+k] = wa[k] * z[pa[p]] + wb[k] * z[pb[p]], NHWC 16-bit. */
int aesr_lerp_pairs(const float* z, const int* pa, const int* pb, const float* wa, const float* wb, void* out_nhwc,
                    int P, int K, int C, int HW, int dtype, void* stream);

/* Interpolation moved BEHIND the decoder's first conv (generate_hr_volumes.py:88 + networks/acai_vanilla.py:87-88):
 *   dec.0(wa*z1 + wb*z2) = wa*conv(z1) + wb*conv(z2) + bias, so the conv runs once per low-resolution slice.
 *   pre  fp32 NHWC [*,HW,C] = aesr_conv3x3_fwd(..., bias NULL, AESR_ACT_NONE, AESR_OUT_SAME_F32) of the encoder latents
 *   out[p*K+k] = LeakyReLU_slope(wa[k] * pre[pa[p]] + wb[k] * pre[pb[p]] + bias[c])   NHWC 16-bit [P*K,HW,C]
 * (slope = 1: no activation; bias may be NULL).  P <= 65535 per call, C % 8 == 0. */
int aesr_lerp_pairs_act(const float* pre, const int* pa, const int* pb, const float* wa, const float* wb,
                        const float* bias, void* out_nhwc, int P, int K, int C, int HW, float slope, int dtype,
                        void* stream);

/* Kept (non-synthesized) slices of the HR volume: dst[out_index[n]] = clamp(src[n], 0, 1)
 * (generate_hr_volumes.py:44,58-67: `recon_volume = images`, the torch.cat chain, the final torch.clamp).
 * src fp32 [N,HW], dst fp32 [*,HW], out_index int32 [N] or NULL (identity).  N <= 65535 per call. */
int aesr_place_slices(const float* src, float* dst, const int* out_index, int N, int HW, int do_clamp, void* stream);

/* Host pipeline utility: `outer` strided 2-D copies (cudaMemcpy2DAsync: `rows` rows of `width` bytes, row pitches dpitch /
 * spitch, consecutive 2-D blocks dst_outer_stride / src_outer_stride bytes apart) between PINNED host memory and the device,
 * stream-ordered.  to_host != 0: device -> host.  Used to download only the SYNTHESIZED slices of the HR volumes
 * (generate_hr_volumes.py:58-66 interleaves them with the kept input slices, which the host already holds): one block per
 * volume, one row per slice pair = A consecutive slices.  No reference counterpart (the reference copies whole tensors). */
int aesr_copy_rows_async(void* dst, size_t dst_outer_stride, size_t dpitch, const void* src, size_t src_outer_stride,
                         size_t spitch, size_t width, size_t rows, size_t outer, int to_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step (kwatsch/cardiac/trainer_ae.py:10-50, kwatsch/brain/trainer_ae.py:92-132).  Gradient tensors are
 * ALWAYS bf16 NHWC (fp32 range); activations are `dtype`; reductions, parameters and optimizer state are fp32.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Merged batches.  One training step runs the SAME modules on two batches with separate BatchNorm statistics: enc(x) and
 * enc(slice_between), dec(z) and dec(z_mix) (kwatsch/cardiac/trainer_ae.py:18-26,165-182).  Convolutions do not care, so the
 * two batches are concatenated (images [0, split) = first pass, [split, N) = second pass) and every conv launches once;
 * the BatchNorm entry points below take `split` and keep per-pass statistics ([2][...] arrays), i.e. the results are those
 * of two module calls in that order.  split <= 0 or >= N: a single pass. */

/* nn.BatchNorm2d in train mode (networks/acai_vanilla.py:58,90).  Per pass p: stats[p][c] = sum a, stats[p][C+c] = sum a^2
 * over count / count1 positions (accumulated by aesr_conv3x3_fwd's epilogue, `stats_split`).  Writes every pass'
 * scale/shift (gamma/sqrt(var+eps), ...), mean, invstd ([passes][C] each), and updates running_mean/var (momentum,
 * unbiased variance) once per pass in order; running_* may be NULL. */
int aesr_bn_finalize(const float* stats, float count, float count1, int passes, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                     float* mean_out, float* invstd_out, int C, void* stream);
/* out = mode(a * scale + shift), mode 0 same / 1 AvgPool2d(2) / 2 Upsample(2) nearest; a, out NHWC `dtype`; images >= split
 * use scale[C..2C) / shift[C..2C). */
int aesr_bn_apply(const void* a, const float* scale, const float* shift, void* out, int N, int H, int W, int C, int mode,
                  int dtype, int split, void* stream);
/* Backward of [LeakyReLU ->] BatchNorm(train) -> pool/upsample: dnext bf16 (gradient of the pooled / upsampled tensor),
 * a = saved post-activation input of the BN; g_out bf16 [N,H,W,C] = gradient w.r.t. the producing conv's
 * pre-activation output; dgamma / dbeta accumulated (over both passes); sums = passes*2*C floats of scratch;
 * mean / invstd [passes][C]; dbias_conv (C floats or NULL) += per-channel sums of g_out = the bias gradient of the
 * producing conv (summed by the apply pass, no extra pass over g_out).
 * phase 0 = reduce + apply; 1 = reduce only (sums[p][c] = sum dy, sums[p][C+c] = sum dy*xhat); 2 = apply only -- a
 * data-parallel caller all-reduces `sums` between 1 and 2 and passes the GLOBAL element counts (`count` / `count1` <= 0:
 * the local ones). */
int aesr_bn_bwd(const void* dnext, const void* a, const float* mean, const float* invstd, const float* gamma,
                float* sums, float slope, void* g_out, float* dgamma, float* dbeta, float* dbias_conv, int N, int H, int W,
                int C, int mode, int dtype, int phase, float count, float count1, int split, void* stream);
/* F.mse_loss(a, b) (kwatsch/base_trainer.py:177): *loss_acc += mean((a-b)^2); d (optional) = grad_scale * 2 (a-b)/n. */
int aesr_mse(const float* a, const float* b, size_t n, float* loss_acc, float* d, float grad_scale, void* stream);
/* Backward of dec.14 + Sigmoid: g_in bf16 (includes LeakyReLU'(a_in)), dw9c[9*C] and dbias accumulated; dbias_in (C floats
 * or NULL) += per-channel sums of g_in = the bias gradient of the conv that produced a_in (dec.12). */
int aesr_head_bwd(const float* dout, const float* out, const void* a_in, const float* w9c, void* g_in, float* dw9c,
                  float* dbias, float* dbias_in, int N, int H, int W, int C, float slope, int dtype, void* stream);
/* Backward of enc.0 (1x1 conv, padding 1): dw[C], db[C] accumulated from g bf16 [N,H+2,W+2,C] and x fp32 [N,1,H,W]. */
int aesr_e0_bwd(const void* g, const float* x, float* dw, float* db, int N, int H, int W, int C, void* stream);
/* Weight gradient of a 3x3 conv: dW fp32 [Cout,Cin,3,3] += g^T * shifted(x), dbias[Cout] += sum g (dbias may be NULL).
 * g bf16 [N,H,W,Cout], x `dtype` [N,H,W,Cin].  algo: 0 auto, 1 tcgen05 (pixels as the GEMM K dimension, MN-major
 * operands straight from NHWC, channels in {32,64,128}), 2 CUDA cores. */
int aesr_wgrad3x3(const void* g, const void* x, float* dW, float* dbias, int N, int H, int W, int Cin, int Cout, int dtype,
                  int algo, void* stream);
/* Backward of z_mix[b] = wa[b] z[b] + wb[b] z[B+b]: g_z[2B] = g_dec[2B] + {wa,wb}[b] * g_mix[b] (bf16 NHWC). */
int aesr_mix_bwd(const void* g_dec, const void* g_mix, const float* wa, const float* wb, void* g_z, int B,
                 size_t per_image, void* stream);
/* torch.optim.Adam step over flat fp32 buffers (kwatsch/trainer_ae.py:29-30); `step` is the 1-based step count. */
int aesr_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream);
/* Same step with the 1-based step count read from device memory (the bias corrections are formed in the kernel) and,
 * when `lr_dev` is not NULL, the learning rate read from device memory too (`lr` is then ignored): the launch carries no
 * per-step host value, so a training step captured in a CUDA graph can be replayed under a per-iteration scheduler
 * (CosineAnnealingLR, kwatsch/base_trainer.py:18-22). */
int aesr_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, const int* step_dev, const float* lr_dev, void* stream);

/* LPIPS-VGG v0.1 (lpips/perceptual.py:19-33, lpips/networks_basic.py:63-110, lpips/pretrained_networks.py:97-135). */
/* conv1_1 with the input pipeline folded in: (2*img-1 if normalize), ScalingLayer 1->3 channels, conv 3->64, ReLU.
 * img fp32 [N,1,H,W]; w fp32 [64,3,3,3]; shift3/scale3 HOST pointers to 3 floats; out NHWC `dtype` [N,H,W,64]. */
int aesr_vgg_conv1_fwd(const float* img, const float* w, const float* b, void* out, int N, int H, int W,
                       const float* shift3, const float* scale3, int normalize, int dtype, void* stream);
/* its backward to the image: g bf16 [N,H,W,64] (ReLU' applied) -> dimg fp32 [N,1,H,W] * out_scale. */
int aesr_vgg_conv1_bwd(const void* g, const float* w, float* dimg, int N, int H, int W, const float* scale3,
                       int normalize, float out_scale, void* stream);
/* MaxPool2d(2) backward + tap gradient + ReLU': g_out = relu'(a) * (route(d_pooled) + g_tap); d_pooled or g_tap may be NULL. */
int aesr_maxpool_bwd(const void* a, const void* d_pooled, const void* g_tap, void* g_out, int N, int H, int W, int C,
                     int dtype, void* stream);
/* Distance head of one VGG tap: val[n] += spatial_mean(sum_c lin_c (f0_c - f1_c)^2) with f = o/(||o||+1e-10);
 * if g1 != NULL also the gradient w.r.t. o1 (bf16), scaled by upstream[n].  o0, o1 NHWC `dtype` [N,HW,C]. */
int aesr_lpips_head(const void* o0, const void* o1, const float* lin, float* val, const float* upstream, void* g1, int N,
                    int HW, int C, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Evaluation / data-path utilities (HBM-bound, one pass over the data)
 * ------------------------------------------------------------------------------------------------------------------ */

/* Per-slice SSIM / squared error / minimum of two fp32 volumes [Z,H,W] (evaluate/metrics.py:111-194, which calls
 * scikit-image structural_similarity(im1, im2) / peak_signal_noise_ratio(true, test) per slice).  win = 7 (or 5),
 * uniform window, sample covariance, K1 .01, K2 .03, border of (win-1)/2 cropped; data_range is the caller's choice
 * (2.0 reproduces the reference's legacy float default, 1.0 the [0,1] images' true range).
 * ssim_sum[z] = sum of S over the (H-win+1)(W-win+1) kept pixels; sqerr_sum[z] = sum (a-b)^2 over the slice;
 * min_key[z] = order-preserving uint32 key of min(a) (key >= 0x80000000 <=> min >= 0).  Z <= 65535. */
int aesr_ssim_psnr(const float* a, const float* b, int Z, int H, int W, int win, double data_range, double* ssim_sum,
                   double* sqerr_sum, unsigned int* min_key, void* stream);

/* Pixel-domain multi-scale VIF of Z slice pairs exactly as evaluate/vifvec.py:7-63 computes it when called with uint8
 * slices (evaluate/metrics.py:65-108): four scales, scipy.ndimage.gaussian_filter semantics on uint8 planes (float64
 * accumulation in scipy's order, truncation to uint8 after every 1-D pass, 'reflect' boundary), uint8 products and
 * variance differences modulo 256.  ref_u8 / dist_u8: uint8 [Z,H,W] (aesr_vif_quantize_u8 = np.uint8(np.clip(x*255,0,255))
 * of fp32 images); weights (device): the four gaussian kernels back to back, radii_host[4] their radii, both formed on
 * the host like scipy's _gaussian_kernel1d; num_den (device) [Z][2] = (numerator, denominator), VIF = num/den (nan if 0). */
size_t aesr_vif_workspace_bytes(int Z, int H, int W);
int aesr_vif_quantize_u8(const float* x, void* out_u8, size_t n, void* stream);
int aesr_vif_mscale(const void* ref_u8, const void* dist_u8, int Z, int H, int W, const double* weights,
                    const int* radii_host, double sigma_nsq, void* workspace, size_t workspace_bytes, double* num_den,
                    void* stream);

/* np.percentile(x, (q_lo, q_hi)) ('linear', float64) by exact 3-pass radix select, then
 * out = clip((x - p_lo) / (p_hi - p_lo), 0, 1) in float64 rounded once to fp32 (generate_hr_volumes.py:130-133,
 * datasets/common.py:408-417).  out may be NULL (percentiles only -> lo_hi_out[2], device doubles, may be NULL). */
size_t aesr_percentile_workspace_bytes(void);
int aesr_percentile_normalize(const float* x, float* out, size_t n, double q_lo, double q_hi, void* workspace,
                              size_t workspace_bytes, double* lo_hi_out, void* stream);

/* Batched zero-pad + crop of fp32 images (datasets/shared_transforms.py AdjustToPatchSize :389-447, CenterCrop
 * :297-363, RandomCrop :48-120): out[b,c,y,x] = in[b,c,y+top[b],x+left[b]] inside the source, 0 outside. */
int aesr_pad_crop_gather(const float* in, float* out, const int* top, const int* left, int B, int C, int Hin, int Win,
                         int Hout, int Wout, void* stream);

/* Fused training augmentation of a batch of fp32 samples [B,C,Hin,Win] -> [B,C,P,P]: composite pad/crop window
 * (AdjustToPatchSize / CenterCrop / RandomCrop, datasets/shared_transforms.py:389-447, 297-363, 48-120; top/left may be
 * negative = zero padding), np.rot90 by rot_k[b] quarter turns (RandomRotation :224-254; NULL = none) and the sigmoid
 * contrast 1/(1+exp(gain[b]*(cutoff[b]-x))) of RandomIntensity (:366-386; gain/cutoff NULL = none) on the channels whose
 * bit is set in chan_mask (`slice_mask`).  The draws come from the caller (numpy RandomState stream preserved on the host). */
int aesr_augment_gather(const float* in, float* out, const int* top, const int* left, const int* rot_k, const float* gain,
                        const float* cutoff, unsigned chan_mask, int B, int C, int Hin, int Win, int P, void* stream);

/* LR-dataset synthesis, datasets/common_brains.py:37-44 (simulate_thick_slices): scipy.ndimage.gaussian_filter1d along
 * axis 0 of a fp32 volume [Z,HW] ('reflect' borders, float64 accumulation in scipy's order, rounded once to fp32).
 * taps = device float64 [2*lw+1], the normalised gaussian (sigma = slice_thickness / 2.355, lw = int(4*sigma + 0.5)).
 * Out of place. */
int aesr_gauss1d_axis0(const float* in, float* out, const double* taps, int lw, int Z, size_t HW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AESR_B200_H */
