// 3x3 / pad-1 / stride-1 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   activations : NHWC, 16-bit (bf16 or fp16, fp32 accumulate), C in {32, 64, 128, 256, 512}
//   weights     : packed [9 taps][Cout][Cin] 16-bit (K-major B operand per tap)
//   tile        : 128 output pixels (16 rows x 8 cols) x BN output channels, accumulator in TMEM (2 buffers)
//
// Zero padding comes from TMA: boxes may start at coordinate -1 / run past the image and out-of-bounds elements are
// zero-filled, so there is no im2col, no padded copy and no masking in the main loop.
//
// Two kernels share the epilogue:
//   conv3x3_halo_kernel   (layers whose filter bank fits in shared memory: every autoencoder layer)
//       One TMA box {KC, 10, 18, 1} per (tile, 64-channel chunk) brings the 18x10 halo window in ONCE; the nine taps
//       are nine UMMA descriptors whose start address is shifted by (dy*10+dx) pixel rows and whose 8-row group
//       stride (SBO) is the halo pitch -- the hardware swizzle is a function of the shared-memory address, so a
//       row-shifted descriptor reads exactly what TMA wrote (measured: tools/gpu_diag.py halo probe).  The whole
//       [9][BN][Cin] filter bank is loaded once per CTA and stays resident.  L2->SMEM traffic per tile drops from
//       9 x (A + B) to 1.4 x A.
//   conv3x3_stream_kernel (large filter banks, e.g. the VGG16 trunk of LPIPS)
//       one K-block = (tap, 64-channel chunk): A box {KC, 8, 16, 1} shifted by the tap offset + the B tile of that tap.
//
// Warp roles (576 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM
// owner, warps 2-17 = four epilogue sets of 4 warps (TMEM -> registers -> fused bias / activation / BN-affine / pool /
// upsample -> global).  Each set owns one TMEM accumulator, so up to four tile epilogues overlap the MMAs of later
// tiles: the epilogue of a thin (32-channel) layer is latency-bound per warp and needs that much thread-level
// parallelism to keep up with 288 tensor-core cycles per tile (measured with ncu, profiles/).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace aesr {

enum ConvAct : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };
enum ConvOut : int {
    OUT_SAME = 0,          // out  = NHWC 16-bit [N,H,W,Cout]
    OUT_AVGPOOL2 = 1,      // out  = NHWC 16-bit [N,H/2,W/2,Cout]  (floor; 2x2 mean of the post-affine activation)
    OUT_UP2 = 2,           // out  = NHWC 16-bit [N,2H,2W,Cout]    (nearest)
    OUT_NCHW_F32 = 3,      // out  = NCHW fp32 [N,Cout,H,W]        (+ optional out2 = NHWC 16-bit copy)
    OUT_SAME_MAXPOOL2 = 4, // out  = NHWC 16-bit full res, out2 = NHWC 16-bit [N,H/2,W/2,Cout] 2x2 max
    // "nearest x2 upsample followed by a 3x3 conv" folded into ONE 3x3 conv at the LOW resolution with 4*C output
    // channels (row = phase * C + c, phase = 2a+b; filters pre-summed per phase, aesr_pack_conv3x3_weight_up2fold)
    // followed by a depth-to-space store: channel block `phase` of low-res pixel (y,x) is pixel (2y+a, 2x+b).
    OUT_SHUFFLE2 = 5,      // out  = NHWC 16-bit [N,2H,2W,Cout/4]
    // same conv, but the C = 32 activations never leave the SM: the decoder head Conv2d(32,1,3,padding=1) is applied
    // to the thread's 2x2 hi-res block and leaves a 4x4 patch of partial sums per LOW-res pixel (head_gather sums
    // the four patches that overlap an output pixel, adds the bias and applies the sigmoid).
    OUT_SHUFFLE2_HEAD = 6, // out  = fp32 [N,H,W,16]
    // un-rounded accumulators (+ bias / activation / affine if given): the decoder's first conv applied to the encoder
    // latents BEFORE the interpolation (conv is linear: conv(a*z1 + b*z2) = a*conv(z1) + b*conv(z2)), lerp_pairs_act
    // then blends these pre-activations per alpha and applies bias + LeakyReLU.
    OUT_SAME_F32 = 7,      // out  = NHWC fp32 [N,H,W,Cout]
    // OUT_SHUFFLE2_HEAD with the head conv itself on the tensor cores (kernel template value only; p.out_mode stays
    // OUT_SHUFFLE2_HEAD): the epilogue writes LeakyReLU(acc + bias) back to TMEM as a 16-bit A operand and one elected
    // thread issues D2[128 px x 16 taps] = A[128 x 32] * Whead[32 x 16] per phase (A from TMEM, B = 1 KB in smem).
    OUT_SHUFFLE2_HEAD_TC = 8,
    // OUT_SHUFFLE2_HEAD with the head conv as warp-level mma.sync on register fragments (kernel template value only):
    // the accumulator is read in the mma fragment distribution (tcgen05.ld.16x256b), bias + LeakyReLU + 16-bit rounding
    // happen in registers and feed the A operand directly; the B operand maps (phase, channel) to the 16 patch entries,
    // so the four phases accumulate into one D fragment and nothing is scattered or shuffled.
    OUT_SHUFFLE2_HEAD_MMA = 9,
};
enum ConvMul : int { MUL_NONE = 0, MUL_LEAKY_GRAD = 1, MUL_RELU_GRAD = 2 };

struct ConvParams {
    int N, H, W, Cin, Cout;
    int BN;                 // output channels per tile (multiple of 32, <= 256, divides Cout)
    int tiles_x, tiles_y, n_blocks, num_tiles;   // num_tiles = spatial tiles (N * tiles_y * tiles_x) * n_blocks
    int num_stages;
    int T, stiles_y;        // halo kernel: M-tiles stacked vertically per super-tile, super-tile rows per image
    int head_smem;          // bytes reserved behind the filter bank for the 16 x 32 head filter block (0 or 1024)
    int nbuf;               // halo kernel: TMEM buffers (super-tiles in flight between the MMA issuer and the epilogue), 2..4
    int st_bufs;            // halo kernel, OUT_SAME / OUT_SHUFFLE2 / OUT_SHUFFLE2_HEAD inference epilogues: staging buffers per epilogue
                            // set for TMA stores (1 or 2; 4 KB = a 16-channel chunk of a tile, 8 KB = a tile of head patches);
                            // 0 = per-thread global stores
    int st_bytes;           // bytes of one staging buffer
    int fp16;               // 1: activations / weights are fp16, 0: bf16
    int debug;              // profiling only (AESR_CONV_DEBUG): bit0 = skip the MMAs, bit1 = skip the activation TMA loads
    // epilogue
    const float* bias;      // [Cout] or null
    const float* scale;     // [Cout] or null: y = act(acc + bias) * scale + shift
    const float* shift;
    float slope;
    int act;
    int out_mode;
    void* out;
    void* out2;
    // optional elementwise multiplier (dgrad): out *= act'(mul_src) with mul_src NHWC 16-bit [N,H,W,Cout]
    const uint16_t* mul_src;
    int mul_mode;
    // optional per-channel statistics of the (post-activation, pre-affine) value: stats[c] += sum, stats[Cout+c] += sum^2;
    // images >= stats_split (second pass of a merged batch, own BatchNorm statistics) accumulate into stats + 2*Cout
    float* stats;
    int stats_split;
    // streamed kernel, split-K: work item = (tile, split s), split s accumulates K-blocks [s*KB/ksplit, (s+1)*KB/ksplit) and
    // stores its raw fp32 accumulators to ws[s][pixel][Cout]; splitk_finish_kernel adds the splits and runs the epilogue.
    int ksplit;
    float* ws;
    int stats_sum_only;     // 1: stats is [Cout], only the sums are taken (bias gradient of the layer whose output gradient this
                            // data-gradient launch produces: dbias[c] = sum over pixels of the epilogue's output)
    // OUT_SHUFFLE2_HEAD: fp32 [9][32] filter of the 32 -> 1 head conv, BY VALUE: kernel parameters live in the constant
    // bank, so the 1152 head MACs per thread read their weights as instruction operands (c[0][..]) instead of through
    // 288 LDS.128 per thread and tile, which made the epilogue shared-memory-bandwidth-bound (191 -> see profiles/).
    alignas(16) float head_wc[9 * 32];
};

constexpr int CONV_TILE_H = 16;
constexpr int CONV_TILE_W = 8;
constexpr int CONV_TILE_M = 128;
constexpr int CONV_EPI_SETS = 4;                          // epilogue warp sets (4 warps = 128 TMEM lanes each)
constexpr int CONV_ISSUERS = 1;                            // MMA-issuing warps (one elected thread each)
constexpr int CONV_FIRST_EPI_WARP = 1 + CONV_ISSUERS;      // warp 0 TMA, warp 1 MMA issuer, then the epilogue sets
constexpr int CONV_THREADS = 32 * CONV_FIRST_EPI_WARP + 128 * CONV_EPI_SETS;   // 576
constexpr int CONV_MAX_STAGES = 8;
// TMEM accumulators in flight: one per epilogue set while they fit in the 512 columns
__host__ __device__ constexpr int conv_num_acc(int BN) { return (CONV_EPI_SETS * BN <= 512) ? CONV_EPI_SETS : 512 / BN; }
constexpr int HALO_H = CONV_TILE_H + 2;     // 18
constexpr int HALO_W = CONV_TILE_W + 2;     // 10
constexpr int CONV_TAIL_BYTES = 256 + 5 * 512 * 4;   // barriers + tmem ptr + per-channel epilogue constants + BN statistics
                                                     // (s_stats: [2 passes][2][Cout], Cout <= 256 when statistics are taken)

// Output tensor maps of the TMA-store epilogue: [0] = the NHWC output (OUT_SAME); OUT_SHUFFLE2: [ph] = the strided view of the
// hi-res output that holds phase ph = 2a+b, i.e. pixels (2y+a, 2x+b), as a {C, W, H, N} tensor of the LOW-res extents.
struct ConvOutMaps {
    CUtensorMap m[4];
};
constexpr int CONV_ST_CHUNK_BYTES = CONV_TILE_M * 32;      // one 16-channel chunk of a tile: 128 pixels x 32 bytes
constexpr int CONV_ST_BAR0 = 8;                             // named barriers 8..11: one per epilogue set

__device__ __forceinline__ uint32_t pack2(float lo, float hi, int fp16) {
    if (fp16) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ uint4 pack8(const float* v, int fp16) {
    return make_uint4(pack2(v[0], v[1], fp16), pack2(v[2], v[3], fp16), pack2(v[4], v[5], fp16),
                      pack2(v[6], v[7], fp16));
}
// sign test valid for both bf16 and fp16 bit patterns: strictly positive <=> sign clear and magnitude non-zero
__device__ __forceinline__ bool pos16(uint32_t h) { return (h & 0x8000u) == 0 && (h & 0x7FFFu) != 0; }

struct ConvBarriers {
    uint64_t* full;        // [CONV_MAX_STAGES]
    uint64_t* empty;       // [CONV_MAX_STAGES]
    uint64_t* tmem_full;   // [CONV_EPI_SETS]
    uint64_t* tmem_empty;  // [CONV_EPI_SETS]
    uint64_t* b_full;      // [1] resident filter bank landed
    uint64_t* head_bar;    // [CONV_EPI_SETS] head MMAs of an epilogue set's tile completed (OUT_SHUFFLE2_HEAD_TC)
    uint32_t* tmem_ptr;
    float *s_bias, *s_scale, *s_shift;
    float* s_stats;        // [2][512] per-CTA sums / sums of squares (training), flushed once at the end of the kernel
    __device__ explicit ConvBarriers(uint8_t* tail) {
        full = reinterpret_cast<uint64_t*>(tail);
        empty = full + CONV_MAX_STAGES;
        tmem_full = empty + CONV_MAX_STAGES;
        tmem_empty = tmem_full + CONV_EPI_SETS;
        b_full = tmem_empty + CONV_EPI_SETS;
        head_bar = b_full + 1;
        tmem_ptr = reinterpret_cast<uint32_t*>(head_bar + CONV_EPI_SETS);
        s_bias = reinterpret_cast<float*>(tail + 256);
        s_scale = s_bias + 512;
        s_shift = s_scale + 512;
        s_stats = s_shift + 512;
    }
};

__device__ __forceinline__ uint32_t conv_tmem_cols(int BN) {
    const int c = conv_num_acc(BN) * BN;
    return (c <= 32) ? 32 : (c <= 64) ? 64 : (c <= 128) ? 128 : (c <= 256) ? 256 : 512;
}

// Common prologue: barrier init, TMEM allocation (warp 1), epilogue constants.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t conv_prologue(const ConvParams& p, const ConvBarriers& bars, int num_stages,
                                                  uint32_t tmem_cols, uint32_t tmem_empty_count) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        for (int a = 0; a < CONV_EPI_SETS; ++a) {
            mbar_init(&bars.tmem_full[a], 1);
            mbar_init(&bars.tmem_empty[a], tmem_empty_count);   // one elected lane per epilogue warp
            mbar_init(&bars.head_bar[a], 1);
        }
        mbar_init(bars.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(bars.tmem_ptr, tmem_cols);
    // folded-upsample modes: the four phase blocks share the C = Cout/4 per-channel constants
    const int cmod = (p.out_mode == OUT_SHUFFLE2 || p.out_mode == OUT_SHUFFLE2_HEAD) ? (p.Cout >> 2) : p.Cout;
    for (int c = threadIdx.x; c < p.Cout; c += CONV_THREADS) {
        const int cc = c % cmod;
        bars.s_bias[c] = p.bias ? p.bias[cc] : 0.f;
        bars.s_scale[c] = p.scale ? p.scale[cc] : 1.f;
        bars.s_shift[c] = p.shift ? p.shift[cc] : 0.f;
    }
    if (p.stats != nullptr)
        for (int c = threadIdx.x; c < 4 * p.Cout; c += CONV_THREADS) bars.s_stats[c] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *bars.tmem_ptr;
}

template <bool WITH_STATS>
__device__ __forceinline__ void conv_teardown(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_base,
                                              uint32_t tmem_cols) {
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (WITH_STATS && p.stats != nullptr)        // every epilogue of this CTA has added its sums: flush them
        for (int c = threadIdx.x; c < (p.stats_sum_only ? 1 : p.stats_split < p.N ? 4 : 2) * p.Cout; c += CONV_THREADS) {
            const float v = bars.s_stats[c];
            if (v != 0.f) atomicAdd(p.stats + c, v);
        }
    if ((threadIdx.x >> 5) == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

struct TileCoord {
    int n, y0, x0, n0;
};
// tile -> (image, tile origin, first output channel).  n-block is the slowest index so that a CTA of the halo kernel
// (grid-strided inside one n-block) keeps one filter bank for its whole life.
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int sp, int nb) {
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    TileCoord t;
    t.n = sp / tiles_per_img;
    const int r = sp - t.n * tiles_per_img;
    const int ty = r / p.tiles_x;
    t.y0 = ty * CONV_TILE_H;
    t.x0 = (r - ty * p.tiles_x) * CONV_TILE_W;
    t.n0 = nb * p.BN;
    return t;
}

// Sum of 16 per-lane values over the 32 lanes of a warp, transposed: 16 shuffles instead of 16 x 5 butterflies.  Each
// round a lane keeps half of its values and sends the other half to its partner (lane ^ 16, 8, 4, 2), adding what it
// receives; a last xor-1 round folds the remaining pair.  On return lane l holds the total of value index
// ((l >> 4) & 1) * 8 + ((l >> 3) & 1) * 4 + ((l >> 2) & 1) * 2 + ((l >> 1) & 1)  (both lanes of a pair hold it).
__device__ __forceinline__ float warp_transpose_sum16(const float (&v)[16], int lane) {
    float a8[8], a4[4], a2[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = h16 ? v[8 + j] : v[j], send = h16 ? v[j] : v[8 + j];
        a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = h8 ? a8[4 + j] : a8[j], send = h8 ? a8[j] : a8[4 + j];
        a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = h4 ? a4[2 + j] : a4[j], send = h4 ? a4[j] : a4[2 + j];
        a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = h2 ? a2[1] : a2[0], send = h2 ? a2[0] : a2[1];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// General epilogue of one tile (training and VGG: act' multiplier, BN statistics, second outputs, every output stage
// selected at run time), executed by the 4 warps of an epilogue set (128 threads = 128 TMEM lanes = 128 pixels);
// 16 accumulator columns per TMEM load to stay inside the 96-register budget of a 576-thread CTA.
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_acc,
                                                   uint64_t* tmem_empty_bar, const TileCoord& t) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;              // pixel index inside the tile
    const int py = row >> 3, px = row & 7;
    const int H = p.H, W = p.W, Cout = p.Cout, fp16 = p.fp16;
    const int n = t.n, y = t.y0 + py, x = t.x0 + px;
    const bool inb = (y < H) && (x < W);
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    uint16_t* out16 = static_cast<uint16_t*>(p.out);
    uint16_t* out2_16 = static_cast<uint16_t*>(p.out2);
    const int out_mode = p.out_mode;
    const size_t pix = (static_cast<size_t>(n) * H + y) * W + x;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
        float v[16];
        const int cg = t.n0 + c0;               // first global output channel of this chunk
        {
            uint32_t raw[16];
            if (p.debug & 8) {
#pragma unroll
                for (int j = 0; j < 16; ++j) raw[j] = 0;
            } else {
                tmem_ld_32x32b_x16(t_addr + c0, raw);
                tmem_ld_wait();
            }
            if (c0 + 16 >= p.BN) {              // accumulator fully read: one elected arrive per warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar);
            }
            if (p.debug & 64) continue;
            // act(a) = max(a, a * neg_slope): neg_slope = 1 (identity), 0.01 (LeakyReLU), 0 (ReLU) -- branch-free
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b = lds_f4(bars.s_bias + cg + 4 * j4);
                const float a0 = __uint_as_float(raw[j4 * 4 + 0]) + b.x, a1 = __uint_as_float(raw[j4 * 4 + 1]) + b.y;
                const float a2 = __uint_as_float(raw[j4 * 4 + 2]) + b.z, a3 = __uint_as_float(raw[j4 * 4 + 3]) + b.w;
                v[j4 * 4 + 0] = fmaxf(a0, a0 * neg_slope);
                v[j4 * 4 + 1] = fmaxf(a1, a1 * neg_slope);
                v[j4 * 4 + 2] = fmaxf(a2, a2 * neg_slope);
                v[j4 * 4 + 3] = fmaxf(a3, a3 * neg_slope);
            }
        }
        if (p.mul_mode != MUL_NONE) {
            if (inb) {
                const uint4* m4 = reinterpret_cast<const uint4*>(p.mul_src + pix * Cout + cg);
                const float neg = (p.mul_mode == MUL_LEAKY_GRAD) ? p.slope : 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const uint4 m = __ldg(m4 + j4);
                    const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[j4 * 8 + u * 2] *= pos16(w[u] & 0xFFFFu) ? 1.f : neg;
                        v[j4 * 8 + u * 2 + 1] *= pos16(w[u] >> 16) ? 1.f : neg;
                    }
                }
            }
        }
        if (p.stats != nullptr) {
            // per-channel sum / sum of squares over the valid pixels of this warp (transposed warp reduction), then one
            // atomic per channel from the even lane of each pair
            float s1[16], s2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                s1[j] = inb ? v[j] : 0.f;
                s2[j] = s1[j] * s1[j];
            }
            const float t1 = warp_transpose_sum16(s1, lane);
            const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if (p.stats_sum_only) {
                if ((lane & 1) == 0) atomicAdd(bars.s_stats + cg + ch, t1);
            } else {
            const float t2 = warp_transpose_sum16(s2, lane);
            if ((lane & 1) == 0) {
                // shared-memory accumulators: one global atomic per channel and CTA at the end of the kernel instead of one
                // per warp and 16-channel chunk (940 k atomics onto 64 addresses made enc.3's forward 95 us against 36 us
                // without statistics, tools/train_layer_times.py)
                float* sst = bars.s_stats + (n >= p.stats_split ? 2 * Cout : 0);      // pass of a merged batch (tile-uniform)
                atomicAdd(sst + cg + ch, t1);
                atomicAdd(sst + Cout + cg + ch, t2);
            }
            }
        }
        if (p.scale != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 sc = lds_f4(bars.s_scale + cg + 4 * j4), sh = lds_f4(bars.s_shift + cg + 4 * j4);
                v[j4 * 4 + 0] = fmaf(v[j4 * 4 + 0], sc.x, sh.x);
                v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], sc.y, sh.y);
                v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], sc.z, sh.z);
                v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], sc.w, sh.w);
            }
        }
        if (out_mode == OUT_SAME || out_mode == OUT_SAME_MAXPOOL2) {
            if (inb && !(p.debug & 4)) {
                uint4* o4 = reinterpret_cast<uint4*>(out16 + pix * Cout + cg);
                o4[0] = pack8(v, fp16);
                o4[1] = pack8(v + 8, fp16);
            }
        }
        if (out_mode == OUT_SAME_F32) {
            if (inb) {
                float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * Cout + cg);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) o4[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
            }
        }
        if (out_mode == OUT_AVGPOOL2 || out_mode == OUT_SAME_MAXPOOL2) {
            // 2x2 window = lanes {l, l^1, l^8} (x neighbour, y neighbour): tile origins are even.
            const bool is_max = (out_mode == OUT_SAME_MAXPOOL2);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = v[j];
                float b = __shfl_xor_sync(0xffffffffu, a, 1);
                a = is_max ? fmaxf(a, b) : a + b;
                b = __shfl_xor_sync(0xffffffffu, a, 8);
                a = is_max ? fmaxf(a, b) : (a + b) * 0.25f;
                v[j] = a;
            }
            const int Ho = H >> 1, Wo = W >> 1;
            const int yo = y >> 1, xo = x >> 1;
            if (((px | py) & 1) == 0 && yo < Ho && xo < Wo) {
                uint16_t* dst = (out_mode == OUT_AVGPOOL2) ? out16 : out2_16;
                uint4* o4 = reinterpret_cast<uint4*>(
                    dst + (static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(yo) * Wo + xo) * Cout + cg);
                o4[0] = pack8(v, fp16);
                o4[1] = pack8(v + 8, fp16);
            }
        } else if (out_mode == OUT_UP2) {
            if (inb) {
                const int Wo = 2 * W;
                const uint4 pk0 = pack8(v, fp16), pk1 = pack8(v + 8, fp16);
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int yo = 2 * y + (d >> 1), xo = 2 * x + (d & 1);
                    uint4* o4 = reinterpret_cast<uint4*>(
                        out16 + ((static_cast<size_t>(n) * 2 * H + yo) * Wo + xo) * Cout + cg);
                    o4[0] = pk0;
                    o4[1] = pk1;
                }
            }
        } else if (out_mode == OUT_SHUFFLE2) {
            if (inb) {
                const int C = Cout >> 2;
                const int ph = cg / C, ch = cg - ph * C;
                const int yo = 2 * y + (ph >> 1), xo = 2 * x + (ph & 1);
                uint4* o4 = reinterpret_cast<uint4*>(
                    out16 + ((static_cast<size_t>(n) * 2 * H + yo) * (2 * W) + xo) * C + ch);
                o4[0] = pack8(v, fp16);
                o4[1] = pack8(v + 8, fp16);
            }
        } else if (out_mode == OUT_NCHW_F32) {
            if (inb) {
                float* o = static_cast<float*>(p.out) + (static_cast<size_t>(n) * Cout + cg) * H * W +
                           static_cast<size_t>(y) * W + x;
#pragma unroll
                for (int j = 0; j < 16; ++j) o[static_cast<size_t>(j) * H * W] = v[j];
                if (p.out2 != nullptr) {
                    uint4* o4 = reinterpret_cast<uint4*>(out2_16 + pix * Cout + cg);
                    o4[0] = pack8(v, fp16);
                    o4[1] = pack8(v + 8, fp16);
                }
            }
        }
    }
}

// Split-K epilogue of the streamed kernel: raw fp32 accumulators -> ws[split][pixel][Cout] (no bias / activation).
__device__ __forceinline__ void conv_epilogue_splitk(const ConvParams& p, uint32_t tmem_acc, uint64_t* tmem_empty_bar,
                                                     const TileCoord& t, int split) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row >> 3, px = row & 7;
    const int y = t.y0 + py, x = t.x0 + px;
    const bool inb = (y < p.H) && (x < p.W);
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const size_t npix = static_cast<size_t>(p.N) * p.H * p.W;
    const size_t pix = (static_cast<size_t>(t.n) * p.H + y) * p.W + x;
    float* dst = p.ws + (static_cast<size_t>(split) * npix + pix) * p.Cout + t.n0;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t raw[16];
        tmem_ld_32x32b_x16(t_addr + c0, raw);
        tmem_ld_wait();
        if (c0 + 16 >= p.BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
        if (inb) {
            float4* o4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
                o4[j4] = make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]),
                                     __uint_as_float(raw[4 * j4 + 2]), __uint_as_float(raw[4 * j4 + 3]));
        }
    }
}

// out[pix][c] = act(sum_s ws[s][pix][c] + bias[c]) * act'(mul_src[pix][c]) -> 16-bit NHWC; one thread = 8 channels of a pixel.
template <bool AF>
__global__ void splitk_finish_kernel(const float* __restrict__ ws, int ksplit, size_t npix, int Cout,
                                     const float* __restrict__ bias, int act, float slope,
                                     const uint16_t* __restrict__ mul_src, int mul_mode, uint16_t* __restrict__ out) {
    const int groups = Cout >> 3;
    const size_t total = npix * groups;
    const float neg_slope = (act == ACT_LEAKY) ? slope : (act == ACT_RELU) ? 0.f : 1.f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = bias ? __ldg(bias + g * 8 + k) : 0.f;
        for (int sidx = 0; sidx < ksplit; ++sidx) {
            const float4* w4 = reinterpret_cast<const float4*>(ws + (static_cast<size_t>(sidx) * npix * Cout) + i * 8);
            const float4 a = __ldg(w4), b = __ldg(w4 + 1);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], v[k] * neg_slope);
        if (mul_mode != MUL_NONE) {
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(mul_src) + i);
            const uint32_t w[4] = {m.x, m.y, m.z, m.w};
            const float neg = (mul_mode == MUL_LEAKY_GRAD) ? slope : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[2 * u] *= pos16(w[u] & 0xFFFFu) ? 1.f : neg;
                v[2 * u + 1] *= pos16(w[u] >> 16) ? 1.f : neg;
            }
        }
        reinterpret_cast<uint4*>(out)[i] = pack8(v, AF ? 1 : 0);
    }
}

// 2x2 pooling of 16 per-lane values over the lanes {l, l^1, l^8, l^9} of a window (x neighbour, y neighbour; tile origins are
// even), transposed: each round a lane keeps half of its values, sends the other half to its partner and combines what it
// receives -- 8 + 4 shuffles instead of 2 x 16 (shuffles are MIO instructions like the issuer's tcgen05.mma).  On return r[0..3]
// = the window's sums (maxima) of channels 8 * (lane & 1) + 4 * ((lane >> 3) & 1) + 0..3: the four lanes of a window hold its
// 16 channels.  Same association as summing the x pair first: bit-identical to the straightforward version.
template <bool IS_MAX>
__device__ __forceinline__ void pool2x2_transposed(const float (&v)[16], int lane, float (&r)[4]) {
    const bool hx = lane & 1, hy = lane & 8;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = hx ? v[8 + j] : v[j], send = hx ? v[j] : v[8 + j];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        a[j] = IS_MAX ? fmaxf(keep, recv) : keep + recv;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = hy ? a[4 + j] : a[j], send = hy ? a[j] : a[4 + j];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        r[j] = IS_MAX ? fmaxf(keep, recv) : keep + recv;
    }
}

// Lean epilogue of the inference output stages (OUT_SAME, OUT_AVGPOOL2, OUT_SHUFFLE2, OUT_SAME_F32; no training extras): 16 accumulator
// columns per TMEM load, so that 16 values + addresses + the role's loop state stay far below the 96 registers a
// 576-thread CTA allows (the 32-column generic epilogue spills its loop state, ncu: LDL stalls in every tile).
// Training variants of the lean epilogue (kernel template values only; p.out_mode stays OUT_SAME / OUT_SAME_MAXPOOL2): the
// fully dynamic epilogue spills (148 bytes of spill stores at 96 registers) and ncu shows its warps waiting on local-memory
// loads (long_scoreboard 5-30 per issue, profiles/r06c_train_conv_ncu.md): VGG conv1_2 took 70 us through it against 45 us
// through the plain OUT_SAME instantiation.  One instantiation per combination the training step uses:
enum ConvLean : int {
    LEAN_SAME_MUL = 16,        // OUT_SAME * act'(mul_src)                      (VGG data gradients)
    LEAN_SAME_STATS = 17,      // OUT_SAME + per-channel sum / sum^2 (2 passes)  (conv in front of a train-mode BatchNorm)
    LEAN_SAME_MUL_SUM = 18,    // OUT_SAME * act'(mul_src) + per-channel sum     (autoencoder data gradients + bias gradient)
};
__host__ __device__ constexpr int lean_out(int mode) { return mode >= 16 ? OUT_SAME : mode; }
__host__ __device__ constexpr bool lean_mul(int mode) { return mode == LEAN_SAME_MUL || mode == LEAN_SAME_MUL_SUM; }
__host__ __device__ constexpr bool lean_stats(int mode) { return mode == LEAN_SAME_STATS || mode == LEAN_SAME_MUL_SUM; }

// TMA-store path (p.st_bufs > 0; OUT_SAME / OUT_SHUFFLE2 without training extras): the 16-channel chunk of the tile goes to a
// 4 KB staging buffer of the epilogue set (row = TMEM lane = pixel, 32 bytes) and ONE elected thread stores it with
// cp.async.bulk.tensor (box {16 ch, 8, 16, 1}, clipped at the image bounds by the hardware) instead of 128 threads x 2 STG.128.
// Measured (profiles/r09_store_cost_stage_isolation.txt): the per-thread global stores cost the 64 / 128-column layers 14-29 %
// (dec.2 0.317 -> 0.246 ms, dec.6 0.540 -> 0.387 ms without them) although the epilogue warps mostly WAIT for the MMAs there;
// the same bytes written to shared memory instead cost ~nothing (dec.2 0.250, dec.6 0.392 ms).  ncu shows the MMA issuer
// stalled on mio_throttle at its UTCHMMAs (profiles/r08_halo_issuer_stalls.txt); staging only the warps of the issuer's
// sub-partition is not enough (per-thread stores from the other three quarters still cost 10-20 %, r10 record).
//   st_set: this set's staging buffers, st_k: running chunk count of this thread's set (buffer = st_k % st_bufs),
//   st_elect: this thread issues (and owns the bulk groups of) the set's stores.
// TMAST is a compile-time property of the kernel instantiation: with the staged path only behind a run-time flag the 32-column
// OUT_SAME instantiation (which never takes it) got 4 % slower (dec.8 0.449 -> 0.468 ms, same box).
template <int KMODE, bool TMAST = false>
__device__ __forceinline__ void conv_epilogue_lean(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_acc,
                                                   uint64_t* tmem_empty_bar, const TileCoord& t,
                                                   const ConvOutMaps* omaps = nullptr, uint8_t* st_set = nullptr,
                                                   uint32_t* st_k = nullptr, bool st_elect = false, int eset = 0) {
    constexpr int MODE = lean_out(KMODE);
    constexpr bool WITH_MUL = lean_mul(KMODE), WITH_STATS = lean_stats(KMODE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;              // pixel index inside the tile
    const int py = row >> 3, px = row & 7;
    const int H = p.H, W = p.W, fp16 = p.fp16;
    const int y = t.y0 + py, x = t.x0 + px;
    const bool inb = (y < H) && (x < W);
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const float2 slope2 = make_float2(neg_slope, neg_slope);
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const bool affine = p.scale != nullptr;
    // element offset of this thread's pixel (channel 0) in the output tensor
    size_t pix;
    bool store;
    int Cpix;                                   // channels per output pixel
    if (MODE == OUT_SAME || MODE == OUT_SAME_F32 || MODE == OUT_SAME_MAXPOOL2) {
        Cpix = p.Cout;
        pix = (static_cast<size_t>(t.n) * H + y) * W + x;
        store = inb;
    } else if (MODE == OUT_AVGPOOL2) {
        Cpix = p.Cout;
        const int Ho = H >> 1, Wo = W >> 1, yo = y >> 1, xo = x >> 1;
        pix = (static_cast<size_t>(t.n) * Ho + yo) * Wo + xo;
        store = yo < Ho && xo < Wo;             // all four lanes of a window store 4 channels each (pool2x2_transposed)
    } else {                                    // OUT_SHUFFLE2: phase offset added per chunk
        Cpix = p.Cout >> 2;
        pix = (static_cast<size_t>(t.n) * 2 * H + 2 * y) * (2 * W) + 2 * x;
        store = inb;
    }
    if (p.debug & 4) store = false;
    uint16_t* out16 = static_cast<uint16_t*>(p.out);
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
        float v[16];
        const int cg = t.n0 + c0;               // first global output channel of this chunk
        {
            uint32_t raw[16];
            if (p.debug & 8) {
#pragma unroll
                for (int j = 0; j < 16; ++j) raw[j] = 0;
            } else {
                tmem_ld_32x32b_x16(t_addr + c0, raw);
                tmem_ld_wait();
            }
            if (c0 + 16 >= p.BN) {              // accumulator fully read: one elected arrive per warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar);
            }
            if (p.debug & 64) continue;
            // act(a) = max(a, a * neg_slope): neg_slope = 1 (identity), 0.01 (LeakyReLU), 0 (ReLU) -- branch-free
            // packed fp32x2 adds / multiplies (same roundings, half the issue slots: the epilogue warps of sub-partition 1 share
            // their scheduler with the MMA issuer)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b = lds_f4(bars.s_bias + cg + 4 * j4);
                const float2 a01 = __fadd2_rn(make_float2(__uint_as_float(raw[j4 * 4 + 0]), __uint_as_float(raw[j4 * 4 + 1])),
                                              make_float2(b.x, b.y));
                const float2 a23 = __fadd2_rn(make_float2(__uint_as_float(raw[j4 * 4 + 2]), __uint_as_float(raw[j4 * 4 + 3])),
                                              make_float2(b.z, b.w));
                const float2 s01 = __fmul2_rn(a01, slope2), s23 = __fmul2_rn(a23, slope2);
                v[j4 * 4 + 0] = fmaxf(a01.x, s01.x);
                v[j4 * 4 + 1] = fmaxf(a01.y, s01.y);
                v[j4 * 4 + 2] = fmaxf(a23.x, s23.x);
                v[j4 * 4 + 3] = fmaxf(a23.y, s23.y);
            }
        }
        if (WITH_MUL) {
            if (inb) {
                const uint4* m4 = reinterpret_cast<const uint4*>(p.mul_src + pix * Cpix + cg);
                const float neg = (p.mul_mode == MUL_LEAKY_GRAD) ? p.slope : 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const uint4 m = __ldg(m4 + j4);
                    const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[j4 * 8 + u * 2] *= pos16(w[u] & 0xFFFFu) ? 1.f : neg;
                        v[j4 * 8 + u * 2 + 1] *= pos16(w[u] >> 16) ? 1.f : neg;
                    }
                }
            }
        }
        if (WITH_STATS) {
            float s1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) s1[j] = inb ? v[j] : 0.f;
            const float t1 = warp_transpose_sum16(s1, lane);
            const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if (KMODE == LEAN_SAME_MUL_SUM) {                 // bias gradient: sums only, one block
                if ((lane & 1) == 0) atomicAdd(bars.s_stats + cg + ch, t1);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) s1[j] *= s1[j];
                const float t2 = warp_transpose_sum16(s1, lane);
                if ((lane & 1) == 0) {
                    float* sst = bars.s_stats + (t.n >= p.stats_split ? 2 * p.Cout : 0);
                    atomicAdd(sst + cg + ch, t1);
                    atomicAdd(sst + p.Cout + cg + ch, t2);
                }
            }
        }
        if (affine) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 sc = lds_f4(bars.s_scale + cg + 4 * j4), sh = lds_f4(bars.s_shift + cg + 4 * j4);
                const float2 r01 = __ffma2_rn(make_float2(v[j4 * 4 + 0], v[j4 * 4 + 1]), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                const float2 r23 = __ffma2_rn(make_float2(v[j4 * 4 + 2], v[j4 * 4 + 3]), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                v[j4 * 4 + 0] = r01.x; v[j4 * 4 + 1] = r01.y; v[j4 * 4 + 2] = r23.x; v[j4 * 4 + 3] = r23.y;
            }
        }
        size_t off;
        if (MODE == OUT_AVGPOOL2) {
            float r[4];
            pool2x2_transposed<false>(v, lane, r);
            if (store)
                *reinterpret_cast<uint2*>(out16 + pix * Cpix + cg + 8 * (lane & 1) + 4 * ((lane >> 3) & 1)) =
                    make_uint2(pack2(r[0] * 0.25f, r[1] * 0.25f, fp16), pack2(r[2] * 0.25f, r[3] * 0.25f, fp16));
            continue;
        } else if (MODE == OUT_SHUFFLE2) {
            const int ph = cg / Cpix, ch = cg - ph * Cpix;
            off = (pix + static_cast<size_t>(ph >> 1) * (2 * W) + (ph & 1)) * Cpix + ch;
        } else {
            off = pix * Cpix + cg;
        }
        if (MODE == OUT_SAME_F32) {
            if (store) {
                float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + off);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) o4[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
            }
        } else if (TMAST && (MODE == OUT_SAME || MODE == OUT_SAME_MAXPOOL2 || MODE == OUT_SHUFFLE2)) {
            const uint4 p0 = pack8(v, fp16), p1 = pack8(v + 8, fp16);
            const uint32_t k = *st_k;
            uint8_t* buf = st_set + (p.st_bufs == 2 ? (k & 1u) : 0u) * p.st_bytes;
            // the store that last used this buffer (chunk k - st_bufs) has been read out of shared memory
            if (st_elect) {
                if (p.st_bufs == 2) bulk_wait_group_read<1>();
                else bulk_wait_group_read<0>();
            }
            named_bar_sync(CONV_ST_BAR0 + eset, 128);
            const int sw = (row >> 2) & 1;             // CU_TENSOR_MAP_SWIZZLE_32B (buffer 1024-byte aligned): conflict-free
            sts_u4(buf + row * 32 + 16 * sw, p0);
            sts_u4(buf + row * 32 + 16 * (sw ^ 1), p1);
            fence_proxy_async();                 // generic-proxy writes -> visible to the async proxy (TMA)
            named_bar_sync(CONV_ST_BAR0 + eset, 128);
            if (st_elect && !(p.debug & 4)) {
                if (MODE != OUT_SHUFFLE2) {
                    tma_store_4d(&omaps->m[0], buf, cg, t.x0, t.y0, t.n);
                } else {
                    const int ph = cg / Cpix;
                    tma_store_4d(&omaps->m[ph], buf, cg - ph * Cpix, t.x0, t.y0, t.n);
                }
                bulk_commit_group();
            }
            *st_k = k + 1;
        } else if ((MODE == OUT_SAME || MODE == OUT_SHUFFLE2 || MODE == OUT_SAME_MAXPOOL2) && p.BN >= 64) {
            // Lane pairs (x, x+1) trade halves so that every store instruction writes whole 32-byte sectors: lane 2i
            // holds pixel P's channels [c0, c0+16) = 16-byte halves A0 A1, lane 2i+1 pixel P+1's B0 B1.  Stored directly
            // (A0 | B0, then A1 | B1) each STG.128 touches 32 half-written sectors; after one exchange the pair writes
            // A0 A1 (one sector of P), then B0 B1 (one sector of P+1): 16 full sectors per instruction.  At launch sizes
            // whose outputs stream to HBM the LSU's sector rate made the stores additive to the MMA time (dec.6:
            // 0.54 ms with, 0.37 ms without stores, profiles/r01i_conv_sweep_tc_head_parked.txt).  Measured (r02a, A/B on one
            // box): 64- and 128-column layers 5-10 % faster, the 32-column dec.8 7 % slower (its short epilogue sits on
            // the MMA -> epilogue -> MMA latency chain and the exchange lengthens it) -- hence BN >= 64 only.
            const uint4 p0 = pack8(v, fp16), p1 = pack8(v + 8, fp16);
            const bool odd = lane & 1;
            const uint4 send = odd ? p0 : p1;
            uint4 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
            recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
            recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
            const int xp = odd ? x - 1 : x + 1;                                  // partner pixel (same row of the tile)
            const bool pstore = (y < H) && (xp < W) && !(p.debug & 4);
            const long long pstep = (MODE == OUT_SHUFFLE2 ? 2 : 1) * static_cast<long long>(Cpix);
            const size_t poff = odd ? off - pstep : off + pstep;
            // instruction 1: the even lane's pixel (A0 from the even lane, A1 from the odd lane)
            if (odd ? pstore : store) *reinterpret_cast<uint4*>(out16 + (odd ? poff + 8 : off)) = odd ? recv : p0;
            // instruction 2: the odd lane's pixel (B0 from the even lane, B1 from the odd lane)
            if (odd ? store : pstore) *reinterpret_cast<uint4*>(out16 + (odd ? off + 8 : poff)) = odd ? p1 : recv;
        } else if (store) {
            uint4* o4 = reinterpret_cast<uint4*>(out16 + off);
            o4[0] = pack8(v, fp16);
            o4[1] = pack8(v + 8, fp16);
        }
        if (MODE == OUT_SAME_MAXPOOL2) {
            float r[4];
            pool2x2_transposed<true>(v, lane, r);
            const int Ho = H >> 1, Wo = W >> 1, yo = y >> 1, xo = x >> 1;
            if (yo < Ho && xo < Wo && !(p.debug & 4))
                *reinterpret_cast<uint2*>(static_cast<uint16_t*>(p.out2) + ((static_cast<size_t>(t.n) * Ho + yo) * Wo + xo) * Cpix + cg +
                                          8 * (lane & 1) + 4 * ((lane >> 3) & 1)) =
                    make_uint2(pack2(r[0], r[1], fp16), pack2(r[2], r[3], fp16));
        }
    }
}

// Epilogue of one tile in OUT_SHUFFLE2_HEAD mode (BN = 128 = 4 phases x 32 channels, one accumulator = one low-res tile).
// Thread = low-res pixel (y,x); accumulator columns [32*ph, 32*ph+32) = the 32 channels of hi-res pixel (2y+a, 2x+b),
// ph = 2a+b.  act = LeakyReLU(acc + bias) stays in fp32 registers; the head filter tap (ky,kx) carries it to output
// pixel (2y+a-ky+1, 2x+b-kx+1) = entry [(a-ky+2)*4 + (b-kx+2)] of this pixel's 4x4 patch (origin (2y-1, 2x-1)).
// 4 x 288 MACs per thread, filter read from the constant bank as FFMA operands.
// Channel-chunk-major: all four phases of a 4-channel chunk are in registers at once, so every filter value
// fetched from the constant bank (LDCU.128 = two FFMA2 operands) feeds four FFMA2s instead of one, and the products
// accumulate straight into the sixteen patch entries as fp32x2 (even / odd channel) partial sums -- no per-phase tap
// sums, no scatter.  Per tile and thread: 576 FFMA2 + 72 LDCU.128 (was 576 + 288) and 16 final adds (was 144 + 36).
// The earlier phase-major version (one phase per iteration, 9 tap sums scattered into the patch) executed ~1750 warp
// instructions per tile at 70 % issue utilisation against 1152-1460 tensor-pipe cycles of the tile's MMAs
// (profiles/r01k_conv_full.md): the layer was bound by this epilogue's issue slots.
// The same loop with 1152 scalar FFMAs (uniform-register filter operand) instead of 576 FFMA2 is SLOWER (1.28 vs 1.07 ms,
// profiles/r02f_head_scalar_vs_packed.txt): a 3-operand FFMA occupies the fp32 pipe for two cycles per warp just like an
// FFMA2, so the packed form is the pipe's full rate and 576 x 2 cycles per tile and scheduler is this epilogue's floor.
constexpr int HEAD_EPI_CW = 4;
template <bool TMAST>
__device__ __forceinline__ void conv_epilogue_head_tile(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_acc,
                                                        uint64_t* tmem_empty_bar, const TileCoord& t, const ConvOutMaps* omaps,
                                                        uint8_t* st_set, uint32_t* st_k, bool st_elect, int eset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row >> 3, px = row & 7;
    const int y = t.y0 + py, x = t.x0 + px;
    const bool inb = (y < p.H) && (x < p.W);
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const float2 slope2 = make_float2(neg_slope, neg_slope);
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    float2 hp2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) hp2[i] = make_float2(0.f, 0.f);
    constexpr int CW = HEAD_EPI_CW;              // channels per chunk (4: 16 + 32 live values, fits the 96-register budget)
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += CW) {
        uint32_t raw[4][CW];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            if (p.debug & 8) {                   // profiling: no TMEM reads
#pragma unroll
                for (int j = 0; j < CW; ++j) raw[ph][j] = 0;
            } else if (CW == 8) tmem_ld_32x32b_x8(t_addr + ph * 32 + c0, reinterpret_cast<uint32_t(&)[8]>(raw[ph]));
            else tmem_ld_32x32b_x4(t_addr + ph * 32 + c0, reinterpret_cast<uint32_t(&)[4]>(raw[ph]));
        }
        if (!(p.debug & 8)) tmem_ld_wait();
        if (c0 == 32 - CW) {                     // accumulator fully read: one elected arrive per warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
        float2 bb[CW / 2];                       // phase blocks share the 32 biases
#pragma unroll
        for (int j4 = 0; j4 < CW / 4; ++j4) {
            const float4 b = lds_f4(bars.s_bias + c0 + 4 * j4);
            bb[2 * j4] = make_float2(b.x, b.y);
            bb[2 * j4 + 1] = make_float2(b.z, b.w);
        }
        float2 v[4][CW / 2];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
#pragma unroll
            for (int j2 = 0; j2 < CW / 2; ++j2) {
                const float2 a = __fadd2_rn(make_float2(__uint_as_float(raw[ph][2 * j2]), __uint_as_float(raw[ph][2 * j2 + 1])),
                                            bb[j2]);
                const float2 s = __fmul2_rn(a, slope2);
                v[ph][j2] = make_float2(fmaxf(a.x, s.x), fmaxf(a.y, s.y));
            }
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int j2 = 0; j2 < CW / 2; ++j2) {
                const float2 w = make_float2(p.head_wc[tap * 32 + c0 + 2 * j2], p.head_wc[tap * 32 + c0 + 2 * j2 + 1]);
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const int e = ((ph >> 1) - tap / 3 + 2) * 4 + ((ph & 1) - tap % 3 + 2);
                    hp2[e] = __ffma2_rn(v[ph][j2], w, hp2[e]);
                }
            }
        }
    }
    if (TMAST) {
        // the tile's 128 patches (64 bytes each) through the set's staging buffer and ONE TMA store (see conv_epilogue_lean):
        // the per-thread stores cost this layer 14 % (1.073 -> 0.918 ms without them, profiles/r01u)
        const uint32_t k = *st_k;
        uint8_t* buf = st_set + (p.st_bufs == 2 ? (k & 1u) : 0u) * p.st_bytes;
        if (st_elect) {
            if (p.st_bufs == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
        }
        named_bar_sync(CONV_ST_BAR0 + eset, 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = make_float4(hp2[4 * i].x + hp2[4 * i].y, hp2[4 * i + 1].x + hp2[4 * i + 1].y,
                                         hp2[4 * i + 2].x + hp2[4 * i + 2].y, hp2[4 * i + 3].x + hp2[4 * i + 3].y);
            sts_u4(buf + row * 64 + 16 * (i ^ ((row >> 1) & 3)),      // CU_TENSOR_MAP_SWIZZLE_64B (buffer 1024-byte aligned): conflict-free
                   make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w)));
        }
        fence_proxy_async();
        named_bar_sync(CONV_ST_BAR0 + eset, 128);
        if (st_elect && !(p.debug & 4)) {
            tma_store_4d(&omaps->m[0], buf, 0, t.x0, t.y0, t.n);
            bulk_commit_group();
        }
        *st_k = k + 1;
    } else if (inb && !(p.debug & 4)) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) +
                                              ((static_cast<size_t>(t.n) * p.H + y) * p.W + x) * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            o[i] = make_float4(hp2[4 * i].x + hp2[4 * i].y, hp2[4 * i + 1].x + hp2[4 * i + 1].y,
                               hp2[4 * i + 2].x + hp2[4 * i + 2].y, hp2[4 * i + 3].x + hp2[4 * i + 3].y);
    }
}

// OUT_SHUFFLE2_HEAD_MMA: the head conv of a tile on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate).
// The CUDA-core versions above need 1152 fp32 MACs per low-res pixel -- ~2300 FMA-pipe cycles per tile and scheduler,
// more than the 1152-1460 tensor-pipe cycles of the tile's own MMAs, so dec.12 ran at the speed of its epilogue
// (profiles/r01s).  Here the fp32 pipe only does bias + LeakyReLU + rounding (as every other layer's output is rounded).
//   A  = act[16 px x 32 ch] of one phase, straight from the accumulator registers (C-fragment == A-fragment distribution)
//   B  = Bp[ph][ch][e] = head filter tap (ky,kx) of channel ch if patch entry e = (a-ky+2)*4 + (b-kx+2) (ph = 2a+b), else 0
//   D  = patch[16 px x 16 entries] summed over the four phases; thread (g,t) ends up with entries 8nt + 2t + {0,1} of
//        pixels g and g+8 -> 8-byte stores, a quad writes one 32-byte sector.
// s_bfrag: shared [4 ph][2 k-steps][2 n-tiles][32 lanes] uint2 = per-lane B fragments (head_mma_fill_bfrag).
__device__ __forceinline__ void head_mma_fill_bfrag(const ConvParams& p, uint2* s_bfrag, int idx /* 0..511 */) {
    const int lane = idx & 31, combo = idx >> 5;                 // combo = (ph * 2 + s) * 2 + nt
    const int nt = combo & 1, ks = (combo >> 1) & 1, ph = combo >> 2;
    const int g = lane >> 2, tq = lane & 3;
    const int e = nt * 8 + g;
    const int ky = (ph >> 1) + 2 - (e >> 2), kx = (ph & 1) + 2 - (e & 3);
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (ky >= 0 && ky <= 2 && kx >= 0 && kx <= 2) {
        const float* wt = p.head_wc + (ky * 3 + kx) * 32 + ks * 16 + 2 * tq;
        w[0] = wt[0]; w[1] = wt[1]; w[2] = wt[8]; w[3] = wt[9];
    }
    s_bfrag[idx] = make_uint2(pack2(w[0], w[1], p.fp16), pack2(w[2], w[3], p.fp16));
}

__device__ __forceinline__ void conv_epilogue_head_mma_tile(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_acc,
                                                            uint64_t* tmem_empty_bar, const TileCoord& t,
                                                            const uint2* s_bfrag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, g = lane >> 2, tq = lane & 3;
    const int fp16 = p.fp16;
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const float2 slope2 = make_float2(neg_slope, neg_slope);
    float2 bb[4];                                                // biases of this thread's channel pairs 8j + 2t + {0,1}
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bb[j].x = bars.s_bias[8 * j + 2 * tq];
        bb[j].y = bars.s_bias[8 * j + 2 * tq + 1];
    }
    const uint32_t sb = smem_u32(s_bfrag) + lane * 8;
    const int x = t.x0 + g;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {                             // two blocks of 16 pixels (TMEM lanes) per warp
        float d[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[0][i] = d[1][i] = 0.f;
        const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32 + hb * 16) << 16);
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            uint32_t raw[16];
            if (p.debug & 8) {                                   // profiling: no TMEM reads
#pragma unroll
                for (int j = 0; j < 16; ++j) raw[j] = 0;
            } else {
                tmem_ld_16x256b_x4(t_addr + ph * 32, raw);
                tmem_ld_wait();
            }
            if (hb == 1 && ph == 3) {                            // accumulator fully read: one elected arrive per warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar);
            }
            uint32_t lo[4], hi[4];                               // 16-bit pairs of pixel g / pixel g + 8
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = __fadd2_rn(make_float2(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1])), bb[j]);
                const float2 b = __fadd2_rn(make_float2(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])), bb[j]);
                const float2 sa = __fmul2_rn(a, slope2), sb2 = __fmul2_rn(b, slope2);
                lo[j] = pack2(fmaxf(a.x, sa.x), fmaxf(a.y, sa.y), fp16);
                hi[j] = pack2(fmaxf(b.x, sb2.x), fmaxf(b.y, sb2.y), fp16);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    uint32_t b0, b1;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(sb + ((ph * 2 + ks) * 2 + nt) * 256));
                    mma_m16n8k16(d[nt], lo[2 * ks], hi[2 * ks], lo[2 * ks + 1], hi[2 * ks + 1], b0, b1, fp16);
                }
            }
        }
        const int y = t.y0 + q * 4 + hb * 2;                     // tile row of pixel g (lane = 8 * row + column); g + 8 = next row
        if (x < p.W && !(p.debug & 4)) {
            float* o = static_cast<float*>(p.out) + ((static_cast<size_t>(t.n) * p.H + y) * p.W + x) * 16 + 2 * tq;
            if (y < p.H) {
                *reinterpret_cast<float2*>(o) = make_float2(d[0][0], d[0][1]);
                *reinterpret_cast<float2*>(o + 8) = make_float2(d[1][0], d[1][1]);
            }
            if (y + 1 < p.H) {
                o += static_cast<size_t>(p.W) * 16;
                *reinterpret_cast<float2*>(o) = make_float2(d[0][2], d[0][3]);
                *reinterpret_cast<float2*>(o + 8) = make_float2(d[1][2], d[1][3]);
            }
        }
    }
}

// OUT_SHUFFLE2_HEAD_TC: same result layout as conv_epilogue_head_tile, but the 32 -> 1 head conv (4 x 288 MACs per low-res
// pixel) runs on the tensor cores.  The CUDA-core version needed ~1650 warp instructions per tile and scheduler --
// more than the 1152 tensor-pipe cycles of the tile's main MMAs (ncu: 56 % issue utilisation, epilogue-bound).
//   phase A  (per phase ph): acc columns [32ph, 32ph+32) -> + bias, LeakyReLU -> 16-bit pairs -> tcgen05.st into
//            columns [32ph, 32ph+16) of the SAME accumulator (K-major A operand: lane = pixel, column j = channels 2j, 2j+1)
//   MMA      one elected thread of the set: D2_ph[128 x 16] = A_ph[128 x 32] * Whead[32 x 16]  (2 K-steps, A from TMEM,
//            B = the [16 taps (9 used)][32] head filter in shared memory) into columns [32ph+16, 32ph+32); commit -> head_bar
//   phase B  D2_ph column tap = the head-conv contribution of hi-res pixel (2y+a, 2x+b) through tap (ky,kx): scatter
//            into the 4x4 patch like the CUDA-core version.
// The activations are rounded to 16 bit here (like every other layer); the head filter is 16-bit.
__device__ __forceinline__ void conv_epilogue_head_tc_tile(const ConvParams& p, const ConvBarriers& bars, uint32_t tmem_acc,
                                                           uint64_t* tmem_empty_bar, uint64_t* head_bar,
                                                           uint32_t head_parity, uint32_t hdesc_lo, uint32_t hdesc_hi,
                                                           int eset, const TileCoord& t) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row >> 3, px = row & 7;
    const int y = t.y0 + py, x = t.x0 + px;
    const bool inb = (y < p.H) && (x < p.W);
    const int fp16 = p.fp16;
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const float2 slope2 = make_float2(neg_slope, neg_slope);
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int ph = 0; ph < 4; ++ph) {
        uint32_t pk[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(t_addr + ph * 32 + half * 16, raw);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b = lds_f4(bars.s_bias + half * 16 + 4 * j4);       // phase blocks share the 32 biases
                const float2 a0 = __fadd2_rn(make_float2(__uint_as_float(raw[j4 * 4 + 0]), __uint_as_float(raw[j4 * 4 + 1])),
                                             make_float2(b.x, b.y));
                const float2 a1 = __fadd2_rn(make_float2(__uint_as_float(raw[j4 * 4 + 2]), __uint_as_float(raw[j4 * 4 + 3])),
                                             make_float2(b.z, b.w));
                const float2 s0 = __fmul2_rn(a0, slope2), s1 = __fmul2_rn(a1, slope2);
                pk[half * 8 + j4 * 2 + 0] = pack2(fmaxf(a0.x, s0.x), fmaxf(a0.y, s0.y), fp16);
                pk[half * 8 + j4 * 2 + 1] = pack2(fmaxf(a1.x, s1.x), fmaxf(a1.y, s1.y), fp16);
            }
        }
        tmem_st_32x32b_x16(t_addr + ph * 32, pk);     // both halves of the phase are in registers: in-place is safe
    }
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1 + eset, 128);                    // all 128 pixels (lanes) of the tile have their A rows in TMEM
    if (q == 0) {
        tc_fence_after();
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_16(CONV_TILE_M, 16, fp16);
#pragma unroll
            for (int ph = 0; ph < 4; ++ph)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16_ts(tmem_acc + ph * 32 + 16, tmem_acc + ph * 32 + 8 * k, hdesc_lo + 2 * k, hdesc_hi, idesc,
                                static_cast<uint32_t>(k));
            umma_commit(head_bar);
        }
        __syncwarp();
    }
    mbar_wait(head_bar, head_parity);
    tc_fence_after();
    float hp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) hp[i] = 0.f;
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        uint32_t raw[16];
        tmem_ld_32x32b_x16(t_addr + ph * 32 + 16, raw);
        tmem_ld_wait();
        if (ph == 3) {                                 // accumulator fully consumed: one elected arrive per warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
        const int a = ph >> 1, b = ph & 1;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) hp[(a - tap / 3 + 2) * 4 + (b - tap % 3 + 2)] += __uint_as_float(raw[tap]);
    }
    if (inb) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) +
                                              ((static_cast<size_t>(t.n) * p.H + y) * p.W + x) * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(hp[4 * i], hp[4 * i + 1], hp[4 * i + 2], hp[4 * i + 3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// halo + resident-filter kernel
//
// Work quantum = a SUPER-TILE: T (4, 2 or 1) M-tiles of 16x8 pixels stacked vertically, fetched by ONE TMA box
// {KC, 10, 16T+2, 1} per K-chunk and accumulated into T TMEM accumulators that are handed to the epilogue with ONE
// commit.  Measured on B200 (tools/conv_debug_sweep.sh, tools/conv_n_sweep.sh): with one M-tile per pipeline step the
// mbarrier hand-offs alone (producer -> MMA -> epilogue -> MMA, everything else stubbed out) cost ~510 cycles per tile
// and do not overlap with the 720 cycles of a 32-channel tile's MMAs; a super-tile pays them once per T tiles.
//   M-tile t of a stage starts 160 pixel rows (16 x pitch 10) further: same row-shifted descriptors as the taps.
//   TMEM: nbuf (2..4) buffers x T accumulators x BN columns (<= 512); super-tile i uses buffer i % nbuf.
//   Epilogue sets (4 x 4 warps): M-tile t of super-tile i belongs to set (i*T + t) & 3 (round-robin in issue order).
// ---------------------------------------------------------------------------------------------------------------
template <int KC>
struct HaloSmem {
    static constexpr int ROW_BYTES = KC * 2;
    __host__ __device__ static constexpr int a_bytes(int T) { return (CONV_TILE_H * T + 2) * HALO_W * ROW_BYTES; }
    __host__ __device__ static constexpr int a_stage(int T) { return ((a_bytes(T) + 1023) / 1024) * 1024; }
    __host__ __device__ static constexpr int b_block(int BN) { return BN * ROW_BYTES; }   // one (tap, chunk) block
    __host__ __device__ static constexpr int b_bytes(int BN, int Cin) { return 9 * (Cin / KC) * b_block(BN); }
    __host__ __device__ static constexpr int total_bytes(int BN, int Cin, int T, int stages) {
        return 1024 + b_bytes(BN, Cin) + stages * a_stage(T) + CONV_TAIL_BYTES;
    }
};
constexpr int HEAD_SMEM_BYTES = 1024;          // 16 rows x 64 bytes, SW64
constexpr int HEAD_MMA_SMEM_BYTES = 4096;      // [4 ph][2 k-steps][2 n-tiles][32 lanes] uint2 B fragments
__host__ __device__ constexpr uint32_t halo_tmem_cols(int T, int BN, int nbuf) {
    const int c = nbuf * T * BN;
    return (c <= 32) ? 32 : (c <= 64) ? 64 : (c <= 128) ? 128 : (c <= 256) ? 256 : 512;
}

template <int KC, int MODE, bool TMAST = false>   // MODE: -1 = run-time epilogue, OUT_* = that output stage compiled in; TMAST:
__global__ void __launch_bounds__(CONV_THREADS, 1)  // staged TMA stores (OUT_SAME / OUT_SHUFFLE2 / OUT_SHUFFLE2_HEAD, p.st_bufs >= 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ ConvOutMaps omaps,
                    const __grid_constant__ ConvParams p) {
    using S = HaloSmem<KC>;
    constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
    constexpr uint32_t ROW_BYTES = S::ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kchunks = p.Cin / KC;
    const int T = p.T;
    const int b_block = S::b_block(p.BN);
    const int a_stage = S::a_stage(T);
    uint8_t* b_smem = smem;                                         // [tap][chunk][BN rows][KC] swizzled
    uint8_t* h_smem = smem + S::b_bytes(p.BN, p.Cin);               // [16 taps][32] head filter, SW64 (head_smem bytes)
    uint8_t* a_smem = h_smem + p.head_smem;                         // [stage][(16T+2)*10 rows][KC] swizzled
    const int num_stages = p.num_stages;
    ConvBarriers bars(a_smem + num_stages * a_stage);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    // epilogue warps attached to one TMEM buffer: T sets of 4 warps (T = 1: one set per buffer)
    const int nbuf = p.nbuf;
    const uint32_t tmem_cols = halo_tmem_cols(T, p.BN, nbuf);
    if (MODE == OUT_SHUFFLE2_HEAD_MMA && threadIdx.x >= 32 * CONV_FIRST_EPI_WARP)      // 512 epilogue threads, one entry each
        head_mma_fill_bfrag(p, reinterpret_cast<uint2*>(h_smem), threadIdx.x - 32 * CONV_FIRST_EPI_WARP);
    const uint32_t tmem_base = conv_prologue(p, bars, num_stages, tmem_cols, 4 * T);

    // CTA -> (n-block, first super-tile, stride): the grid is split evenly between the n-blocks.
    const int st_per_img = p.tiles_x * p.stiles_y;
    const int st_total = p.N * st_per_img;
    const int ctas_per_nb = gridDim.x / p.n_blocks;
    const int nb = blockIdx.x / ctas_per_nb;
    const int first = blockIdx.x - nb * ctas_per_nb;
    const bool active = nb < p.n_blocks;

    if (warp == 0) {
        // TMA producer: the whole warp runs the (uniform) loop, one elected lane issues
        if (active) {
            if (elect_one_sync()) {
                // resident filter bank of this n-block: 9 * kchunks TMA boxes {KC, BN} on one barrier
                mbar_arrive_expect_tx(bars.b_full, 9 * kchunks * b_block + (MODE == OUT_SHUFFLE2_HEAD_TC ? p.head_smem : 0));
                if (MODE == OUT_SHUFFLE2_HEAD_TC) tma_load_2d(h_smem, &tmap_h, bars.b_full, 0, 0);
                for (int tap = 0; tap < 9; ++tap)
                    for (int kc = 0; kc < kchunks; ++kc)
                        tma_load_2d(b_smem + (tap * kchunks + kc) * b_block, &tmap_w, bars.b_full, kc * KC,
                                    tap * p.Cout + nb * p.BN);
            }
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t a_bytes = S::a_bytes(T);
            // super-tile coordinates advance incrementally (no div/mod per step)
            int n = first / st_per_img, sy = (first % st_per_img) / p.tiles_x, tx = first % p.tiles_x;
            const int dn = ctas_per_nb / st_per_img, dsy = (ctas_per_nb % st_per_img) / p.tiles_x,
                      dtx = ctas_per_nb % p.tiles_x;
            for (int st = first; st < st_total; st += ctas_per_nb) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    if (p.debug & 256) mbar_wait_parked(&bars.empty[stage], phase ^ 1, 20000);
                    else mbar_wait(&bars.empty[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        if (p.debug & 2) {
                            mbar_arrive(&bars.full[stage]);
                        } else {
                            mbar_arrive_expect_tx(&bars.full[stage], a_bytes);
                            tma_load_4d(a_smem + stage * a_stage, &tmap_x, &bars.full[stage], kc * KC,
                                        tx * CONV_TILE_W - 1, sy * T * CONV_TILE_H - 1, n);
                        }
                    }
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                tx += dtx;
                if (tx >= p.tiles_x) { tx -= p.tiles_x; ++sy; }
                sy += dsy;
                if (sy >= p.stiles_y) { sy -= p.stiles_y; ++n; }
                n += dn;
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform loop, every tcgen05 instruction under elect.sync (see elect_one_sync)
        if (active) {
            const uint32_t idesc = make_idesc_16(CONV_TILE_M, p.BN, p.fp16);
            int stage = 0;
            uint32_t phase = 0;
            int buf = 0;
            uint32_t buf_phase = 0;
            mbar_wait(bars.b_full, 0);
            // Descriptors: only the 14-bit start-address field (bits 0..13 of the low word, address >> 4) changes
            // between MMAs, so each MMA costs one 32-bit add per operand.
            //   A rows: group g (output row g of M-tile t) starts at halo pixel (16 t + g + dy) * 10 + dx  =>  start
            //   address advanced by (160 t + dy*10 + dx) rows, 8-row group stride (SBO) = halo pitch.
            const uint64_t a_tmpl = make_smem_desc(smem_u32(a_smem), HALO_W * ROW_BYTES, LAYOUT);
            const uint64_t b_tmpl = make_smem_desc(smem_u32(b_smem), 8 * ROW_BYTES, LAYOUT);
            const uint32_t a_hi = static_cast<uint32_t>(a_tmpl >> 32), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
            const uint32_t a_lo0 = static_cast<uint32_t>(a_tmpl), b_lo0 = static_cast<uint32_t>(b_tmpl);
            const uint32_t b_blk16 = static_cast<uint32_t>(b_block) >> 4;
            const uint32_t b_tap16 = b_blk16 * kchunks;
            constexpr uint32_t kRow16 = ROW_BYTES >> 4;
            constexpr uint32_t kTile16 = CONV_TILE_H * HALO_W * kRow16;
            int sy = (first % st_per_img) / p.tiles_x, tx = first % p.tiles_x;
            const int dsy = (ctas_per_nb % st_per_img) / p.tiles_x, dtx = ctas_per_nb % p.tiles_x;
            for (int st = first; st < st_total; st += ctas_per_nb) {
                const int rows_left = p.tiles_y - sy * T;
                const int nvalid = rows_left < T ? rows_left : T;          // M-tiles of this super-tile inside the image
                mbar_wait(&bars.tmem_empty[buf], buf_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem0 = tmem_base + buf * T * p.BN;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bars.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_stage_lo = a_lo0 + stage * (static_cast<uint32_t>(a_stage) >> 4);
                    const uint32_t b_kc = b_lo0 + kc * b_blk16;
                    if (!(p.debug & 32)) {
                        for (int t = 0; t < nvalid; ++t) {
                            const uint32_t d_tmem = d_tmem0 + t * p.BN;
                            const uint32_t a_lo = a_stage_lo + t * kTile16;
                            if (elect_one_sync()) {
                                if (KC == 64) {
                                    // filter rows rolled, the 3 x 4 MMAs of a row unrolled: with all 9 taps unrolled the 2 x 36
                                    // descriptor pairs exceed the 63 uniform registers and ptxas spills them through vector registers
                                    // (156 MOV.SPILL / R2UR.FILL and 176 R2UR between the UTCHMMAs of a tile).  Same box: 64->64
                                    // pooled 210 -> 183 us, 128->128 274 -> 214 us, dec.2 / dec.6 -3 % (profiles/r09_issuer_loop_ab.txt);
                                    // the KC = 32 loop (18 pairs, 9 spill moves) is 2-3 % FASTER fully unrolled and stays so.
                                    uint32_t b_lo = b_kc, a_row = a_lo;
#pragma unroll 1
                                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                                        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                                            for (int k = 0; k < KC / 16; ++k)
                                                umma_f16_split(d_tmem, a_row + dx * kRow16 + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc,
                                                               (dy | dx | k) != 0 ? 1u : static_cast<uint32_t>(kc != 0));
                                            b_lo += b_tap16;
                                        }
                                        a_row += HALO_W * kRow16;
                                    }
                                } else {
                                    uint32_t b_lo = b_kc;
#pragma unroll
                                    for (int tap = 0; tap < 9; ++tap) {
                                        const uint32_t a_tap = a_lo + ((tap / 3) * HALO_W + (tap % 3)) * kRow16;
#pragma unroll
                                        for (int k = 0; k < KC / 16; ++k)
                                            umma_f16_split(d_tmem, a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc,
                                                           (tap | k) != 0 ? 1u : static_cast<uint32_t>(kc != 0));
                                        b_lo += b_tap16;
                                    }
                                }
                            }
                            __syncwarp();
                        }
                    }
                    if (elect_one_sync()) umma_commit(&bars.empty[stage]);
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) umma_commit(&bars.tmem_full[buf]);
                __syncwarp();
                if (++buf == nbuf) { buf = 0; buf_phase ^= 1; }
                tx += dtx;
                if (tx >= p.tiles_x) { tx -= p.tiles_x; ++sy; }
                sy += dsy;
                if (sy >= p.stiles_y) sy -= p.stiles_y;
            }
        }
    } else if (active) {
        const int eset = (warp - CONV_FIRST_EPI_WARP) >> 2;
        // M-tiles are dealt to the four epilogue sets round-robin in issue order: M-tile t of super-tile i (running index
        // i*T + t) belongs to set (i*T + t) % nsets.  T = 4: set e <-> M-tile e of every super-tile; T = 2: set e <-> M-tile
        // e & 1 of the super-tiles with parity e >> 1; T = 1: set e <-> every 4th super-tile.  Super-tile i lives in TMEM
        // buffer i % nbuf (its (i / nbuf)-th use), so with nbuf = 4 the issuer runs up to three super-tiles ahead of
        // the slowest epilogue (measured: with two buffers the store-heavy epilogues of the 64/128-column layers stalled
        // the MMAs, profiles/r01f_debug_sweep_large_n.txt).
        // A set must see EVERY use of a buffer it touches (an mbarrier parity wait cannot skip a phase), hence only
        // min(4, nbuf * T) sets take part: T = 1 with two buffers runs on sets 0 / 1.
        const int nsets = (nbuf * T < CONV_EPI_SETS) ? nbuf * T : CONV_EPI_SETS;       // 2 or 4
        int i = 0, buf = 0;
        uint32_t buf_phase = 0, head_uses = 0;
        // TMA-store staging of this set (behind the tail), the set's issuing thread, its running chunk count
        uint8_t* st_set = a_smem + num_stages * a_stage + ((CONV_TAIL_BYTES + 1023) & ~1023) + eset * p.st_bufs * p.st_bytes;
        const bool st_elect = ((warp - CONV_FIRST_EPI_WARP) & 3) == 0 && lane == 0;
        uint32_t st_k = 0;
        const uint64_t h_tmpl = make_smem_desc(smem_u32(h_smem), 8 * 64, UMMA_LAYOUT_SW64);
        const uint32_t hdesc_lo = static_cast<uint32_t>(h_tmpl), hdesc_hi = static_cast<uint32_t>(h_tmpl >> 32);
        if (MODE == OUT_SHUFFLE2_HEAD_TC) mbar_wait(bars.b_full, 0);       // the head filter block landed with the bank
        for (int st = first; st < st_total; st += ctas_per_nb, ++i) {
            const int t = (eset - i * T) & (nsets - 1);
            if (eset < nsets && t < T) {
                const int n = st / st_per_img;
                const int r = st - n * st_per_img;
                const int sy = r / p.tiles_x;
                TileCoord tc;
                tc.n = n;
                tc.y0 = (sy * T + t) * CONV_TILE_H;
                tc.x0 = (r - sy * p.tiles_x) * CONV_TILE_W;
                tc.n0 = nb * p.BN;
                if (p.debug & 256) mbar_wait_parked(&bars.tmem_full[buf], buf_phase, 20000);
                else mbar_wait(&bars.tmem_full[buf], buf_phase);
                tc_fence_after();
                if (sy * T + t < p.tiles_y) {
                    const uint32_t acc = tmem_base + (buf * T + t) * p.BN;
                    if (MODE == OUT_SHUFFLE2_HEAD_TC) {
                        conv_epilogue_head_tc_tile(p, bars, acc, &bars.tmem_empty[buf], &bars.head_bar[eset], head_uses & 1,
                                                   hdesc_lo, hdesc_hi, eset, tc);
                        ++head_uses;
                    } else if (MODE == OUT_SHUFFLE2_HEAD_MMA) {
                        conv_epilogue_head_mma_tile(p, bars, acc, &bars.tmem_empty[buf], tc, reinterpret_cast<const uint2*>(h_smem));
                    } else if (MODE == OUT_SHUFFLE2_HEAD) {
                        conv_epilogue_head_tile<TMAST>(p, bars, acc, &bars.tmem_empty[buf], tc, &omaps, st_set, &st_k, st_elect, eset);
                    }
                    else if (MODE >= 0) conv_epilogue_lean<MODE, TMAST>(p, bars, acc, &bars.tmem_empty[buf], tc, &omaps, st_set, &st_k, st_elect, eset);
                    else conv_epilogue_tile(p, bars, acc, &bars.tmem_empty[buf], tc);
                } else {                       // M-tile below the image: nothing to read, release the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.tmem_empty[buf]);
                }
            }
            if (++buf == nbuf) { buf = 0; buf_phase ^= 1; }
        }
        if (TMAST && st_elect) bulk_wait_group_all();             // stores issued by this thread
    }
    conv_teardown<(MODE < 0) || lean_stats(MODE)>(p, bars, tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// streamed kernel: one K-block = (tap, KC-channel chunk), A and B both through the stage ring
// ---------------------------------------------------------------------------------------------------------------
template <int KC>
struct StreamSmem {
    static constexpr int A_BYTES = CONV_TILE_M * KC * 2;
    __host__ __device__ static constexpr int b_bytes(int BN) { return ((BN * KC * 2 + 1023) / 1024) * 1024; }
    __host__ __device__ static constexpr int stage_bytes(int BN) { return A_BYTES + b_bytes(BN); }
    __host__ __device__ static constexpr int total_bytes(int BN, int stages) {
        return 1024 + stages * stage_bytes(BN) + CONV_TAIL_BYTES;
    }
};

template <int KC, int MODE>        // MODE: -1 = run-time epilogue, else a compiled-in lean epilogue (OUT_SAME, OUT_SAME_MAXPOOL2, ConvLean)
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ ConvParams p) {
    using S = StreamSmem<KC>;
    constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
    constexpr uint32_t SBO = 8 * KC * 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = S::stage_bytes(p.BN);
    const int num_stages = p.num_stages;
    ConvBarriers bars(smem + num_stages * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    const uint32_t tmem_base = conv_prologue(p, bars, num_stages, conv_tmem_cols(p.BN), 4);
    const int kchunks = p.Cin / KC;
    const int KB = 9 * kchunks;

    // work item = (tile, K split): item / ksplit = tile, item % ksplit = split; ksplit = 1: the whole K range per tile
    const int ksplit = p.ksplit;
    const int num_items = p.num_tiles * ksplit;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int tile = item / ksplit, sp = item - tile * ksplit;
                const TileCoord t = tile_coord(p, tile / p.n_blocks, tile % p.n_blocks);
                const int kb0 = sp * KB / ksplit, kb1 = (sp + 1) * KB / ksplit;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tap = kb / kchunks, kc = kb - tap * kchunks;
                    const int dy = tap / 3, dx = tap - dy * 3;
                    mbar_wait(&bars.empty[stage], phase ^ 1);
                    uint8_t* a_dst = smem + stage * stage_bytes;
                    mbar_arrive_expect_tx(&bars.full[stage], S::A_BYTES + p.BN * KC * 2);
                    tma_load_4d(a_dst, &tmap_x, &bars.full[stage], kc * KC, t.x0 + dx - 1, t.y0 + dy - 1, t.n);
                    tma_load_2d(a_dst + S::A_BYTES, &tmap_w, &bars.full[stage], kc * KC, tap * p.Cout + t.n0);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_16(CONV_TILE_M, p.BN, p.fp16);
            int stage = 0;
            uint32_t phase = 0;
            const int num_acc = conv_num_acc(p.BN);
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t s_tmpl = make_smem_desc(smem_u32(smem), SBO, LAYOUT);
            const uint32_t s_hi = static_cast<uint32_t>(s_tmpl >> 32), s_lo0 = static_cast<uint32_t>(s_tmpl);
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int sp = item % ksplit;
                const int kb0 = sp * KB / ksplit, kb1 = (sp + 1) * KB / ksplit;
                mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bars.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = s_lo0 + stage * (static_cast<uint32_t>(stage_bytes) >> 4);
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)   // +32 bytes along K inside the swizzle row
                        umma_f16_split(d_tmem, a_lo + 2 * k, s_hi, a_lo + (S::A_BYTES >> 4) + 2 * k, s_hi, idesc,
                                       k != 0 ? 1u : static_cast<uint32_t>(kb != kb0));
                    umma_commit(&bars.empty[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars.tmem_full[acc]);
                if (++acc == num_acc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= CONV_FIRST_EPI_WARP) {
        const int num_acc = conv_num_acc(p.BN);
        const int eset = (warp - CONV_FIRST_EPI_WARP) >> 2;
        if (eset < num_acc) {
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x + eset * gridDim.x; item < num_items; item += gridDim.x * num_acc) {
                const int tile = item / ksplit, sp = item - tile * ksplit;
                const TileCoord t = tile_coord(p, tile / p.n_blocks, tile % p.n_blocks);
                mbar_wait(&bars.tmem_full[eset], acc_phase);
                tc_fence_after();
                if (ksplit > 1) conv_epilogue_splitk(p, tmem_base + eset * p.BN, &bars.tmem_empty[eset], t, sp);
                else if (MODE >= 0) conv_epilogue_lean<MODE>(p, bars, tmem_base + eset * p.BN, &bars.tmem_empty[eset], t);
                else conv_epilogue_tile(p, bars, tmem_base + eset * p.BN, &bars.tmem_empty[eset], t);
                acc_phase ^= 1;
            }
        }
    }
    conv_teardown<true>(p, bars, tmem_base, conv_tmem_cols(p.BN));
}

}  // namespace aesr
