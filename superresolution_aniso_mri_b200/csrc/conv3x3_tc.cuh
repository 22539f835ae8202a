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
};
enum ConvMul : int { MUL_NONE = 0, MUL_LEAKY_GRAD = 1, MUL_RELU_GRAD = 2 };

struct ConvParams {
    int N, H, W, Cin, Cout;
    int BN;                 // output channels per tile (multiple of 32, <= 256, divides Cout)
    int tiles_x, tiles_y, n_blocks, num_tiles;   // num_tiles = spatial tiles (N * tiles_y * tiles_x) * n_blocks
    int num_stages;
    int num_issuers;        // MMA-issuing threads in use (halo kernel): 2 when the stage ring is deeper than one tile's
                            // K-chunks, else 1 (an issuer one full ring ahead would alias the mbarrier phase parity)
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
    // optional per-channel statistics of the (post-activation, pre-affine) value: stats[c] += sum, stats[Cout+c] += sum^2
    float* stats;
};

constexpr int CONV_TILE_H = 16;
constexpr int CONV_TILE_W = 8;
constexpr int CONV_TILE_M = 128;
constexpr int CONV_EPI_SETS = 4;                          // epilogue warp sets (4 warps = 128 TMEM lanes each)
constexpr int CONV_ISSUERS = 2;                            // MMA-issuing threads (warps 1..CONV_ISSUERS)
constexpr int CONV_FIRST_EPI_WARP = 1 + CONV_ISSUERS;      // warp 0 TMA, warps 1-2 MMA issuers, then the epilogue sets
constexpr int CONV_THREADS = 32 * CONV_FIRST_EPI_WARP + 128 * CONV_EPI_SETS;   // 608
constexpr int CONV_MAX_STAGES = 8;
// TMEM accumulators in flight: one per epilogue set while they fit in the 512 columns
__host__ __device__ constexpr int conv_num_acc(int BN) { return (CONV_EPI_SETS * BN <= 512) ? CONV_EPI_SETS : 512 / BN; }
constexpr int HALO_H = CONV_TILE_H + 2;     // 18
constexpr int HALO_W = CONV_TILE_W + 2;     // 10
constexpr int CONV_TAIL_BYTES = 256 + 3 * 512 * 4;   // barriers + tmem ptr + per-channel epilogue constants

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
    uint32_t* tmem_ptr;
    float *s_bias, *s_scale, *s_shift;
    __device__ explicit ConvBarriers(uint8_t* tail) {
        full = reinterpret_cast<uint64_t*>(tail);
        empty = full + CONV_MAX_STAGES;
        tmem_full = empty + CONV_MAX_STAGES;
        tmem_empty = tmem_full + CONV_EPI_SETS;
        b_full = tmem_empty + CONV_EPI_SETS;
        tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);
        s_bias = reinterpret_cast<float*>(tail + 256);
        s_scale = s_bias + 512;
        s_shift = s_scale + 512;
    }
};

__device__ __forceinline__ uint32_t conv_tmem_cols(int BN) {
    const int c = conv_num_acc(BN) * BN;
    return (c <= 32) ? 32 : (c <= 64) ? 64 : (c <= 128) ? 128 : (c <= 256) ? 256 : 512;
}

// Common prologue: barrier init, TMEM allocation (warp 1), epilogue constants.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t conv_prologue(const ConvParams& p, const ConvBarriers& bars, int num_stages) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        for (int a = 0; a < CONV_EPI_SETS; ++a) {
            mbar_init(&bars.tmem_full[a], 1);
            mbar_init(&bars.tmem_empty[a], 128);
        }
        mbar_init(bars.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(bars.tmem_ptr, conv_tmem_cols(p.BN));
    for (int c = threadIdx.x; c < p.Cout; c += CONV_THREADS) {
        bars.s_bias[c] = p.bias ? p.bias[c] : 0.f;
        bars.s_scale[c] = p.scale ? p.scale[c] : 1.f;
        bars.s_shift[c] = p.shift ? p.shift[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *bars.tmem_ptr;
}

__device__ __forceinline__ void conv_teardown(const ConvParams& p, uint32_t tmem_base) {
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, conv_tmem_cols(p.BN));
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

// Epilogue of one tile, executed by the 4 epilogue warps (128 threads = 128 TMEM lanes = 128 pixels).
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
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t raw[32];
        if (p.debug & 8) {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = 0;
        } else {
            tmem_ld_32x32b_x32(t_addr + c0, raw);
            tmem_ld_wait();
        }
        if (c0 + 32 >= p.BN) {                  // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(tmem_empty_bar);
        }
        if (p.debug & 64) continue;
        float v[32];
        const int cg = t.n0 + c0;               // first global output channel of this chunk
        {
            // act(a) = max(a, a * neg_slope): neg_slope = 1 (identity), 0.01 (LeakyReLU), 0 (ReLU) -- branch-free
            const float4* b4 = reinterpret_cast<const float4*>(bars.s_bias + cg);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b = b4[j4];
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
                const uint4* m4 = reinterpret_cast<const uint4*>(
                    p.mul_src + (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
                const float neg = (p.mul_mode == MUL_LEAKY_GRAD) ? p.slope : 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
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
            // per-channel sum / sum of squares over the valid pixels of this warp, then one atomic per channel
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float s1 = inb ? v[j] : 0.f;
                float s2 = s1 * s1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == j) {
                    atomicAdd(p.stats + cg + j, s1);
                    atomicAdd(p.stats + Cout + cg + j, s2);
                }
            }
        }
        if (p.scale != nullptr) {
            const float4* sc4 = reinterpret_cast<const float4*>(bars.s_scale + cg);
            const float4* sh4 = reinterpret_cast<const float4*>(bars.s_shift + cg);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 sc = sc4[j4], sh = sh4[j4];
                v[j4 * 4 + 0] = fmaf(v[j4 * 4 + 0], sc.x, sh.x);
                v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], sc.y, sh.y);
                v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], sc.z, sh.z);
                v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], sc.w, sh.w);
            }
        }
        if (p.out_mode == OUT_SAME || p.out_mode == OUT_SAME_MAXPOOL2) {
            if (inb && !((p.debug & 4) && v[0] != 12345.f)) {
                uint4* o4 = reinterpret_cast<uint4*>(
                    out16 + (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) o4[j4] = pack8(v + j4 * 8, fp16);
            }
        }
        if (p.out_mode == OUT_AVGPOOL2 || p.out_mode == OUT_SAME_MAXPOOL2) {
            // 2x2 window = lanes {l, l^1, l^8} (x neighbour, y neighbour): tile origins are even.
            const bool is_max = (p.out_mode == OUT_SAME_MAXPOOL2);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
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
                uint16_t* dst = (p.out_mode == OUT_AVGPOOL2) ? out16 : out2_16;
                uint4* o4 = reinterpret_cast<uint4*>(
                    dst + (static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(yo) * Wo + xo) * Cout + cg);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) o4[j4] = pack8(v + j4 * 8, fp16);
            }
        } else if (p.out_mode == OUT_UP2) {
            if (inb) {
                const int Ho = 2 * H, Wo = 2 * W;
                uint4 pk[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) pk[j4] = pack8(v + j4 * 8, fp16);
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int yo = 2 * y + (d >> 1), xo = 2 * x + (d & 1);
                    uint4* o4 = reinterpret_cast<uint4*>(
                        out16 + (static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(yo) * Wo + xo) * Cout + cg);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) o4[j4] = pk[j4];
                }
            }
        } else if (p.out_mode == OUT_NCHW_F32) {
            if (inb) {
                float* o = static_cast<float*>(p.out) + (static_cast<size_t>(n) * Cout + cg) * H * W +
                           static_cast<size_t>(y) * W + x;
#pragma unroll
                for (int j = 0; j < 32; ++j) o[static_cast<size_t>(j) * H * W] = v[j];
                if (p.out2 != nullptr) {
                    uint4* o4 = reinterpret_cast<uint4*>(
                        out2_16 + (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) o4[j4] = pack8(v + j4 * 8, fp16);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// halo + resident-filter kernel
// ---------------------------------------------------------------------------------------------------------------
template <int KC>
struct HaloSmem {
    static constexpr int ROW_BYTES = KC * 2;
    static constexpr int A_BYTES = HALO_H * HALO_W * ROW_BYTES;                  // 23040 (KC=64) / 11520 (KC=32)
    static constexpr int A_STAGE = ((A_BYTES + 1023) / 1024) * 1024;
    __host__ __device__ static constexpr int b_block(int BN) { return BN * ROW_BYTES; }   // one (tap, chunk) block
    __host__ __device__ static constexpr int b_bytes(int BN, int Cin) { return 9 * (Cin / KC) * b_block(BN); }
    __host__ __device__ static constexpr int total_bytes(int BN, int Cin, int stages) {
        return 1024 + b_bytes(BN, Cin) + stages * A_STAGE + CONV_TAIL_BYTES;
    }
};

template <int KC>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const ConvParams p) {
    using S = HaloSmem<KC>;
    constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
    constexpr uint32_t ROW_BYTES = S::ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kchunks = p.Cin / KC;
    const int b_block = S::b_block(p.BN);
    uint8_t* b_smem = smem;                                         // [tap][chunk][BN rows][KC] swizzled
    uint8_t* a_smem = smem + S::b_bytes(p.BN, p.Cin);               // [stage][18*10 rows][KC] swizzled (1024-aligned)
    const int num_stages = p.num_stages;
    ConvBarriers bars(a_smem + num_stages * S::A_STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    const uint32_t tmem_base = conv_prologue(p, bars, num_stages);

    // CTA -> (n-block, first spatial tile, stride): the grid is split evenly between the n-blocks.
    const int sp_tiles = p.num_tiles / p.n_blocks;
    const int ctas_per_nb = gridDim.x / p.n_blocks;
    const int nb = blockIdx.x / ctas_per_nb;
    const int first = blockIdx.x - nb * ctas_per_nb;
    const bool active = nb < p.n_blocks;

    if (warp == 0) {
        if (lane == 0 && active) {
            // resident filter bank of this n-block: 9 * kchunks TMA boxes {KC, BN} on one barrier
            mbar_arrive_expect_tx(bars.b_full, 9 * kchunks * b_block);
            for (int tap = 0; tap < 9; ++tap)
                for (int kc = 0; kc < kchunks; ++kc)
                    tma_load_2d(b_smem + (tap * kchunks + kc) * b_block, &tmap_w, bars.b_full, kc * KC,
                                tap * p.Cout + nb * p.BN);
            int stage = 0;
            uint32_t phase = 0;
            // tile coordinates advance incrementally (one producer thread: no div/mod per tile)
            const int tiles_per_img = p.tiles_x * p.tiles_y;
            int n = first / tiles_per_img, ty = (first % tiles_per_img) / p.tiles_x, tx = first % p.tiles_x;
            const int dn = ctas_per_nb / tiles_per_img, dty = (ctas_per_nb % tiles_per_img) / p.tiles_x,
                      dtx = ctas_per_nb % p.tiles_x;
            for (int sp = first; sp < sp_tiles; sp += ctas_per_nb) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bars.empty[stage], phase ^ 1);
                    if (p.debug & 2) {
                        mbar_arrive(&bars.full[stage]);
                    } else {
                        mbar_arrive_expect_tx(&bars.full[stage], S::A_BYTES);
                        tma_load_4d(a_smem + stage * S::A_STAGE, &tmap_x, &bars.full[stage], kc * KC,
                                    tx * CONV_TILE_W - 1, ty * CONV_TILE_H - 1, n);
                    }
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                tx += dtx;
                if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
                ty += dty;
                if (ty >= p.tiles_y) { ty -= p.tiles_y; ++n; }
                n += dn;
            }
        }
    } else if (warp < CONV_FIRST_EPI_WARP) {
        // CONV_ISSUERS MMA-issuing threads alternate tiles (issuer j takes this CTA's tiles j, j+2, ...).  A tile's
        // barrier round trip (2 waits, 2 fences, 2 commits: ~760 cycles measured with everything else stubbed out) is
        // serial in the issuing thread and the tensor-core queue is too shallow to cover it, so with ONE issuer a
        // thin-layer tile costs overhead + MMA time (1900 cycles vs ~850 of MMA); with two, one thread feeds the tensor
        // core while the other does its bookkeeping (1400 cycles; four issuers were slower again: register pressure at
        // 672 threads).  Stages and accumulators are assigned by tile index: each thread steps them by two tiles.
        const int issuer = warp - 1;
        const int num_acc = conv_num_acc(p.BN);
        const int n_iss = p.num_issuers;
        if (lane == 0 && active && issuer < n_iss) {
            const uint32_t idesc = make_idesc_16(CONV_TILE_M, p.BN, p.fp16);
            int stage = 0;
            uint32_t phase = 0;
            int acc = issuer % num_acc;
            uint32_t acc_phase = 0;
            for (int i = 0; i < issuer * kchunks; ++i)
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            mbar_wait(bars.b_full, 0);
            // Descriptors: only the 14-bit start-address field (bits 0..13 of the low word, address >> 4) changes
            // between MMAs, so each MMA costs one 32-bit add per operand.
            //   A rows: group g (output row g of the tile) starts at halo pixel (g + dy) * 10 + dx  =>  start address
            //   advanced by (dy*10+dx) rows, 8-row group stride (SBO) = halo pitch.
            const uint64_t a_tmpl = make_smem_desc(smem_u32(a_smem), HALO_W * ROW_BYTES, LAYOUT);
            const uint64_t b_tmpl = make_smem_desc(smem_u32(b_smem), 8 * ROW_BYTES, LAYOUT);
            const uint32_t a_hi = static_cast<uint32_t>(a_tmpl >> 32), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
            const uint32_t a_lo0 = static_cast<uint32_t>(a_tmpl), b_lo0 = static_cast<uint32_t>(b_tmpl);
            const uint32_t b_blk16 = static_cast<uint32_t>(b_block) >> 4;
            const uint32_t b_tap16 = b_blk16 * kchunks;
            for (int sp = first + issuer * ctas_per_nb; sp < sp_tiles; sp += n_iss * ctas_per_nb) {
                mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bars.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + stage * (S::A_STAGE >> 4);
                    uint32_t b_lo = b_lo0 + kc * b_blk16;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if ((p.debug & 1) && tap > 0) break;
                        if (p.debug & 32) break;
                        constexpr uint32_t kRow16 = ROW_BYTES >> 4;
                        const uint32_t a_tap = a_lo + ((tap / 3) * HALO_W + (tap % 3)) * kRow16;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            umma_f16_split(d_tmem, a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc,
                                           (tap | k) != 0 ? 1u : static_cast<uint32_t>(kc != 0));
                        b_lo += b_tap16;
                    }
                    if (p.debug & 16) mbar_arrive(&bars.empty[stage]); else umma_commit(&bars.empty[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                if (p.debug & 32) mbar_arrive(&bars.tmem_full[acc]); else umma_commit(&bars.tmem_full[acc]);
                acc += n_iss;                               // skip the other issuer's tile
                if (acc >= num_acc) { acc -= num_acc; acc_phase ^= 1; }
                for (int i = 0; i < (n_iss - 1) * kchunks; ++i)
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (active) {
        // epilogue set `eset` owns accumulator `eset`: it handles this CTA's tiles eset, eset + num_acc, ...
        const int num_acc = conv_num_acc(p.BN);
        const int eset = (warp - CONV_FIRST_EPI_WARP) >> 2;
        if (eset < num_acc) {
            uint32_t acc_phase = 0;
            for (int sp = first + eset * ctas_per_nb; sp < sp_tiles; sp += ctas_per_nb * num_acc) {
                const TileCoord t = tile_coord(p, sp, nb);
                mbar_wait(&bars.tmem_full[eset], acc_phase);
                tc_fence_after();
                conv_epilogue_tile(p, bars, tmem_base + eset * p.BN, &bars.tmem_empty[eset], t);
                acc_phase ^= 1;
            }
        }
    }
    conv_teardown(p, tmem_base);
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

template <int KC>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const ConvParams p) {
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
    const uint32_t tmem_base = conv_prologue(p, bars, num_stages);
    const int kchunks = p.Cin / KC;
    const int KB = 9 * kchunks;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord(p, tile / p.n_blocks, tile % p.n_blocks);
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&bars.empty[stage], phase ^ 1);
                        uint8_t* a_dst = smem + stage * stage_bytes;
                        mbar_arrive_expect_tx(&bars.full[stage], S::A_BYTES + p.BN * KC * 2);
                        tma_load_4d(a_dst, &tmap_x, &bars.full[stage], kc * KC, t.x0 + dx - 1, t.y0 + dy - 1, t.n);
                        tma_load_2d(a_dst + S::A_BYTES, &tmap_w, &bars.full[stage], kc * KC, tap * p.Cout + t.n0);
                        if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    }
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
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&bars.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = s_lo0 + stage * (static_cast<uint32_t>(stage_bytes) >> 4);
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)   // +32 bytes along K inside the swizzle row
                        umma_f16_split(d_tmem, a_lo + 2 * k, s_hi, a_lo + (S::A_BYTES >> 4) + 2 * k, s_hi, idesc,
                                       k != 0 ? 1u : static_cast<uint32_t>(kb != 0));
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
            for (int tile = blockIdx.x + eset * gridDim.x; tile < p.num_tiles; tile += gridDim.x * num_acc) {
                const TileCoord t = tile_coord(p, tile / p.n_blocks, tile % p.n_blocks);
                mbar_wait(&bars.tmem_full[eset], acc_phase);
                tc_fence_after();
                conv_epilogue_tile(p, bars, tmem_base + eset * p.BN, &bars.tmem_empty[eset], t);
                acc_phase ^= 1;
            }
        }
    }
    conv_teardown(p, tmem_base);
}

}  // namespace aesr
