// 3x3 / pad-1 / stride-1 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   activations : NHWC bf16, C in {32, 64, 128, 256, 512}
//   weights     : packed [9 taps][Cout][Cin] bf16 (K-major B operand per tap)
//   accumulate  : fp32 in TMEM, 128 output pixels (16 rows x 8 cols) x BN output channels per tile
//
// One K-block = one filter tap x KC input channels.  The A tile of a K-block is the 16x8 pixel window shifted by the
// tap offset, fetched by a single 4-D TMA box {KC, 8, 16, 1} whose start coordinate may be -1 / run past the image:
// TMA zero-fills out-of-bounds elements, which *is* the conv's zero padding (no im2col, no halo buffers, no masks).
// The box lands in shared memory as 128 rows (pixels) of KC*2 bytes with the hardware 128B/64B swizzle, i.e. exactly
// the canonical K-major UMMA operand layout.
//
// Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM
// owner, warps 2-5 = epilogue (TMEM -> registers -> fused bias / activation / BN-affine / pool / upsample -> global).
// Two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace aesr {

enum ConvAct : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };
enum ConvOut : int {
    OUT_SAME = 0,          // out  = NHWC bf16 [N,H,W,Cout]
    OUT_AVGPOOL2 = 1,      // out  = NHWC bf16 [N,H/2,W/2,Cout]  (floor; 2x2 mean of the post-affine activation)
    OUT_UP2 = 2,           // out  = NHWC bf16 [N,2H,2W,Cout]    (nearest)
    OUT_NCHW_F32 = 3,      // out  = NCHW fp32 [N,Cout,H,W]      (+ optional out2 = NHWC bf16 copy)
    OUT_SAME_MAXPOOL2 = 4, // out  = NHWC bf16 full res, out2 = NHWC bf16 [N,H/2,W/2,Cout] 2x2 max
};
enum ConvMul : int { MUL_NONE = 0, MUL_LEAKY_GRAD = 1, MUL_RELU_GRAD = 2 };

struct ConvParams {
    int N, H, W, Cin, Cout;
    int BN;                 // output channels per tile (multiple of 32, <= 256, divides Cout)
    int tiles_x, tiles_y, n_blocks, num_tiles;
    int num_stages;
    // epilogue
    const float* bias;      // [Cout] or null
    const float* scale;     // [Cout] or null: y = act(acc + bias) * scale + shift
    const float* shift;
    float slope;
    int act;
    int out_mode;
    void* out;
    void* out2;
    // optional elementwise multiplier (dgrad): out *= act'(mul_src) with mul_src NHWC bf16 [N,H,W,Cout]
    const __nv_bfloat16* mul_src;
    int mul_mode;
    // optional per-channel statistics of the stored (post-activation, pre-affine) value: sums[c], sums[Cout + c]
    float* stats;
};

constexpr int CONV_TILE_H = 16;
constexpr int CONV_TILE_W = 8;
constexpr int CONV_TILE_M = 128;
constexpr int CONV_THREADS = 192;
constexpr int CONV_MAX_STAGES = 8;

template <int KC>
struct ConvSmem {
    static constexpr int A_BYTES = CONV_TILE_M * KC * 2;
    __host__ __device__ static constexpr int b_bytes(int BN) { return ((BN * KC * 2 + 1023) / 1024) * 1024; }
    __host__ __device__ static constexpr int stage_bytes(int BN) { return A_BYTES + b_bytes(BN); }
    // barriers + tmem ptr + epilogue constants (3 * 512 floats)
    static constexpr int TAIL_BYTES = 256 + 3 * 512 * 4;
    __host__ __device__ static constexpr int total_bytes(int BN, int stages) {
        return 1024 /*alignment slack*/ + stages * stage_bytes(BN) + TAIL_BYTES;
    }
};

template <int KC>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const ConvParams p) {
    using S = ConvSmem<KC>;
    constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
    constexpr uint32_t ROW_BYTES = KC * 2;
    constexpr uint32_t SBO = 8 * ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = S::stage_bytes(p.BN);
    const int num_stages = p.num_stages;
    uint8_t* tail = smem + num_stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);                 // [CONV_MAX_STAGES]
    uint64_t* empty_bar = full_bar + CONV_MAX_STAGES;                       // [CONV_MAX_STAGES]
    uint64_t* tmem_full_bar = empty_bar + CONV_MAX_STAGES;                  // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;                           // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(tail + 256);
    float* s_scale = s_bias + 512;
    float* s_shift = s_scale + 512;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_cols = (2 * p.BN <= 32) ? 32 : (2 * p.BN <= 64) ? 64 : (2 * p.BN <= 128) ? 128
                             : (2 * p.BN <= 256) ? 256 : 512;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, tmem_cols);
    for (int c = threadIdx.x; c < p.Cout; c += CONV_THREADS) {
        s_bias[c] = p.bias ? p.bias[c] : 0.f;
        s_scale[c] = p.scale ? p.scale[c] : 1.f;
        s_shift[c] = p.shift ? p.shift[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int kchunks = p.Cin / KC;
    const int KB = 9 * kchunks;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int nb = tile % p.n_blocks;
                const int sp = tile / p.n_blocks;
                const int n = sp / tiles_per_img;
                const int r = sp - n * tiles_per_img;
                const int ty = r / p.tiles_x;
                const int tx = r - ty * p.tiles_x;
                const int y0 = ty * CONV_TILE_H, x0 = tx * CONV_TILE_W, n0 = nb * p.BN;
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = smem + stage * stage_bytes;
                        uint8_t* b_dst = a_dst + S::A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], S::A_BYTES + p.BN * KC * 2);
                        tma_load_4d(a_dst, &tmap_x, &full_bar[stage], kc * KC, x0 + dx - 1, y0 + dy - 1, n);
                        tma_load_2d(b_dst, &tmap_w, &full_bar[stage], kc * KC, tap * p.Cout + n0);
                        if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(CONV_TILE_M, p.BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                    const uint32_t b_addr = a_addr + S::A_BYTES;
                    const uint64_t a_desc = make_smem_desc(a_addr, SBO, LAYOUT);
                    const uint64_t b_desc = make_smem_desc(b_addr, SBO, LAYOUT);
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                        umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (4 warps = 128 TMEM lanes) =================
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;              // pixel index inside the tile
        const int py = row >> 3, px = row & 7;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int H = p.H, W = p.W, Cout = p.Cout;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nb = tile % p.n_blocks;
            const int sp = tile / p.n_blocks;
            const int n = sp / tiles_per_img;
            const int r = sp - n * tiles_per_img;
            const int ty = r / p.tiles_x;
            const int tx = r - ty * p.tiles_x;
            const int y = ty * CONV_TILE_H + py, x = tx * CONV_TILE_W + px;
            const int n0 = nb * p.BN;
            const bool inb = (y < H) && (x < W);
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.BN;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32b_x32(t_addr + c0, raw);
                tmem_ld_wait();
                if (c0 + 32 >= p.BN) {              // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&tmem_empty_bar[acc]);
                }
                float v[32];
                const int cg = n0 + c0;             // first global output channel of this chunk
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = __uint_as_float(raw[j]) + s_bias[cg + j];
                    if (p.act == ACT_LEAKY) a = a > 0.f ? a : a * p.slope;
                    else if (p.act == ACT_RELU) a = fmaxf(a, 0.f);
                    v[j] = a;
                }
                if (p.mul_mode != MUL_NONE) {
                    if (inb) {
                        const uint4* m4 = reinterpret_cast<const uint4*>(
                            p.mul_src + (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const uint4 m = __ldg(m4 + j4);
                            const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float lo = bf16_lo(w[u]), hi = bf16_hi(w[u]);
                                const float neg = (p.mul_mode == MUL_LEAKY_GRAD) ? p.slope : 0.f;
                                v[j4 * 8 + u * 2] *= (lo > 0.f) ? 1.f : neg;
                                v[j4 * 8 + u * 2 + 1] *= (hi > 0.f) ? 1.f : neg;
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
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_scale[cg + j], s_shift[cg + j]);
                }
                if (p.out_mode == OUT_SAME || p.out_mode == OUT_SAME_MAXPOOL2) {
                    if (inb) {
                        uint4* o4 = reinterpret_cast<uint4*>(
                            static_cast<__nv_bfloat16*>(p.out) +
                            (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            o4[j4] = make_uint4(pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]),
                                                pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]),
                                                pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]),
                                                pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]));
                    }
                }
                if (p.out_mode == OUT_AVGPOOL2 || p.out_mode == OUT_SAME_MAXPOOL2) {
                    // 2x2 window = lanes {l, l^1, l^8} (x neighbour, y neighbour): tile origin is even-aligned.
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
                        void* dst = (p.out_mode == OUT_AVGPOOL2) ? p.out : p.out2;
                        uint4* o4 = reinterpret_cast<uint4*>(
                            static_cast<__nv_bfloat16*>(dst) +
                            (static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(yo) * Wo + xo) * Cout + cg);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            o4[j4] = make_uint4(pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]),
                                                pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]),
                                                pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]),
                                                pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]));
                    }
                } else if (p.out_mode == OUT_UP2) {
                    if (inb) {
                        const int Ho = 2 * H, Wo = 2 * W;
                        uint4 pk[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            pk[j4] = make_uint4(pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]),
                                                pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]),
                                                pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]),
                                                pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]));
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            const int yo = 2 * y + (d >> 1), xo = 2 * x + (d & 1);
                            uint4* o4 = reinterpret_cast<uint4*>(
                                static_cast<__nv_bfloat16*>(p.out) +
                                (static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(yo) * Wo + xo) * Cout + cg);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) o4[j4] = pk[j4];
                        }
                    }
                } else if (p.out_mode == OUT_NCHW_F32) {
                    if (inb) {
                        float* o = static_cast<float*>(p.out) +
                                   (static_cast<size_t>(n) * Cout + cg) * H * W + static_cast<size_t>(y) * W + x;
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[static_cast<size_t>(j) * H * W] = v[j];
                        if (p.out2 != nullptr) {
                            uint4* o4 = reinterpret_cast<uint4*>(
                                static_cast<__nv_bfloat16*>(p.out2) +
                                (static_cast<size_t>(n) * H * W + static_cast<size_t>(y) * W + x) * Cout + cg);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4)
                                o4[j4] = make_uint4(pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]),
                                                    pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]),
                                                    pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]),
                                                    pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]));
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace aesr
