// 3x3 / pad-1 convolution of the 32 -> 32 channel layers (enc.3, dec.8, dec.10; acai_vanilla.py:55,92,94) with the three
// HORIZONTAL filter taps folded into the GEMM's N dimension.
//
// Why.  An SS-mode tcgen05.mma (M = 128, K = 16) is bound by its shared-memory operand bytes at 128 B/clk until N = 256
// (tools/umma_rate.py: 40.1 / 48 / 56 / 64 / 96 cycles at N = 32 / 64 / 96 / 128 / 192).  With 32 output channels the tap-by-tap
// formulation of conv3x3_halo_kernel re-reads the 4 KB A tile for every 1 KB of B: 18 MMAs x 40 cycles = 720 cycles per 128
// pixels, 40 % of the tensor pipe's math rate at best (measured in the kernel: ~1150 cycles per tile).  Here
//     E_dx[y][x'] = sum_{dy,ci} in[y+dy-1][x'][ci] * W[dy][dx][ci][:]          (x' = INPUT column)
// is computed for dx = 0, 1, 2 by ONE MMA per (dy, K-step): B = the three taps of filter row dy stacked along N (96 rows -- in
// the packed [tap][Cout][Cin] filter these are CONTIGUOUS rows, so the existing weight packing is used as it is) and
//     out[y][x] = E_0[y][x-1] + E_1[y][x] + E_2[y][x+1]
// is formed by the epilogue with two warp shuffles per channel (TMEM lane = pixel, neighbouring columns = neighbouring lanes).
// 6 MMAs x 56 cycles per tile instead of 18 x 40, and the A tile is read from shared memory 6 instead of 18 times.
//
// Geometry.  M-tile = 8 rows x 16 INPUT columns, TMEM lane = 16 py + pxi; a warp's 32 lanes are two rows, so the column
// neighbours of lanes 1..14 of each row are in the same warp and NOTHING is exchanged between warps or tiles: the tile's first
// and last column (pxi = 0, 15) only serve as neighbours, 14 x 8 outputs per tile (tiles overlap by two input columns; the
// left neighbour of image column 0 is the TMA zero fill).  The 16-pixel pitch makes the 8-row groups of the A operand dense:
// no horizontal halo, vertical tap dy = start address + dy * 1024 bytes (swizzle-atom aligned).
//   activation box {32 ch, 16, 8 T + 2, 1} at (x0 - 1, y0 - 1) per super-tile of T vertically stacked M-tiles, SW64
//   TMEM: nbuf buffers x T accumulators x 96 columns (T = 1: five tiles in flight; T = 2: two buffers of two)
//   epilogue: 4 sets of 4 warps as in conv3x3_halo_kernel; avg-pool partners are lanes (l, l + 1) for odd pxi and l ^ 16.
// TMEM reads are not the limit (tools/tmem_ld_probe.py: >= 860 B/clk per SM with 16 warps; a 96-column tile is 48 KB), the
// shuffle unit is next (1.0 clk per warp shuffle SM-wide: 256 cycles per tile) -- profiles/r08_*.
#pragma once
#include "conv3x3_tc.cuh"

namespace aesr {

constexpr int FOLD_TILE_H = 8;
constexpr int FOLD_TILE_W = 16;                 // input columns per tile
constexpr int FOLD_VALID_W = 14;                // output columns per tile
constexpr int FOLD_KC = 32;
constexpr int FOLD_ROW_BYTES = FOLD_KC * 2;     // 64: SW64
constexpr int FOLD_N = 96;                      // three horizontal taps x 32 output channels
constexpr int FOLD_B_BYTES = 9 * 32 * FOLD_ROW_BYTES;      // resident filter bank [9 taps][32][32]
constexpr int FOLD_MAX_BUF = 5;
constexpr int FOLD_TAIL_BYTES = 512 + 3 * 32 * 4 + 4 * 32 * 4;   // barriers + tmem ptr | bias, scale, shift | BN statistics
__host__ __device__ constexpr int fold_a_stage(int T) { return (FOLD_TILE_H * T + 2) * FOLD_TILE_W * FOLD_ROW_BYTES; }   // multiple of 1024
__host__ __device__ constexpr int fold_total_bytes(int T, int stages) {
    return 1024 + FOLD_B_BYTES + stages * fold_a_stage(T) + FOLD_TAIL_BYTES;
}

struct FoldBarriers {
    uint64_t *full, *empty, *tmem_full, *tmem_empty, *b_full;
    uint32_t* tmem_ptr;
    float *s_bias, *s_scale, *s_shift, *s_stats;
    __device__ explicit FoldBarriers(uint8_t* tail) {
        full = reinterpret_cast<uint64_t*>(tail);
        empty = full + CONV_MAX_STAGES;
        tmem_full = empty + CONV_MAX_STAGES;
        tmem_empty = tmem_full + 8;
        b_full = tmem_empty + 8;
        tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);
        s_bias = reinterpret_cast<float*>(tail + 512);
        s_scale = s_bias + 32;
        s_shift = s_scale + 32;
        s_stats = s_shift + 32;                 // [2 passes][2][32]
    }
};

// lds_f4 as a volatile asm: the tile loop is fully unrolled here, and a plain asm load of the 96 loop-invariant epilogue
// constants is hoisted out of the persistent loop and SPILLED (24 STL.64 + LDLs per tile in the first build's SASS).
__device__ __forceinline__ float4 lds_f4v(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// Epilogue of one M-tile.  KMODE: OUT_SAME, OUT_AVGPOOL2 or a ConvLean training variant (conv3x3_tc.cuh).
template <int KMODE>
__device__ __forceinline__ void fold_epilogue_tile(const ConvParams& p, const FoldBarriers& bars, uint32_t tmem_acc,
                                                   uint64_t* tmem_empty_bar, int n, int y0, int x0) {
    constexpr int MODE = lean_out(KMODE);
    constexpr bool WITH_MUL = lean_mul(KMODE), WITH_STATS = lean_stats(KMODE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int py = row >> 4, pxi = row & 15;    // pxi = input column of the tile; output column = x0 + pxi - 1
    const int H = p.H, W = p.W, fp16 = p.fp16;
    const int y = y0 + py, x = x0 + pxi - 1;
    const bool valid = pxi >= 1 && pxi <= FOLD_VALID_W && y < H && x < W;
    const float neg_slope = (p.act == ACT_LEAKY) ? p.slope : (p.act == ACT_RELU) ? 0.f : 1.f;
    const uint32_t t_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const bool affine = p.scale != nullptr;
    size_t pix;
    bool store;
    if (MODE == OUT_AVGPOOL2) {
        const int Ho = H >> 1, Wo = W >> 1, yo = y >> 1, xo = x >> 1;
        pix = (static_cast<size_t>(n) * Ho + yo) * Wo + xo;
        // 2x2 window = lanes {l, l + 1} (x0 is even: odd pxi = even image column) x {l, l ^ 16} (y0 is a multiple of 8)
        store = (pxi & 1) && pxi < FOLD_VALID_W && !(py & 1) && yo < Ho && xo < Wo;
    } else {
        pix = (static_cast<size_t>(n) * H + y) * W + x;
        store = valid;
    }
    uint16_t* out16 = static_cast<uint16_t*>(p.out);
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
        float v[16];
        {
            // three loads of 16 columns, software-pipelined over two register sets (48 live values, not 64)
            uint32_t ra[16], rb[16];
            tmem_ld_32x32b_x16(t_addr + 32 + c0, ra);       // tap dx = 1: this column
            tmem_ld_32x32b_x16(t_addr + c0, rb);            // tap dx = 0: belongs to the output column on the right
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(ra[j]);
            tmem_ld_32x32b_x16(t_addr + 64 + c0, ra);       // tap dx = 2: belongs to the output column on the left
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __shfl_up_sync(0xffffffffu, __uint_as_float(rb[j]), 1);
            tmem_ld_wait();
            if (c0 == 16) {                                 // accumulator fully read: one elected arrive per warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __shfl_down_sync(0xffffffffu, __uint_as_float(ra[j]), 1);
        }
        // act(a) = max(a, a * neg_slope): neg_slope = 1 (identity), 0.01 (LeakyReLU), 0 (ReLU) -- branch-free
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = lds_f4v(bars.s_bias + c0 + 4 * j4);
            const float a0 = v[j4 * 4 + 0] + b.x, a1 = v[j4 * 4 + 1] + b.y, a2 = v[j4 * 4 + 2] + b.z, a3 = v[j4 * 4 + 3] + b.w;
            v[j4 * 4 + 0] = fmaxf(a0, a0 * neg_slope);
            v[j4 * 4 + 1] = fmaxf(a1, a1 * neg_slope);
            v[j4 * 4 + 2] = fmaxf(a2, a2 * neg_slope);
            v[j4 * 4 + 3] = fmaxf(a3, a3 * neg_slope);
        }
        if (WITH_MUL) {
            if (valid) {
                const uint4* m4 = reinterpret_cast<const uint4*>(p.mul_src + pix * 32 + c0);
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
            for (int j = 0; j < 16; ++j) s1[j] = valid ? v[j] : 0.f;
            const float t1 = warp_transpose_sum16(s1, lane);
            const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if (KMODE == LEAN_SAME_MUL_SUM) {                 // bias gradient: sums only, one block
                if ((lane & 1) == 0) atomicAdd(bars.s_stats + c0 + ch, t1);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) s1[j] *= s1[j];
                const float t2 = warp_transpose_sum16(s1, lane);
                if ((lane & 1) == 0) {
                    float* sst = bars.s_stats + (n >= p.stats_split ? 64 : 0);
                    atomicAdd(sst + c0 + ch, t1);
                    atomicAdd(sst + 32 + c0 + ch, t2);
                }
            }
        }
        if (affine) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 sc = lds_f4v(bars.s_scale + c0 + 4 * j4), sh = lds_f4v(bars.s_shift + c0 + 4 * j4);
                v[j4 * 4 + 0] = fmaf(v[j4 * 4 + 0], sc.x, sh.x);
                v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], sc.y, sh.y);
                v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], sc.z, sh.z);
                v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], sc.w, sh.w);
            }
        }
        if (MODE == OUT_AVGPOOL2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = v[j];
                a += __shfl_down_sync(0xffffffffu, a, 1);
                a += __shfl_xor_sync(0xffffffffu, a, 16);
                v[j] = a * 0.25f;
            }
        }
        if (store) {
            uint4* o4 = reinterpret_cast<uint4*>(out16 + pix * 32 + c0);
            o4[0] = pack8(v, fp16);
            o4[1] = pack8(v + 8, fp16);
        }
    }
}

template <int MODE>        // OUT_SAME, OUT_AVGPOOL2 or a ConvLean training variant
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_fold_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int T = p.T, nbuf = p.nbuf, num_stages = p.num_stages;
    const int a_stage = fold_a_stage(T);
    uint8_t* b_smem = smem;                                         // [9 taps][32 rows][32 ch] swizzled, taps contiguous
    uint8_t* a_smem = smem + FOLD_B_BYTES;                          // [stage][(8T+2)*16 pixel rows][32 ch] swizzled
    FoldBarriers bars(a_smem + num_stages * a_stage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool WITH_STATS = lean_stats(MODE);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        for (int a = 0; a < nbuf; ++a) {
            mbar_init(&bars.tmem_full[a], 1);
            mbar_init(&bars.tmem_empty[a], 4 * T);                  // one elected lane per epilogue warp and M-tile
        }
        mbar_init(bars.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(bars.tmem_ptr, 512);
    if (threadIdx.x < 32) {
        bars.s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
        bars.s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
        bars.s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
    }
    if (WITH_STATS && threadIdx.x < 128) bars.s_stats[threadIdx.x] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *bars.tmem_ptr;

    const int st_per_img = p.tiles_x * p.stiles_y;
    const int st_total = p.N * st_per_img;
    const int first = blockIdx.x, stride = gridDim.x;

    if (warp == 0) {
        // TMA producer: warp-uniform loop, one elected lane issues
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(bars.b_full, FOLD_B_BYTES);
            for (int dy = 0; dy < 3; ++dy) tma_load_2d(b_smem + dy * (FOLD_N * FOLD_ROW_BYTES), &tmap_w, bars.b_full, 0, dy * FOLD_N);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        int n = first / st_per_img, sy = (first % st_per_img) / p.tiles_x, tx = first % p.tiles_x;
        const int dn = stride / st_per_img, dsy = (stride % st_per_img) / p.tiles_x, dtx = stride % p.tiles_x;
        for (int st = first; st < st_total; st += stride) {
            mbar_wait(&bars.empty[stage], phase ^ 1);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&bars.full[stage], a_stage);
                tma_load_4d(a_smem + stage * a_stage, &tmap_x, &bars.full[stage], 0, tx * FOLD_VALID_W - 1,
                            sy * T * FOLD_TILE_H - 1, n);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
            tx += dtx;
            if (tx >= p.tiles_x) { tx -= p.tiles_x; ++sy; }
            sy += dsy;
            if (sy >= p.stiles_y) { sy -= p.stiles_y; ++n; }
            n += dn;
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform loop, every tcgen05 instruction under elect.sync
        const uint32_t idesc = make_idesc_16(CONV_TILE_M, FOLD_N, p.fp16);
        int stage = 0, buf = 0;
        uint32_t phase = 0, buf_phase = 0;
        mbar_wait(bars.b_full, 0);
        // only the start-address field (bits 0..13 of the low word, address >> 4) changes between MMAs
        const uint64_t a_tmpl = make_smem_desc(smem_u32(a_smem), 8 * FOLD_ROW_BYTES, UMMA_LAYOUT_SW64);
        const uint64_t b_tmpl = make_smem_desc(smem_u32(b_smem), 8 * FOLD_ROW_BYTES, UMMA_LAYOUT_SW64);
        const uint32_t a_hi = static_cast<uint32_t>(a_tmpl >> 32), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(a_tmpl), b_lo0 = static_cast<uint32_t>(b_tmpl);
        constexpr uint32_t kRowDy16 = (FOLD_TILE_W * FOLD_ROW_BYTES) >> 4;            // one image row of the tile: 1024 B
        constexpr uint32_t kTile16 = FOLD_TILE_H * kRowDy16;                           // one M-tile: 8 rows
        constexpr uint32_t kBDy16 = (FOLD_N * FOLD_ROW_BYTES) >> 4;                    // one filter row: 96 x 64 B
        int sy = (first % st_per_img) / p.tiles_x, tx = first % p.tiles_x;
        const int dsy = (stride % st_per_img) / p.tiles_x, dtx = stride % p.tiles_x;
        for (int st = first; st < st_total; st += stride) {
            const int rows_left = p.tiles_y - sy * T;
            const int nvalid = rows_left < T ? rows_left : T;
            mbar_wait(&bars.tmem_empty[buf], buf_phase ^ 1);
            mbar_wait(&bars.full[stage], phase);
            tc_fence_after();
            const uint32_t a_stage_lo = a_lo0 + stage * (static_cast<uint32_t>(a_stage) >> 4);
            for (int t = 0; t < nvalid; ++t) {
                const uint32_t d_tmem = tmem_base + (buf * T + t) * FOLD_N;
                const uint32_t a_lo = a_stage_lo + t * kTile16;
                if (elect_one_sync()) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int k = 0; k < FOLD_KC / 16; ++k)
                            umma_f16_split(d_tmem, a_lo + dy * kRowDy16 + 2 * k, a_hi, b_lo0 + dy * kBDy16 + 2 * k, b_hi, idesc,
                                           (dy | k) != 0 ? 1u : 0u);
                }
                __syncwarp();
            }
            if (elect_one_sync()) {
                umma_commit(&bars.empty[stage]);
                umma_commit(&bars.tmem_full[buf]);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
            if (++buf == nbuf) { buf = 0; buf_phase ^= 1; }
            tx += dtx;
            if (tx >= p.tiles_x) { tx -= p.tiles_x; ++sy; }
            sy += dsy;
            if (sy >= p.stiles_y) sy -= p.stiles_y;
        }
    } else {
        // Epilogue sets: M-tile t of super-tile i (running index i*T + t) belongs to set (i*T + t) & 3; super-tile i lives in
        // TMEM buffer i % nbuf.  A set does not see every use of a buffer when nbuf = 5, but the use it waits for (tile j) can
        // only be preceded by the use of tile j - 5 * (tiles per buffer), whose MMAs completed before those of the set's own
        // previous tile (commits complete in issue order) -- so the barrier is never a whole phase behind the waiter.
        const int eset = (warp - CONV_FIRST_EPI_WARP) >> 2;
        int i = 0, buf = 0;
        uint32_t buf_phase = 0;
        for (int st = first; st < st_total; st += stride, ++i) {
            const int t = (eset - i * T) & 3;
            if (t < T) {
                const int n = st / st_per_img;
                const int r = st - n * st_per_img;
                const int sy = r / p.tiles_x;
                const int tx = r - sy * p.tiles_x;
                mbar_wait(&bars.tmem_full[buf], buf_phase);
                tc_fence_after();
                if (sy * T + t < p.tiles_y) {
                    fold_epilogue_tile<MODE>(p, bars, tmem_base + (buf * T + t) * FOLD_N, &bars.tmem_empty[buf], n,
                                             (sy * T + t) * FOLD_TILE_H, tx * FOLD_VALID_W);
                } else {                       // M-tile below the image: nothing to read, release the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.tmem_empty[buf]);
                }
            }
            if (++buf == nbuf) { buf = 0; buf_phase ^= 1; }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (WITH_STATS && p.stats != nullptr)        // every epilogue of this CTA has added its sums: flush them
        for (int c = threadIdx.x; c < (p.stats_sum_only ? 1 : p.stats_split < p.N ? 4 : 2) * 32; c += CONV_THREADS) {
            const float v = bars.s_stats[c];
            if (v != 0.f) atomicAdd(p.stats + c, v);
        }
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace aesr
