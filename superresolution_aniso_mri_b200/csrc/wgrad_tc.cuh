// Weight gradient of the 3x3 / pad-1 conv on the tensor cores (tcgen05 + TMEM), pixels as the GEMM K dimension:
//
//     dW[co][ci][tap] = sum_{n,y,x} g[n,y,x,co] * X[n, y+dy-1, x+dx-1, ci]
//
// Both operands are NHWC, i.e. [pixel][channel] = "MN-major" for the MMA (channels contiguous, K = pixels strided by
// one row): A = g tile (128 pixels x Cout), B = X halo (18x10 pixels x Cin) seen through a descriptor shifted by the
// tap offset -- the same halo trick as the forward kernel, with the 8-pixel K groups strided by the halo pitch.  TMA
// zero-fills out-of-image pixels of both operands (= the conv's zero padding / ragged tile edges).
// A CTA walks its share of the pixel tiles and keeps accumulating into the SAME TMEM columns (one N-wide accumulator
// per tap of its tap group); only at the very end the 4 epilogue warps add the partial dW to global memory with fp32
// atomics.  M is always 128: for Cout < 128 the A descriptor's MN-block stride (LBO) is 0, so TMEM lanes >= Cout hold
// duplicates that are simply not read back.
#pragma once
#include "common.cuh"

namespace aesr {

struct WgradParams {
    int N, H, W, Cg, Cx;          // g: [N,H,W,Cg] bf16 (Cg = Cout), x: [N,H,W,Cx] 16-bit (Cx = Cin)
    int tiles_x, tiles_y, num_tiles;
    int taps_per_group, num_groups;
    int ctas_per_group;
    int x_fp16;                   // x operand format: 1 fp16, 0 bf16 (g is always bf16)
    int stages;                   // pipeline depth (2..WG_MAX_STAGES, as many as fit in shared memory)
    int fold_dx;                  // Cin = 32: one N = 96 MMA per filter row (tuning knob, default on)
    float* dW;                    // [Cg][Cx][9] fp32, accumulated
};

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 6;       // (g tile, x halo) pairs in flight: as many as fit (one tile's MMAs are ~1.3-5 k cycles, an
                                       // L2 -> SMEM round trip under load about as much: two stages left the producer exposed)

__host__ __device__ constexpr int wg_g_chunk_bytes(int Cg) { return 128 * (Cg < 64 ? Cg : 64) * 2; }
__host__ __device__ constexpr int wg_x_chunk_bytes(int Cx) {
    return ((18 * 10 * (Cx < 64 ? Cx : 64) * 2 + 1023) / 1024) * 1024;
}
__host__ __device__ constexpr int wg_stage_bytes(int Cg, int Cx) {
    return wg_g_chunk_bytes(Cg) * ((Cg + 63) / 64) + wg_x_chunk_bytes(Cx) * ((Cx + 63) / 64);
}
__host__ __device__ constexpr int wg_smem_bytes(int Cg, int Cx, int stages) { return 1024 + stages * wg_stage_bytes(Cg, Cx) + 256; }

// MN-major shared-memory descriptor: LBO = stride between 64-element (swizzle-row) blocks along M/N,
// SBO = stride between groups of 8 K-rows.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                      uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type & 0x7) << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                   const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int gKC = p.Cg < 64 ? p.Cg : 64, xKC = p.Cx < 64 ? p.Cx : 64;
    const int g_chunks = (p.Cg + 63) / 64, x_chunks = (p.Cx + 63) / 64;
    const int g_chunk_bytes = wg_g_chunk_bytes(p.Cg), x_chunk_bytes = wg_x_chunk_bytes(p.Cx);
    const int g_bytes = g_chunk_bytes * g_chunks;
    const int stage_bytes = wg_stage_bytes(p.Cg, p.Cx);
    const int num_stages = p.stages;
    uint8_t* tail = smem + num_stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
    uint64_t* done_bar = empty_bar + WG_MAX_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NN = p.Cx;                                   // MMA N (<= 128)
    const int ncols = p.taps_per_group * NN;
    const uint32_t tmem_cols = ncols <= 32 ? 32 : ncols <= 64 ? 64 : ncols <= 128 ? 128 : ncols <= 256 ? 256 : 512;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_g);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int group = blockIdx.x / p.ctas_per_group;
    const int first = blockIdx.x - group * p.ctas_per_group;
    const int tap0 = group * p.taps_per_group;
    const int ntaps = (9 - tap0) < p.taps_per_group ? (9 - tap0) : p.taps_per_group;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    // Producer and issuer run WARP-UNIFORM loops and issue under elect.sync (see elect_one_sync in common.cuh): issued
    // from `if (lane == 0)` code every TMA / tcgen05 instruction is wrapped in an ELECT + broadcast + branch loop, and with
    // the 64-bit descriptors rebuilt per MMA the single issuing thread needed ~150 cycles per MMA -- 10.6 k cycles per
    // 128-pixel tile against 72 MMAs x 40-48 tensor-pipe cycles (tools/train_layer_times.py: 138 us for 32->32 @130^2).
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = first; t < p.num_tiles; t += p.ctas_per_group) {
            const int n = t / tiles_per_img, r = t - n * tiles_per_img;
            const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one_sync()) {
                uint8_t* base = smem + stage * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], 128 * p.Cg * 2 + 180 * p.Cx * 2);
                for (int c = 0; c < g_chunks; ++c)
                    tma_load_4d(base + c * g_chunk_bytes, &tmap_g, &full_bar[stage], c * 64, x0, y0, n);
                for (int c = 0; c < x_chunks; ++c)
                    tma_load_4d(base + g_bytes + c * x_chunk_bytes, &tmap_x, &full_bar[stage], c * 64, x0 - 1, y0 - 1, n);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // idesc: D f32, A = bf16 (g), B = x format, both MN-major, M = 128, N = Cx
        const uint32_t bfmt = p.x_fp16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (bfmt << 10) | (1u << 15) | (1u << 16) |
                               ((static_cast<uint32_t>(NN) >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_layout = gKC == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
        const uint32_t b_layout = xKC == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
        const uint32_t a_row = gKC * 2, b_row = xKC * 2;
        const uint32_t a_lbo = g_chunks > 1 ? g_chunk_bytes : 0;     // Cout < 128: duplicate the block (rows unused)
        const uint32_t b_lbo = x_chunks > 1 ? x_chunk_bytes : 0;
        // only the 14-bit start-address field (address >> 4, low word) changes between MMAs
        const uint64_t a_tmpl = make_smem_desc_mn(smem_u32(smem), a_lbo, 8 * a_row, a_layout);
        const uint64_t b_tmpl = make_smem_desc_mn(smem_u32(smem) + g_bytes, b_lbo, 10 * b_row, b_layout);
        const uint32_t a_hi = static_cast<uint32_t>(a_tmpl >> 32), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(a_tmpl), b_lo0 = static_cast<uint32_t>(b_tmpl);
        const uint32_t a_ks16 = (16 * a_row) >> 4, b_row16 = b_row >> 4, stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
        // dx-folded variant (Cin = 32, all nine taps in this CTA): same descriptor with LBO = one pixel, N = 96
        const bool fold_dx = p.Cx == 32 && ntaps == 9 && p.fold_dx;
        const uint32_t b_lo96 = static_cast<uint32_t>(make_smem_desc_mn(smem_u32(smem) + g_bytes, b_row, 10 * b_row, b_layout));
        const uint32_t idesc96 = (idesc & ~(0x3Fu << 17)) | ((96u >> 3) << 17);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t acc0 = 0;                                           // 0 for the very first K-step of every accumulator
        for (int t = first; t < p.num_tiles; t += p.ctas_per_group) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_st = a_lo0 + stage * stage16, b_st = b_lo0 + stage * stage16;
            const uint32_t b_st96 = b_lo96 + stage * stage16;
            if (elect_one_sync()) {
                if (fold_dx) {
                    // Cin = 32: the three dx taps of a filter row are ONE MMA with N = 96 -- the B operand's 32-channel
                    // N-blocks are strided by LBO = one halo pixel (64 bytes), so block j reads the window shifted by
                    // j pixels, and its 32 accumulator columns are exactly tap (dy, j)'s.  24 MMAs of N = 96 per tile
                    // instead of 72 of N = 32 (each A read of 4 KB now feeds three taps).
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t b_tap = b_st96 + dy * 10 * b_row16;
                        const uint32_t d_tmem = tmem_base + dy * 3 * NN;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_f16_split(d_tmem, a_st + ks * a_ks16, a_hi, b_tap + ks * 20 * b_row16, b_hi, idesc96,
                                           ks == 0 ? acc0 : 1u);
                    }
                } else {
                    for (int tl = 0; tl < ntaps; ++tl) {
                        const int tap = tap0 + tl, dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t b_tap = b_st + (dy * 10 + dx) * b_row16;
                        const uint32_t d_tmem = tmem_base + tl * NN;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)                       // K = 128 pixels = 8 x 16
                            umma_f16_split(d_tmem, a_st + ks * a_ks16, a_hi, b_tap + ks * 20 * b_row16, b_hi, idesc,
                                           ks == 0 ? acc0 : 1u);
                    }
                }
                umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            acc0 = 1u;
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(done_bar);
        __syncwarp();
    } else {
        // final epilogue: TMEM lane = output channel co, columns = [tap][ci]
        mbar_wait(done_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int co = q * 32 + lane;
        const bool has_tiles = first < p.num_tiles;
        if (has_tiles) {
            for (int tl = 0; tl < ntaps; ++tl) {
                for (int c0 = 0; c0 < NN; c0 += 32) {
                    uint32_t raw[32];
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * NN + c0, raw);
                    tmem_ld_wait();
                    if (co < p.Cg) {
                        float* dst = p.dW + (static_cast<size_t>(co) * p.Cx + c0) * 9 + (tap0 + tl);
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j * 9, __uint_as_float(raw[j]));
                    }
                }
            }
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

// dbias[c] += sum over pixels of g[p][c]  (g bf16 [P][C], C % 8 == 0, C <= 512).  block = (C/8 channel groups) x
// (256 / (C/8) pixel rows): a thread sums 8 channels (one 16-byte load) over a strip of pixels, the rows are folded through
// shared memory, one atomic per channel and block.  (2-byte loads per lane before: ~20 us per layer, 13 layers per step.)
__global__ void colsum_bf16_kernel(const uint16_t* __restrict__ g, float* __restrict__ out, size_t P, int C) {
    extern __shared__ float s_col[];                 // [rows][C]
    const int groups = C >> 3, rows = blockDim.x / groups;
    const int cg = threadIdx.x % groups, row = threadIdx.x / groups;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t p = blockIdx.x * static_cast<size_t>(rows) + row; p < P; p += static_cast<size_t>(gridDim.x) * rows) {
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(g + p * C) + cg);
        const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[2 * k] += bf16_lo(u[k]);
            acc[2 * k + 1] += bf16_hi(u[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_col[row * C + cg * 8 + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += s_col[r * C + c];
        atomicAdd(out + c, t);
    }
}

}  // namespace aesr
