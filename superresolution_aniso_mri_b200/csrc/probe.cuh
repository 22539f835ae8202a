// Hardware probe (diagnostic entry point, not on the product path): does a tcgen05 shared-memory descriptor whose
// start address is shifted by whole 128-byte rows inside a TMA-written, 128B-swizzled halo tile address the right
// data?  If yes, one halo load per tile can feed all 9 filter taps (no 9x re-fetch of the activation window).
//
// One CTA, one tile: Cin = 64, Cout = 64, 16x8 output pixels.  Halo tile = 18 rows x `pitch` pixels x 128 B.
//   variant bit 0: set the descriptor's base-offset field to (start_address >> 7) & 7
//   pitch = 10 (dense TMA box {64,10,18,1}) or 16 (box {64,16,18,1}: 8-row groups 1024-byte aligned)
#pragma once
#include "common.cuh"

namespace aesr {

__global__ void __launch_bounds__(128, 1)
halo_probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  float* __restrict__ out /*[128][64]*/, int x0, int y0, int n, int pitch, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_halo = smem;                       // 18 * 16 * 128 = 36864 B max
    uint8_t* b_all = smem + 36864;                // 9 * 64 * 128 = 73728 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_all + 73728);
    uint64_t* done_bar = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 18 * pitch * 128 + 9 * 64 * 128);
        tma_load_4d(a_halo, &tmap_x, bar, 0, x0 - 1, y0 - 1, n);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_all + tap * 8192, &tmap_w, bar, 0, tap * 64);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 64);
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint32_t a_addr = smem_u32(a_halo) + (dy * pitch + dx) * 128;
            uint64_t a_desc = make_smem_desc(a_addr, pitch * 128, UMMA_LAYOUT_SW128);
            if (variant & 1) a_desc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
            const uint64_t b_desc = make_smem_desc(smem_u32(b_all + tap * 8192), 1024, UMMA_LAYOUT_SW128);
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (tap | k) != 0);
        }
        umma_commit(done_bar);
    }
    __syncwarp();
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, raw);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[row * 64 + c0 + j] = __uint_as_float(raw[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// Tensor-core issue-rate probe (diagnostic): `iters` back-to-back tcgen05.mma (M = 128, N, K = 16, fp16) on whatever is
// in shared memory, A descriptor = {start row shift, 8-row-group stride `pitch_rows`}, row = KC*2 bytes (SW64 / SW128).
// cycles[0] = clock64 from the first issue to the commit's mbarrier completion.  Answers: what does an SS-mode MMA cost
// when its A operand rows are (a) dense and swizzle-atom aligned, (b) row-shifted / pitched like the halo tile?
template <int NACC>
__global__ void __launch_bounds__(128, 1)
umma_rate_probe_kernel(long long* __restrict__ cycles, int N, int kc, int pitch_rows, int shift_rows, int iters,
                       int a_advance_rows, int fill_random) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    // operands: zeros, or (fill_random) pseudo-random fp16 values in (-2, 2) -- data toggling changes the tensor-pipe power
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (i + 1 + blockIdx.x * 7919u) * 2654435761u;
        h ^= h >> 15;
        reinterpret_cast<uint32_t*>(smem)[i] = fill_random ? ((h & 0x83FF83FFu) | 0x3C003C00u) : 0u;
    }
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 1) {
        const uint32_t row_bytes = kc * 2;
        const uint32_t layout = (kc == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
        const uint64_t a_desc0 = make_smem_desc(smem_u32(smem) + shift_rows * row_bytes, pitch_rows * row_bytes, layout);
        const uint64_t b_desc = make_smem_desc(smem_u32(smem) + 128 * 1024, 8 * row_bytes, layout);
        const uint32_t idesc = make_idesc_16(128, N, 1);
        const uint32_t adv = (a_advance_rows * row_bytes) >> 4;
        uint32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = tmem_base + (j % NACC) * N;
        long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {            // 8 MMAs per trip: loop overhead amortised
            if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) umma_f16(d[j], a_desc0 + j * adv, b_desc, idesc, 1);
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&done_bar);
        __syncwarp();
        mbar_wait(&done_bar, 0);
        const long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (elect_one_sync()) {                          // per CTA: SM cycles and wall-clock ns of the same interval
            cycles[2 * blockIdx.x] = t1 - t0;
            cycles[2 * blockIdx.x + 1] = g1 - g0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Issue-pattern probe (diagnostic): the MMA issue loop of conv3x3_halo_kernel in isolation -- T M-tiles x 9 taps x KC/16
// K-steps per "super-tile" with the kernel's descriptor arithmetic (row-shifted A per tap, one B block per tap, first MMA
// of a tile overwrites), no TMA, no epilogue.  variant bit 0: tcgen05.commit to a barrier after every super-tile (like
// empty[stage] + tmem_full); bit 1: additionally WAIT for the commit of the super-tile `lag` iterations back before
// issuing (like the tmem_empty / empty[stage] round trips, without the other warps); bit 2: 17 extra warps spin on a
// never-completing mbarrier (the parked producer / epilogue warps of the real kernel).
// cycles[2b] = SM cycles, cycles[2b+1] = ns for `iters` super-tiles on CTA b.
// variant bits 3..5 switch single features of the pattern OFF: bit 3: same B block for every tap, bit 4: no K-step
// advance inside the swizzle row, bit 5: accumulate flag always 1.  KC compile-time and the tap / K loops unrolled like
// the real kernel (a rolled loop with run-time div/mod makes the issuing thread itself the bound: ~57 cycles per MMA).
template <int KC>
__global__ void __launch_bounds__(576, 1)
umma_pattern_probe_kernel(long long* __restrict__ cycles, int BN, int T, int iters, int variant, int lag,
                          int fill_random) {
    constexpr int kc = KC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[8];
    __shared__ uint64_t never_bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (i + 1 + blockIdx.x * 7919u) * 2654435761u;
        h ^= h >> 15;
        reinterpret_cast<uint32_t*>(smem)[i] = fill_random ? ((h & 0x83FF83FFu) | 0x3C003C00u) : 0u;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        mbar_init(&never_bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    __shared__ volatile int done_flag;
    if (threadIdx.x == 0) done_flag = 0;
    __syncthreads();
    if (warp == 1) {
        const uint32_t row_bytes = kc * 2;
        const uint32_t layout = (kc == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
        const uint64_t a_tmpl = make_smem_desc(smem_u32(smem), 10 * row_bytes, layout);
        const uint64_t b_tmpl = make_smem_desc(smem_u32(smem) + 48 * 1024, 8 * row_bytes, layout);
        const uint32_t a_hi = static_cast<uint32_t>(a_tmpl >> 32), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(a_tmpl), b_lo0 = static_cast<uint32_t>(b_tmpl);
        const uint32_t idesc = make_idesc_16(128, BN, 1);
        const uint32_t kRow16 = row_bytes >> 4, kTile16 = 16 * 10 * kRow16, b_tap16 = (BN * row_bytes) >> 4;
        const int nbuf = (512 / (T * BN)) < 4 ? (512 / (T * BN)) : 4;
        const uint32_t b_step = (variant & 8) ? 0u : b_tap16;
        const uint32_t k_step = (variant & 16) ? 0u : 2u;
        const uint32_t first_acc = (variant & 32) ? 1u : 0u;
        long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if ((variant & 2) && i >= lag) mbar_wait(&bars[(i - lag) & 7], ((i - lag) >> 3) & 1);
            const uint32_t d0 = tmem_base + (i % nbuf) * T * BN;
            for (int t = 0; t < T; ++t) {
                if (elect_one_sync()) {
                    uint32_t b_lo = b_lo0;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t a_tap = a_lo0 + t * kTile16 + ((tap / 3) * 10 + (tap % 3)) * kRow16;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            umma_f16_split(d0 + t * BN, a_tap + k_step * k, a_hi, b_lo + k_step * k, b_hi, idesc,
                                           (tap | k) != 0 ? 1u : first_acc);
                        b_lo += b_step;
                    }
                }
                __syncwarp();
            }
            if (variant & 1) {
                if (elect_one_sync()) umma_commit(&bars[i & 7]);
                __syncwarp();
            }
        }
        if (elect_one_sync()) umma_commit(&never_bar);
        __syncwarp();
        mbar_wait(&never_bar, 0);
        const long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (elect_one_sync()) {
            cycles[2 * blockIdx.x] = t1 - t0;
            cycles[2 * blockIdx.x + 1] = g1 - g0;
            done_flag = 1;
        }
    } else if (variant & 4) {
        // parked warps: poll a barrier word in shared memory like the waiting roles of the real kernel
        while (!done_flag) {
            mbar_try_wait(&bars[7], 1);          // phase 1 of a barrier nobody completes twice: returns false quickly
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// mbarrier hop-latency probe (diagnostic): warp 1 signals barrier A (plain arrive, or tcgen05.commit when mode & 1) and
// waits on barrier B; warp 0 (or, mode & 4, all of warps 2-3 as well, 32 lanes each) waits on A and arrives on B.
// mode & 2: poll with mbarrier.test_wait instead of the (possibly suspending) try_wait.  cycles[0] / iters = round trip.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(128, 1) sync_probe_kernel(long long* __restrict__ cycles, int iters, int mode) {
    __shared__ uint64_t bar_a, bar_b;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    const int waiters = (mode & 4) ? 3 : 1;
    if (threadIdx.x == 0) {
        mbar_init(&bar_a, 1);
        mbar_init(&bar_b, waiters);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_ptr, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const bool poll = mode & 2;
    if (warp == 1) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one_sync()) {
                if (mode & 1) umma_commit(&bar_a); else mbar_arrive(&bar_a);
            }
            __syncwarp();
            if (poll) { while (!mbar_test_wait(&bar_b, i & 1)) {} } else mbar_wait(&bar_b, i & 1);
        }
        const long long t1 = clock64();
        if (elect_one_sync()) cycles[0] = t1 - t0;
    } else if (warp == 0 || ((mode & 4) && warp >= 2)) {
        for (int i = 0; i < iters; ++i) {
            if (poll) { while (!mbar_test_wait(&bar_a, i & 1)) {} } else mbar_wait(&bar_a, i & 1);
            __syncwarp();
            if (elect_one_sync()) mbar_arrive(&bar_b);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_ptr, 32);
    }
}

// TMEM read-rate / shuffle-rate probe (diagnostic): `nwarps` warps (multiple of 4: warp w reads TMEM lane quarter w % 4) each
// issue `iters` x 4 back-to-back loads of 16 accumulator columns (tcgen05.ld.32x32b.x16 = 2 KB per warp instruction), one
// tcgen05.wait::ld per four loads; mode 1: 64 fp32 shuffles (shfl.sync.down) per iteration instead; mode 2: both.
// cycles[0] = SM cycles of the slowest warp from a common start to its end.  Answers: is the TMEM read path 64 B/clk per SM
// or per lane quarter (a 96-column accumulator is 48 KB per tile), and what do 64 shuffles per tile and warp cost?
__global__ void __launch_bounds__(512, 1) tmem_ld_probe_kernel(long long* __restrict__ cycles, int iters, int mode,
                                                               float* __restrict__ sink) {
    __shared__ uint32_t tmem_ptr;
    __shared__ long long t_end[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = static_cast<float>(lane);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode != 1) {
            uint32_t r0[16], r1[16], r2[16], r3[16];
            const uint32_t c = static_cast<uint32_t>((i * 64 + (warp >> 2) * 128) & 448);
            tmem_ld_32x32b_x16(base + c, r0);
            tmem_ld_32x32b_x16(base + c + 16, r1);
            tmem_ld_32x32b_x16(base + c + 32, r2);
            tmem_ld_32x32b_x16(base + c + 48, r3);
            tmem_ld_wait();
            acc += __uint_as_float(r0[0] ^ r1[5] ^ r2[9] ^ r3[15]);
        }
        if (mode != 0) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = acc + static_cast<float>(j);
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __shfl_down_sync(0xffffffffu, v[(j + 1) & 15], 1);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += v[j];
        }
    }
    const long long t1 = clock64();
    if (lane == 0) t_end[warp] = t1 - t0;
    if (acc == 123.456f) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) m = t_end[w] > m ? t_end[w] : m;
        cycles[blockIdx.x] = m;
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_ptr, 512);
    }
}

// Launch-gap probe: a chain of dependent kernels of `ctas` CTAs x 128 threads, each spinning `spin` clock cycles after the
// programmatic-dependency wait (griddepcontrol.wait is a no-op for a launch without the attribute).
__global__ void launch_gap_probe_kernel(int* sink, int spin) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1);
}

}  // namespace aesr
