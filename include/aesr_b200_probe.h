/* Diagnostic entry points (micro-benchmarks of tcgen05 / mbarrier behaviour behind the design decisions in DESIGN.md 4.1).
 * NOT part of the product library: libaesr_b200.so is built without them; `python -m superresolution_aniso_mri_b200.build
 * --probes` builds lib/libaesr_b200_probe.so (the same sources with -DAESR_WITH_PROBES) for tools/umma_rate.py,
 * tools/sync_probe.py and tools/gpu_diag.py. */
#ifndef AESR_B200_PROBE_H
#define AESR_B200_PROBE_H
#include "aesr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* DIAGNOSTIC (not on the product path): one 16x8 tile of a 64->64 bf16 conv computed from a single TMA halo load with
 * row-shifted UMMA descriptors; used by tools/gpu_diag.py to establish what the hardware's swizzle addressing does.
 * out fp32 [128][64] raw accumulators. */
int aesr_probe_halo_conv(const void* x, const void* w_packed, float* out, int N, int H, int W, int x0, int y0, int n,
                         int pitch, int variant, void* stream);

/* Diagnostic: `iters` back-to-back tcgen05.mma (M=128, N, K=16) per CTA on `grid` CTAs at once, A descriptor start shifted
 * by `shift_rows` rows, 8-row-group stride `pitch_rows`, advancing `a_advance_rows` rows between MMAs, rotating over `nacc`
 * TMEM accumulators; operands all-zero or (fill_random) pseudo-random fp16.  cycles[2*b] = SM cycles (clock64),
 * cycles[2*b+1] = wall-clock ns (globaltimer) of CTA b (tools/umma_rate.py). */
int aesr_probe_umma_rate(long long* cycles, int N, int kc, int pitch_rows, int shift_rows, int iters,
                         int a_advance_rows, int nacc, int grid, int fill_random, void* stream);

/* Diagnostic: the MMA issue loop of the halo conv kernel in isolation (T M-tiles x 9 taps x kc/16 K-steps per super-tile,
 * the kernel's descriptor arithmetic, no TMA, no epilogue).  variant bit 0: tcgen05.commit after every super-tile, bit 1:
 * wait for the commit `lag` super-tiles back before issuing, bit 2: 17 more warps polling an mbarrier.
 * cycles[2*b] = SM cycles, cycles[2*b+1] = ns for `iters` super-tiles on CTA b (tools/umma_rate.py --pattern). */
int aesr_probe_umma_pattern(long long* cycles, int BN, int kc, int T, int iters, int variant, int lag, int grid,
                            int fill_random, void* stream);

/* Diagnostic: mbarrier round-trip latency between two warps (mode bit 0: signal with tcgen05.commit, bit 1: poll with
 * test_wait instead of try_wait, bit 2: three waiting warps).  cycles[0] = total cycles for `iters` round trips. */
int aesr_probe_sync(long long* cycles, int iters, int mode, void* stream);

/* Diagnostic: TMEM read rate and shuffle rate.  `nwarps` (4, 8, 12, 16) warps per CTA each issue `iters` x 4 loads of 16
 * accumulator columns (2 KB per warp instruction) from their lane quarter (mode 0), `iters` x 64 fp32 shuffles (mode 1) or both
 * (mode 2); cycles[b] = SM cycles of the slowest warp of CTA b; sink = any device buffer of >= 512 floats (never written). */
int aesr_probe_tmem_ld(long long* cycles, int nwarps, int iters, int mode, int grid, float* sink, void* stream);

/* Diagnostic: device time per kernel of a CUDA graph holding a chain of `n_kernels` dependent launches (`ctas` CTAs of 128
 * threads, each spinning `spin_cycles`), with (`pdl` = 1) or without programmatic dependent launch; HOST pointer result. */
int aesr_probe_launch_gap(float* us_per_kernel, int n_kernels, int ctas, int spin_cycles, int pdl, int replays);

#ifdef __cplusplus
}
#endif
#endif /* AESR_B200_PROBE_H */
