// Memory-bound kernels around the tensor-core convs: network head/tail convs with 1 input or 1 output channel
// (CUDA cores, nothing for a tensor core to do there), latent interpolation + layout change, and the
// image-domain utilities.  All coalesced / 16-byte vectorised; grids sized from the problem, no smem unless it buys
// coalescing (the NCHW <-> NHWC transposes).  FP16 template flag = 16-bit activation format (fp16 or bf16).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace aesr {

template <bool FP16>
__device__ __forceinline__ uint32_t pack2_t(float lo, float hi) {
    if (FP16) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16x2(lo, hi);
}
template <bool FP16>
__device__ __forceinline__ float2 unpack2_t(uint32_t u) {
    if (FP16) return __half22float2(*reinterpret_cast<__half2*>(&u));
    return make_float2(bf16_lo(u), bf16_hi(u));
}
template <bool FP16>
__device__ __forceinline__ uint16_t cvt16_t(float v) {
    if (FP16) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

// ---------------------------------------------------------------------------------------------------------------
// enc.0 : Conv2d(1, C, kernel 1, padding 1)   (networks/acai_vanilla.py:51)
// x fp32 [N,1,H,W]  ->  out 16-bit NHWC [N,H+2,W+2,C];  out = w[c] * xpad + b[c]  (ring pixels = bias)
// one thread = one output pixel x 8 channels (one 16-byte store)
// ---------------------------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void e0_conv1x1_pad1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                       const float* __restrict__ b, uint16_t* __restrict__ out, int N, int H, int W,
                                       int C) {
    const int Ho = H + 2, Wo = W + 2;
    const int groups = C >> 3;
    const size_t total = static_cast<size_t>(N) * Ho * Wo * groups;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        const size_t pix = i / groups;
        const int xo = static_cast<int>(pix % Wo);
        const int yo = static_cast<int>((pix / Wo) % Ho);
        const int n = static_cast<int>(pix / (static_cast<size_t>(Wo) * Ho));
        const int yi = yo - 1, xi = xo - 1;
        const float xv = (yi >= 0 && yi < H && xi >= 0 && xi < W)
                             ? __ldg(x + (static_cast<size_t>(n) * H + yi) * W + xi) : 0.f;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(__ldg(w + g * 8 + j), xv, __ldg(b + g * 8 + j));
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2_t<FP16>(v[0], v[1]), pack2_t<FP16>(v[2], v[3]),
                                                      pack2_t<FP16>(v[4], v[5]), pack2_t<FP16>(v[6], v[7]));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dec.14 + dec.15 : Conv2d(C, 1, 3, padding 1) + Sigmoid  (networks/acai_vanilla.py:98), C = 32
// in 16-bit NHWC [N,H,W,32] -> out fp32 image n at out + slot(n) * stride, clamped to [0,1] like
// generate_hr_volumes.py:67 (a no-op after the sigmoid).  apply_sigmoid = 0 writes the raw logit.
// One thread per output pixel; the 3x3x32 filter sits in shared memory as fp32.
// ---------------------------------------------------------------------------------------------------------------
template <int C, bool FP16>
__global__ void head_conv3x3_sigmoid_kernel(const uint16_t* __restrict__ in, const float* __restrict__ w /*[9][C]*/,
                                            const float* __restrict__ bias_ptr, float* __restrict__ out, const int* __restrict__ out_index,
                                            int N, int H, int W, size_t out_image_stride, int apply_sigmoid) {
    __shared__ float sw[9 * C];
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const size_t total = static_cast<size_t>(N) * H * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int n = static_cast<int>(i / (static_cast<size_t>(W) * H));
        float acc = __ldg(bias_ptr);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int xx = x + dx - 1;
                if (xx < 0 || xx >= W) continue;
                const uint4* src = reinterpret_cast<const uint4*>(
                    in + ((static_cast<size_t>(n) * H + yy) * W + xx) * C);
                const float* wt = sw + (dy * 3 + dx) * C;
#pragma unroll
                for (int j4 = 0; j4 < C / 8; ++j4) {
                    const uint4 m = __ldg(src + j4);
                    const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = unpack2_t<FP16>(u[k]);
                        acc = fmaf(f.x, wt[j4 * 8 + k * 2], acc);
                        acc = fmaf(f.y, wt[j4 * 8 + k * 2 + 1], acc);
                    }
                }
            }
        }
        float r = acc;
        if (apply_sigmoid) {
            r = 1.f / (1.f + __expf(-acc));
            r = fminf(fmaxf(r, 0.f), 1.f);
        }
        const size_t slot = out_index ? static_cast<size_t>(__ldg(out_index + n)) : static_cast<size_t>(n);
        out[slot * out_image_stride + static_cast<size_t>(y) * W + x] = r;
    }
}

// Shared-memory tiled version of the same head conv: one block = 16x16 output pixels of one image.  The 18x18x32 input
// window is staged once (coalesced 16-byte loads) with an 80-byte pixel pitch (64 B of channels + 16 B pad: the eight
// lanes of an LDS.128 phase then hit eight distinct 4-bank groups), so each input byte is fetched from L2 ~1.27x instead
// of 9x and the inner loop runs from conflict-free shared memory.
template <bool FP16>
__global__ void __launch_bounds__(256)
head_conv3x3_tiled_kernel(const uint16_t* __restrict__ in, const float* __restrict__ w /*[9][32]*/,
                          const float* __restrict__ bias_ptr, float* __restrict__ out,
                          const int* __restrict__ out_index, int H, int W, size_t out_image_stride,
                          int apply_sigmoid) {
    constexpr int C = 32, TW = 16, TH = 16, PITCH = 80;
    __shared__ __align__(16) uint8_t tile[(TH + 2) * (TW + 2) * PITCH];
    __shared__ float sw[9 * C];
    const int n = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    for (int i = threadIdx.x; i < 9 * C; i += 256) sw[i] = w[i];
    const uint16_t* img = in + static_cast<size_t>(n) * H * W * C;
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2) * 4; i += 256) {
        const int chunk = i & 3, px = i >> 2;
        const int hx = px % (TW + 2), hy = px / (TW + 2);
        const int gx = x0 + hx - 1, gy = y0 + hy - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gx >= 0 && gx < W && gy >= 0 && gy < H)
            v = __ldg(reinterpret_cast<const uint4*>(img + (static_cast<size_t>(gy) * W + gx) * C) + chunk);
        *reinterpret_cast<uint4*>(tile + px * PITCH + chunk * 16) = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x = x0 + tx, y = y0 + ty;
    float acc0 = __ldg(bias_ptr), acc1 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const uint8_t* src = tile + ((ty + t / 3) * (TW + 2) + tx + t % 3) * PITCH;
        const float* wt = sw + t * C;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 m = *reinterpret_cast<const uint4*>(src + j4 * 16);
            const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2_t<FP16>(u[k]);
                acc0 = fmaf(f.x, wt[j4 * 8 + k * 2], acc0);
                acc1 = fmaf(f.y, wt[j4 * 8 + k * 2 + 1], acc1);
            }
        }
    }
    if (x < W && y < H) {
        float r = acc0 + acc1;
        if (apply_sigmoid) {
            r = 1.f / (1.f + __expf(-r));
            r = fminf(fmaxf(r, 0.f), 1.f);
        }
        const size_t slot = out_index ? static_cast<size_t>(__ldg(out_index + n)) : static_cast<size_t>(n);
        out[slot * out_image_stride + static_cast<size_t>(y) * W + x] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// encoder stem: enc.0 Conv2d(1,32,1,padding=1) and enc.1 Conv2d(32,32,3,padding=1) are both linear with nothing in
// between (networks/acai_vanilla.py:51,55), so their composition is ONE 3x3 conv on the single input channel:
//   a1[y,x,co] = b1[co] + sum_tap inside(tap) * ( Weff[tap][co] * xpad[tap] + Beff[tap][co] )
//   Weff[tap][co] = sum_ci W1[co][ci][tap] * w0[ci],   Beff[tap][co] = sum_ci W1[co][ci][tap] * b0[ci]
// on the (H+2)x(W+2) grid enc.0 produces; inside(tap) = the tap's pixel lies on that grid (enc.1's zero padding),
// xpad = the image zero-padded by one pixel (enc.0's padding: ring pixels carry only the bias b0).
// Replaces a 64 B/px activation round trip and a 32->32 tensor-core conv (smem-bandwidth-bound at N = 32) by
// 4 B/px read + 64 B/px write.  stem_fold_kernel forms Weff/Beff once per parameter version.
// ---------------------------------------------------------------------------------------------------------------
__global__ void stem_fold_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                 const float* __restrict__ w1 /*[32][32][3][3]*/, float* __restrict__ weff /*[9][32]*/,
                                 float* __restrict__ beff /*[9][32]*/, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * C) return;
    const int co = i % C, tap = i / C;
    float sw = 0.f, sb = 0.f;
    for (int ci = 0; ci < C; ++ci) {
        const float w = w1[(static_cast<size_t>(co) * C + ci) * 9 + tap];
        sw = fmaf(w, w0[ci], sw);
        sb = fmaf(w, b0[ci], sb);
    }
    weff[i] = sw;
    beff[i] = sb;
}

// Folded stem parameters BY VALUE (kernel parameters live in the constant bank: the 288 MACs per pixel read their
// weights as instruction operands, no shared-memory traffic).  ball = b1 + sum of all nine beff taps (interior pixels).
struct StemParams {
    alignas(16) float weff[9 * 32];
    alignas(16) float beff[9 * 32];
    alignas(16) float b1[32];
    alignas(16) float ball[32];
};

// One thread = one output pixel position in STEM_P = 4 images x 32 channels (two passes of 16 channels).
//  * filter reuse: the folded filter reaches the FMAs through uniform registers (LDCU from the constant bank); with one
//    pixel per thread the 72 LDCU.128 per pixel bounded the kernel (measured 1.5 TB/s, LDCU : FFMA2 = 1.3 : 1 in the
//    SASS) -- four images at the same pixel position share every loaded filter value (and all border predicates).
//  * coalesced stores: the 32 pixels of a warp are one contiguous 2 KB block of the NHWC output per image; each warp
//    transposes its blocks through shared memory (XOR-swizzled, conflict-free both ways) and stores fully coalesced
//    512-byte rows instead of 16 bytes per lane at a 64-byte stride.
constexpr int STEM_P = 4;
template <bool FP16>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ x, const __grid_constant__ StemParams sp, uint16_t* __restrict__ out, int N,
                 int H, int W, float slope) {
    constexpr int C = 32, P = STEM_P;
    __shared__ uint4 stage[4][P][128];
    const int Ho = H + 2, Wo = W + 2;
    const int total = Ho * Wo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int wpix0 = blockIdx.x * blockDim.x + warp * 32;          // first pixel of this warp
    if (wpix0 >= total) return;                                      // whole warp out of range
    const bool valid = pix < total;
    const int yo = valid ? pix / Wo : 0, xo = valid ? pix - yo * Wo : 0;
    const bool interior = yo >= 1 && yo <= H && xo >= 1 && xo <= W;
    const int n_chunks = min(32, total - wpix0) * 4;                 // 16-byte chunks this warp owns per image
    const int sw = (lane >> 1) & 3;
    bool on_grid[9];
    int off[9];                                                      // image offset of the tap, -1 = outside the image
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = yo + tap / 3 - 1, xx = xo + tap % 3 - 1;            // position on the (H+2)x(W+2) grid
        on_grid[tap] = yy >= 0 && yy < Ho && xx >= 0 && xx < Wo;          // else: enc.1 zero padding
        const int yi = yy - 1, xi = xx - 1;                                // position in the image
        off[tap] = (valid && yi >= 0 && yi < H && xi >= 0 && xi < W) ? yi * W + xi : -1;
    }
    for (int n0 = blockIdx.y * P; n0 < N; n0 += gridDim.y * P) {
        const int np = min(P, N - n0);
        float2 xv[P][9];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const float* img = x + static_cast<size_t>(n0 + i) * H * W;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float v = (i < np && off[tap] >= 0) ? __ldg(img + off[tap]) : 0.f;
                xv[i][tap] = make_float2(v, v);
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            constexpr int CH = C / 2;
            const int c0 = half * CH;
            float2 acc[P][CH / 2];
            {
                float2 a0[CH / 2];
                if (interior) {
#pragma unroll
                    for (int j = 0; j < CH / 2; ++j) a0[j] = make_float2(sp.ball[c0 + 2 * j], sp.ball[c0 + 2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < CH / 2; ++j) a0[j] = make_float2(sp.b1[c0 + 2 * j], sp.b1[c0 + 2 * j + 1]);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (on_grid[tap]) {
#pragma unroll
                            for (int j = 0; j < CH / 2; ++j)
                                a0[j] = __fadd2_rn(a0[j], make_float2(sp.beff[tap * C + c0 + 2 * j],
                                                                      sp.beff[tap * C + c0 + 2 * j + 1]));
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j < CH / 2; ++j) acc[i][j] = a0[j];
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int j = 0; j < CH / 2; ++j) {
                    const float2 wv = make_float2(sp.weff[tap * C + c0 + 2 * j], sp.weff[tap * C + c0 + 2 * j + 1]);
#pragma unroll
                    for (int i = 0; i < P; ++i) acc[i][j] = __ffma2_rn(wv, xv[i][tap], acc[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                uint32_t pk[CH / 2];
#pragma unroll
                for (int j = 0; j < CH / 2; ++j)
                    pk[j] = pack2_t<FP16>(fmaxf(acc[i][j].x, acc[i][j].x * slope), fmaxf(acc[i][j].y, acc[i][j].y * slope));
                uint4* st = stage[warp][i];
                st[lane * 4 + ((2 * half) ^ sw)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                st[lane * 4 + ((2 * half + 1) ^ sw)] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        __syncwarp();
        for (int i = 0; i < np; ++i) {
            uint4* o4 = reinterpret_cast<uint4*>(out + (static_cast<size_t>(n0 + i) * total + wpix0) * C);
            const uint4* st = stage[warp][i];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 32 + lane;                               // linear 16-byte chunk of the warp's block
                const int px = c >> 2, j = c & 3;
                if (c < n_chunks) o4[c] = st[px * 4 + (j ^ ((px >> 1) & 3))];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same stem on warp-level tensor-core fragments (mma.sync m16n8k8, tf32 operands, fp32 accumulate).  The CUDA-core
// kernel above executes 288 fp32 MACs per output pixel and was bound by instruction issue at 38 % of the HBM roofline
// (profiles/r01q_memory_kernels_full.md); the work is a [pixels x 9] x [9 x 32] product, far too thin for a tcgen05 tile,
// but four HMMA steps per 16 pixels make it disappear behind the 64 B/px store stream.
//   D[16 px x 32 ch] = bias(px) + A[16 x 32] * B[32 x 32],  K = 32 slots = { xhi*Whi, xlo*Whi, xhi*Wlo } over the 9 taps
//   (x = xhi + xlo, W = Whi + Wlo split into tf32 halves: the dropped xlo*Wlo term is ~2^-22 relative, i.e. fp32-exact
//   for a result that is rounded to 16 bit anyway -- no range restriction on x, unlike an fp16 operand).
//   K slots are dealt to the four threads of a quad so that thread t only needs taps 2t, 2t+1 (+ tap 8):
//     k-step 0: xhi[2t], xhi[2t+1] * Whi     k-step 1: xlo[2t], xlo[2t+1] * Whi     k-step 2: xhi[2t], xhi[2t+1] * Wlo
//     k-step 3: tap 8 -- t = 0: xhi*Whi, t = 1: xlo*Whi, t = 2: xhi*Wlo, t = 3: unused; upper half unused
//   bias(px) = b1 + sum of beff over the taps on the (H+2)x(W+2) grid: nine border classes (first / inner / last row x
//   column), formed in fp32 on the host and used as the accumulator's initial value -- exact.
//   Channels are permuted across the four n-tiles (column c of n-tile nt = channel 8*(c/2) + 2*nt + c%2), so that thread
//   (g, t) of the accumulator fragment owns channels 8t..8t+7 of pixels g and g+8: one 16-byte store per pixel, a quad
//   writes a pixel's 64 bytes, a warp 2 x 512 contiguous bytes.
// One warp = 16 consecutive output pixels (flat index on the (H+2)x(W+2) grid) of every STEM_MMA-strided image; all
// pixel geometry (tap offsets, predicates, bias) is loop-invariant, the next image's taps are prefetched.
// ---------------------------------------------------------------------------------------------------------------
struct StemMmaParams {
    alignas(16) float weff[9 * 32];
    alignas(16) float bias_tab[9 * 32];      // [row class * 3 + column class][channel], class 0 first, 1 inner, 2 last
};

__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                 uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// tf32 split: hi = the 19 bits the tensor core reads, lo = the (exact) remainder, itself cut to tf32
__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) & 0xFFFFE000u;
}

template <bool FP16, int TERMS>      // TERMS = 3: xhi*Whi + xlo*Whi + xhi*Wlo (fp32-exact); 1: xhi*Whi only (profiling)
__global__ void __launch_bounds__(128)
stem_mma_kernel(const float* __restrict__ x, const __grid_constant__ StemMmaParams sp, uint16_t* __restrict__ out, int N,
                int H, int W, float slope) {
    const int Ho = H + 2, Wo = W + 2;
    const int total = Ho * Wo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int blk = blockIdx.x * 4 + warp;                 // 16-pixel block of the output grid
    if (blk * 16 >= total) return;
    // B fragments (loop-invariant): b0 <-> k = t, b1 <-> k = t + 4 of each k-step; n = g of n-tile nt
    uint32_t bhiA[4], bhiB[4], bloA[4], bloB[4], b8[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int ch = 8 * (g >> 1) + 2 * nt + (g & 1);
        uint32_t h, l;
        tf32_split(sp.weff[(2 * tq) * 32 + ch], h, l);
        bhiA[nt] = h; bloA[nt] = l;
        tf32_split(sp.weff[(2 * tq + 1) * 32 + ch], h, l);
        bhiB[nt] = h; bloB[nt] = l;
        tf32_split(sp.weff[8 * 32 + ch], h, l);
        b8[nt] = (tq <= 1) ? h : (tq == 2) ? l : 0u;
    }
    // pixel geometry of this thread's two pixels (g and g + 8 of the block)
    int off[2][3];                                          // image offsets of taps 2t, 2t+1, 8 (-1: reads zero)
    bool valid[2];
    float binit[2][8];                                      // accumulator start = bias of the pixel's border class, channels 8t..8t+7
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int pix = blk * 16 + g + 8 * r;
        valid[r] = pix < total;
        const int yo = valid[r] ? pix / Wo : 1, xo = valid[r] ? pix - yo * Wo : 1;
        const int taps[3] = {2 * tq, 2 * tq + 1, 8};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int yi = yo + taps[i] / 3 - 2, xi = xo + taps[i] % 3 - 2;        // position in the image
            off[r][i] = (valid[r] && yi >= 0 && yi < H && xi >= 0 && xi < W) ? yi * W + xi : -1;
        }
        const int cls = ((yo == 0) ? 0 : (yo == Ho - 1) ? 2 : 1) * 3 + ((xo == 0) ? 0 : (xo == Wo - 1) ? 2 : 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) binit[r][j] = sp.bias_tab[cls * 32 + 8 * tq + j];
    }
    const size_t img = static_cast<size_t>(H) * W;
    float xv[2][3];
    int n = blockIdx.y;
    if (n < N) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 3; ++i) xv[r][i] = off[r][i] >= 0 ? __ldg(x + n * img + off[r][i]) : 0.f;
    }
    for (; n < N; n += gridDim.y) {
        uint32_t hi[2][3], lo[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 3; ++i) tf32_split(xv[r][i], hi[r][i], lo[r][i]);
        const int nn = n + gridDim.y;                       // next image's taps into registers ...
        if (nn < N) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 3; ++i) xv[r][i] = off[r][i] >= 0 ? __ldg(x + nn * img + off[r][i]) : 0.f;
        }
        // ... and the image four iterations ahead into L2: one iteration of work (~600 cycles) does not cover a DRAM
        // miss, ncu showed 45 % of the warp samples waiting on the first use of xv (profiles/r02a_stem_mma_full.md)
        const int np = n + 4 * gridDim.y;
        if (np < N) {
            if (off[0][0] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + np * img + off[0][0]));
            if (off[1][2] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + np * img + off[1][2]));
        }
        const uint32_t a8_0 = (tq == 1) ? lo[0][2] : (tq == 3) ? 0u : hi[0][2];
        const uint32_t a8_1 = (tq == 1) ? lo[1][2] : (tq == 3) ? 0u : hi[1][2];
        float d[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            d[nt][0] = binit[0][2 * nt]; d[nt][1] = binit[0][2 * nt + 1];
            d[nt][2] = binit[1][2 * nt]; d[nt][3] = binit[1][2 * nt + 1];
            mma_m16n8k8_tf32(d[nt], hi[0][0], hi[1][0], hi[0][1], hi[1][1], bhiA[nt], bhiB[nt]);
            if (TERMS == 3) {
                mma_m16n8k8_tf32(d[nt], lo[0][0], lo[1][0], lo[0][1], lo[1][1], bhiA[nt], bhiB[nt]);
                mma_m16n8k8_tf32(d[nt], hi[0][0], hi[1][0], hi[0][1], hi[1][1], bloA[nt], bloB[nt]);
            }
            mma_m16n8k8_tf32(d[nt], a8_0, a8_1, 0u, 0u, b8[nt], 0u);
        }
        uint16_t* o = out + (static_cast<size_t>(n) * total + blk * 16 + g) * 32 + 8 * tq;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t pk[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float v0 = d[nt][2 * r], v1 = d[nt][2 * r + 1];
                pk[nt] = pack2_t<FP16>(fmaxf(v0, v0 * slope), fmaxf(v1, v1 * slope));
            }
            if (valid[r]) *reinterpret_cast<uint4*>(o + r * 8 * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// head gather: second half of the decoder tail when dec.12 ran in OUT_SHUFFLE2_HEAD mode.  part fp32 [N,h,w,16] holds,
// per LOW-res pixel (yl,xl), the 4x4 patch (origin (2yl-1, 2xl-1)) of head-conv partial sums of its own 2x2 hi-res
// block.  Output pixel (Y,X) is covered by the patches of 2x2 low-res pixels; sum them, add the bias, sigmoid, clamp
// (networks/acai_vanilla.py:98, generate_hr_volumes.py:67) and write image n at out + slot(n) * stride.
// One thread = one low-res pixel = a 2x2 block of outputs.  16 B/hi-res px read + 4 B/px written.
// ---------------------------------------------------------------------------------------------------------------
// Block = 8 x 32 low-res pixels of one image.  The (8+2) x (32+2) window of patches is staged once through shared
// memory with fully coalesced 16-byte loads (a patch row of the window is one contiguous run of 64-byte patches) at a
// 17-float pixel pitch, so that the 16 scattered patch entries a thread needs come from conflict-free LDS instead of
// 16 global loads touching 32 sectors each (measured: the per-pixel version ran at 40 % of the HBM roofline).
constexpr int HG_TW = 32, HG_TH = 8, HG_PITCH = 17;
__global__ void __launch_bounds__(256)
head_gather_kernel(const float* __restrict__ part, const float* __restrict__ bias_ptr, float* __restrict__ out,
                   const int* __restrict__ out_index, int N, int h, int w, size_t out_image_stride, int apply_sigmoid) {
    __shared__ float tile[(HG_TH + 2) * (HG_TW + 2) * HG_PITCH];
    const float bias = __ldg(bias_ptr);
    const int W = 2 * w;
    const int x0 = blockIdx.x * HG_TW, y0 = blockIdx.y * HG_TH;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int xl = x0 + tx, yl = y0 + ty;
    for (int n = blockIdx.z; n < N; n += gridDim.z) {
        const float4* pn = reinterpret_cast<const float4*>(part + static_cast<size_t>(n) * h * w * 16);
        // stage the window: (HG_TH+2) rows x (HG_TW+2) pixels x 4 float4
        for (int i = threadIdx.x; i < (HG_TH + 2) * (HG_TW + 2) * 4; i += 256) {
            const int q = i & 3, px = i >> 2;
            const int hx = px % (HG_TW + 2), hy = px / (HG_TW + 2);
            const int gx = x0 + hx - 1, gy = y0 + hy - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = __ldg(pn + (static_cast<size_t>(gy) * w + gx) * 4 + q);
            float* d = tile + px * HG_PITCH + q * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        if (xl < w && yl < h) {
            // window pixel (hy, hx) = (ty + 1 + dy, tx + 1 + dx); out-of-image neighbours were staged as zeros
            const float* c = tile + ((ty + 1) * (HG_TW + 2) + tx + 1) * HG_PITCH;
            constexpr int ROW = (HG_TW + 2) * HG_PITCH;
            float r[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    // rows: own patch row 1+a, plus the neighbour above (its row 3) for a = 0 / below (row 0) for a = 1
                    const int dy = a ? 1 : -1, rn = a ? 0 : 3;
                    const int dx = b ? 1 : -1, cn = b ? 0 : 3;
                    float s = bias + c[(1 + a) * 4 + (1 + b)];
                    s += c[dy * ROW + rn * 4 + (1 + b)];
                    s += c[dx * HG_PITCH + (1 + a) * 4 + cn];
                    s += c[dy * ROW + dx * HG_PITCH + rn * 4 + cn];
                    if (apply_sigmoid) {
                        s = 1.f / (1.f + __expf(-s));
                        s = fminf(fmaxf(s, 0.f), 1.f);
                    }
                    r[a * 2 + b] = s;
                }
            }
            const size_t slot = out_index ? static_cast<size_t>(__ldg(out_index + n)) : static_cast<size_t>(n);
            float* o = out + slot * out_image_stride + static_cast<size_t>(2 * yl) * W + 2 * xl;
            *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
            *reinterpret_cast<float2*>(o + W) = make_float2(r[2], r[3]);
        }
        __syncthreads();
    }
}

// head gather, TMA version: the (8+2) x (32+2) window of 64-byte patches of a block comes in by ONE cp.async.bulk.tensor load
// (box {16 floats, 34, 10, 1}, SWIZZLE_64B, zero fill outside the image = the missing neighbours of border pixels), double-buffered
// over the images a block walks through, so nothing is staged by the threads (the version above spends ~27 of its ~150
// instructions per low-res pixel on LDG.128 + 4 scalar STS per patch, and 16 scalar LDS on reading it back).  A thread reads whole
// 16-byte patch rows: 12 LDS.128, conflict-free under the 64-byte swizzle (the 8 lanes of a quarter warp hit 8 distinct 16-byte
// slots).  Same summation order as head_gather_kernel: bit-identical output.
constexpr int HGT_WIN_BYTES = (HG_TH + 2) * (HG_TW + 2) * 64;                   // 21760
constexpr int HGT_BUF_BYTES = ((HGT_WIN_BYTES + 1023) / 1024) * 1024;           // 22528
constexpr int HGT_SMEM_BYTES = 1024 + 2 * HGT_BUF_BYTES + 64;
__device__ __forceinline__ float4 hgt_row(const uint8_t* win, int r, int c) {    // patch r of the window, patch row c (4 floats)
    return lds_f4(reinterpret_cast<const float*>(win + r * 64 + 16 * (c ^ ((r >> 1) & 3))));
}
__global__ void __launch_bounds__(256)
head_gather_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ bias_ptr, float* __restrict__ out,
                       const int* __restrict__ out_index, int N, int h, int w, size_t out_image_stride, int apply_sigmoid) {
    extern __shared__ uint8_t hgt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hgt_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * HGT_BUF_BYTES);
    const float bias = __ldg(bias_ptr);
    const int W = 2 * w;
    const int x0 = blockIdx.x * HG_TW, y0 = blockIdx.y * HG_TH;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int xl = x0 + tx, yl = y0 + ty;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int n = blockIdx.z;
    if (threadIdx.x == 0 && n < N) {
        mbar_arrive_expect_tx(&full[0], HGT_WIN_BYTES);
        tma_load_4d(smem, &tmap, &full[0], 0, x0 - 1, y0 - 1, n);
    }
    constexpr int ROW = HG_TW + 2;
    const int r0 = (ty + 1) * ROW + tx + 1;                      // this thread's own patch in the window
    for (int it = 0; n < N; n += gridDim.z, ++it) {
        const int b = it & 1;
        if (threadIdx.x == 0 && n + static_cast<int>(gridDim.z) < N) {          // the other buffer was read in iteration it - 1
            mbar_arrive_expect_tx(&full[b ^ 1], HGT_WIN_BYTES);
            tma_load_4d(smem + (b ^ 1) * HGT_BUF_BYTES, &tmap, &full[b ^ 1], 0, x0 - 1, y0 - 1, n + gridDim.z);
        }
        mbar_wait(&full[b], (it >> 1) & 1);
        if (xl < w && yl < h) {
            const uint8_t* win = smem + b * HGT_BUF_BYTES;
            // own rows 1, 2; the neighbours' rows / columns that overlap this pixel's 2x2 outputs (patch origin (2y-1, 2x-1))
            const float4 o1 = hgt_row(win, r0, 1), o2 = hgt_row(win, r0, 2);
            const float4 u3 = hgt_row(win, r0 - ROW, 3), d0 = hgt_row(win, r0 + ROW, 0);
            const float4 l1 = hgt_row(win, r0 - 1, 1), l2 = hgt_row(win, r0 - 1, 2);
            const float4 q1 = hgt_row(win, r0 + 1, 1), q2 = hgt_row(win, r0 + 1, 2);
            const float4 ul = hgt_row(win, r0 - ROW - 1, 3), ur = hgt_row(win, r0 - ROW + 1, 3);
            const float4 dl = hgt_row(win, r0 + ROW - 1, 0), dr = hgt_row(win, r0 + ROW + 1, 0);
            float r[4];
            // (a, b) = (0,0): own [1][1], up [3][1], left [1][3], up-left [3][3]   (same order as head_gather_kernel)
            r[0] = ((bias + o1.y) + u3.y + l1.w) + ul.w;
            r[1] = ((bias + o1.z) + u3.z + q1.x) + ur.x;        // (0,1): own [1][2], up [3][2], right [1][0], up-right [3][0]
            r[2] = ((bias + o2.y) + d0.y + l2.w) + dl.w;        // (1,0): own [2][1], down [0][1], left [2][3], down-left [0][3]
            r[3] = ((bias + o2.z) + d0.z + q2.x) + dr.x;        // (1,1): own [2][2], down [0][2], right [2][0], down-right [0][0]
            if (apply_sigmoid) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s = 1.f / (1.f + __expf(-r[i]));
                    r[i] = fminf(fmaxf(s, 0.f), 1.f);
                }
            }
            const size_t slot = out_index ? static_cast<size_t>(__ldg(out_index + n)) : static_cast<size_t>(n);
            float* o = out + slot * out_image_stride + static_cast<size_t>(2 * yl) * W + 2 * xl;
            *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
            *reinterpret_cast<float2*>(o + W) = make_float2(r[2], r[3]);
        }
        __syncthreads();                                         // the window may be refilled
    }
}

// ---------------------------------------------------------------------------------------------------------------
// latent interpolation + layout change
//   z    fp32 NCHW [*, C, HW]  (public latent layout)
//   out  16-bit NHWC [M, HW, C]  (decoder input), optionally also fp32 NCHW [M, C, HW] (the public z_mix)
//   out[m] = wa[m] * z[ia[m]] + wb[m] * z[ib[m]]   -- three separately rounded fp32 ops (mul, mul, add), exactly
//   what `alpha * latent_1 + (1 - alpha) * latent_2` does in torch (generate_hr_volumes.py:88); ib[m] < 0 => copy.
// 32 pixels x 32 channels per block through a padded smem tile: reads coalesced along pixels, writes along channels.
// ---------------------------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void lerp_nchw_to_nhwc_kernel(const float* __restrict__ z, const int* __restrict__ ia,
                                         const int* __restrict__ ib, const float* __restrict__ wa,
                                         const float* __restrict__ wb, uint16_t* __restrict__ out_nhwc,
                                         float* __restrict__ out_nchw, int C, int HW) {
    __shared__ float tile[32][33];
    const int m = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int a = ia[m], b = ib[m];
    const float fa = wa[m], fb = (b >= 0) ? wb[m] : 0.f;
    const float* za = z + static_cast<size_t>(a) * C * HW;
    const float* zb = z + static_cast<size_t>(b >= 0 ? b : a) * C * HW;
    for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
        const int c = c0 + cy, p = p0 + threadIdx.x;
        float v = 0.f;
        if (c < C && p < HW) {
            const size_t off = static_cast<size_t>(c) * HW + p;
            if (b >= 0) v = __fadd_rn(__fmul_rn(fa, za[off]), __fmul_rn(fb, zb[off]));
            else v = za[off];
            if (out_nchw) out_nchw[static_cast<size_t>(m) * C * HW + off] = v;
        }
        tile[cy][threadIdx.x] = v;
    }
    __syncthreads();
    for (int py = threadIdx.y; py < 32; py += blockDim.y) {
        const int p = p0 + py, c = c0 + threadIdx.x;
        if (p < HW && c < C)
            out_nhwc[(static_cast<size_t>(m) * HW + p) * C + c] = cvt16_t<FP16>(tile[threadIdx.x][py]);
    }
}

// All K alpha steps of one slice pair in one pass (volume synthesis): the two fp32 latents are read ONCE and K blended
// 16-bit NHWC latents are written, out[(p*K + k)] = wa[k] * z[pa[p]] + wb[k] * z[pb[p]]  (same three roundings).
// Algorithmic bytes per pair: 2 * 4 * C*HW read + K * 2 * C*HW written (vs 8 B read per OUTPUT element unfused).
template <bool FP16>
__global__ void lerp_pairs_kernel(const float* __restrict__ z, const int* __restrict__ pa, const int* __restrict__ pb,
                                  const float* __restrict__ wa, const float* __restrict__ wb,
                                  uint16_t* __restrict__ out_nhwc, int K, int C, int HW) {
    __shared__ float tile[32][33];
    const int p = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* za = z + static_cast<size_t>(pa[p]) * C * HW;
    const float* zb = z + static_cast<size_t>(pb[p]) * C * HW;
    float va[4], vb[4];                              // blockDim.y == 8: 4 channel rows per thread
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = c0 + threadIdx.y + 8 * r, px = p0 + threadIdx.x;
        const bool ok = c < C && px < HW;
        va[r] = ok ? za[static_cast<size_t>(c) * HW + px] : 0.f;
        vb[r] = ok ? zb[static_cast<size_t>(c) * HW + px] : 0.f;
    }
    for (int k = 0; k < K; ++k) {
        const float fa = wa[k], fb = wb[k];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            tile[threadIdx.y + 8 * r][threadIdx.x] = __fadd_rn(__fmul_rn(fa, va[r]), __fmul_rn(fb, vb[r]));
        __syncthreads();
        uint16_t* o = out_nhwc + (static_cast<size_t>(p) * K + k) * HW * C;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int px = p0 + threadIdx.y + 8 * r, c = c0 + threadIdx.x;
            if (px < HW && c < C) o[static_cast<size_t>(px) * C + c] = cvt16_t<FP16>(tile[threadIdx.x][threadIdx.y + 8 * r]);
        }
        __syncthreads();
    }
}

// Interpolation AFTER the decoder's first conv (volume synthesis, fused pipeline).  dec.0 is linear, so
//   dec.0(wa*z1 + wb*z2) = wa*dec.0_nobias(z1) + wb*dec.0_nobias(z2) + bias
// and the conv runs once per low-resolution slice instead of once per synthesized slice (Z vs (Z-1)*A launches of the
// same work).  pre = un-rounded fp32 NHWC accumulators [*, HW*C] of conv3x3(OUT_SAME_F32, no bias); this kernel forms
//   out[p*K + k] = LeakyReLU(wa[k] * pre[pa[p]] + wb[k] * pre[pb[p]] + bias[c])   -> NHWC 16-bit
// with the same three separately rounded lerp operations as lerp_pairs.  One thread = 8 consecutive channels of one
// pixel: two 32-byte reads per operand, one 16-byte store per alpha.
// Algorithmic bytes per pair: 2 * 4 * HW*C read + K * 2 * HW*C written.
template <bool FP16>
__global__ void __launch_bounds__(256)
lerp_pairs_act_kernel(const float* __restrict__ pre, const int* __restrict__ pa, const int* __restrict__ pb,
                      const float* __restrict__ wa, const float* __restrict__ wb, const float* __restrict__ bias,
                      uint16_t* __restrict__ out, int K, int C, int HWC, float slope) {
    const int p = blockIdx.y;
    const float4* a4 = reinterpret_cast<const float4*>(pre + static_cast<size_t>(__ldg(pa + p)) * HWC);
    const float4* b4 = reinterpret_cast<const float4*>(pre + static_cast<size_t>(__ldg(pb + p)) * HWC);
    uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * K * HWC);
    const int groups = HWC >> 3;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const float4 a0 = __ldg(a4 + 2 * g), a1 = __ldg(a4 + 2 * g + 1);
        const float4 b0 = __ldg(b4 + 2 * g), b1 = __ldg(b4 + 2 * g + 1);
        const int c = (g << 3) % C;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (bias != nullptr) {
            s0 = __ldg(reinterpret_cast<const float4*>(bias + c));
            s1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
        }
        const float va[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float vb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float sb[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        for (int k = 0; k < K; ++k) {
            const float fa = __ldg(wa + k), fb = __ldg(wb + k);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = __fadd_rn(__fadd_rn(__fmul_rn(fa, va[j]), __fmul_rn(fb, vb[j])), sb[j]);
                v[j] = fmaxf(t, t * slope);
            }
            o4[static_cast<size_t>(k) * groups + g] =
                make_uint4(pack2_t<FP16>(v[0], v[1]), pack2_t<FP16>(v[2], v[3]), pack2_t<FP16>(v[4], v[5]),
                           pack2_t<FP16>(v[6], v[7]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// kept (original) slices of the HR volume: dst[out_index[n]] = clamp(src[n], 0, 1)   (generate_hr_volumes.py:44,58-67)
// fp32 images of HW pixels, float4 vectorised when HW % 4 == 0.
// ---------------------------------------------------------------------------------------------------------------
__global__ void place_slices_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                    const int* __restrict__ out_index, int N, int HW, int do_clamp) {
    const int n = blockIdx.y;
    const size_t slot = out_index ? static_cast<size_t>(out_index[n]) : static_cast<size_t>(n);
    const float* s = src + static_cast<size_t>(n) * HW;
    float* d = dst + slot * HW;
    if ((HW & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(s);
        float4* d4 = reinterpret_cast<float4*>(d);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (HW >> 2); i += gridDim.x * blockDim.x) {
            float4 v = __ldg(s4 + i);
            if (do_clamp) {
                v.x = fminf(fmaxf(v.x, 0.f), 1.f); v.y = fminf(fmaxf(v.y, 0.f), 1.f);
                v.z = fminf(fmaxf(v.z, 0.f), 1.f); v.w = fminf(fmaxf(v.w, 0.f), 1.f);
            }
            d4[i] = v;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            float v = __ldg(s + i);
            if (do_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
            d[i] = v;
        }
    }
}

}  // namespace aesr
