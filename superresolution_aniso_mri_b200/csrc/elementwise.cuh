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
