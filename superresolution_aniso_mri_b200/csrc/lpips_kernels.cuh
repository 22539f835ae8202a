// LPIPS-VGG v0.1 pieces that are not 3x3 tensor-core convs (lpips/networks_basic.py:63-110, lpips/perceptual.py:19-33,
// lpips/pretrained_networks.py:97-135): the input preparation fused into conv1_1 (1 -> 3 -> 64 channels on CUDA
// cores), its backward, the max-pool backward, and the fused distance head forward / backward.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "elementwise.cuh"
#include "train_kernels.cuh"

namespace aesr {

// 8 consecutive 16-bit channels (one 16-byte load) as floats
template <bool AF>
__device__ __forceinline__ void load8(const uint16_t* p, float (&v)[8]) {
    const uint4 m = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = unpack2_t<AF>(u[k]);
        v[2 * k] = f.x;
        v[2 * k + 1] = f.y;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// conv1_1 with the LPIPS input pipeline folded in:
//   u = 2*img - 1 (perceptual.py:30-31, normalize=True) ; v_c = (u - shift_c) / scale_c for c in 0..2
//   (ScalingLayer, networks_basic.py:99-100: a 1-channel image broadcasts against the [1,3,1,1] buffers) ;
//   y = ReLU(conv3x3(v, W[64,3,3,3]) + b), zero padding applies to v (after the scaling).
// img fp32 [N,1,H,W] -> out 16-bit NHWC [N,H,W,64].  One thread = one pixel x 8 output channels.
// ---------------------------------------------------------------------------------------------------------------
// The three input channels are affine images of ONE image (v_c = (u - shift_c) / scale_c), so the 3 -> 64 conv collapses to
// a 1 -> 64 conv on img with a border-dependent bias (zero padding applies to v, i.e. an off-image tap contributes
// nothing, not even its shift term):
//   y[co] = bcls[class(y,x)][co] + sum_{t on image} W2[t][co] * img_t ,   W2 = s * sum_c W[co][c][t] / scale_c   (s = 2 if normalize)
//   bcls[class][co] = b[co] - sum_{t valid for the class} ( sum_c W[co][c][t] * shift_c / scale_c  [+ sum_c W/scale_c if normalize] )
// with nine classes (first / inner / last row x column) like the encoder stem.  72 instead of 216 MACs per thread and the
// filter read as 16-byte shared-memory words (the 3-channel version did one LDS per MAC: 231 us for 24 images).
template <bool AF>
__global__ void vgg_conv1_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w /*[64][3][3][3]*/,
                                     const float* __restrict__ b, uint16_t* __restrict__ out, int N, int H, int W,
                                     float sh0, float sh1, float sh2, float sc0, float sc1, float sc2, int normalize) {
    __shared__ __align__(16) float sW[9][64];       // W2[t][co]
    __shared__ float sK[9][64];                      // per-tap constant that an on-image tap subtracts
    __shared__ __align__(16) float sB[9][64];       // bias per border class
    const float shf[3] = {sh0, sh1, sh2}, scv[3] = {sc0, sc1, sc2};
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int t = i / 64, co = i % 64;
        float we = 0.f, ws = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float wv = w[co * 27 + c * 9 + t];
            we += wv / scv[c];
            ws += wv * shf[c] / scv[c];
        }
        sW[t][co] = normalize ? 2.f * we : we;
        sK[t][co] = normalize ? ws + we : ws;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int cls = i / 64, co = i % 64, cy = cls / 3, cx = cls % 3;
        float acc = b[co];
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
            // H == 1 / W == 1: a pixel is first and last at once; class 0 then also drops the far tap (handled below)
            const bool ok = !(cy == 0 && dy == 0) && !(cy == 2 && dy == 2) && !(cx == 0 && dx == 0) && !(cx == 2 && dx == 2);
            if (ok) acc -= sK[t][co];
        }
        sB[cls][co] = acc;
    }
    __syncthreads();
    const size_t total = static_cast<size_t>(N) * H * W * 8;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int grp = static_cast<int>(i & 7);
        const size_t p = i >> 3;
        const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
        const size_t nb = (p / (static_cast<size_t>(W) * H)) * H * W;
        float o[8];
        const bool degenerate = H < 2 || W < 2;          // single-row / single-column images: generic per-tap bias
        if (!degenerate) {
            const int cls = ((y == 0) ? 0 : (y == H - 1) ? 2 : 1) * 3 + ((x == 0) ? 0 : (x == W - 1) ? 2 : 1);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[cls][grp * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[cls][grp * 8 + 4]);
            o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = b[grp * 8 + j];
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            if (!in) continue;
            const float u = img[nb + static_cast<size_t>(yy) * W + xx];
            const float4 w0 = *reinterpret_cast<const float4*>(&sW[t][grp * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&sW[t][grp * 8 + 4]);
            o[0] = fmaf(w0.x, u, o[0]); o[1] = fmaf(w0.y, u, o[1]); o[2] = fmaf(w0.z, u, o[2]); o[3] = fmaf(w0.w, u, o[3]);
            o[4] = fmaf(w1.x, u, o[4]); o[5] = fmaf(w1.y, u, o[5]); o[6] = fmaf(w1.z, u, o[6]); o[7] = fmaf(w1.w, u, o[7]);
            if (degenerate) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] -= sK[t][grp * 8 + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2_t<AF>(o[0], o[1]), pack2_t<AF>(o[2], o[3]),
                                                      pack2_t<AF>(o[4], o[5]), pack2_t<AF>(o[6], o[7]));
    }
}
// Same conv, one thread = 8 output channels of FOUR consecutive pixels of a row: the 3 x 6 image window is loaded once
// (18 instead of 36 loads) and every 16-byte filter word read from shared memory feeds four pixels (18 instead of 72 LDS.128
// per quad) -- the one-pixel version was bound by its shared-memory reads (77 us for 24 images at 128^2, 0.4 % of the DRAM
// bandwidth, ncu profiles/r04b_train_elementwise_ncu.md).  H, W >= 2 (the host falls back to the kernel above otherwise).
template <bool AF>
__global__ void __launch_bounds__(256)
vgg_conv1_fwd_quad_kernel(const float* __restrict__ img, const float* __restrict__ w /*[64][3][3][3]*/,
                          const float* __restrict__ b, uint16_t* __restrict__ out, int N, int H, int W,
                          float sh0, float sh1, float sh2, float sc0, float sc1, float sc2, int normalize) {
    __shared__ __align__(16) float sW[9][64];       // W2[t][co]
    __shared__ float sK[9][64];                      // per-tap constant that an on-image tap subtracts
    __shared__ __align__(16) float sB[9][64];       // bias per border class
    const float shf[3] = {sh0, sh1, sh2}, scv[3] = {sc0, sc1, sc2};
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int t = i / 64, co = i % 64;
        float we = 0.f, ws = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float wv = w[co * 27 + c * 9 + t];
            we += wv / scv[c];
            ws += wv * shf[c] / scv[c];
        }
        sW[t][co] = normalize ? 2.f * we : we;
        sK[t][co] = normalize ? ws + we : ws;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int cls = i / 64, co = i % 64, cy = cls / 3, cx = cls % 3;
        float acc = b[co];
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
            const bool ok = !(cy == 0 && dy == 0) && !(cy == 2 && dy == 2) && !(cx == 0 && dx == 0) && !(cx == 2 && dx == 2);
            if (ok) acc -= sK[t][co];
        }
        sB[cls][co] = acc;
    }
    __syncthreads();
    const int qx = (W + 3) >> 2;
    const size_t total = static_cast<size_t>(N) * H * qx * 8;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int grp = static_cast<int>(i & 7);
        const size_t q = i >> 3;
        const int x0 = static_cast<int>(q % qx) * 4, y = static_cast<int>((q / qx) % H);
        const size_t nb = (q / (static_cast<size_t>(qx) * H)) * H * W;
        float u[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = y + r - 1;
            const bool row_in = yy >= 0 && yy < H;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int xx = x0 + c - 1;
                u[r][c] = (row_in && xx >= 0 && xx < W) ? __ldg(img + nb + static_cast<size_t>(yy) * W + xx) : 0.f;
            }
        }
        float o[4][8];
        const int cy = (y == 0) ? 0 : (y == H - 1) ? 2 : 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + j;
            const int cls = cy * 3 + ((x == 0) ? 0 : (x >= W - 1) ? 2 : 1);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[cls][grp * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[cls][grp * 8 + 4]);
            o[j][0] = b0.x; o[j][1] = b0.y; o[j][2] = b0.z; o[j][3] = b0.w;
            o[j][4] = b1.x; o[j][5] = b1.y; o[j][6] = b1.z; o[j][7] = b1.w;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 w0 = *reinterpret_cast<const float4*>(&sW[t][grp * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&sW[t][grp * 8 + 4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float uv = u[t / 3][j + t % 3];          // off-image taps hold 0: they contribute nothing
                o[j][0] = fmaf(w0.x, uv, o[j][0]); o[j][1] = fmaf(w0.y, uv, o[j][1]);
                o[j][2] = fmaf(w0.z, uv, o[j][2]); o[j][3] = fmaf(w0.w, uv, o[j][3]);
                o[j][4] = fmaf(w1.x, uv, o[j][4]); o[j][5] = fmaf(w1.y, uv, o[j][5]);
                o[j][6] = fmaf(w1.z, uv, o[j][6]); o[j][7] = fmaf(w1.w, uv, o[j][7]);
            }
        }
        uint16_t* orow = out + (nb + static_cast<size_t>(y) * W + x0) * 64 + grp * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (x0 + j >= W) break;
#pragma unroll
            for (int k = 0; k < 8; ++k) o[j][k] = fmaxf(o[j][k], 0.f);
            *reinterpret_cast<uint4*>(orow + j * 64) = make_uint4(pack2_t<AF>(o[j][0], o[j][1]), pack2_t<AF>(o[j][2], o[j][3]),
                                                                  pack2_t<AF>(o[j][4], o[j][5]), pack2_t<AF>(o[j][6], o[j][7]));
        }
    }
}
// backward of the above w.r.t. the image: g bf16 [N,H,W,64] is dL/d(pre-ReLU conv1_1 output) (ReLU' already applied)
//   dimg[p] = (normalize ? 2 : 1) * sum_c (1/scale_c) * sum_{tap,co} W[co][c][tap] * g[p - off(tap)][co]
// Eight lanes per pixel, 8 channels (one 16-byte load) per lane and tap, three shuffles to fold the eight partial sums.
__global__ void vgg_conv1_bwd_kernel(const uint16_t* __restrict__ g, const float* __restrict__ w,
                                     float* __restrict__ dimg, int N, int H, int W, float sc0, float sc1, float sc2,
                                     int normalize, float out_scale) {
    __shared__ __align__(16) float swe[9 * 64];          // effective 1-channel filter: sum_c W[co][c][tap] / scale_c
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int t = i / 64, co = i % 64;
        swe[i] = w[co * 27 + 0 * 9 + t] / sc0 + w[co * 27 + 1 * 9 + t] / sc1 + w[co * 27 + 2 * 9 + t] / sc2;
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const size_t grp_id = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 3;
    const size_t ngrp = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 3;
    const size_t total = static_cast<size_t>(N) * H * W;
    const size_t rounds = (total + ngrp - 1) / ngrp;                          // warp-uniform trip count (shuffles inside)
    for (size_t it = 0; it < rounds; ++it) {
        const size_t p = it * ngrp + grp_id;
        const bool live = p < total;
        float acc = 0.f;
        if (live) {
            const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
            const size_t nb = (p / (static_cast<size_t>(W) * H)) * H * W;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                float gv[8];
                load8<false>(g + (nb + static_cast<size_t>(yy) * W + xx) * 64 + sub * 8, gv);
                const float4 w0 = *reinterpret_cast<const float4*>(&swe[t * 64 + sub * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&swe[t * 64 + sub * 8 + 4]);
                acc = fmaf(w0.x, gv[0], acc); acc = fmaf(w0.y, gv[1], acc); acc = fmaf(w0.z, gv[2], acc); acc = fmaf(w0.w, gv[3], acc);
                acc = fmaf(w1.x, gv[4], acc); acc = fmaf(w1.y, gv[5], acc); acc = fmaf(w1.z, gv[6], acc); acc = fmaf(w1.w, gv[7], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (live && sub == 0) dimg[p] = acc * (normalize ? 2.f : 1.f) * out_scale;
    }
}
// Quad version of the backward: one thread = 8 channels (lane & 7) of FOUR consecutive output pixels; the effective filter
// of the lane's 8 channels lives in 72 registers, the 3 x 6 window of g is loaded once (18 x 16 bytes for 4 outputs instead
// of 36), no shared-memory reads in the loop.
__global__ void __launch_bounds__(256, 2)
vgg_conv1_bwd_quad_kernel(const uint16_t* __restrict__ g, const float* __restrict__ w, float* __restrict__ dimg, int N,
                          int H, int W, float sc0, float sc1, float sc2, int normalize, float out_scale) {
    const int sub = threadIdx.x & 7;
    __shared__ float swe[9 * 64];                        // effective 1-channel filter, formed once per block
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int t = i / 64, co = i % 64;
        swe[i] = w[co * 27 + 0 * 9 + t] / sc0 + w[co * 27 + 1 * 9 + t] / sc1 + w[co * 27 + 2 * 9 + t] / sc2;
    }
    __syncthreads();
    float we[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k) we[t][k] = swe[t * 64 + sub * 8 + k];
    const int qx = (W + 3) >> 2;
    const size_t grp_id = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 3;
    const size_t ngrp = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 3;
    const size_t total = static_cast<size_t>(N) * H * qx;
    const size_t rounds = (total + ngrp - 1) / ngrp;                          // warp-uniform trip count (shuffles inside)
    const float osc = (normalize ? 2.f : 1.f) * out_scale;
    for (size_t it = 0; it < rounds; ++it) {
        const size_t q = it * ngrp + grp_id;
        const bool live = q < total;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int x0 = 0, y = 0;
        size_t nb = 0;
        if (live) {
            x0 = static_cast<int>(q % qx) * 4;
            y = static_cast<int>((q / qx) % H);
            nb = (q / (static_cast<size_t>(qx) * H)) * H * W;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int yy = y + r - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int xx = x0 + c - 1;
                    if (xx < 0 || xx >= W) continue;
                    float gv[8];
                    load8<false>(g + (nb + static_cast<size_t>(yy) * W + xx) * 64 + sub * 8, gv);
                    // g pixel (yy, xx) reaches output (y, x0 + j) through tap (2 - r, j - c + 2)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int tx = j - c + 2;
                        if (tx < 0 || tx > 2) continue;
                        const int t = (2 - r) * 3 + tx;
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[j] = fmaf(we[t][k], gv[k], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
        }
        if (live && sub < 4 && x0 + sub < W) {
            const float v = sub == 0 ? acc[0] : sub == 1 ? acc[1] : sub == 2 ? acc[2] : acc[3];
            dimg[nb + static_cast<size_t>(y) * W + x0 + sub] = v * osc;
        }
    }
}
// ---------------------------------------------------------------------------------------------------------------
// MaxPool2d(2) backward fused with the tap gradient and the ReLU' of the producing conv:
//   g_out[y,x,c] = relu'(a) * ( first_argmax(a over its 2x2 window) ? d_pooled[y/2,x/2,c] : 0  +  g_tap[y,x,c] )
// a: post-ReLU activation 16-bit [N,H,W,C]; d_pooled bf16 [N,H/2,W/2,C]; g_tap bf16 [N,H,W,C] or null.
// ---------------------------------------------------------------------------------------------------------------
// One thread = 8 channels (16 bytes) of one pixel: the element-per-thread version spent its time on 2-byte accesses and
// three integer divisions per element (0.23 ms of a 5 ms training step, bench.py kernel_ms).
template <bool AF>
__global__ void maxpool_bwd_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ d_pooled,
                                   const uint16_t* __restrict__ g_tap, uint16_t* __restrict__ g_out, int N, int H,
                                   int W, int C) {
    const int Ho = H / 2, Wo = W / 2, groups = C >> 3;
    const size_t total = static_cast<size_t>(N) * H * W * groups;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        const size_t p = i / groups;
        const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
        const int n = static_cast<int>(p / (static_cast<size_t>(W) * H));
        float av[8], gs[8];
        load8<AF>(a + i * 8, av);
        if (g_tap) load8<false>(g_tap + i * 8, gs);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) gs[k] = 0.f;
        }
        const int yo = y >> 1, xo = x >> 1;
        if (d_pooled != nullptr && yo < Ho && xo < Wo) {
            // torch picks the first maximum in window scan order (0,0),(0,1),(1,0),(1,1)
            const uint16_t* base = a + ((static_cast<size_t>(n) * H + 2 * yo) * W + 2 * xo) * C + g * 8;
            float w0[8], w1[8], w2[8], w3[8], dp[8];
            load8<AF>(base, w0);
            load8<AF>(base + C, w1);
            load8<AF>(base + static_cast<size_t>(W) * C, w2);
            load8<AF>(base + static_cast<size_t>(W) * C + C, w3);
            load8<false>(d_pooled + ((static_cast<size_t>(n) * Ho + yo) * Wo + xo) * C + g * 8, dp);
            const int me = (y & 1) * 2 + (x & 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int arg = 0;
                float best = w0[k];
                if (w1[k] > best) { best = w1[k]; arg = 1; }
                if (w2[k] > best) { best = w2[k]; arg = 2; }
                if (w3[k] > best) { best = w3[k]; arg = 3; }
                if (arg == me) gs[k] += dp[k];
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = pack_bf16x2(av[2 * k] > 0.f ? gs[2 * k] : 0.f, av[2 * k + 1] > 0.f ? gs[2 * k + 1] : 0.f);
        reinterpret_cast<uint4*>(g_out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LPIPS distance head of one tap (networks_basic.py:70-77 + lpips/common.py:12-14):
//   f = o / (||o||_2 + 1e-10) over channels ; val[n] += (1/HW) * sum_px sum_c lin_c (f0_c - f1_c)^2
// f0 = features of image set 0 (reference), f1 = set 1 (synthesized); both 16-bit NHWC [N,HW,C], C % 8 == 0, C <= 512.
// ---------------------------------------------------------------------------------------------------------------
// G = min(C/8, 32) lanes share a pixel, each lane owns 16-byte chunks of 8 channels (chunk index lane, lane+G, ...: at most
// two for C <= 512, kept in registers for both passes); 32/G pixels per warp.  The per-image sums are collected in shared
// memory and leave the block as one atomic per image (the one-warp-per-pixel version issued 2-byte loads and one global
// atomic per PIXEL onto N addresses: 0.53 ms of a 5 ms training step for ten launches, bench.py kernel_ms).
constexpr int LPIPS_MAX_IMG_SMEM = 256;
__device__ __forceinline__ float group_sum(float v, int G) {
    for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
template <bool AF>
__global__ void lpips_head_fwd_kernel(const uint16_t* __restrict__ o0, const uint16_t* __restrict__ o1,
                                      const float* __restrict__ lin, float* __restrict__ val, int N, int HW, int C) {
    __shared__ float s_val[LPIPS_MAX_IMG_SMEM];
    const bool use_smem = N <= LPIPS_MAX_IMG_SMEM;
    if (use_smem) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_val[i] = 0.f;
        __syncthreads();
    }
    const int chunks = C >> 3;
    const int G = chunks < 32 ? chunks : 32;
    const int lane = threadIdx.x & 31, sub = lane % G, ppw = 32 / G;
    const size_t warp_id = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    const size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t total = static_cast<size_t>(N) * HW;
    const float inv_hw = 1.f / static_cast<float>(HW);
    for (size_t base = warp_id * ppw; base < total; base += nwarps * ppw) {          // warp-uniform trip count
        const size_t p = base + lane / G;
        const bool live = p < total;
        float x0[2][8], x1[2][8];
        float n0 = 0.f, n1 = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ch = sub + r * G;
            if (live && ch < chunks) {
                load8<AF>(o0 + p * C + ch * 8, x0[r]);
                load8<AF>(o1 + p * C + ch * 8, x1[r]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) x0[r][k] = x1[r][k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                n0 = fmaf(x0[r][k], x0[r][k], n0);
                n1 = fmaf(x1[r][k], x1[r][k], n1);
            }
        }
        n0 = group_sum(n0, G);
        n1 = group_sum(n1, G);
        const float i0 = 1.f / (sqrtf(n0) + 1e-10f), i1 = 1.f / (sqrtf(n1) + 1e-10f);
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ch = sub + r * G;
            if (ch < chunks) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float d = x0[r][k] * i0 - x1[r][k] * i1;
                    acc = fmaf(__ldg(lin + ch * 8 + k) * d, d, acc);
                }
            }
        }
        acc = group_sum(acc, G);
        if (live && sub == 0) {
            const int n = static_cast<int>(p / HW);
            if (use_smem) atomicAdd(&s_val[n], acc * inv_hw);
            else atomicAdd(val + n, acc * inv_hw);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            if (s_val[i] != 0.f) atomicAdd(val + i, s_val[i]);
    }
}
// backward w.r.t. o1:  with e_c = 2 lin_c (f1_c - f0_c), s = sum_c e_c o1_c, r = ||o1||, q = r + eps
//   d val / d o1_j = e_j / q - o1_j * s / (r q^2)       ; times upstream[n] / HW.   Output bf16 NHWC.
// val != NULL: the forward value is accumulated in the same pass (lin_c d_c^2 = e_c^2 / (4 lin_c) needs nothing new:
// d_c = x0_c i0 - x1_c i1 is formed anyway) -- training calls forward and backward together, one read of the features.
// R = 16-byte chunks per lane: 1 for C <= 256 (the three large taps; halves the live registers: 72 -> fewer, more warps in
// flight for a kernel that waits on its loads), 2 for C = 512.
template <bool AF, int R>
__global__ void lpips_head_bwd_kernel(const uint16_t* __restrict__ o0, const uint16_t* __restrict__ o1,
                                      const float* __restrict__ lin, const float* __restrict__ upstream /*[N]*/,
                                      uint16_t* __restrict__ g1, float* __restrict__ val, int N, int HW, int C) {
    __shared__ float s_val[LPIPS_MAX_IMG_SMEM];
    const bool use_smem = val != nullptr && N <= LPIPS_MAX_IMG_SMEM;
    if (use_smem) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_val[i] = 0.f;
        __syncthreads();
    }
    const float inv_hw = 1.f / static_cast<float>(HW);
    const int chunks = C >> 3;
    const int G = chunks < 32 ? chunks : 32;
    const int lane = threadIdx.x & 31, sub = lane % G, ppw = 32 / G;
    const size_t warp_id = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    const size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t total = static_cast<size_t>(N) * HW;
    for (size_t base = warp_id * ppw; base < total; base += nwarps * ppw) {
        const size_t p = base + lane / G;
        const bool live = p < total;
        float x0[R][8], x1[R][8];
        float n0 = 0.f, n1 = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ch = sub + r * G;
            if (live && ch < chunks) {
                load8<AF>(o0 + p * C + ch * 8, x0[r]);
                load8<AF>(o1 + p * C + ch * 8, x1[r]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) x0[r][k] = x1[r][k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                n0 = fmaf(x0[r][k], x0[r][k], n0);
                n1 = fmaf(x1[r][k], x1[r][k], n1);
            }
        }
        n0 = group_sum(n0, G);
        n1 = group_sum(n1, G);
        const float r_ = sqrtf(n1);
        const float i0 = 1.f / (sqrtf(n0) + 1e-10f), q = r_ + 1e-10f, i1 = 1.f / q;
        float e[R][8];
        float sdot = 0.f, acc = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ch = sub + r * G;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float l = ch < chunks ? __ldg(lin + ch * 8 + k) : 0.f;
                const float d = x0[r][k] * i0 - x1[r][k] * i1;                 // same expression as the forward kernel
                acc = fmaf(l * d, d, acc);
                e[r][k] = 2.f * l * (x1[r][k] * i1 - x0[r][k] * i0);
                sdot = fmaf(e[r][k], x1[r][k], sdot);
            }
        }
        sdot = group_sum(sdot, G);
        if (val != nullptr) {
            acc = group_sum(acc, G);
            if (live && sub == 0) {
                const int n = static_cast<int>(p / HW);
                if (use_smem) atomicAdd(&s_val[n], acc * inv_hw);
                else atomicAdd(val + n, acc * inv_hw);
            }
        }
        if (!live) continue;                       // no shuffles below this point
        const float up = upstream[p / HW] / HW;
        const float kq = (r_ > 0.f) ? sdot / (r_ * q * q) : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ch = sub + r * G;
            if (ch < chunks) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    o[k] = pack_bf16x2(up * (e[r][2 * k] * i1 - x1[r][2 * k] * kq),
                                       up * (e[r][2 * k + 1] * i1 - x1[r][2 * k + 1] * kq));
                *reinterpret_cast<uint4*>(g1 + p * C + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            if (s_val[i] != 0.f) atomicAdd(val + i, s_val[i]);
    }
}

}  // namespace aesr
