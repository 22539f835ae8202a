// Training-step kernels around the tensor-core convs (all memory-bound or tiny): train-mode BatchNorm forward /
// backward fused with the pool / upsample that follows it, MSE, the 1-channel head / tail convs' backward, the
// weight-gradient conv (CUDA-core version), the latent mix backward and the fused Adam update.
//
// dtypes: activations A16 = fp16 or bf16 (template flag AF: true = fp16), gradients are ALWAYS bf16 (fp32 range:
// dL/dx of a mean-reduced loss is ~1e-6..1e-9, below fp16's normal range), reductions / parameters fp32.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "elementwise.cuh"

namespace aesr {

enum BnMode : int { BN_SAME = 0, BN_POOL = 1, BN_UP = 2 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float bf16_to_f(uint16_t u) { return __uint_as_float(static_cast<uint32_t>(u) << 16); }
template <bool AF>
__device__ __forceinline__ float a16_to_f(uint16_t u) {
    return AF ? __half2float(__ushort_as_half(u)) : bf16_to_f(u);
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm2d (training) -- networks/acai_vanilla.py:58,90.  The producing conv's epilogue accumulated
// stats[c] = sum a, stats[C+c] = sum a^2 over the `count` = N*H*W positions of the batch.
//   scale = gamma / sqrt(var_biased + eps), shift = beta - mean * scale          (normalisation of this pass)
//   running_mean = (1-m) rm + m mean ; running_var = (1-m) rv + m var * count/(count-1)   (torch semantics)
// ---------------------------------------------------------------------------------------------------------------
// `groups` (1 or 2) consecutive passes of the SAME BatchNorm module computed by one conv launch (images [0, split) and
// [split, N) of a merged batch, e.g. enc(x) and enc(slice_between) of one training step): stats / outputs are
// [groups][...]; the running statistics are updated once per group IN ORDER, exactly like two module calls.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, float count0, float count1, int groups,
                                   const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int g = 0; g < groups; ++g) {
        const float count = g == 0 ? count0 : count1;
        const float* st = stats + static_cast<size_t>(g) * 2 * C;
        const double mean = static_cast<double>(st[c]) / count;
        double var = static_cast<double>(st[C + c]) / count - mean * mean;
        if (var < 0) var = 0;
        const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        const float sc = gamma[c] * invstd;
        scale[g * C + c] = sc;
        shift[g * C + c] = beta[c] - static_cast<float>(mean) * sc;
        mean_out[g * C + c] = static_cast<float>(mean);
        invstd_out[g * C + c] = invstd;
        if (running_mean != nullptr) {
            const double unbiased = count > 1.f ? var * count / (count - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
        }
    }
}

// out = pool/up/same( a * scale + shift ): a [N,H,W,C] 16-bit -> out 16-bit.  One thread = 8 channels of one output px.
template <bool AF>
__global__ void bn_apply_kernel(const uint16_t* __restrict__ a, const float* __restrict__ scale,
                                const float* __restrict__ shift, uint16_t* __restrict__ out, int N, int H, int W, int C,
                                int mode, int split) {
    // images >= split belong to the second pass of a merged batch: scale / shift are [2][C]
    const int Ho = mode == BN_POOL ? H / 2 : mode == BN_UP ? 2 * H : H;
    const int Wo = mode == BN_POOL ? W / 2 : mode == BN_UP ? 2 * W : W;
    const int groups = C >> 3;
    const size_t total = static_cast<size_t>(N) * Ho * Wo * groups;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        const size_t pix = i / groups;
        const int xo = static_cast<int>(pix % Wo), yo = static_cast<int>((pix / Wo) % Ho);
        const int n = static_cast<int>(pix / (static_cast<size_t>(Wo) * Ho));
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int taps = mode == BN_POOL ? 4 : 1;
        for (int t = 0; t < taps; ++t) {
            const int yi = mode == BN_POOL ? 2 * yo + (t >> 1) : mode == BN_UP ? (yo >> 1) : yo;
            const int xi = mode == BN_POOL ? 2 * xo + (t & 1) : mode == BN_UP ? (xo >> 1) : xo;
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a + ((static_cast<size_t>(n) * H + yi) * W + xi) * C) + g);
            const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2_t<AF>(u[k]);
                v[2 * k] += f.x;
                v[2 * k + 1] += f.y;
            }
        }
        const float inv = mode == BN_POOL ? 0.25f : 1.f;
        const int go = (n >= split ? C : 0) + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] * inv, __ldg(scale + go + j), __ldg(shift + go + j));
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2_t<AF>(v[0], v[1]), pack2_t<AF>(v[2], v[3]),
                                                      pack2_t<AF>(v[4], v[5]), pack2_t<AF>(v[6], v[7]));
    }
}

// gradient w.r.t. the BN output at full resolution, derived from the gradient of the pooled / upsampled tensor
__device__ __forceinline__ float bn_dy_at(const uint16_t* __restrict__ dnext, int mode, int n, int y, int x, int c,
                                          int H, int W, int C) {
    if (mode == BN_POOL) {
        const int Ho = H / 2, Wo = W / 2, yo = y >> 1, xo = x >> 1;
        if (yo >= Ho || xo >= Wo) return 0.f;        // row / column dropped by the floor of AvgPool2d(2)
        return 0.25f * bf16_to_f(dnext[((static_cast<size_t>(n) * Ho + yo) * Wo + xo) * C + c]);
    }
    if (mode == BN_UP) {
        const int Wo = 2 * W, Ho = 2 * H;
        const uint16_t* p = dnext + ((static_cast<size_t>(n) * Ho + 2 * y) * Wo + 2 * x) * C + c;
        return bf16_to_f(p[0]) + bf16_to_f(p[C]) + bf16_to_f(p[static_cast<size_t>(Wo) * C]) +
               bf16_to_f(p[static_cast<size_t>(Wo) * C + C]);
    }
    return bf16_to_f(dnext[((static_cast<size_t>(n) * H + y) * W + x) * C + c]);
}

// the same for 8 consecutive channels (16-byte accesses)
__device__ __forceinline__ void bn_dy8_at(const uint16_t* __restrict__ dnext, int mode, int n, int y, int x, int c0, int H,
                                          int W, int C, float (&dy)[8]) {
    auto ld8 = [](const uint16_t* p, float (&v)[8]) {
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[2 * k] = bf16_lo(u[k]); v[2 * k + 1] = bf16_hi(u[k]); }
    };
    if (mode == BN_POOL) {
        const int Ho = H / 2, Wo = W / 2, yo = y >> 1, xo = x >> 1;
        if (yo >= Ho || xo >= Wo) {                  // row / column dropped by the floor of AvgPool2d(2)
#pragma unroll
            for (int k = 0; k < 8; ++k) dy[k] = 0.f;
            return;
        }
        ld8(dnext + ((static_cast<size_t>(n) * Ho + yo) * Wo + xo) * C + c0, dy);
#pragma unroll
        for (int k = 0; k < 8; ++k) dy[k] *= 0.25f;
        return;
    }
    if (mode == BN_UP) {
        const int Wo = 2 * W, Ho = 2 * H;
        const uint16_t* p = dnext + ((static_cast<size_t>(n) * Ho + 2 * y) * Wo + 2 * x) * C + c0;
        float t0[8], t1[8], t2[8], t3[8];
        ld8(p, t0);
        ld8(p + C, t1);
        ld8(p + static_cast<size_t>(Wo) * C, t2);
        ld8(p + static_cast<size_t>(Wo) * C + C, t3);
#pragma unroll
        for (int k = 0; k < 8; ++k) dy[k] = t0[k] + t1[k] + t2[k] + t3[k];       // same order as the scalar version
        return;
    }
    ld8(dnext + ((static_cast<size_t>(n) * H + y) * W + x) * C + c0, dy);
}

// pass 1: sums[c] += sum dy, sums[C+c] += sum dy * xhat   (xhat = (a - mean) * invstd)
// block = 256 threads = (C/8 channel groups) x (256 / (C/8) pixel rows); a thread owns 8 channels (16-byte loads) of a
// strip of pixels, the rows are folded through shared memory and the block issues one atomic per channel and sum.
// (The lane-per-channel version moved 2 bytes per lane and load: 0.48 ms of a 5 ms training step together with pass 2.
//  Three more mask-weighted sums here would give the producing conv's bias gradient in closed form, but 24 more
//  accumulators took the kernel from 64 to 110 registers and every launch got 15-30 % slower -- measured, reverted.)
constexpr int BN_SUMS = 2;
template <bool AF>
__global__ void bn_bwd_reduce_kernel(const uint16_t* __restrict__ dnext, const uint16_t* __restrict__ a,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     float* __restrict__ sums, int N, int H, int W, int C, int mode, int split, float slope) {
    // blockIdx.y = pass of a merged batch: images [0, split) or [split, N); mean / invstd are [2][C], sums [2][2][C]
    extern __shared__ float s_red[];                 // [rows][2][C]
    const int groups = C >> 3, rows = blockDim.x / groups;
    const int g = threadIdx.x % groups, row = threadIdx.x / groups;
    const int pass = blockIdx.y;
    const size_t p_begin = pass == 0 ? 0 : static_cast<size_t>(split) * H * W;
    const size_t npix = pass == 0 && gridDim.y > 1 ? static_cast<size_t>(split) * H * W : static_cast<size_t>(N) * H * W;
    mean += pass * C;
    invstd += pass * C;
    sums += pass * BN_SUMS * C;
    float mu[8], is[8], a1[8], a2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = mean[g * 8 + k];
        is[k] = invstd[g * 8 + k];
        a1[k] = a2[k] = 0.f;
    }
    if (row < rows) {
        for (size_t p = p_begin + blockIdx.x * static_cast<size_t>(rows) + row; p < npix; p += static_cast<size_t>(gridDim.x) * rows) {
            const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
            const int n = static_cast<int>(p / (static_cast<size_t>(W) * H));
            float dy[8], av[8];
            bn_dy8_at(dnext, mode, n, y, x, g * 8, H, W, C, dy);
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a + p * C + g * 8));
            const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2_t<AF>(u[k]);
                av[2 * k] = f.x;
                av[2 * k + 1] = f.y;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a1[k] += dy[k];
                a2[k] = fmaf(dy[k], (av[k] - mu[k]) * is[k], a2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s_red[(row * 2 + 0) * C + g * 8 + k] = a1[k];
            s_red[(row * 2 + 1) * C + g * 8 + k] = a2[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += s_red[r * 2 * C + i];
        atomicAdd(sums + i, t);                       // sums layout [2][C] = s_red's inner layout
    }
}

// pass 2: g = gamma * invstd * (dy - S1/M - xhat * S2/M) * act'(a)   (act = LeakyReLU(slope): a > 0 ? 1 : slope)
//         dgamma[c] += S2, dbeta[c] += S1 (done once, by block 0).  One thread = 8 channels of one pixel.
template <bool AF>
__global__ void bn_bwd_apply_kernel(const uint16_t* __restrict__ dnext, const uint16_t* __restrict__ a,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums, float count,
                                    float count1, int split,
                                    float slope, uint16_t* __restrict__ g_out, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, float* __restrict__ dbias_conv, int N, int H, int W, int C,
                                    int mode) {
    const int groups = C >> 3;
    const size_t total = static_cast<size_t>(N) * H * W * groups;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {      // both passes of a merged batch share gamma / beta
            const float* s2 = sums + BN_SUMS * C;
            if (dgamma != nullptr) {
                atomicAdd(dgamma + c, sums[C + c] + (split < N ? s2[C + c] : 0.f));
                atomicAdd(dbeta + c, sums[c] + (split < N ? s2[c] : 0.f));
            }
        }
    // dbias_conv: this thread's channel group is the same in every iteration (the grid stride is a multiple of C/8), so the
    // sums of the stored (bf16-rounded) gradients stay in 8 registers; blocks fold them through shared memory at the end.
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        const size_t p = i / groups;
        const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
        const int n = static_cast<int>(p / (static_cast<size_t>(W) * H));
        float dy[8], av[8], o[8];
        bn_dy8_at(dnext, mode, n, y, x, g * 8, H, W, C, dy);
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(a) + i);
        const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack2_t<AF>(u[k]);
            av[2 * k] = f.x;
            av[2 * k + 1] = f.y;
        }
        const int pass = n >= split ? 1 : 0;
        const float cnt = pass ? count1 : count;
        const float* mean_p = mean + pass * C;
        const float* invstd_p = invstd + pass * C;
        const float* sums_p = sums + pass * BN_SUMS * C;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            const float is = __ldg(invstd_p + c);
            const float xh = (av[k] - __ldg(mean_p + c)) * is;
            float gr = __ldg(gamma + c) * is * (dy[k] - __ldg(sums_p + c) / cnt - xh * __ldg(sums_p + C + c) / cnt);
            o[k] = gr * (av[k] > 0.f ? 1.f : slope);
        }
        const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                    pack_bf16x2(o[6], o[7]));
        reinterpret_cast<uint4*>(g_out)[i] = pk;
        if (dbias_conv != nullptr) {
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { bsum[2 * k] += bf16_lo(pw[k]); bsum[2 * k + 1] += bf16_hi(pw[k]); }
        }
    }
    if (dbias_conv != nullptr) {                     // block-uniform
        __shared__ float s_b[512];
        for (int c = threadIdx.x; c < C; c += blockDim.x) s_b[c] = 0.f;
        __syncthreads();
        const int g = static_cast<int>((blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) % groups);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = bsum[k];
            // lanes that share a channel group: lane % groups (groups = 4 .. 64 -> fold the lane bits above log2(groups))
            for (int off = 16; off >= groups && off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (groups >= 32 || (threadIdx.x & 31) < groups) atomicAdd(&s_b[g * 8 + k], v);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            if (s_b[c] != 0.f) atomicAdd(dbias_conv + c, s_b[c]);
    }
}

// SyncBN: dgamma / dbeta come from the LOCAL sums (the gradient all-reduce averages them over ranks), while the apply
// pass uses the all-reduced sums.  This tiny kernel banks the local sums between the two.
__global__ void bn_bwd_accum_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int C, int passes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        dgamma[c] += sums[C + c] + (passes > 1 ? sums[BN_SUMS * C + C + c] : 0.f);
        dbeta[c] += sums[c] + (passes > 1 ? sums[BN_SUMS * C + c] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// MSE (F.mse_loss(reduction='mean'), kwatsch/base_trainer.py:177): loss_acc += sum (a-b)^2 * inv_n (fp32 atomics),
// optionally d = (a - b) * 2 * inv_n * grad_scale.
// ---------------------------------------------------------------------------------------------------------------
__global__ void mse_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float inv_n,
                           float* __restrict__ loss_acc, float* __restrict__ d, float grad_scale) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float diff = a[i] - b[i];
        acc = fmaf(diff, diff, acc);
        if (d) d[i] = diff * (2.f * inv_n * grad_scale);
    }
    acc = warp_sum(acc);
    __shared__ float sm[32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_acc, v * inv_n);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// head (dec.14 + sigmoid) backward.  dlogit = dout * out * (1 - out).
//   data:   g_in[p][c] = act'(a_in[p][c]) * sum_tap w[tap][c] * dlogit[p - off(tap)]          (bf16 out)
//   weight: dW[tap][c] += sum_q dlogit[q] * a_in[q + off(tap)][c] ; dbias += sum_q dlogit[q]
// ---------------------------------------------------------------------------------------------------------------
template <int C, bool AF>
__global__ void head_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                     const uint16_t* __restrict__ a_in, const float* __restrict__ w /*[9][C]*/,
                                     uint16_t* __restrict__ g_in, int N, int H, int W, float slope) {
    __shared__ float sw[9 * C];
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const size_t total = static_cast<size_t>(N) * H * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W), y = static_cast<int>((i / W) % H);
        const size_t nb = (i / (static_cast<size_t>(W) * H)) * H * W;
        float dl[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            // forward tap t reads a[q + (t/3-1, t%3-1)]; pixel p receives from q = p - off(t)
            const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t q = nb + static_cast<size_t>(yy) * W + xx;
                const float o = out[q];
                v = dout[q] * o * (1.f - o);
            }
            dl[t] = v;
        }
        const uint16_t* ap = a_in + i * C;
        uint16_t* gp = g_in + i * C;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) s = fmaf(sw[t * C + c], dl[t], s);
            s *= a16_to_f<AF>(ap[c]) > 0.f ? 1.f : slope;
            gp[c] = __bfloat16_as_ushort(__float2bfloat16_rn(s));
        }
    }
}

// Fused data + weight gradient of the head (one pass over a_in): both need the SAME nine neighbours dl[p - off(t)]:
//   g_in[p][c]  = act'(a_in[p][c]) * sum_t w[t][c] * dl[p - off(t)]
//   dW[t][c]   += a_in[p][c] * dl[p - off(t)]        (sum over q of dl[q] * a[q + off(t)] re-indexed by p = q + off(t))
// thread = (pixel, 8-channel group): one 16-byte load of a_in, one 16-byte store of g_in (both coalesced), 72 + 72 FMAs,
// 72 register accumulators reduced through shared-memory atomics once per block.  (The two-kernel version stored g_in
// with 2-byte scattered stores and gathered a_in nine times: 1.44 ms of an 8 ms step.)
// Version 2 (this one): thread = (pixel, 4-channel group), 8 threads per pixel, a tile of 8 rows x 32 pixels per block
// iteration.  The 10 x 34 window of dlogit = dout * out * (1 - out) that the tile needs is formed ONCE per iteration in shared
// memory (the first version recomputed the nine neighbours in each of a pixel's four threads from two global loads
// each), the filter is read from shared memory (16-byte broadcasts) instead of 72 registers, and 36 instead of 72 dW
// accumulators per thread let three blocks of 256 threads share an SM: the first version ran one block of 190-register
// threads per SM, 2 warps per scheduler, and took 109 us for 36 images at 12.7 % occupancy (ncu, profiles/r04b_*).
// Also accumulates dbias_in[c] += sum_p g_in[p][c]: the bias gradient of the conv that produced a_in (dec.12), so that no
// separate column-sum pass over g_in is needed.
constexpr int HB_PIX = 32;                       // tile width (pixels of a row; W % 32 == 0 not required)
constexpr int HB_ROWS = 8;                       // tile height: one barrier pair and eight independent a_in loads per thread
template <bool AF>
__global__ void __launch_bounds__(256, 3)
head_bwd_fused_kernel(const float* __restrict__ dout, const float* __restrict__ out, const uint16_t* __restrict__ a_in,
                      const float* __restrict__ w /*[9][32]*/, uint16_t* __restrict__ g_in,
                      float* __restrict__ dw /*[32][9]*/, float* __restrict__ dbias, float* __restrict__ dbias_in /*[32] or null*/,
                      int N, int H, int W, float slope) {
    constexpr int C = 32;
    __shared__ __align__(16) float sw[9 * C];
    __shared__ float sdw[9 * C + 1 + C];
    __shared__ float sdl[HB_ROWS + 2][HB_PIX + 2];
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < 9 * C + 1 + C; i += blockDim.x) sdw[i] = 0.f;
    const int g = threadIdx.x & 7;                       // 4-channel group
    const int pl = threadIdx.x >> 3;                     // pixel column inside the tile
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    float accb = 0.f, accg[4] = {0.f, 0.f, 0.f, 0.f};
    const int tiles_x = (W + HB_PIX - 1) / HB_PIX, tiles_y = (H + HB_ROWS - 1) / HB_ROWS;
    const long long tiles = static_cast<long long>(N) * tiles_y * tiles_x;
    for (long long tidx = blockIdx.x; tidx < tiles; tidx += gridDim.x) {
        const int tx = static_cast<int>(tidx % tiles_x);
        const long long nty = tidx / tiles_x;
        const int y0 = static_cast<int>(nty % tiles_y) * HB_ROWS;
        const size_t nb = static_cast<size_t>(nty / tiles_y) * H * W;
        const int x0 = tx * HB_PIX;
        __syncthreads();                                  // previous iteration's readers are done (also covers the init above)
        // (HB_ROWS + 2) x (HB_PIX + 2) window of dlogit = dout * out * (1 - out), formed once per tile
        for (int e = threadIdx.x; e < (HB_ROWS + 2) * (HB_PIX + 2); e += blockDim.x) {
            const int r = e / (HB_PIX + 2), cidx = e - r * (HB_PIX + 2);
            const int yy = y0 + r - 1, xx = x0 + cidx - 1;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t q = nb + static_cast<size_t>(yy) * W + xx;
                const float o = __ldg(out + q);
                v = __ldg(dout + q) * o * (1.f - o);
            }
            sdl[r][cidx] = v;
        }
        const int x = x0 + pl;
        uint2 araw[HB_ROWS];                              // the tile's activations: all loads in flight before the barrier
#pragma unroll
        for (int r = 0; r < HB_ROWS; ++r) {
            const bool in = (x < W) && (y0 + r < H);
            araw[r] = in ? __ldg(reinterpret_cast<const uint2*>(a_in + (nb + static_cast<size_t>(y0 + r) * W + x) * C) + g)
                         : make_uint2(0u, 0u);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < HB_ROWS; ++r) {
            if (x >= W || y0 + r >= H) continue;
            // forward tap t reads a[q + (t/3-1, t%3-1)]; pixel p receives from q = p - off(t): window entry (r + 2 - dy, pl + 2 - dx)
            float dl[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) dl[t] = sdl[r + 2 - t / 3][pl + 2 - t % 3];
            if (g == 0) accb += dl[4];
            float a[4];
            a[0] = a16_to_f<AF>(static_cast<uint16_t>(araw[r].x & 0xFFFFu));
            a[1] = a16_to_f<AF>(static_cast<uint16_t>(araw[r].x >> 16));
            a[2] = a16_to_f<AF>(static_cast<uint16_t>(araw[r].y & 0xFFFFu));
            a[3] = a16_to_f<AF>(static_cast<uint16_t>(araw[r].y >> 16));
            float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[t * C + g * 4]);
                gv[0] = fmaf(wv.x, dl[t], gv[0]); gv[1] = fmaf(wv.y, dl[t], gv[1]);
                gv[2] = fmaf(wv.z, dl[t], gv[2]); gv[3] = fmaf(wv.w, dl[t], gv[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[t][j] = fmaf(a[j], dl[t], acc[t][j]);
            }
            // the stored (bf16-rounded) value is what the weight-gradient GEMM of dec.12 multiplies: sum the rounded values
            uint32_t pk[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float v0 = gv[2 * k] * (a[2 * k] > 0.f ? 1.f : slope), v1 = gv[2 * k + 1] * (a[2 * k + 1] > 0.f ? 1.f : slope);
                pk[k] = pack_bf16x2(v0, v1);
                accg[2 * k] += bf16_lo(pk[k]);
                accg[2 * k + 1] += bf16_hi(pk[k]);
            }
            reinterpret_cast<uint2*>(g_in + (nb + static_cast<size_t>(y0 + r) * W + x) * C)[g] = make_uint2(pk[0], pk[1]);
        }
    }
    // lanes with the same channel group (lane & 7): butterfly over lane bits 3..4, then one shared atomic per value
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[t][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if ((threadIdx.x & 31) < 8) atomicAdd(&sdw[t * C + g * 4 + j], v);
        }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = accg[j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((threadIdx.x & 31) < 8) atomicAdd(&sdw[9 * C + 1 + g * 4 + j], v);
    }
    accb = warp_sum(accb);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sdw[9 * C], accb);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
        const int t = i / C, c = i - t * C;
        atomicAdd(dw + c * 9 + t, sdw[i]);               // nn.Conv2d layout [1][C][3][3]
    }
    if (threadIdx.x == 0) atomicAdd(dbias, sdw[9 * C]);
    if (dbias_in != nullptr && threadIdx.x < C) atomicAdd(dbias_in + threadIdx.x, sdw[9 * C + 1 + threadIdx.x]);
}

// one warp handles a strip of pixels; lane = channel (C = 32); 9 tap accumulators per lane.
template <int C, bool AF>
__global__ void head_bwd_weight_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                       const uint16_t* __restrict__ a_in, float* __restrict__ dw /*[9][C]*/,
                                       float* __restrict__ dbias, int N, int H, int W) {
    static_assert(C == 32, "lane == channel");
    const int lane = threadIdx.x & 31;
    const size_t warp_id = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    const size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t total = static_cast<size_t>(N) * H * W;
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float accb = 0.f;
    for (size_t q = warp_id; q < total; q += nwarps) {
        const int x = static_cast<int>(q % W), y = static_cast<int>((q / W) % H);
        const size_t nb = (q / (static_cast<size_t>(W) * H)) * H * W;
        const float o = out[q];
        const float dl = dout[q] * o * (1.f - o);
        accb += dl;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                acc[t] = fmaf(dl, a16_to_f<AF>(a_in[(nb + static_cast<size_t>(yy) * W + xx) * C + lane]), acc[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw + lane * 9 + t, acc[t]);   // nn.Conv2d layout [1][C][3][3]
    if (lane == 0) atomicAdd(dbias, accb);
}

// ---------------------------------------------------------------------------------------------------------------
// enc.0 backward: a0[p][c] = w0[c] * xpad[p] + b0[c]  =>  dw0[c] += sum_p g[p][c] * xpad[p], db0[c] += sum_p g[p][c]
// g bf16 [N,H+2,W+2,C] (C = 32), x fp32 [N,1,H,W].  Four threads per pixel, 8 channels (one 16-byte load) each; the 64
// pixel rows of a block are folded through shared memory and the block issues 64 atomics.  (One warp per pixel with
// 2-byte loads and 64 atomics per WARP onto the same 64 addresses: 89 us, profiles/r02u_train_launches.md.)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
e0_bwd_kernel(const uint16_t* __restrict__ g, const float* __restrict__ x, float* __restrict__ dw,
              float* __restrict__ db, int N, int H, int W, int C) {
    __shared__ float s_w[64][33], s_b[64][33];
    const int cg = threadIdx.x & 3, row = threadIdx.x >> 2;          // 4 channel groups x 64 pixel rows
    const int Ho = H + 2, Wo = W + 2;
    const size_t total = static_cast<size_t>(N) * Ho * Wo;
    float aw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ab[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t p = blockIdx.x * static_cast<size_t>(64) + row; p < total; p += static_cast<size_t>(gridDim.x) * 64) {
        const int xo = static_cast<int>(p % Wo), yo = static_cast<int>((p / Wo) % Ho);
        const int n = static_cast<int>(p / (static_cast<size_t>(Wo) * Ho));
        const int yi = yo - 1, xi = xo - 1;
        const float xv = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? __ldg(x + (static_cast<size_t>(n) * H + yi) * W + xi) : 0.f;
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(g + p * C) + cg);
        const uint32_t u[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lo = bf16_lo(u[k]), hi = bf16_hi(u[k]);
            aw[2 * k] = fmaf(lo, xv, aw[2 * k]);
            aw[2 * k + 1] = fmaf(hi, xv, aw[2 * k + 1]);
            ab[2 * k] += lo;
            ab[2 * k + 1] += hi;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s_w[row][cg * 8 + k] = aw[k];
        s_b[row][cg * 8 + k] = ab[k];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = threadIdx.x & 31;
        const bool bias = threadIdx.x >= 32;
        float t = 0.f;
        for (int r = 0; r < 64; ++r) t += bias ? s_b[r][c] : s_w[r][c];
        atomicAdd((bias ? db : dw) + c, t);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 / pad 1 conv (CUDA cores, fp32 accumulate):
//   dW[co][ci][tap] += sum_{n,y,x} g[n,y,x,co] * X[n, y+dy-1, x+dx-1, ci] ;  dbias[co] += sum g
// g bf16 [N,H,W,Cout], X 16-bit [N,H,W,Cin].  Block = (32 ci) x (8 co groups): thread (tx, ty) owns input channel
// ci0 + tx and the CO_PER output channels co0 + ty*CO_PER ..; it walks a strip of pixels with 9 taps in registers.
// grid = (pixel strips, Cin/32, Cout/(8*CO_PER)); partial sums are atomically added to dW (fp32).
// ---------------------------------------------------------------------------------------------------------------
template <bool AF, int CO_PER>
__global__ void __launch_bounds__(256)
wgrad3x3_kernel(const uint16_t* __restrict__ g, const uint16_t* __restrict__ X, float* __restrict__ dW,
                float* __restrict__ dbias, int N, int H, int W, int Cin, int Cout) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int ci = blockIdx.y * 32 + tx;
    const int co0 = (blockIdx.z * 8 + ty) * CO_PER;
    const size_t npix = static_cast<size_t>(N) * H * W;
    float acc[CO_PER][9];
    float accb[CO_PER];
#pragma unroll
    for (int j = 0; j < CO_PER; ++j) {
        accb[j] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
    }
    for (size_t p = blockIdx.x; p < npix; p += gridDim.x) {
        const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
        const size_t nb = (p / (static_cast<size_t>(W) * H)) * H * W;
        float gv[CO_PER];
        bool any = false;
#pragma unroll
        for (int j = 0; j < CO_PER; ++j) {
            gv[j] = bf16_to_f(__ldg(g + p * Cout + co0 + j));      // warp-uniform address: broadcast
            accb[j] += gv[j];
            any |= gv[j] != 0.f;
        }
        if (!any) continue;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float xv = a16_to_f<AF>(__ldg(X + (nb + static_cast<size_t>(yy) * W + xx) * Cin + ci));
#pragma unroll
            for (int j = 0; j < CO_PER; ++j) acc[j][t] = fmaf(gv[j], xv, acc[j][t]);
        }
    }
#pragma unroll
    for (int j = 0; j < CO_PER; ++j) {
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(dW + (static_cast<size_t>(co0 + j) * Cin + ci) * 9 + t, acc[j][t]);
        if (blockIdx.y == 0 && tx == 0 && dbias) atomicAdd(dbias + co0 + j, accb[j]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// latent-mix backward (kwatsch/cardiac/trainer_ae.py:173 / brain :265-266): z_mix[b] = wa[b] z[b] + wb[b] z[B+b]
//   g_z[b]   = g_dec[b]   + wa[b] * g_mix[b]
//   g_z[B+b] = g_dec[B+b] + wb[b] * g_mix[b]          all bf16 NHWC [., HW*C]
// ---------------------------------------------------------------------------------------------------------------
__global__ void mix_bwd_kernel(const uint16_t* __restrict__ g_dec, const uint16_t* __restrict__ g_mix,
                               const float* __restrict__ wa, const float* __restrict__ wb, uint16_t* __restrict__ g_z,
                               int B, size_t per_image) {
    const size_t total = 2 * static_cast<size_t>(B) * per_image;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(i / per_image);
        const size_t r = i - static_cast<size_t>(n) * per_image;
        const int b = n < B ? n : n - B;
        const float w = n < B ? wa[b] : wb[b];
        const float v = bf16_to_f(g_dec[i]) + w * bf16_to_f(g_mix[static_cast<size_t>(b) * per_image + r]);
        g_z[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused Adam over flat fp32 buffers (torch.optim.Adam semantics, kwatsch/trainer_ae.py:29-30: L2-in-grad weight decay)
// ---------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, const int* __restrict__ step_dev,
                            const float* __restrict__ lr_dev) {
    if (lr_dev != nullptr) lr = *lr_dev;      // learning rate in device memory (per-iteration schedulers under graph replay)
    if (step_dev != nullptr) {      // step count in device memory (a captured CUDA graph replays with a new count)
        const float t = static_cast<float>(*step_dev);
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float gi = g[i];
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

}  // namespace aesr
