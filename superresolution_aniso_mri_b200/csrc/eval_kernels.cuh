// Image-domain utilities of the evaluation / data path (all HBM-bound, one pass over the data):
//   * per-slice SSIM + squared error + min      (evaluate/metrics.py:111-194 -> scikit-image structural_similarity /
//                                                peak_signal_noise_ratio, restated from the published algorithm)
//   * exact order statistics by radix select     (np.percentile inside generate_hr_volumes.normalize_img :130-133 and
//                                                datasets/common.rescale_intensities :408-417) + affine / clip apply
//   * batched zero-pad + crop gather             (datasets/shared_transforms.py AdjustToPatchSize / CenterCrop /
//                                                RandomCrop :389-447, 297-363, 48-120)
#pragma once
#include "common.cuh"

namespace aesr {

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// ---------------------------------------------------------------------------------------------------------------
// SSIM with a win x win uniform window, sample covariance (NP/(NP-1)), K1 = .01, K2 = .03, mean over the image with a
// (win-1)/2 border cropped -- every window of a kept pixel lies inside the image, so the filter's boundary mode never
// enters.  Also accumulates sum (a-b)^2 and min(a) over the FULL slice (PSNR: data_range 1 if min >= 0 else 2).
// grid = (tiles_x, tiles_y, Z), block = 16 x 16; float64 window sums like skimage.
// ---------------------------------------------------------------------------------------------------------------
template <int WIN>
__global__ void __launch_bounds__(256)
ssim_psnr_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, double c1, double c2,
                 double* __restrict__ ssim_sum, double* __restrict__ sqerr_sum, unsigned int* __restrict__ min_key) {
    constexpr int T = 16, R = WIN / 2, HT = T + 2 * R;
    __shared__ float sa[HT][HT + 1], sb[HT][HT + 1];
    __shared__ double red[3][8];
    const int z = blockIdx.z, x0 = blockIdx.x * T, y0 = blockIdx.y * T;
    const float* pa = a + static_cast<size_t>(z) * H * W;
    const float* pb = b + static_cast<size_t>(z) * H * W;
    for (int i = threadIdx.x; i < HT * HT; i += 256) {
        const int hy = i / HT, hx = i % HT;
        const int gy = y0 + hy - R, gx = x0 + hx - R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sa[hy][hx] = in ? pa[static_cast<size_t>(gy) * W + gx] : 0.f;
        sb[hy][hx] = in ? pb[static_cast<size_t>(gy) * W + gx] : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x = x0 + tx, y = y0 + ty;
    double s_val = 0.0, e_val = 0.0;
    float mn = 3.4e38f;
    if (x < W && y < H) {
        const float av = sa[ty + R][tx + R], bv = sb[ty + R][tx + R];
        const double d = static_cast<double>(av) - static_cast<double>(bv);
        e_val = d * d;
        mn = av;
        if (x >= R && x < W - R && y >= R && y < H - R) {
            double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
            for (int dy = 0; dy < WIN; ++dy)
#pragma unroll
                for (int dx = 0; dx < WIN; ++dx) {
                    const double u = sa[ty + dy][tx + dx], v = sb[ty + dy][tx + dx];
                    sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
                }
            constexpr double NP = WIN * WIN, cov_norm = NP / (NP - 1.0);
            const double ux = sx / NP, uy = sy / NP;
            const double vx = cov_norm * (sxx / NP - ux * ux), vy = cov_norm * (syy / NP - uy * uy);
            const double vxy = cov_norm * (sxy / NP - ux * uy);
            s_val = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
        }
    }
    // block reduction (warp shuffles on doubles, then 8 warps through smem)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
        e_val += __shfl_xor_sync(0xffffffffu, e_val, o);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s_val; red[1][warp] = e_val; red[2][warp] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0, e = 0, m = 3.4e38;
        for (int w = 0; w < 8; ++w) { s += red[0][w]; e += red[1][w]; m = fmin(m, red[2][w]); }
        atomicAdd(ssim_sum + z, s);
        atomicAdd(sqerr_sum + z, e);
        atomicMin(min_key + z, f2key(static_cast<float>(m)));      // order-preserving uint key; host decodes
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Exact k-th order statistics of an fp32 array by 3-pass radix select on the order-preserving uint32 key
// (sign flip trick), for up to 4 ranks at once.  Pass p histograms the 11/11/10 key bits of the elements whose higher
// bits match each rank's prefix; a tiny kernel then narrows prefix and rank.  key -> float at the end.
// ---------------------------------------------------------------------------------------------------------------
struct SelectState {
    uint32_t prefix[4];      // key bits decided so far (high bits)
    uint64_t rank[4];        // remaining rank inside the current prefix bucket
    float result[4];
};

template <int SHIFT, int BITS>
__global__ void select_hist_kernel(const float* __restrict__ x, size_t n, const SelectState* __restrict__ st, int nranks,
                                   unsigned int* __restrict__ hist /*[4][2048]*/) {
    __shared__ unsigned int sh[4][1 << BITS];
    for (int i = threadIdx.x; i < 4 * (1 << BITS); i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t pre[4];
    for (int r = 0; r < 4; ++r) pre[r] = r < nranks ? st->prefix[r] : 0;
    constexpr uint32_t HIGH_MASK = (SHIFT + BITS >= 32) ? 0u : ~((1u << (SHIFT + BITS)) - 1u);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint32_t k = f2key(x[i]);
        const uint32_t bin = (k >> SHIFT) & ((1u << BITS) - 1u);
        for (int r = 0; r < nranks; ++r)
            if ((k & HIGH_MASK) == pre[r]) atomicAdd(&sh[r][bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nranks * (1 << BITS); i += blockDim.x) {
        const unsigned int v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + (i >> BITS) * 2048 + (i & ((1 << BITS) - 1)), v);
    }
}

template <int SHIFT, int BITS>
__global__ void select_narrow_kernel(SelectState* st, int nranks, unsigned int* hist, int last) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    uint64_t rank = st->rank[r];
    unsigned int* h = hist + r * 2048;
    uint32_t bin = 0;
    for (; bin < (1u << BITS); ++bin) {
        const unsigned int c = h[bin];
        if (rank < c) break;
        rank -= c;
    }
    st->rank[r] = rank;
    st->prefix[r] |= bin << SHIFT;
    if (last) st->result[r] = key2f(st->prefix[r]);
    for (uint32_t i = 0; i < (1u << BITS); ++i) h[i] = 0;      // ready for the next pass
}

// out = clip((x - lo) / (hi - lo), 0, 1) with float64 arithmetic then one rounding to fp32, exactly what numpy >= 2
// does for  ((img - min_val) / (max_val - min_val)).clip(0, 1)  with float64 percentiles.
__global__ void normalize_apply_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n,
                                       const double* __restrict__ lo_hi) {
    const double lo = lo_hi[0], inv_den = lo_hi[1] - lo_hi[0];
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        double v = __ddiv_rn(__dsub_rn(static_cast<double>(x[i]), lo), inv_den);
        v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        out[i] = static_cast<float>(v);
    }
}

// lo_hi[j] = linear interpolation between the two order statistics around the fractional rank (numpy 'linear'):
// v = x_k + frac * (x_{k+1} - x_k), in float64.  st->result = {x_klo, x_klo+1, x_khi, x_khi+1}.
__global__ void percentile_finish_kernel(const SelectState* st, double frac_lo, double frac_hi, double* lo_hi) {
    // numpy's _lerp: a + (b-a)*t, but b - (b-a)*(1-t) when t >= 0.5 (float64)
    // (b - a) is formed in the ARRAY dtype (fp32 subtraction, one rounding) before the float64 lerp, as numpy does.
    const double a0 = st->result[0], a1 = st->result[1], b0 = st->result[2], b1 = st->result[3];
    const double da = static_cast<double>(__fsub_rn(st->result[1], st->result[0]));
    const double db = static_cast<double>(__fsub_rn(st->result[3], st->result[2]));
    // separately rounded multiply and add (numpy has no fused multiply-add here; nvcc would contract a + b*c)
    lo_hi[0] = frac_lo >= 0.5 ? __dsub_rn(a1, __dmul_rn(da, __dsub_rn(1.0, frac_lo))) : __dadd_rn(a0, __dmul_rn(da, frac_lo));
    lo_hi[1] = frac_hi >= 0.5 ? __dsub_rn(b1, __dmul_rn(db, __dsub_rn(1.0, frac_hi))) : __dadd_rn(b0, __dmul_rn(db, frac_hi));
}

// ---------------------------------------------------------------------------------------------------------------
// Batched zero-pad + crop: out[b, c, y, x] = in[b, c, y + top[b], x + left[b]] if inside the source image else 0.
// (top/left may be negative = padding on that side.)  fp32, coalesced along x.
// ---------------------------------------------------------------------------------------------------------------
__global__ void pad_crop_gather_kernel(const float* __restrict__ in, float* __restrict__ out,
                                       const int* __restrict__ top, const int* __restrict__ left, int C, int Hin,
                                       int Win, int Hout, int Wout) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int t = top[b], l = left[b];
    const float* src = in + (static_cast<size_t>(b) * C + c) * Hin * Win;
    float* dst = out + (static_cast<size_t>(b) * C + c) * Hout * Wout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Hout * Wout; i += gridDim.x * blockDim.x) {
        const int y = i / Wout, x = i - y * Wout;
        const int sy = y + t, sx = x + l;
        dst[i] = (sy >= 0 && sy < Hin && sx >= 0 && sx < Win) ? __ldg(src + static_cast<size_t>(sy) * Win + sx) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused training augmentation (datasets/shared_transforms.py: AdjustToPatchSize :389-447 / CenterCrop :297-363 /
// RandomCrop :48-120 as one composite window, RandomRotation :224-254, RandomIntensity :366-386):
//   A[y][x]   = in[b, c, y + top[b], x + left[b]] inside the source image, else 0          (P x P window)
//   R         = np.rot90(A, k[b]):  k=1: R[i][j] = A[j][P-1-i],  k=2: A[P-1-i][P-1-j],  k=3: A[P-1-j][i]
//   out[b,c]  = 1 / (1 + exp(gain[b] * (cutoff[b] - R)))  for the channels of chan_mask (every fp32 operation rounded
//               separately like numpy: sub, mul, exp, add, div; exp correctly rounded through double), else R.
// The random draws are made on the host from the caller's numpy RandomState in the reference's order.  Writes are
// coalesced along x; the rotated reads of k = 1, 3 walk a column of the 4 B/px source (L1/L2 resident: one sample's
// window is P*P*4 bytes).  8 B per output pixel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void augment_gather_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ top,
                                      const int* __restrict__ left, const int* __restrict__ rot_k,
                                      const float* __restrict__ gain, const float* __restrict__ cutoff, unsigned chan_mask,
                                      int C, int Hin, int Win, int P) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int t = top[b], l = left[b], k = rot_k ? (rot_k[b] & 3) : 0;
    const bool contrast = gain != nullptr && ((chan_mask >> (c & 31)) & 1u);
    const float g = contrast ? gain[b] : 0.f, co = contrast ? cutoff[b] : 0.f;
    const float* src = in + (static_cast<size_t>(b) * C + c) * Hin * Win;
    float* dst = out + (static_cast<size_t>(b) * C + c) * P * P;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P * P; idx += gridDim.x * blockDim.x) {
        const int i = idx / P, j = idx - i * P;
        const int ay = (k == 0) ? i : (k == 1) ? j : (k == 2) ? P - 1 - i : P - 1 - j;
        const int ax = (k == 0) ? j : (k == 1) ? P - 1 - i : (k == 2) ? P - 1 - j : i;
        const int sy = ay + t, sx = ax + l;
        float v = (sy >= 0 && sy < Hin && sx >= 0 && sx < Win) ? __ldg(src + static_cast<size_t>(sy) * Win + sx) : 0.f;
        if (contrast) {
            const float u = __fmul_rn(g, __fsub_rn(co, v));
            const float e = static_cast<float>(exp(static_cast<double>(u)));
            v = __fdiv_rn(1.f, __fadd_rn(1.f, e));
        }
        dst[idx] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LR-dataset synthesis (datasets/common_brains.py:37-44, simulate_thick_slices): scipy.ndimage.gaussian_filter1d along z
// for every (y, x) column -- 'reflect' borders (half-sample symmetric, any overhang), float64 accumulation in scipy's
// order (centre tap, then symmetric pairs from the farthest tap inwards, separately rounded multiply and add), result
// rounded once to fp32.  w = the 2*lw+1 normalised taps formed on the host with scipy's formula.  One thread per output
// voxel, coalesced along the in-plane index; the 2*lw+1 planes a block reads stay in L1/L2.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int m = 2 * n;
    i %= m;
    if (i < 0) i += m;
    return i < n ? i : m - 1 - i;
}
__global__ void gauss1d_axis0_kernel(const float* __restrict__ in, float* __restrict__ out, const double* __restrict__ w,
                                     int lw, int Z, size_t HW) {
    const int z = blockIdx.y;
    for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < HW;
         p += static_cast<size_t>(gridDim.x) * blockDim.x) {
        double acc = __dmul_rn(static_cast<double>(__ldg(in + static_cast<size_t>(z) * HW + p)), w[lw]);
        for (int jj = -lw; jj < 0; ++jj) {
            const double a = static_cast<double>(__ldg(in + static_cast<size_t>(reflect_index(z + jj, Z)) * HW + p));
            const double b = static_cast<double>(__ldg(in + static_cast<size_t>(reflect_index(z - jj, Z)) * HW + p));
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), w[jj + lw]));
        }
        out[static_cast<size_t>(z) * HW + p] = static_cast<float>(acc);
    }
}

}  // namespace aesr
