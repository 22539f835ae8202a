// Pixel-domain multi-scale VIF exactly as the reference computes it (evaluate/vifvec.py:7-63, called with UINT8 slices by
// evaluate/metrics.py:65-108).  All planes are uint8 like the reference's numpy arrays:
//   * scipy.ndimage.gaussian_filter on uint8 data returns uint8, one separable pass at a time: float64 accumulation
//     (centre tap, then symmetric pairs from the farthest tap inwards -- the order of ni_filters.c NI_Correlate1D, no FMA
//     contraction), 'reflect' boundary, C truncation to uint8 after EACH pass;
//   * `ref * ref`, `mu1 * mu1` and the variance subtractions wrap modulo 256 (numpy uint8 arithmetic).
// HBM-trivial (a slice is 16 KB); what matters is bit-exactness of the integer planes: every pass is one kernel over all
// slices, a thread per output byte, filter weights (computed on the host exactly like scipy) read through __ldg.
#pragma once
#include "common.cuh"

namespace aesr {

__device__ __forceinline__ int vif_reflect(int i, int n) {       // scipy 'reflect': d c b a | a b c d | d c b a
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// fp32 image -> uint8: np.uint8(np.clip(x * 255., 0, 255)) with the product in float32 (evaluate/metrics.py:72-73)
__global__ void vif_quantize_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = fminf(fmaxf(__fmul_rn(x[i], 255.f), 0.f), 255.f);
        out[i] = static_cast<uint8_t>(v);
    }
}

// one separable pass over `planes` images of h x w bytes; axis 0 = along y (rows), 1 = along x.
// out[p] = (uint8) ( in[c] * w[lw] + sum_{jj=-lw}^{-1} (in[c+jj] + in[c-jj]) * w[jj+lw] )   in float64, scipy's order
__global__ void vif_gauss1d_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int planes, int h, int w,
                                      int axis, const double* __restrict__ wt, int lw) {
    const size_t total = static_cast<size_t>(planes) * h * w;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const uint8_t* img = in + (i / (static_cast<size_t>(w) * h)) * static_cast<size_t>(w) * h;
        const int n = axis == 0 ? h : w, c = axis == 0 ? y : x;
        const int stride = axis == 0 ? w : 1;
        const uint8_t* line = img + (axis == 0 ? x : static_cast<size_t>(y) * w);
        double tmp = __dmul_rn(static_cast<double>(line[static_cast<size_t>(c) * stride]), __ldg(wt + lw));
        for (int jj = -lw; jj < 0; ++jj) {
            const double a = static_cast<double>(line[static_cast<size_t>(vif_reflect(c + jj, n)) * stride]);
            const double b = static_cast<double>(line[static_cast<size_t>(vif_reflect(c - jj, n)) * stride]);
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), __ldg(wt + jj + lw)));
        }
        out[i] = static_cast<uint8_t>(tmp);          // values lie in [0, 255]: truncation like the C cast in scipy
    }
}

// [::2, ::2] of `planes` images
__global__ void vif_subsample_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int planes, int h, int w) {
    const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
    const size_t total = static_cast<size_t>(planes) * h2 * w2;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w2), y = static_cast<int>((i / w2) % h2);
        const size_t p = i / (static_cast<size_t>(w2) * h2);
        out[i] = in[(p * h + 2 * y) * w + 2 * x];
    }
}

// uint8 products (modulo 256): rr = r*r, dd = d*d, rd = r*d
__global__ void vif_products_kernel(const uint8_t* __restrict__ r, const uint8_t* __restrict__ d, uint8_t* __restrict__ rr,
                                    uint8_t* __restrict__ dd, uint8_t* __restrict__ rd, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const unsigned a = r[i], b = d[i];
        rr[i] = static_cast<uint8_t>(a * a);
        dd[i] = static_cast<uint8_t>(b * b);
        rd[i] = static_cast<uint8_t>(a * b);
    }
}

// Per-slice accumulation of one scale: num_den[z][0] += sum log10(1 + g^2 s1 / (sv + nsq)), [z][1] += sum log10(1 + s1 / nsq).
// One block per slice, fixed summation order (deterministic); the scales are added in launch order like the reference's
// `num += num_det`.
__global__ void __launch_bounds__(256)
vif_accumulate_kernel(const uint8_t* __restrict__ mu1, const uint8_t* __restrict__ mu2, const uint8_t* __restrict__ grr,
                      const uint8_t* __restrict__ gdd, const uint8_t* __restrict__ grd, int hw, double sigma_nsq,
                      double* __restrict__ num_den) {
    __shared__ double red[2][256];
    const int z = blockIdx.x;
    const size_t base = static_cast<size_t>(z) * hw;
    const double eps = 1e-10;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < hw; i += 256) {
        const unsigned m1 = mu1[base + i], m2 = mu2[base + i];
        const uint8_t s1u = static_cast<uint8_t>(grr[base + i] - static_cast<uint8_t>(m1 * m1));
        const uint8_t s2u = static_cast<uint8_t>(gdd[base + i] - static_cast<uint8_t>(m2 * m2));
        const uint8_t s12u = static_cast<uint8_t>(grd[base + i] - static_cast<uint8_t>(m1 * m2));
        const double s1 = s1u, s2 = s2u, s12 = s12u;
        double g = __ddiv_rn(s12, __dadd_rn(s1, eps));
        double sv = __dsub_rn(s2, __dmul_rn(g, s12));
        if (s1 < eps) { g = 0.0; sv = s2; }
        if (s2 < eps) { g = 0.0; sv = 0.0; }
        if (g < 0.0) { sv = s2; g = 0.0; }
        if (sv <= eps) sv = eps;
        num += log10(__dadd_rn(1.0, __ddiv_rn(__dmul_rn(__dmul_rn(g, g), s1), __dadd_rn(sv, sigma_nsq))));
        den += log10(__dadd_rn(1.0, __ddiv_rn(s1, sigma_nsq)));
    }
    red[0][threadIdx.x] = num;
    red[1][threadIdx.x] = den;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        num_den[2 * z] += red[0][0];
        num_den[2 * z + 1] += red[1][0];
    }
}

}  // namespace aesr
