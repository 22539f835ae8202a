// extern "C" entry points of libaesr_b200.so (declared in include/aesr_b200.h).
// Host-side glue only: argument validation, TMA descriptor encoding, grid sizing, launches.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/aesr_b200.h"
#include "conv3x3_tc.cuh"
#include "conv3x3_fold.cuh"
#include "elementwise.cuh"
#include "eval_kernels.cuh"
#include "lpips_kernels.cuh"
#ifdef AESR_WITH_PROBES
#include "../../include/aesr_b200_probe.h"
#include "probe.cuh"
#endif
#include "train_kernels.cuh"
#include "vif_kernels.cuh"
#include "wgrad_tc.cuh"

using namespace aesr;

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::once_flag g_init_once;
int g_init_status = AESR_ERR_CUDA;
int g_sm_count = 0;
int g_max_smem_optin = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return fail(AESR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AESR_ERR_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
    return AESR_OK;
}

void do_init(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        g_init_status = fail(AESR_ERR_CUDA, "no usable CUDA device %d: %s", device,
                             cudaGetErrorString(cudaGetLastError()));
        return;
    }
    if (prop.major != 10) {
        g_init_status = fail(AESR_ERR_ARCH,
                             "aesr_b200 needs a compute-capability 10.x (B200, sm_100a) device, got %d.%d (%s); "
                             "there is no fallback path", prop.major, prop.minor, prop.name);
        return;
    }
    g_sm_count = prop.multiProcessorCount;
    g_max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        g_init_status = fail(AESR_ERR_CUDA, "cannot resolve cuTensorMapEncodeTiled from the driver");
        return;
    }
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    g_init_status = AESR_OK;
}

int ensure_init() {
    if (g_init_status != AESR_OK) {
        int dev = 0;
        cudaGetDevice(&dev);
        return aesr_init(dev);
    }
    return AESR_OK;
}

// NHWC 16-bit activation [N,H,W,C] viewed as a 4-D tensor {C, W, H, N}; box = {KC, box_w, box_h, 1}.
// (bf16 and fp16 are both plain 2-byte copies for TMA; out-of-bounds elements are zero-filled.)
int make_act_tmap(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int KC, int box_w, int box_h) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE,
                                KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AESR_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
    return AESR_OK;
}

// packed weights 16-bit [9*Cout rows][Cin]; box = {KC, BN}.
int make_wgt_tmap(CUtensorMap* m, const void* ptr, int rows, int Cin, int KC, int BN) {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE,
                                KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AESR_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return AESR_OK;
}

// 16-bit NHWC output (or a strided view of it) as a 4-D tensor {C, W, H, N} with explicit byte strides {W, H, N}; box = one
// 16-channel chunk of a 16 x 8 pixel tile (32-byte rows, SWIZZLE_32B) or, box_c = 32, its 64-byte rows (SWIZZLE_64B): the
// TMA-store epilogues of the halo kernel.
int make_out_tmap(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, size_t sw, size_t sh, size_t sn, int box_c = 16) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sw, (cuuint64_t)sh, (cuuint64_t)sn};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)CONV_TILE_W, (cuuint32_t)CONV_TILE_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AESR_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
    return AESR_OK;
}

template <bool FP16>
__global__ void pack_conv3x3_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cout, int Cin,
                                           int transpose_flip) {
    const int total = 9 * Cout * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int tap, row, col;     // destination index: [tap][row][col]
        float v;
        if (!transpose_flip) {
            col = i % Cin; row = (i / Cin) % Cout; tap = i / (Cin * Cout);
            v = w[(static_cast<size_t>(row) * Cin + col) * 9 + tap];
        } else {
            // dgrad: dX = conv(dY, W') with W'[tap][ci][co] = W[co][ci][8 - tap]
            col = i % Cout; row = (i / Cout) % Cin; tap = i / (Cin * Cout);
            v = w[(static_cast<size_t>(col) * Cin + row) * 9 + (8 - tap)];
        }
        out[i] = cvt16_t<FP16>(v);
    }
}

// All 3x3 filters of a model in ONE launch (training re-packs every filter after every optimizer step: 24 launches of
// ~2 us each otherwise).  jobs[j] = {src offset (floats), dst offset (16-bit elements), Cout, Cin, transpose_flip, 0}.
template <bool FP16>
__global__ void pack_conv3x3_weight_batch_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                                 const long long* __restrict__ jobs) {
    const long long* jb = jobs + 6 * blockIdx.y;
    const float* w = src + jb[0];
    uint16_t* out = dst + jb[1];
    const int Cout = static_cast<int>(jb[2]), Cin = static_cast<int>(jb[3]), transpose_flip = static_cast<int>(jb[4]);
    const int total = 9 * Cout * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int tap, row, col;
        float v;
        if (!transpose_flip) {
            col = i % Cin; row = (i / Cin) % Cout; tap = i / (Cin * Cout);
            v = w[(static_cast<size_t>(row) * Cin + col) * 9 + tap];
        } else {
            col = i % Cout; row = (i / Cout) % Cin; tap = i / (Cin * Cout);
            v = w[(static_cast<size_t>(col) * Cin + row) * 9 + (8 - tap)];
        }
        out[i] = cvt16_t<FP16>(v);
    }
}

// "nearest x2 upsample -> 3x3 conv" folded to the low resolution (conv3x3_tc.cuh OUT_SHUFFLE2): destination
// [tap = (dy,dx) low-res offset][phase = 2a+b][co][ci] = sum of the original taps (ky,kx) that read low-res offset
// (dy,dx) when producing hi-res pixel (2y+a, 2x+b):  a = 0: dy=0 <- {ky=0}, dy=1 <- {1,2};  a = 1: dy=1 <- {0,1}, dy=2 <- {2}.
template <bool FP16>
__global__ void pack_conv3x3_weight_up2fold_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cout,
                                                   int Cin) {
    const int total = 9 * 4 * Cout * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin;
        const int co = (i / Cin) % Cout;
        const int ph = (i / (Cin * Cout)) & 3;
        const int tap = i / (Cin * Cout * 4);
        const int a = ph >> 1, b = ph & 1, dy = tap / 3, dx = tap % 3;
        const int my = a == 0 ? (dy == 0 ? 1 : dy == 1 ? 6 : 0) : (dy == 0 ? 0 : dy == 1 ? 3 : 4);   // bit ky set
        const int mx = b == 0 ? (dx == 0 ? 1 : dx == 1 ? 6 : 0) : (dx == 0 ? 0 : dx == 1 ? 3 : 4);
        const float* wk = w + (static_cast<size_t>(co) * Cin + ci) * 9;
        float v = 0.f;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                if (((my >> ky) & 1) && ((mx >> kx) & 1)) v += wk[ky * 3 + kx];
        out[i] = cvt16_t<FP16>(v);
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: `configured` is a bit mask over device ordinals
// (a process that drives a second GPU, like the reference's cuda:1 loss offload, configures each kernel once per device).
template <typename K>
int set_max_smem(K kernel, int* configured) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const int bit = 1 << (dev & 31);
    if (!(*configured & bit)) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem_optin));
        *configured |= bit;
    }
    return AESR_OK;
}

// Split-K workspace: one buffer per device, grown on demand OUTSIDE stream capture (cudaMalloc is not capturable; the
// training engine's first, eager step of a configuration sizes it before the step is captured).  Kernels of one stream use it
// one after the other; the weight-gradient side stream never runs split-K convs.
float* g_splitk_ws[32] = {nullptr};
size_t g_splitk_bytes[32] = {0};
int tune(int key);
int splitk_workspace(size_t need, cudaStream_t stream, float** out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    dev &= 31;
    if (g_splitk_bytes[dev] < need) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        CUDA_TRY(cudaStreamIsCapturing(stream, &st));
        if (st != cudaStreamCaptureStatusNone)
            return fail(AESR_ERR_INVALID, "conv3x3_fwd: the split-K workspace must grow (%zu bytes) inside a stream capture; "
                                          "run the same shapes once outside the capture first", need);
        if (g_splitk_ws[dev]) {
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaFree(g_splitk_ws[dev]));
            g_splitk_ws[dev] = nullptr;
            g_splitk_bytes[dev] = 0;
        }
        size_t bytes = need < (size_t(64) << 20) ? (size_t(64) << 20) : need;
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&g_splitk_ws[dev]), bytes));
        g_splitk_bytes[dev] = bytes;
    }
    *out = g_splitk_ws[dev];
    return AESR_OK;
}

template <int KC>
int launch_stream(const void* x, const void* w, ConvParams p, cudaStream_t stream) {
    using S = StreamSmem<KC>;
    p.BN = p.Cout <= 256 ? p.Cout : 256;
    if (p.Cout % p.BN != 0) return fail(AESR_ERR_INVALID, "conv3x3_fwd: Cout=%d not a multiple of %d", p.Cout, p.BN);
    p.n_blocks = p.Cout / p.BN;
    p.num_tiles = p.N * p.tiles_x * p.tiles_y * p.n_blocks;
    int stages = CONV_MAX_STAGES;
    while (stages > 2 && S::total_bytes(p.BN, stages) > g_max_smem_optin) --stages;
    p.num_stages = stages;
    CUtensorMap tx, tw;
    int rc = make_act_tmap(&tx, x, p.N, p.H, p.W, p.Cin, KC, CONV_TILE_W, CONV_TILE_H);
    if (rc != AESR_OK) return rc;
    rc = make_wgt_tmap(&tw, w, 9 * p.Cout, p.Cin, KC, p.BN);
    if (rc != AESR_OK) return rc;
    // Split-K: layers with fewer tiles than half the SMs (the 16x16 / 8x8 VGG layers of a 12-24 image batch: 12-96 tiles,
    // each streaming its whole 9 x Cin filter slab through one SM's L2 port) are cut along K so that ~all SMs take part;
    // the splits leave raw fp32 accumulators in a workspace and a small kernel adds them and runs the epilogue.
    p.ksplit = 1;
    const int KB = 9 * (p.Cin / KC);
    const bool simple = p.out_mode == OUT_SAME && p.stats == nullptr && p.scale == nullptr && p.out2 == nullptr;
    if (simple && tune(8) != 1 && p.num_tiles * 2 <= g_sm_count) {
        int sk = g_sm_count / p.num_tiles;
        if (sk > KB / 4) sk = KB / 4;                      // at least four K-blocks per split
        if (sk > 16) sk = 16;
        if (sk >= 2) p.ksplit = sk;
    }
    if (p.ksplit > 1) {
        const size_t npix = static_cast<size_t>(p.N) * p.H * p.W;
        const size_t need = static_cast<size_t>(p.ksplit) * npix * p.Cout * sizeof(float);
        rc = splitk_workspace(need, stream, &p.ws);
        if (rc != AESR_OK) return rc;
    }
    const int items = p.num_tiles * p.ksplit;
    const int grid = items < g_sm_count ? items : g_sm_count;
    void* out_final = p.out;
    const int smem = S::total_bytes(p.BN, stages);
#define AESR_STREAM(MODE)                                                                            \
    {                                                                                                \
        static int configured = 0;                                                                   \
        rc = set_max_smem(conv3x3_stream_kernel<KC, MODE>, &configured);                             \
        if (rc != AESR_OK) return rc;                                                                \
        conv3x3_stream_kernel<KC, MODE><<<grid, CONV_THREADS, smem, stream>>>(tx, tw, p);            \
    }
    const bool bare = !p.scale && !p.stats;
    if (p.ksplit > 1) AESR_STREAM(OUT_SAME)                     // split-K: raw accumulators, the epilogue runs in splitk_finish
    else if (bare && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode == MUL_NONE) AESR_STREAM(OUT_SAME)
    else if (bare && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode != MUL_NONE) AESR_STREAM(LEAN_SAME_MUL)
    else if (bare && p.out_mode == OUT_SAME_MAXPOOL2 && p.mul_mode == MUL_NONE) AESR_STREAM(OUT_SAME_MAXPOOL2)
    else AESR_STREAM(-1)
#undef AESR_STREAM
    rc = check_launch("conv3x3_stream");
    if (rc != AESR_OK || p.ksplit == 1) return rc;
    const size_t npix = static_cast<size_t>(p.N) * p.H * p.W;
    const size_t total = npix * (p.Cout / 8);
    size_t fg = (total + 255) / 256;
    if (fg > static_cast<size_t>(g_sm_count) * 8) fg = static_cast<size_t>(g_sm_count) * 8;
    if (p.fp16)
        splitk_finish_kernel<true><<<static_cast<int>(fg), 256, 0, stream>>>(p.ws, p.ksplit, npix, p.Cout, p.bias, p.act, p.slope, p.mul_src, p.mul_mode, static_cast<uint16_t*>(out_final));
    else
        splitk_finish_kernel<false><<<static_cast<int>(fg), 256, 0, stream>>>(p.ws, p.ksplit, npix, p.Cout, p.bias, p.act, p.slope, p.mul_src, p.mul_mode, static_cast<uint16_t*>(out_final));
    return check_launch("splitk_finish");
}

// largest N tile (multiple of 32 dividing Cout, <= 256) whose resident filter bank + >= 2 halo stages fit; 0 = none
template <int KC>
int halo_pick_bn(int Cin, int Cout) {
    using S = HaloSmem<KC>;
    for (int bn = Cout <= 256 ? Cout : 256; bn >= 32; bn -= 32) {
        if (Cout % bn != 0) continue;
        if (S::total_bytes(bn, Cin, 1, 2) <= g_max_smem_optin) return bn;
    }
    return 0;
}

// Profiling / tuning knobs (aesr_set_tuning; initial values from the environment): 0 = AESR_CONV_DEBUG stage mask,
// 1 = AESR_CONV_T forced M-tiles per super-tile, 2 = AESR_CONV_NBUF forced TMEM buffers, 3 = AESR_CONV_STAGES cap,
// 4 = AESR_HEAD_MMA decoder head behind dec.12 on warp-level mma.sync fragments (16-bit activations / filter).
// 5 = AESR_STEM_CUDA_CORES encoder stem on the CUDA cores (fp32 FMAs) instead of warp-level tf32 mma.sync.
// 6 = AESR_WGRAD_NO_FOLD weight gradient of Cin = 32 layers with one MMA per tap instead of one per filter row.
// 7 = AESR_WGRAD_CTAS cap on the CTAs of a weight-gradient launch (each CTA ends with Cout x Cin x taps atomics).
// 8 = AESR_NO_SPLITK streamed conv kernel without split-K (A/B measurements).
// 9 = AESR_FOLD 32 -> 32 layers on conv3x3_fold_kernel (horizontal taps folded into N = 96; measured SLOWER than the tap-by-tap
//     halo kernel, profiles/r08_fold_sweep.txt: kept as an opt-in experiment with its tests).
// 10 = AESR_NO_TMA_STORE inference epilogues with per-thread global stores instead of staged TMA stores (A/B measurements).
// 11 = AESR_HEAD_GATHER_NO_TMA head_gather with thread-staged windows instead of TMA loads (A/B measurements).
int g_tune[12] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
int tune(int key) {
    static const char* names[12] = {"AESR_CONV_DEBUG", "AESR_CONV_T", "AESR_CONV_NBUF", "AESR_CONV_STAGES", "AESR_HEAD_MMA",
                                    "AESR_STEM_CUDA_CORES", "AESR_WGRAD_NO_FOLD", "AESR_WGRAD_CTAS", "AESR_NO_SPLITK", "AESR_FOLD",
                                    "AESR_NO_TMA_STORE", "AESR_HEAD_GATHER_NO_TMA"};
    if (g_tune[key] < 0) g_tune[key] = getenv(names[key]) ? atoi(getenv(names[key])) : 0;
    return g_tune[key];
}

// Super-tile shape: T M-tiles per pipeline step and nbuf TMEM buffers of T accumulators (nbuf * T * BN <= 512 columns).
// Four buffers whenever the accumulators allow it (T * BN <= 128: the issuer then runs three super-tiles ahead of the
// epilogue), T as large as that allows; the activation ring must still hold >= 3 stages next to the resident filter
// bank (2 * kchunks for multi-chunk layers, 2 for T = 1).
template <int KC>
void halo_pick_T(int BN, int Cin, int tiles_y, int extra, int* T_out, int* stages_out, int* nbuf_out, bool staged = false) {
    using S = HaloSmem<KC>;
    const int kchunks = Cin / KC;
    const int forced = tune(1), forced_nbuf = tune(2), stage_cap = tune(3);
    for (int pass = 0; pass < 2; ++pass) {          // pass 0: four buffers; pass 1: whatever fits
        for (int T = 4; T >= 1; T >>= 1) {
            if (forced && T != forced && T != 1) continue;
            if (T > 1 && tiles_y <= T / 2) continue;
            int nbuf = 512 / (T * BN);
            if (nbuf > 4) nbuf = 4;
            if (forced_nbuf >= 2 && forced_nbuf <= nbuf) nbuf = forced_nbuf;
            if (nbuf == 3) nbuf = 2;                       // power of two: buffer <-> epilogue-set mapping
            if (nbuf < 2) continue;
            if (pass == 0 && nbuf < 4 && !forced && !forced_nbuf && T > 1) continue;
            int stages = CONV_MAX_STAGES;
            if (stage_cap >= 2 && stage_cap < stages) stages = stage_cap;
            while (stages > 1 && S::total_bytes(BN, Cin, T, stages) + extra > g_max_smem_optin) --stages;
            // (with the TMA-store staging behind the tail two stages of a T = 2 super-tile measure like three,
            //  profiles/r01i: dec.2 0.364 vs 0.369 ms, and beat T = 1 with more stages)
            const int need = (T == 1) ? 2 : (kchunks > 1 ? 2 * kchunks : (staged ? 2 : 3));
            if (S::total_bytes(BN, Cin, T, stages) + extra <= g_max_smem_optin && (stages >= need || T == 1)) {
                *T_out = T;
                *stages_out = stages;
                *nbuf_out = nbuf;
                return;
            }
        }
    }
    *T_out = 1;
    *stages_out = 2;
    *nbuf_out = (512 / BN) >= 4 ? 4 : 2;
}

// 32 -> 32 layers: horizontal taps folded into N = 96 (conv3x3_fold.cuh).  Returns -1 when the launch is not one of the
// compiled-in epilogue combinations (FOLD_NOT_APPLICABLE: the caller then takes the tap-by-tap halo kernel).
constexpr int FOLD_NOT_APPLICABLE = 1;
int launch_fold(const void* x, const void* w, ConvParams p, int He, int We, cudaStream_t stream) {
    const bool plain = p.mul_mode == MUL_NONE && p.stats == nullptr && p.out2 == nullptr;
    int mode = -1;
    if (plain && p.out_mode == OUT_SAME) mode = OUT_SAME;
    else if (plain && p.out_mode == OUT_AVGPOOL2) mode = OUT_AVGPOOL2;
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode != MUL_NONE && !p.stats) mode = LEAN_SAME_MUL;
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode != MUL_NONE && p.stats && p.stats_sum_only) mode = LEAN_SAME_MUL_SUM;
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode == MUL_NONE && p.stats && !p.stats_sum_only) mode = LEAN_SAME_STATS;
    if (mode < 0) return FOLD_NOT_APPLICABLE;
    p.BN = 32;
    p.n_blocks = 1;
    p.tiles_x = (We + FOLD_VALID_W - 1) / FOLD_VALID_W;
    p.tiles_y = (He + FOLD_TILE_H - 1) / FOLD_TILE_H;
    // T M-tiles per TMA box / commit and nbuf TMEM buffers of T accumulators: T * nbuf * 96 <= 512 columns
    int T = 1, nbuf = FOLD_MAX_BUF;
    const int forced = tune(1), forced_nbuf = tune(2), stage_cap = tune(3);
    if (forced == 2 && p.tiles_y >= 2) { T = 2; nbuf = 2; }
    // T = 1: at least one buffer per epilogue set (a set may only wait on a buffer use whose predecessor it has seen complete)
    if (T == 1 && forced_nbuf >= CONV_EPI_SETS && forced_nbuf <= FOLD_MAX_BUF) nbuf = forced_nbuf;
    int stages = CONV_MAX_STAGES;
    if (stage_cap >= 2 && stage_cap < stages) stages = stage_cap;
    p.T = T;
    p.nbuf = nbuf;
    p.stiles_y = (p.tiles_y + T - 1) / T;
    p.num_stages = stages;
    const int st_total = p.N * p.tiles_x * p.stiles_y;
    p.num_tiles = st_total;
    CUtensorMap tx, tw;
    int rc = make_act_tmap(&tx, x, p.N, p.H, p.W, 32, FOLD_KC, FOLD_TILE_W, FOLD_TILE_H * T + 2);
    if (rc != AESR_OK) return rc;
    rc = make_wgt_tmap(&tw, w, 9 * 32, 32, FOLD_KC, FOLD_N);
    if (rc != AESR_OK) return rc;
    const int grid = st_total < g_sm_count ? st_total : g_sm_count;
    const int smem = fold_total_bytes(T, stages);
#define AESR_FOLD(MODE)                                                                              \
    case MODE: {                                                                                     \
        static int configured = 0;                                                                   \
        rc = set_max_smem(conv3x3_fold_kernel<MODE>, &configured);                                   \
        if (rc != AESR_OK) return rc;                                                                \
        conv3x3_fold_kernel<MODE><<<grid, CONV_THREADS, smem, stream>>>(tx, tw, p);                  \
    } break;
    switch (mode) {
        AESR_FOLD(OUT_SAME)
        AESR_FOLD(OUT_AVGPOOL2)
        AESR_FOLD(LEAN_SAME_MUL)
        AESR_FOLD(LEAN_SAME_MUL_SUM)
        AESR_FOLD(LEAN_SAME_STATS)
    }
#undef AESR_FOLD
    return check_launch("conv3x3_fold");
}

template <int KC>
int launch_halo(const void* x, const void* w, const void* head_w16, ConvParams p, cudaStream_t stream) {
    using S = HaloSmem<KC>;
    p.BN = halo_pick_bn<KC>(p.Cin, p.Cout);
    if (p.BN == 0) return fail(AESR_ERR_INVALID, "conv3x3_fwd: filter bank %dx%d does not fit the halo kernel", p.Cout, p.Cin);
    p.n_blocks = p.Cout / p.BN;
    const bool head_tc = p.out_mode == OUT_SHUFFLE2_HEAD && head_w16 != nullptr;
    const bool head_mma = p.out_mode == OUT_SHUFFLE2_HEAD && !head_tc && tune(4) != 0;
    p.head_smem = head_tc ? HEAD_SMEM_BYTES : head_mma ? HEAD_MMA_SMEM_BYTES : 0;
    int T = 1, stages = 2, nbuf = 2;
    // inference instantiations (output stage compiled in, no training extras) vs the fully dynamic one
    const bool plain = p.mul_mode == MUL_NONE && p.stats == nullptr && p.out2 == nullptr;
    // TMA-store epilogue (conv3x3_tc.cuh conv_epilogue_lean): 4 epilogue sets x st_bufs x 4 KB behind the tail; two buffers per
    // set when >= 3 activation stages still fit next to them, else one
    // BN >= 64 only: the 32-column layers' short epilogue sits on the MMA -> epilogue -> MMA latency chain and the two named
    // barriers per chunk lengthen it (dec.8 437 -> 457 us with staged stores, profiles/r09_layer_times_tma_store.txt)
    const bool head_cc = p.out_mode == OUT_SHUFFLE2_HEAD && !head_tc && !head_mma;      // CUDA-core head: one 8 KB store per tile
    // training instantiations with an OUT_SAME-layout main output (ConvLean variants, max-pool second output) take it as well
    const bool lean_train = !p.scale && ((p.out_mode == OUT_SAME_MAXPOOL2 && p.mul_mode == MUL_NONE && !p.stats) ||
                                         (p.out_mode == OUT_SAME && !p.out2 && (p.mul_mode != MUL_NONE || p.stats) &&
                                          !(p.mul_mode == MUL_NONE && p.stats && p.stats_sum_only) &&
                                          !(p.mul_mode != MUL_NONE && p.stats && !p.stats_sum_only)));
    const bool tma_st = (((plain && (p.out_mode == OUT_SAME || p.out_mode == OUT_SHUFFLE2)) || lean_train) && p.BN >= 64 || head_cc) &&
                        tune(10) != 1;
    p.st_bufs = 0;
    p.st_bytes = head_cc ? 2 * CONV_ST_CHUNK_BYTES : CONV_ST_CHUNK_BYTES;
    int staging = 0;
    if (tma_st) {
        for (int bufs = 2; bufs >= 1; --bufs) {
            staging = CONV_EPI_SETS * bufs * p.st_bytes + 1024;           // + alignment of the staging area to 1024 bytes
            halo_pick_T<KC>(p.BN, p.Cin, p.tiles_y, p.head_smem + staging, &T, &stages, &nbuf, true);
            p.st_bufs = bufs;
            if (stages >= 3 || bufs == 1) break;
        }
        if (S::total_bytes(p.BN, p.Cin, T, stages) + p.head_smem + staging > g_max_smem_optin || stages < 2) {
            p.st_bufs = 0;                     // no room (largest resident banks): per-thread stores
            staging = 0;
        }
    }
    if (p.st_bufs == 0) halo_pick_T<KC>(p.BN, p.Cin, p.tiles_y, p.head_smem, &T, &stages, &nbuf);
    p.T = T;
    p.nbuf = nbuf;
    p.stiles_y = (p.tiles_y + T - 1) / T;
    p.num_stages = stages;
    const int st_total = p.N * p.tiles_x * p.stiles_y;
    p.num_tiles = st_total * p.n_blocks;
    CUtensorMap tx, tw, th;
    int rc = make_act_tmap(&tx, x, p.N, p.H, p.W, p.Cin, KC, HALO_W, CONV_TILE_H * T + 2);
    if (rc != AESR_OK) return rc;
    rc = make_wgt_tmap(&tw, w, 9 * p.Cout, p.Cin, KC, p.BN);
    if (rc != AESR_OK) return rc;
    th = tw;
    if (head_tc) {      // [16 taps (9 used)][32 channels] 16-bit, one SW64 box
        rc = make_wgt_tmap(&th, head_w16, 16, 32, 32, 16);
        if (rc != AESR_OK) return rc;
    }
    int per_nb = g_sm_count / p.n_blocks;
    if (per_nb < 1) per_nb = 1;
    if (per_nb > st_total) per_nb = st_total;
    const int smem = S::total_bytes(p.BN, p.Cin, T, stages) + p.head_smem + staging;
    const int grid = per_nb * p.n_blocks;
    ConvOutMaps om;
    memset(&om, 0, sizeof(om));
    if (p.st_bufs > 0) {
        if (p.out_mode == OUT_SHUFFLE2_HEAD) {
            // fp32 [N,H,W,16] patches as a 16-bit tensor of 32 "channels": one 64-byte row per low-res pixel
            rc = make_out_tmap(&om.m[0], p.out, p.N, p.H, p.W, 32, 64, static_cast<size_t>(p.W) * 64, static_cast<size_t>(p.H) * p.W * 64, 32);
            if (rc != AESR_OK) return rc;
        } else if (p.out_mode == OUT_SAME || p.out_mode == OUT_SAME_MAXPOOL2) {
            const size_t C = static_cast<size_t>(p.Cout);
            rc = make_out_tmap(&om.m[0], p.out, p.N, p.H, p.W, p.Cout, C * 2, p.W * C * 2, static_cast<size_t>(p.H) * p.W * C * 2);
            if (rc != AESR_OK) return rc;
        } else {
            // depth-to-space: phase ph = 2a+b of low-res pixel (y,x) is hi-res pixel (2y+a, 2x+b) of the [N,2H,2W,C] output
            const size_t C = static_cast<size_t>(p.Cout >> 2), W2 = 2 * static_cast<size_t>(p.W), H2 = 2 * static_cast<size_t>(p.H);
            for (int ph = 0; ph < 4; ++ph) {
                const uint16_t* base = static_cast<const uint16_t*>(p.out) + ((ph >> 1) * W2 + (ph & 1)) * C;
                rc = make_out_tmap(&om.m[ph], base, p.N, p.H, p.W, static_cast<int>(C), 2 * C * 2, 2 * W2 * C * 2, H2 * W2 * C * 2);
                if (rc != AESR_OK) return rc;
            }
        }
    }
#define AESR_HALO_T(MODE, TM)                                                                        \
    {                                                                                                \
        static int configured = 0;                                                                   \
        rc = set_max_smem(conv3x3_halo_kernel<KC, MODE, TM>, &configured);                           \
        if (rc != AESR_OK) return rc;                                                                \
        conv3x3_halo_kernel<KC, MODE, TM><<<grid, CONV_THREADS, smem, stream>>>(tx, tw, th, om, p);  \
    }
#define AESR_HALO(MODE) AESR_HALO_T(MODE, false)
#define AESR_HALO_ST(MODE)                                                                           \
    {                                                                                                \
        if (p.st_bufs > 0) AESR_HALO_T(MODE, true) else AESR_HALO_T(MODE, false)                     \
    }
    if (head_tc) AESR_HALO(OUT_SHUFFLE2_HEAD_TC)
    else if (head_mma) AESR_HALO(OUT_SHUFFLE2_HEAD_MMA)
    else if (p.out_mode == OUT_SHUFFLE2_HEAD) AESR_HALO_ST(OUT_SHUFFLE2_HEAD)
    else if (plain && p.out_mode == OUT_SAME) AESR_HALO_ST(OUT_SAME)
    else if (plain && p.out_mode == OUT_AVGPOOL2) AESR_HALO(OUT_AVGPOOL2)
    else if (plain && p.out_mode == OUT_SHUFFLE2) AESR_HALO_ST(OUT_SHUFFLE2)
    else if (plain && p.out_mode == OUT_SAME_F32) AESR_HALO(OUT_SAME_F32)
    // training instantiations (conv3x3_tc.cuh ConvLean): no eval-BatchNorm affine, no second output except the max-pool
    else if (!p.scale && p.out_mode == OUT_SAME_MAXPOOL2 && p.mul_mode == MUL_NONE && !p.stats) AESR_HALO_ST(OUT_SAME_MAXPOOL2)
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode != MUL_NONE && !p.stats) AESR_HALO_ST(LEAN_SAME_MUL)
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode != MUL_NONE && p.stats && p.stats_sum_only) AESR_HALO_ST(LEAN_SAME_MUL_SUM)
    else if (!p.scale && p.out_mode == OUT_SAME && !p.out2 && p.mul_mode == MUL_NONE && p.stats && !p.stats_sum_only) AESR_HALO_ST(LEAN_SAME_STATS)
    else AESR_HALO(-1)
#undef AESR_HALO_ST
#undef AESR_HALO_T
#undef AESR_HALO
    return check_launch("conv3x3_halo");
}

}  // namespace

extern "C" {

// The calling thread's current device is NOT changed: every entry point launches on the device that is current when it
// is called (PyTorch's device guard / the caller's cudaSetDevice).  SM count and shared-memory limits are taken from the
// first device initialised; every device of a node must be the same B200 part (checked below).
int aesr_init(int device) {
    std::call_once(g_init_once, do_init, device);
    if (g_init_status != AESR_OK) return g_init_status;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(AESR_ERR_CUDA, "no usable CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
    if (prop.major != 10 || prop.multiProcessorCount != g_sm_count)
        return fail(AESR_ERR_ARCH, "aesr_b200: device %d (%s, cc %d.%d, %d SMs) differs from the initialised B200 (%d SMs)",
                    device, prop.name, prop.major, prop.minor, prop.multiProcessorCount, g_sm_count);
    return AESR_OK;
}

const char* aesr_last_error(void) { return g_err; }

int aesr_set_tuning(int key, int value) {
    if (key < 0 || key > 11 || value < 0) return fail(AESR_ERR_INVALID, "set_tuning: key %d value %d", key, value);
    g_tune[key] = value;
    return AESR_OK;
}
int aesr_sm_count(void) { return g_sm_count; }
int64_t aesr_launch_count(void) { return g_launches.load(); }

int aesr_pack_conv3x3_weight(const float* w, void* packed, int Cout, int Cin, int transpose_flip, int dtype,
                             void* stream) {
    if (!w || !packed || Cout <= 0 || Cin <= 0) return fail(AESR_ERR_INVALID, "pack_conv3x3_weight: bad arguments");
    const int total = 9 * Cout * Cin;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        pack_conv3x3_weight_kernel<true><<<(total + 255) / 256, 256, 0, s>>>(w, static_cast<uint16_t*>(packed), Cout, Cin, transpose_flip);
    else
        pack_conv3x3_weight_kernel<false><<<(total + 255) / 256, 256, 0, s>>>(w, static_cast<uint16_t*>(packed), Cout, Cin, transpose_flip);
    return check_launch("pack_conv3x3_weight");
}

int aesr_pack_conv3x3_weight_up2fold(const float* w, void* packed, int Cout, int Cin, int dtype, void* stream) {
    if (!w || !packed || Cout <= 0 || Cin <= 0) return fail(AESR_ERR_INVALID, "pack_conv3x3_weight_up2fold: bad arguments");
    const int total = 36 * Cout * Cin;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        pack_conv3x3_weight_up2fold_kernel<true><<<(total + 255) / 256, 256, 0, s>>>(w, static_cast<uint16_t*>(packed), Cout, Cin);
    else
        pack_conv3x3_weight_up2fold_kernel<false><<<(total + 255) / 256, 256, 0, s>>>(w, static_cast<uint16_t*>(packed), Cout, Cin);
    return check_launch("pack_conv3x3_weight_up2fold");
}

}  // extern "C"

namespace {
// shared by the conv entry points: validate, fill ConvParams, pick the kernel
int conv3x3_dispatch(const void* x, const void* w_packed, const float* bias, const float* scale, const float* shift,
                     void* out, void* out2, const void* mul_src, float* stats, const float* head_w, const void* head_w16,
                     int N, int H, int W,
                     int Cin, int Cout, int act, float slope, int out_mode, int mul_mode, int dtype, int algo,
                     void* stream, int stats_split = 0) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!x || !w_packed || !out) return fail(AESR_ERR_INVALID, "conv3x3_fwd: null tensor");
    if (N <= 0 || H <= 0 || W <= 0) return fail(AESR_ERR_INVALID, "conv3x3_fwd: empty shape N=%d H=%d W=%d", N, H, W);
    if (Cin % 32 != 0 || Cin < 32 || Cin > 512 || (Cin > 32 && Cin % 64 != 0))
        return fail(AESR_ERR_INVALID, "conv3x3_fwd: Cin=%d unsupported (32, 64, 128, 256, 512)", Cin);
    if (Cout % 32 != 0 || Cout < 32 || Cout > 512) return fail(AESR_ERR_INVALID, "conv3x3_fwd: Cout=%d unsupported", Cout);
    if ((scale == nullptr) != (shift == nullptr)) return fail(AESR_ERR_INVALID, "conv3x3_fwd: scale/shift must come together");
    if (out_mode < 0 || out_mode > 7) return fail(AESR_ERR_INVALID, "conv3x3_fwd: out_mode=%d", out_mode);
    if (out_mode == AESR_OUT_SAME_MAXPOOL2 && !out2) return fail(AESR_ERR_INVALID, "conv3x3_fwd: maxpool needs out2");
    if (out_mode == AESR_OUT_SHUFFLE2 && Cout % 128 != 0)
        return fail(AESR_ERR_INVALID, "conv3x3_fwd: OUT_SHUFFLE2 needs Cout = 4*C with C a multiple of 32, got %d", Cout);
    if (out_mode == OUT_SHUFFLE2_HEAD && (Cout != 128 || !head_w || scale || stats || mul_mode != AESR_MUL_NONE))
        return fail(AESR_ERR_INVALID, "conv3x3_up2_head_fwd: needs Cout = 4*32, a head filter and a plain epilogue");
    if (mul_mode != AESR_MUL_NONE && !mul_src) return fail(AESR_ERR_INVALID, "conv3x3_fwd: mul_mode without mul_src");
    if (dtype != AESR_DT_BF16 && dtype != AESR_DT_FP16) return fail(AESR_ERR_INVALID, "conv3x3_fwd: dtype=%d", dtype);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out)) & 15)
        return fail(AESR_ERR_INVALID, "conv3x3_fwd: tensors must be 16-byte aligned");

    ConvParams p{};
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    // AvgPool2d(2) floors: the last row / column of an odd-sized map never reaches the output (65x65 -> 32x32), so
    // tiles only cover the even extent (the halo loads still read the real tensor bounds).  Not when per-channel
    // statistics of the full map are wanted.
    int He = H, We = W;
    if (out_mode == AESR_OUT_AVGPOOL2 && stats == nullptr) {
        He = (H / 2) * 2;
        We = (W / 2) * 2;
        if (He == 0 || We == 0) return AESR_OK;      // empty pooled output
    }
    p.tiles_x = (We + CONV_TILE_W - 1) / CONV_TILE_W;
    p.tiles_y = (He + CONV_TILE_H - 1) / CONV_TILE_H;
    p.fp16 = (dtype == AESR_DT_FP16);
    p.debug = tune(0);                                          // profiling only
    p.bias = bias; p.scale = scale; p.shift = shift; p.slope = slope; p.act = act; p.out_mode = out_mode;
    p.out = out; p.out2 = out2; p.mul_src = static_cast<const uint16_t*>(mul_src); p.mul_mode = mul_mode;
    p.stats = stats;
    p.stats_split = (stats_split > 0 && stats_split < N) ? stats_split : N;
    p.stats_sum_only = stats_split < 0;
    if (stats && Cout > 256) return fail(AESR_ERR_INVALID, "conv3x3_fwd: statistics need Cout <= 256, got %d", Cout);
    if (head_w) memcpy(p.head_wc, head_w, sizeof(p.head_wc));      // HOST pointer: travels as a kernel parameter

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int KC = (Cin >= 64) ? 64 : 32;
    bool halo = false;
    if (algo != AESR_ALGO_STREAM) {
        const int bn = (KC == 64) ? halo_pick_bn<64>(Cin, Cout) : halo_pick_bn<32>(Cin, Cout);
        halo = bn > 0;
        if (!halo && algo == AESR_ALGO_HALO)
            return fail(AESR_ERR_INVALID, "conv3x3_fwd: filter bank %dx%d too large for AESR_ALGO_HALO", Cout, Cin);
        if (out_mode == OUT_SHUFFLE2_HEAD && bn != Cout) halo = false;     // all four phases must sit in one tile
        // Cin = 256: the resident bank only fits with 32-column tiles (40 % MMA ceiling, the activation re-read by Cout / 32
        // n-blocks); the streamed kernel with 256-column tiles is 1.4-1.5x faster on the VGG conv3/conv4 shapes
        // (profiles/r06d_vgg_halo_vs_stream_sweep.txt: 256->256 @32^2 n24 60 vs 40 us, 256->512 @16^2 37 vs 26 us)
        if (halo && algo == AESR_ALGO_AUTO && bn < 64 && Cout >= 128 && out_mode != AESR_OUT_SHUFFLE2) halo = false;
    }
    if (out_mode == OUT_SHUFFLE2_HEAD && !halo)
        return fail(AESR_ERR_INVALID, "conv3x3_up2_head_fwd: Cin=%d: the 128-row folded bank must fit the resident-filter kernel", Cin);
    if (halo && Cin == 32 && Cout == 32 && algo == AESR_ALGO_AUTO && tune(9) == 1) {
        rc = launch_fold(x, w_packed, p, He, We, s);
        if (rc != FOLD_NOT_APPLICABLE) return rc;
    }
    if (halo) return KC == 64 ? launch_halo<64>(x, w_packed, head_w16, p, s) : launch_halo<32>(x, w_packed, head_w16, p, s);
    return KC == 64 ? launch_stream<64>(x, w_packed, p, s) : launch_stream<32>(x, w_packed, p, s);
}
}  // namespace

extern "C" {

int aesr_pack_conv3x3_weight_batch(const float* src_base, void* dst_base, const long long* jobs_dev, int n_jobs,
                                   int max_elems, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!src_base || !dst_base || !jobs_dev || n_jobs <= 0 || n_jobs > 65535 || max_elems <= 0)
        return fail(AESR_ERR_INVALID, "pack_conv3x3_weight_batch: bad arguments");
    int gx = (max_elems + 255) / 256;
    if (gx > 64) gx = 64;
    const dim3 grid(gx, n_jobs);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16) pack_conv3x3_weight_batch_kernel<true><<<grid, 256, 0, s>>>(src_base, static_cast<uint16_t*>(dst_base), jobs_dev);
    else pack_conv3x3_weight_batch_kernel<false><<<grid, 256, 0, s>>>(src_base, static_cast<uint16_t*>(dst_base), jobs_dev);
    return check_launch("pack_conv3x3_weight_batch");
}

int aesr_conv3x3_fwd(const void* x, const void* w_packed, const float* bias, const float* scale, const float* shift,
                     void* out, void* out2, const void* mul_src, float* stats, int N, int H, int W, int Cin, int Cout,
                     int act, float slope, int out_mode, int mul_mode, int dtype, int algo, int stats_split, void* stream) {
    if (out_mode == OUT_SHUFFLE2_HEAD) return fail(AESR_ERR_INVALID, "conv3x3_fwd: use aesr_conv3x3_up2_head_fwd");
    return conv3x3_dispatch(x, w_packed, bias, scale, shift, out, out2, mul_src, stats, nullptr, nullptr, N, H, W, Cin, Cout,
                            act, slope, out_mode, mul_mode, dtype, algo, stream, stats_split);
}

int aesr_conv3x3_up2_head_fwd(const void* x, const void* w_folded, const float* bias, const float* head_w9c_host,
                              const void* head_w16, float* partial, int N, int H, int W, int Cin, int act, float slope,
                              int dtype, int algo, void* stream) {
    if (head_w16 && (reinterpret_cast<uintptr_t>(head_w16) & 15))
        return fail(AESR_ERR_INVALID, "conv3x3_up2_head_fwd: head_w16 must be 16-byte aligned");
    return conv3x3_dispatch(x, w_folded, bias, nullptr, nullptr, partial, nullptr, nullptr, nullptr, head_w9c_host, head_w16,
                            N, H, W, Cin, 128, act, slope, OUT_SHUFFLE2_HEAD, AESR_MUL_NONE, dtype, algo, stream);
}

int aesr_head_gather(const float* partial, const float* bias, float* out, const int* out_index, int N, int h, int w,
                     size_t out_image_stride, int apply_sigmoid, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!partial || !bias || !out || N <= 0 || h <= 0 || w <= 0) return fail(AESR_ERR_INVALID, "head_gather: bad arguments");
    if ((out_image_stride & 1) || (reinterpret_cast<uintptr_t>(out) & 7))
        return fail(AESR_ERR_INVALID, "head_gather: output images must be 8-byte aligned (even stride)");
    if (static_cast<size_t>(h) * w > (1u << 28)) return fail(AESR_ERR_INVALID, "head_gather: image too large");
    dim3 grid((w + HG_TW - 1) / HG_TW, (h + HG_TH - 1) / HG_TH, N < 65535 ? N : 65535);
    if (grid.y > 65535) return fail(AESR_ERR_INVALID, "head_gather: image too tall");
    if (tune(11) != 1 && !(reinterpret_cast<uintptr_t>(partial) & 15)) {
        // TMA version: fp32 [N,h,w,16] as {16, w, h, N}, one (8+2) x (32+2) window of patches per load, zero fill outside the image;
        // each block walks through the images of its window position (double-buffered loads): ~5 resident blocks per SM
        CUtensorMap tm;
        cuuint64_t dims[4] = {16, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)N};
        cuuint64_t strides[3] = {64, (cuuint64_t)w * 64, (cuuint64_t)h * w * 64};
        cuuint32_t box[4] = {16, (cuuint32_t)(HG_TW + 2), (cuuint32_t)(HG_TH + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(partial), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AESR_ERR_CUDA, "cuTensorMapEncodeTiled(head patches) failed: %d", (int)r);
        static int configured = 0;
        {
            int dev = 0;
            CUDA_TRY(cudaGetDevice(&dev));
            const int bit = 1 << (dev & 31);
            if (!(configured & bit)) {
                CUDA_TRY(cudaFuncSetAttribute(head_gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HGT_SMEM_BYTES));
                configured |= bit;
            }
        }
        const int per_pos = grid.x * grid.y;
        int gz = (5 * g_sm_count + per_pos - 1) / per_pos;
        if (gz < 1) gz = 1;
        if (gz > N) gz = N;
        if (gz > 65535) gz = 65535;
        grid.z = gz;
        head_gather_tma_kernel<<<grid, 256, HGT_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
            tm, bias, out, out_index, N, h, w, out_image_stride, apply_sigmoid);
        return check_launch("head_gather_tma");
    }
    head_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        partial, bias, out, out_index, N, h, w, out_image_stride, apply_sigmoid);
    return check_launch("head_gather");
}

int aesr_stem_fold(const float* w0, const float* b0, const float* w1, float* weff, float* beff, int C, void* stream) {
    if (!w0 || !b0 || !w1 || !weff || !beff || C != 32) return fail(AESR_ERR_INVALID, "stem_fold: bad arguments (C = 32)");
    stem_fold_kernel<<<(9 * C + 95) / 96, 96, 0, static_cast<cudaStream_t>(stream)>>>(w0, b0, w1, weff, beff, C);
    return check_launch("stem_fold");
}

int aesr_stem_fwd(const float* x, const float* weff_beff_b1_host, void* out, int N, int H, int W, int C, float slope,
                  int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!x || !weff_beff_b1_host || !out || N <= 0 || H <= 0 || W <= 0 || C != 32)
        return fail(AESR_ERR_INVALID, "stem_fwd: bad arguments (C = 32)");
    if (static_cast<size_t>(H + 2) * (W + 2) > (1u << 28)) return fail(AESR_ERR_INVALID, "stem_fwd: image too large");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float *weff = weff_beff_b1_host, *beff = weff_beff_b1_host + 9 * 32, *b1 = weff_beff_b1_host + 18 * 32;
    if (tune(5) != 1) {
        // warp-level tensor-core stem: bias of the nine border classes (first / inner / last row x column) in fp32
        StemMmaParams mp;
        memcpy(mp.weff, weff, sizeof(mp.weff));
        for (int cy = 0; cy < 3; ++cy)
            for (int cx = 0; cx < 3; ++cx)
                for (int c = 0; c < 32; ++c) {
                    float t = b1[c];
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap % 3;
                        const bool on_grid = !(cy == 0 && dy == 0) && !(cy == 2 && dy == 2) && !(cx == 0 && dx == 0) &&
                                             !(cx == 2 && dx == 2);
                        if (on_grid) t += beff[tap * 32 + c];
                    }
                    mp.bias_tab[(cy * 3 + cx) * 32 + c] = t;
                }
        const int blocks = ((H + 2) * (W + 2) + 15) / 16;
        const int gx = (blocks + 3) / 4;
        // images are dealt to gridDim.y CTAs per pixel block; pick the split whose CTA count fills whole waves of the
        // resident capacity (5 CTAs of 128 threads per SM at 91 registers) -- 265 x 5 CTAs left the second wave 79 % full
        const long cap = 5L * g_sm_count;
        int gy = 1;
        double best = -1.0;
        for (int c = 1; c <= 16 && c <= N; ++c) {
            const long total = static_cast<long>(gx) * c;
            const double fill = static_cast<double>(total) / (static_cast<double>((total + cap - 1) / cap) * cap);
            if (fill > best + 0.02) { best = fill; gy = c; }
        }
        const dim3 grid(gx, gy);
        const int variant = tune(5);
        if (variant == 2) {
            if (dtype == AESR_DT_FP16) stem_mma_kernel<true, 1><<<grid, 128, 0, s>>>(x, mp, static_cast<uint16_t*>(out), N, H, W, slope);
            else stem_mma_kernel<false, 1><<<grid, 128, 0, s>>>(x, mp, static_cast<uint16_t*>(out), N, H, W, slope);
        } else if (dtype == AESR_DT_FP16)
            stem_mma_kernel<true, 3><<<grid, 128, 0, s>>>(x, mp, static_cast<uint16_t*>(out), N, H, W, slope);
        else
            stem_mma_kernel<false, 3><<<grid, 128, 0, s>>>(x, mp, static_cast<uint16_t*>(out), N, H, W, slope);
        return check_launch("stem_mma");
    }
    StemParams sp;
    memcpy(sp.weff, weff, sizeof(sp.weff));
    memcpy(sp.beff, beff, sizeof(sp.beff));
    memcpy(sp.b1, b1, sizeof(sp.b1));
    for (int c = 0; c < 32; ++c) {
        float t = sp.b1[c];
        for (int tap = 0; tap < 9; ++tap) t += sp.beff[tap * 32 + c];
        sp.ball[c] = t;
    }
    const int per_img = (H + 2) * (W + 2);
    const int groups = (N + STEM_P - 1) / STEM_P;
    const dim3 grid((per_img + 127) / 128, groups < 65535 ? groups : 65535);
    if (dtype == AESR_DT_FP16)
        stem_conv_kernel<true><<<grid, 128, 0, s>>>(x, sp, static_cast<uint16_t*>(out), N, H, W, slope);
    else
        stem_conv_kernel<false><<<grid, 128, 0, s>>>(x, sp, static_cast<uint16_t*>(out), N, H, W, slope);
    return check_launch("stem_conv");
}

int aesr_e0_fwd(const float* x, const float* w, const float* b, void* out, int N, int H, int W, int C, int dtype,
                void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!x || !w || !b || !out || N <= 0 || H <= 0 || W <= 0 || C % 8 != 0)
        return fail(AESR_ERR_INVALID, "e0_fwd: bad arguments");
    const size_t total = static_cast<size_t>(N) * (H + 2) * (W + 2) * (C / 8);
    const int block = 256;
    size_t grid = (total + block - 1) / block;
    const size_t cap = static_cast<size_t>(g_sm_count) * 16;
    if (grid > cap) grid = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        e0_conv1x1_pad1_kernel<true><<<static_cast<int>(grid), block, 0, s>>>(x, w, b, static_cast<uint16_t*>(out), N, H, W, C);
    else
        e0_conv1x1_pad1_kernel<false><<<static_cast<int>(grid), block, 0, s>>>(x, w, b, static_cast<uint16_t*>(out), N, H, W, C);
    return check_launch("e0_conv1x1_pad1");
}

int aesr_head_fwd(const void* in, const float* w9c, const float* bias, float* out, const int* out_index, int N, int H, int W,
                  int C, size_t out_image_stride, int apply_sigmoid, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!in || !w9c || !bias || !out || N <= 0 || H <= 0 || W <= 0) return fail(AESR_ERR_INVALID, "head_fwd: bad arguments");
    if (C != 32) return fail(AESR_ERR_INVALID, "head_fwd: C=%d unsupported (32)", C);
    const size_t total = static_cast<size_t>(N) * H * W;
    const int block = 128;
    size_t grid = (total + block - 1) / block;
    const size_t cap = static_cast<size_t>(g_sm_count) * 32;
    if (grid > cap) grid = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint16_t* in16 = static_cast<const uint16_t*>(in);
    if (N <= 65535) {        // shared-memory tiled kernel: one block per 16x16 output pixels of one image
        dim3 tgrid((W + 15) / 16, (H + 15) / 16, N);
        if (dtype == AESR_DT_FP16)
            head_conv3x3_tiled_kernel<true><<<tgrid, 256, 0, s>>>(in16, w9c, bias, out, out_index, H, W, out_image_stride, apply_sigmoid);
        else
            head_conv3x3_tiled_kernel<false><<<tgrid, 256, 0, s>>>(in16, w9c, bias, out, out_index, H, W, out_image_stride, apply_sigmoid);
        return check_launch("head_conv3x3_tiled");
    }
    if (dtype == AESR_DT_FP16)
        head_conv3x3_sigmoid_kernel<32, true><<<static_cast<int>(grid), block, 0, s>>>(in16, w9c, bias, out, out_index, N, H, W, out_image_stride, apply_sigmoid);
    else
        head_conv3x3_sigmoid_kernel<32, false><<<static_cast<int>(grid), block, 0, s>>>(in16, w9c, bias, out, out_index, N, H, W, out_image_stride, apply_sigmoid);
    return check_launch("head_conv3x3_sigmoid");
}

int aesr_lerp_latents(const float* z, const int* ia, const int* ib, const float* wa, const float* wb, void* out_nhwc,
                      float* out_nchw, int M, int C, int HW, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!z || !ia || !ib || !wa || !wb || !out_nhwc || C <= 0 || HW <= 0) return fail(AESR_ERR_INVALID, "lerp_latents: bad arguments");
    if (M == 0) return AESR_OK;
    if (M < 0 || M > 65535) return fail(AESR_ERR_INVALID, "lerp_latents: M=%d out of range (1..65535 per call)", M);
    dim3 grid((HW + 31) / 32, (C + 31) / 32, M), block(32, 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        lerp_nchw_to_nhwc_kernel<true><<<grid, block, 0, s>>>(z, ia, ib, wa, wb, static_cast<uint16_t*>(out_nhwc), out_nchw, C, HW);
    else
        lerp_nchw_to_nhwc_kernel<false><<<grid, block, 0, s>>>(z, ia, ib, wa, wb, static_cast<uint16_t*>(out_nhwc), out_nchw, C, HW);
    return check_launch("lerp_nchw_to_nhwc");
}

int aesr_lerp_pairs(const float* z, const int* pa, const int* pb, const float* wa, const float* wb, void* out_nhwc,
                    int P, int K, int C, int HW, int dtype, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!z || !pa || !pb || !wa || !wb || !out_nhwc || C <= 0 || HW <= 0 || K <= 0) return fail(AESR_ERR_INVALID, "lerp_pairs: bad arguments");
    if (P == 0) return AESR_OK;
    if (P < 0 || P > 65535) return fail(AESR_ERR_INVALID, "lerp_pairs: P=%d out of range (1..65535 per call)", P);
    dim3 grid((HW + 31) / 32, (C + 31) / 32, P), block(32, 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        lerp_pairs_kernel<true><<<grid, block, 0, s>>>(z, pa, pb, wa, wb, static_cast<uint16_t*>(out_nhwc), K, C, HW);
    else
        lerp_pairs_kernel<false><<<grid, block, 0, s>>>(z, pa, pb, wa, wb, static_cast<uint16_t*>(out_nhwc), K, C, HW);
    return check_launch("lerp_pairs");
}

int aesr_lerp_pairs_act(const float* pre, const int* pa, const int* pb, const float* wa, const float* wb,
                        const float* bias, void* out_nhwc, int P, int K, int C, int HW, float slope, int dtype,
                        void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!pre || !pa || !pb || !wa || !wb || !out_nhwc || C <= 0 || HW <= 0 || K <= 0)
        return fail(AESR_ERR_INVALID, "lerp_pairs_act: bad arguments");
    if (C % 8 != 0) return fail(AESR_ERR_INVALID, "lerp_pairs_act: C=%d must be a multiple of 8", C);
    if (static_cast<size_t>(HW) * C > (1u << 30)) return fail(AESR_ERR_INVALID, "lerp_pairs_act: slice too large");
    if ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(out_nhwc) | reinterpret_cast<uintptr_t>(bias)) & 15)
        return fail(AESR_ERR_INVALID, "lerp_pairs_act: tensors must be 16-byte aligned");
    if (P == 0) return AESR_OK;
    if (P < 0 || P > 65535) return fail(AESR_ERR_INVALID, "lerp_pairs_act: P=%d out of range (1..65535 per call)", P);
    const int HWC = HW * C;
    int gx = ((HWC >> 3) + 255) / 256;
    if (gx > 32) gx = 32;
    dim3 grid(gx, P);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AESR_DT_FP16)
        lerp_pairs_act_kernel<true><<<grid, 256, 0, s>>>(pre, pa, pb, wa, wb, bias, static_cast<uint16_t*>(out_nhwc), K, C, HWC, slope);
    else
        lerp_pairs_act_kernel<false><<<grid, 256, 0, s>>>(pre, pa, pb, wa, wb, bias, static_cast<uint16_t*>(out_nhwc), K, C, HWC, slope);
    return check_launch("lerp_pairs_act");
}

int aesr_place_slices(const float* src, float* dst, const int* out_index, int N, int HW, int do_clamp, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!src || !dst || HW <= 0) return fail(AESR_ERR_INVALID, "place_slices: bad arguments");
    if (N == 0) return AESR_OK;
    if (N < 0 || N > 65535) return fail(AESR_ERR_INVALID, "place_slices: N=%d out of range (1..65535 per call)", N);
    const int block = 256;
    int gx = ((HW >> 2) + block - 1) / block;
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    place_slices_kernel<<<dim3(gx, N), block, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, out_index, N, HW, do_clamp);
    return check_launch("place_slices");
}

int aesr_copy_rows_async(void* dst, size_t dst_outer_stride, size_t dpitch, const void* src, size_t src_outer_stride,
                         size_t spitch, size_t width, size_t rows, size_t outer, int to_host, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!dst || !src || width == 0 || width > dpitch || width > spitch)
        return fail(AESR_ERR_INVALID, "copy_rows_async: bad arguments (width %zu, pitches %zu / %zu)", width, dpitch, spitch);
    const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    for (size_t o = 0; o < outer; ++o) {
        const cudaError_t e = cudaMemcpy2DAsync(static_cast<char*>(dst) + o * dst_outer_stride, dpitch,
                                                static_cast<const char*>(src) + o * src_outer_stride, spitch, width, rows,
                                                kind, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return fail(AESR_ERR_CUDA, "copy_rows_async: %s", cudaGetErrorString(e));
    }
    return AESR_OK;
}

#ifdef AESR_WITH_PROBES
int aesr_probe_halo_conv(const void* x, const void* w_packed, float* out, int N, int H, int W, int x0, int y0, int n,
                         int pitch, int variant, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (pitch != 10 && pitch != 16) return fail(AESR_ERR_INVALID, "probe: pitch must be 10 or 16");
    CUtensorMap tx, tw;
    rc = make_act_tmap(&tx, x, N, H, W, 64, 64, pitch, 18);
    if (rc != AESR_OK) return rc;
    rc = make_wgt_tmap(&tw, w_packed, 9 * 64, 64, 64, 64);
    if (rc != AESR_OK) return rc;
    const int smem = 1024 + 36864 + 73728 + 64;
    CUDA_TRY(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    halo_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tx, tw, out, x0, y0, n, pitch, variant);
    return check_launch("halo_probe");
}

int aesr_probe_umma_rate(long long* cycles, int N, int kc, int pitch_rows, int shift_rows, int iters,
                         int a_advance_rows, int nacc, int grid, int fill_random, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!cycles || N < 16 || N > 256 || N % 16 || (kc != 32 && kc != 64) || iters <= 0 || nacc < 1 || nacc * N > 512 ||
        grid < 1 || grid > 1024)
        return fail(AESR_ERR_INVALID, "probe_umma_rate: bad arguments");
    const int smem = 1024 + 160 * 1024;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define AESR_PROBE(NA)                                                                                             \
    case NA:                                                                                                       \
        CUDA_TRY(cudaFuncSetAttribute(umma_rate_probe_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        umma_rate_probe_kernel<NA><<<grid, 128, smem, s>>>(cycles, N, kc, pitch_rows, shift_rows, iters, a_advance_rows, \
                                                          fill_random);                                            \
        break;
    switch (nacc) {
        AESR_PROBE(1) AESR_PROBE(2) AESR_PROBE(4) AESR_PROBE(8)
        default: return fail(AESR_ERR_INVALID, "probe_umma_rate: nacc must be 1, 2, 4 or 8");
    }
#undef AESR_PROBE
    return check_launch("umma_rate_probe");
}

int aesr_probe_umma_pattern(long long* cycles, int BN, int kc, int T, int iters, int variant, int lag, int grid,
                            int fill_random, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!cycles || BN < 32 || BN > 256 || BN % 32 || (kc != 32 && kc != 64) || T < 1 || T > 4 || T * BN > 256 ||
        iters <= 0 || lag < 1 || lag > 7 || grid < 1 || grid > 1024 || 9 * BN * kc * 2 > 148 * 1024 ||
        (16 * T + 2) * 10 * kc * 2 > 48 * 1024)
        return fail(AESR_ERR_INVALID, "probe_umma_pattern: bad arguments");
    const int smem = 1024 + 196 * 1024;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (kc == 64) {
        CUDA_TRY(cudaFuncSetAttribute(umma_pattern_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umma_pattern_probe_kernel<64><<<grid, 576, smem, s>>>(cycles, BN, T, iters, variant, lag, fill_random);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(umma_pattern_probe_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umma_pattern_probe_kernel<32><<<grid, 576, smem, s>>>(cycles, BN, T, iters, variant, lag, fill_random);
    }
    return check_launch("umma_pattern_probe");
}

int aesr_probe_tmem_ld(long long* cycles, int nwarps, int iters, int mode, int grid, float* sink, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!cycles || !sink || nwarps < 4 || nwarps > 16 || nwarps % 4 || iters <= 0 || mode < 0 || mode > 2 || grid < 1 || grid > 1024)
        return fail(AESR_ERR_INVALID, "probe_tmem_ld: bad arguments");
    tmem_ld_probe_kernel<<<grid, 32 * nwarps, 0, static_cast<cudaStream_t>(stream)>>>(cycles, iters, mode, sink);
    return check_launch("tmem_ld_probe");
}

int aesr_probe_sync(long long* cycles, int iters, int mode, void* stream) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!cycles || iters <= 0) return fail(AESR_ERR_INVALID, "probe_sync: bad arguments");
    sync_probe_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(cycles, iters, mode);
    return check_launch("sync_probe");
}

int aesr_probe_launch_gap(float* us_per_kernel, int n_kernels, int ctas, int spin_cycles, int pdl, int replays) {
    int rc = ensure_init();
    if (rc != AESR_OK) return rc;
    if (!us_per_kernel || n_kernels <= 0 || ctas <= 0 || replays <= 0) return fail(AESR_ERR_INVALID, "probe_launch_gap: bad arguments");
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreate(&s));
    int* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, sizeof(int)));
    CUDA_TRY(cudaMemset(sink, 0, sizeof(int)));
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n_kernels; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(128);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, launch_gap_probe_kernel, sink, spin_cycles));
    }
    CUDA_TRY(cudaStreamEndCapture(s, &graph));
    CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaGraphLaunch(exec, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaEventRecord(e0, s));
    for (int r = 0; r < replays; ++r) CUDA_TRY(cudaGraphLaunch(exec, s));
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *us_per_kernel = ms * 1e3f / (static_cast<float>(replays) * n_kernels);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaStreamDestroy(s);
    return AESR_OK;
}
#endif  // AESR_WITH_PROBES

}  // extern "C"

#include "api_train.cuh"
