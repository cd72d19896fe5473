// Shared declarations for the sunb200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing: no C++ exceptions cross the C ABI; every entry point returns an int code and
// leaves a message for sunb_last_error().
// ---------------------------------------------------------------------------------------------
enum SunbStatus { SUNB_OK = 0, SUNB_ERR_ARG = -1, SUNB_ERR_CUDA = -2, SUNB_ERR_WORKSPACE = -3, SUNB_ERR_DRIVER = -4 };

void sunb_set_error(const char* fmt, ...);

#define SUNB_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            sunb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,               \
                           cudaGetErrorString(_e));                                            \
            return SUNB_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define SUNB_REQUIRE(cond, ...)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            sunb_set_error(__VA_ARGS__);                                                       \
            return SUNB_ERR_ARG;                                                               \
        }                                                                                      \
    } while (0)

#define SUNB_TRY(expr)                                                                         \
    do {                                                                                       \
        int _s = (expr);                                                                       \
        if (_s != SUNB_OK) return _s;                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------
// GEMM / implicit-GEMM problem description shared by the tcgen05 kernel and the SIMT checker kernel
//   C[m, g*c_goff + n] = epilogue( sum_{tap, k} A_tap[m, g*a_goff + k] * Wt[(g*taps + tap)*N + n, k] )
// a_mode 0: A_tap[m,:] = A[m,:]                        (1x1 conv / linear; taps == 1)
// a_mode 1: 3x3 conv, pad 1, stride 1 over an NHWC image [B,H,W,C]; A_tap[m,:] = pixel shifted by
//           (tap/3-1, tap%3-1), zero outside the image.  Output tiles are built from (bw x bh)-pixel
//           sub-boxes so that every 128-row tile is a set of complete spatial boxes.
// ---------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_GELU = 2 };
enum { MAP_IDENT = 0, MAP_S2D = 1 };   // output row mapping: identity, or 2x2 space-to-depth of an oH x oW raster

struct GemmParams {
    int M, N, K;          // rows, output columns per group, reduction length per tap (real, un-padded)
    int taps, groups;
    int a_goff, c_goff;
    int a_mode, H, W, bw, bh;
    // operands (raw pointers: SIMT kernel + tensor-map creation)
    const bf16* A; int lda;          // 2-D mode: row stride in elements.  conv mode: lda = channels per pixel
    const bf16* Wt; int ldw;         // weights, K-major rows of ldw elements
    // epilogue
    const float* bias; int bias_mod;      // bias[(m % bias_mod) * bias_ld + col]; bias_mod == 1 -> plain per-column bias
    int bias_ld;
    int act;
    const bf16* resid; int ldr;           // added before the activation; indexed with the raster row m
    const float* row_scale; int rows_per_img;   // optional per-image scale of the accumulator (DropPath)
    bf16* out; int ldc;                   // bf16 output (nullable)
    float* out_f32; int ldc_f32;          // fp32 output (nullable)
    int out_map, oH, oW;
    // training extras
    bf16* out2; int ldc2;                 // pre-activation copy (value before `act`), bf16, nullable
    const bf16* dact_aux; int ld_aux;     // backward: multiply the result by act'(aux[m, col]) of kind `dact`
    int dact;
    // fused 2x2 max-pool + position table behind the activation (conv_slab.cu only; eval stem tail, visformer.py:237,431):
    //   pool_out[img, y/2, x/2, col] = max over the 2x2 window of act(...) + pool_pos[(y/2)*(W/2) + x/2, col]
    bf16* pool_out; const float* pool_pos;
};

// MUFU approximations (rel. error ~2^-22), one instruction each
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// erf via Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 output resolution).
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
    float y = fmaf(1.061405429f, t, -1.453152027f);
    y = fmaf(y, t, 1.421413741f);
    y = fmaf(y, t, -0.284496736f);
    y = fmaf(y, t, 0.254829592f);
    y = 1.f - y * t * ex2_approx(-1.4426950408889634f * ax * ax);
    return copysignf(y, x);
}

// Phi(v) = 0.5 * (1 + erf(v / sqrt 2)) through the A&S 7.1.26 polynomial with the 1/sqrt2 and 0.5 factors folded into the
// constants: h = 0.5 * poly(t) * exp(-v^2 / 2), Phi = v >= 0 ? 1 - h : h (abs err 1.5e-7).  `e` returns exp(-v^2 / 2), which the
// derivative needs as well (Phi' = e / sqrt(2 pi)): 2 MUFU for Phi and Phi' together.
__device__ __forceinline__ float gelu_phi_e(float v, float& e) {
    const float t = rcp_approx(fmaf(0.2316418882f, fabsf(v), 1.f));          // 0.3275911 / sqrt(2)
    float y = fmaf(0.5307027145f, t, -0.7265760135f);                         // 0.5 * A&S coefficients
    y = fmaf(y, t, 0.7107068705f);
    y = fmaf(y, t, -0.142248368f);
    y = fmaf(y, t, 0.127414796f);
    e = ex2_approx(-0.7213475204444817f * v * v);                             // exp(-v^2/2)
    const float h = y * t * e;
    return v >= 0.f ? 1.f - h : h;
}
__device__ __forceinline__ float gelu_phi(float v) {
    float e;
    return gelu_phi_e(v, e);
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Forward GELU of nn.GELU() (exact-erf form, reference visformer.py:139).  The GELU epilogues are bound by the MUFU pipe
// (a quarter-rate unit): the A&S form costs 2 MUFU + 14 ALU per element, this one 1 MUFU + 7 ALU.
//   v * Phi(v) = 0.5 v (1 + tanh(P(v))),  P(v) = v (a + b w + c w^2),  w = min(v^2, 64)
// with (a, b, c) a minimax fit of P to atanh(erf(v / sqrt 2)): |deviation from the erf form| <= 2.5e-5 for every v, plus the
// tanh.approx error (<= 2^-11 relative on tanh).  Both are below the bf16 rounding of the stored activation; the measured
// device error is asserted in tests/test_gpu_kernels.py::test_gelu_device_accuracy.  -DSUNB_GELU_ERF restores the A&S form.
__device__ __forceinline__ float gelu_fast(float v) {
#ifdef SUNB_GELU_ERF
    return v * gelu_phi(v);
#else
    const float w = fminf(v * v, 64.f);
    const float p = fmaf(fmaf(-3.5151678e-4f, w, 3.7005646e-2f), w, 0.797507884f);
    const float hv = 0.5f * v;
    return fmaf(hv, tanh_approx(v * p), hv);
#endif
}

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == ACT_LRELU) return v > 0.f ? v : 0.1f * v;
    if (act == ACT_GELU) return gelu_fast(v);
    return v;
}

// derivative of the activation evaluated at the saved tensor (GELU: pre-activation; LeakyReLU: either side of it)
__device__ __forceinline__ float gelu_grad(float x) {       // Phi(x) + x * phi(x), erf form (training gradients)
    float e;
    const float phi = gelu_phi_e(x, e);
    return fmaf(x * 0.3989422804014327f, e, phi);
}
__device__ __forceinline__ float act_grad(float x, int act) {
    if (act == ACT_LRELU) return x > 0.f ? 1.f : 0.1f;
    if (act == ACT_GELU) return gelu_grad(x);
    return 1.f;
}

// raster row of tile-row r.  A conv-mode tile is one (bw x bh) spatial block of 128/(bw*bh) consecutive images, so the
// whole A tile is a single 4-D TMA box (channels, bw, bh, images).  Images past the batch map to M (dropped).
__device__ __forceinline__ int conv_tile_row_to_pixel(const GemmParams& p, int tile, int r) {
    const int box = p.bw * p.bh;
    const int nimg = 128 / box;
    const int tiles_x = p.W / p.bw, spi = tiles_x * (p.H / p.bh);
    const int blk = tile % spi, img = (tile / spi) * nimg + r / box;
    const int rr = r % box;
    const int y = (blk / tiles_x) * p.bh + rr / p.bw;
    const int x = (blk % tiles_x) * p.bw + rr % p.bw;
    const int m = (img * p.H + y) * p.W + x;
    return m < p.M ? m : p.M;
}

__device__ __forceinline__ int map_out_row(const GemmParams& p, int m) {
    if (p.out_map == MAP_S2D) {
        const int hw = p.oH * p.oW;
        const int img = m / hw, rem = m % hw;
        const int y = rem / p.oW, x = rem % p.oW;
        return ((img * (p.oH / 2) + y / 2) * (p.oW / 2) + x / 2) * 4 + (y & 1) * 2 + (x & 1);
    }
    return m;
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per thread-instruction, so the row-strided
// epilogue traffic is made of whole sectors instead of 16-byte halves.
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// 16 consecutive bf16 <-> 16 floats through one 256-bit access
__device__ __forceinline__ void load16_bf16(const bf16* p, float* f) {
    uint32_t r[8];
    ld_global_256(p, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&r[j]);
        f[2 * j] = __bfloat162float(h.x);
        f[2 * j + 1] = __bfloat162float(h.y);
    }
}
__device__ __forceinline__ void store16_bf16(bf16* p, const float* f) {
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    st_global_256(p, r);
}

// Epilogue for NC consecutive columns [col, col+NC) of raster row m of group g.  v = fp32 accumulators.
// Every per-element loop is branch-free inside (the activation switch is hoisted) so the unrolled bodies interleave.
template <int NC>
__device__ __forceinline__ void epilogue_row(const GemmParams& p, int g, int m, int col, float* v) {
    if (m >= p.M) return;
    const int ncol = min(NC, p.N - col);
    if (ncol <= 0) return;
    const int gcol = g * p.c_goff + col;
    const bool full = (ncol == NC) && (NC % 8 == 0);
    if (p.row_scale) {
        const float rs = p.row_scale[m / p.rows_per_img];
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] *= rs;
    }
    if (p.resid) {
        const bf16* res = p.resid + (size_t)m * p.ldr + gcol;
        if (full && (NC % 16 == 0) && ((((size_t)res) & 31) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 16) {
                float f[16];
                load16_bf16(res + i, f);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[i + j] += f[j];
            }
        } else if (full && ((((size_t)res) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(res + i);
                const bf16* h = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i + j] += __bfloat162float(h[j]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < ncol) v[i] += __bfloat162float(res[i]);
        }
    }
    if (p.bias) {
        const float* bias = p.bias + (size_t)(m % p.bias_mod) * p.bias_ld + gcol;
        if (full && ((((size_t)bias) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(bias + i);
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < ncol) v[i] += bias[i];
        }
    }
    const int orow = map_out_row(p, m);
    if (p.out2) {                                   // training forward: keep the pre-activation for the backward pass
        bf16* o2 = p.out2 + (size_t)orow * p.ldc2 + gcol;
        if (full && (NC % 16 == 0) && ((((size_t)o2) & 31) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 16) store16_bf16(o2 + i, v + i);
        } else if (full && ((((size_t)o2) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                *reinterpret_cast<uint4*>(o2 + i) = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < ncol) o2[i] = __float2bfloat16(v[i]);
        }
    }
    if (p.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = gelu_fast(v[i]);
    } else if (p.act == ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = v[i] > 0.f ? v[i] : 0.1f * v[i];
    }
    if (p.dact_aux) {                               // backward: chain through the activation of the producing layer
        const bf16* ax = p.dact_aux + (size_t)m * p.ld_aux + gcol;
        float a[NC];
        if (full && (NC % 16 == 0) && ((((size_t)ax) & 31) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 16) load16_bf16(ax + i, a + i);
        } else if (full && ((((size_t)ax) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(ax + i);
                const bf16* h = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i + j] = __bfloat162float(h[j]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i) a[i] = (i < ncol) ? __bfloat162float(ax[i]) : 0.f;
        }
        if (p.dact == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] *= gelu_grad(a[i]);
        } else if (p.dact == ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] *= a[i] > 0.f ? 1.f : 0.1f;
        }
    }
    if (p.out) {
        bf16* o = p.out + (size_t)orow * p.ldc + gcol;
        if (full && (NC % 16 == 0) && ((((size_t)o) & 31) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 16) store16_bf16(o + i, v + i);
        } else if (full && ((((size_t)o) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                *reinterpret_cast<uint4*>(o + i) = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < ncol) o[i] = __float2bfloat16(v[i]);
        }
    }
    if (p.out_f32) {
        float* o = p.out_f32 + (size_t)orow * p.ldc_f32 + gcol;
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < ncol) o[i] = v[i];
    }
}

// Straight-line epilogue for the hottest GELU GEMMs of the eval forward (conv1 of every MLP: folded-BN bias + GELU -> bf16, no
// residual / DropPath / pre-activation copy / row map).  The general epilogue_row above spends ~100 of its ~415 warp
// instructions per 32-column chunk on uniform flag tests and parameter reloads (ncu source page of gemm_tc_kernel<256>, 57 %
// issue utilisation with the ready warps queueing): this one keeps only the math, the loads and the stores.
__device__ __forceinline__ bool epilogue_is_bias_gelu(const GemmParams& p) {
    return p.act == ACT_GELU && p.bias && p.bias_mod == 1 && p.out && !p.row_scale && !p.resid && !p.out2 && !p.dact_aux &&
           !p.out_f32 && p.out_map == MAP_IDENT && p.groups == 1 && (p.N % 32) == 0 && (p.ldc % 16) == 0 &&
           ((((size_t)p.out) & 31) == 0) && ((((size_t)p.bias) & 15) == 0);
}
template <int NC>
__device__ __forceinline__ void epilogue_row_bias_gelu(const float* __restrict__ bias, bf16* __restrict__ out, int ldc, int M, int m,
                                                       int col, float* v) {
    if (m >= M) return;
    const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
        const float4 b = __ldg(b4 + i / 4);
        v[i] = gelu_fast(v[i] + b.x);
        v[i + 1] = gelu_fast(v[i + 1] + b.y);
        v[i + 2] = gelu_fast(v[i + 2] + b.z);
        v[i + 3] = gelu_fast(v[i + 3] + b.w);
    }
    bf16* o = out + (size_t)m * ldc + col;
#pragma unroll
    for (int i = 0; i < NC; i += 16) store16_bf16(o + i, v + i);
}

struct WgradParams {
    int P;                    // reduction length: pixel rows (2-D mode) or B*H*W (conv mode)
    int Ma, Nb;               // output rows (= dY channels) and columns (= X channels) per group
    int Ca, Cb;               // channel extents of the dY / X buffers (TMA zero-fills beyond them)
    int groups, a_goff, b_goff;
    int taps;                 // 1 or 9
    int mode, H, W, bw, bh;   // mode 1: X is shifted per tap over an NHWC [B,H,W,ldx] image
    const bf16* dY; int ldy;
    const bf16* X; int ldx;
    float* out; int ldo;      // [groups][taps][Ma][ldo]
    int ksplit;
};

// launchers (gemm_tc.cu / conv_slab.cu).  There is one implementation per operation: no run-time dispatch.
int sunb_launch_gemm_tc(const GemmParams& p, cudaStream_t stream);
int sunb_conv_slab_supported(const GemmParams& p);
int sunb_launch_conv_slab(const GemmParams& p, cudaStream_t stream);
inline int sunb_launch_gemm(const GemmParams& p, cudaStream_t stream) { return sunb_launch_gemm_tc(p, stream); }

// Opt a kernel in to `bytes` (> 48 KB) of dynamic shared memory on the CURRENT device.  The attribute is per device, so the
// bookkeeping is per (kernel, device): a process that drives several GPUs configures each of them (api.cu).
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every step is a chain of 40 (eval forward) to ~230 (training step) dependent kernels on
// one stream; at data-parallel shard sizes each kernel runs for 5-15 us and the launch + prologue latency between two
// kernels is a quarter of the step.  Kernels launched through sunb_launch carry the programmatic-stream-serialization
// attribute and follow one protocol:
//   * set up everything that does not touch global memory (barrier init, TMEM allocation, tensor-map prefetch),
//   * pdl_trigger(): once EVERY CTA of this grid has got here the next kernel of the stream may start launching -- its CTAs
//     become resident as SM resources free up and run their own set-up,
//   * pdl_wait(): returns when the PREVIOUS grid has completed and its writes are visible; no global memory is read or
//     written before it.  Completion is transitive (a grid cannot complete before its own wait has returned), and since
//     all CTAs of a grid are resident (or done) when its trigger fires, a waiting successor never holds a resource an
//     unscheduled predecessor CTA needs.
// In a kernel launched without the attribute both instructions are no-ops.  -DSUNB_NO_PDL builds without the attribute.
// The overlap pays when kernels are short.  The eval engine at throughput batch sizes runs 40 kernels of 50-1000 us each,
// where early-resident successors measured 1-2 % SLOWER, so sunb_encoder_forward suspends the attribute above
// SUNB_PDL_MAX_IMAGES images for the duration of the call (a scheduling hint scoped to the calling thread, not a kernel switch).
// ------------------------------------------------------------------------------------------------
constexpr int SUNB_PDL_MAX_IMAGES = 640;
bool sunb_pdl_allowed();                 // api.cu: false while a large-batch encoder forward is being enqueued on this thread
void sunb_pdl_allow(bool on);
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t sunb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
#ifndef SUNB_NO_PDL
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (sunb_pdl_allowed()) {
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
#endif
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

int sunb_opt_in_smem(const void* kernel, int bytes);
int sunb_num_sms();      // SM count of the current device
