// Stem entry on tensor cores (reference: test_phase/models/visformer.py:209-210,216,221-223,232).
//   conv1 3x3 s2 (3->64) [+ folded BN + LeakyReLU]  -> a1  [B*1600, 64]  bf16 NHWC
//   downsample 3x3 s2 (3->128) [+ folded BN]        -> idn [B*1600, 128] bf16 NHWC
// Both convolutions read the same 27-value input patch, so they are one GEMM: M = B*1600 pixels, N = 192, K = 27 (padded
// to 32).  The fp32 NCHW image cannot be fetched by TMA into a K-major tile, so every CTA gathers its 128-pixel im2col
// tile with ordinary loads, writes it to shared memory in the 128-byte-swizzled K-major layout the UMMA descriptor expects
// (fence.proxy.async makes it visible to the tensor core), issues two tcgen05.mma (K = 32) into TMEM and runs the
// bias / LeakyReLU / bf16-store epilogue from tcgen05.ld.  The weight tile [192][32] is built once per CTA.
// Persistent warp-specialised CTA per SM: gather warps, an MMA warp and drain warps work on consecutive tiles concurrently.
#include "common.cuh"

namespace {

constexpr int IMG = 80, OUTP = 40, NPIX = OUTP * OUTP, NOUT = 192, KREAL = 27;
constexpr int GATHER_THREADS = 256, MMA_WARP = GATHER_THREADS / 32, THREADS = GATHER_THREADS + 32 + 256;   // 8 gather warps, MMA warp, 8 epilogue warps
constexpr int A_STAGES = 3;
constexpr int B_BYTES = NOUT * 128, A_BYTES = 128 * 128;
constexpr int SMEM_BYTES = B_BYTES + A_STAGES * A_BYTES + 128 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {       // byte offset of element (row, k) in a SW128 K-major tile
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a broken pipeline traps instead of hanging
    if (mbar_try_wait(bar, parity)) return;     // fast path: no clock read (CS2R costs ~100+ cycles on the issuer's serial path)
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 4096; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("sunb stem_in_tc: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Persistent, warp-specialised: warps 0-7 gather im2col tiles into a 3-deep ring, warp 8 issues the MMAs into
// double-buffered TMEM (2 x 192 columns), warps 9-16 drain (bias / LeakyReLU / bf16 stores).  The three phases of
// consecutive tiles overlap; the first version ran them back to back in one CTA and was limited to two CTAs per SM by its
// 256-column TMEM allocation.
__global__ void __launch_bounds__(THREADS, 1) stem_in_tc_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                                const float* __restrict__ b1, const float* __restrict__ wd,
                                                                const float* __restrict__ bd, bf16* __restrict__ a1,
                                                                bf16* __restrict__ idn, int B, int lrelu, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sB = gen;                                  // [192][64] bf16, SW128
    const uint32_t bars = base + B_BYTES + A_STAGES * A_BYTES;
    auto a_full = [&](int s) { return bars + 8u * s; };
    auto a_empty = [&](int s) { return bars + 8u * (A_STAGES + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * A_STAGES + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * A_STAGES + 2 + a); };
    const uint32_t tmem_slot_addr = bars + 8u * (2 * A_STAGES + 4);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + B_BYTES + A_STAGES * A_BYTES + 8 * (2 * A_STAGES + 4));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < A_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_full(s)), "r"(GATHER_THREADS) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a_empty(s)) : "memory");
        }
        for (int a = 0; a < 2; ++a) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(acc_full(a)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(acc_empty(a)) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot_addr) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();                   // the master weights below may have been updated by an earlier kernel of the stream
    // weight tile: rows 0..63 conv1, 64..191 downsample; k >= 27 zero (k in [32,64) is never read: K = 32)
    for (int i = tid; i < NOUT * 32; i += THREADS) {
        const int n = i >> 5, k = i & 31;
        float v = 0.f;
        if (k < KREAL) v = n < 64 ? w1[n * KREAL + k] : wd[(n - 64) * KREAL + k];
        *reinterpret_cast<bf16*>(sB + sw128_off(n, k)) = __float2bfloat16(v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    pdl_trigger();                // TMEM is allocated: the next kernel may start its set-up
    // kind::f16: D fp32, A/B bf16 K-major, M = 128, N = 192
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int total = B * NPIX;

    if (warp < GATHER_THREADS / 32) {
        // ================================================================ gather: element (r, k) = x[img][ci][2*oy-1+ky][2*ox-1+kx],
        // k = (ci*3+ky)*3+kx.  thread = (pixel row r, 16-wide k half): the pixel decode happens once, the 16 values leave as two
        // 16-byte swizzled chunks; lanes walk consecutive pixels so the stride-2 image reads stay within a few sectors.
        const int r = tid & 127, kh = tid >> 7;
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int s = lt % A_STAGES, ph = (lt / A_STAGES) & 1;
            const int m = tile * 128 + r;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
            if (m < total) {
                const int img = m / NPIX, rem = m - img * NPIX, oy = rem / OUTP, ox = rem - oy * OUTP;
                const float* xi = x + (size_t)img * 3 * IMG * IMG;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = kh * 16 + j;
                    if (k < KREAL) {
                        const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
                        const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                        if (iy >= 0 && iy < IMG && ix >= 0 && ix < IMG) v[j] = __ldg(xi + (ci * IMG + iy) * IMG + ix);
                    }
                }
            }
            mbar_wait(a_empty(s), ph ^ 1);               // loads are already in flight / landed: only the smem slot is awaited
            uint8_t* sA = gen + B_BYTES + s * A_BYTES;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[c * 8 + 2 * j], v[c * 8 + 2 * j + 1]);
                *reinterpret_cast<uint4*>(sA + sw128_off(r, kh * 16 + c * 8)) = u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
            mbar_arrive(a_full(s));
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int s = lt % A_STAGES, ph = (lt / A_STAGES) & 1, acc = lt & 1, aph = (lt >> 1) & 1;
            mbar_wait(acc_empty(acc), aph ^ 1);
            mbar_wait(a_full(s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t a_desc = make_sw128_desc(base + B_BYTES + s * A_BYTES), b_desc = make_sw128_desc(base);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem + acc * 256), "l"(a_desc + 2 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"((uint32_t)k) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_empty(s)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(acc_full(acc)) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ================================================================ epilogue: thread = one pixel row, 3 chunks of 32 channels
        const int q = warp & 3, half = (warp - MMA_WARP - 1) >> 2;
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1, aph = (lt >> 1) & 1;
            mbar_wait(acc_full(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = tile * 128 + q * 32 + lane;
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {
                const int c = half * 3 + cc;                 // chunk 0,1 -> conv1 channels; 2..5 -> downsample channels
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + c * 32), v);
                if (cc == 2) {                               // accumulator in registers: release the TMEM buffer
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(acc));
                }
                if (m < total) {
                    const bool is1 = c < 2;
                    const float* bias = is1 ? b1 + c * 32 : bd + (c - 2) * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias + i);
                        v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                    }
                    if (is1 && lrelu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.1f * v[i];
                    }
                    bf16* o = is1 ? a1 + (size_t)m * 64 + c * 32 : idn + (size_t)m * 128 + (c - 2) * 32;
                    store16_bf16(o, v);                      // rows are 128 / 256 B and chunks 64 B: always 32-byte aligned
                    store16_bf16(o + 16, v + 16);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace

int sunb_launch_stem_in_tc(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, bf16* a1,
                           bf16* idn, int B, int lrelu, cudaStream_t stream) {
    SUNB_REQUIRE(B > 0, "stem_in: B must be positive");
    SUNB_REQUIRE((((size_t)b1) & 15) == 0 && (((size_t)bd) & 15) == 0, "stem_in: biases must be 16-byte aligned");
    SUNB_REQUIRE((((size_t)a1) & 31) == 0 && (((size_t)idn) & 31) == 0, "stem_in: outputs must be 32-byte aligned");
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&stem_in_tc_kernel), SMEM_BYTES));
    const int n_tiles = (B * NPIX + 127) / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < sms ? n_tiles : sms;      // persistent: one CTA per SM
    SUNB_CHECK_CUDA(sunb_launch(&stem_in_tc_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, stream, x, w1, b1, wd, bd, a1, idn, B, lrelu, n_tiles));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, bf16* a1,
                        bf16* idn, int B, int lrelu, cudaStream_t stream) {
    return sunb_launch_stem_in_tc(x, w1, b1, wd, bd, a1, idn, B, lrelu, stream);
}
