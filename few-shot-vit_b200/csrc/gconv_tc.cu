// Grouped 3x3 convolution (8 groups x 32 channels, pad 1, 20x20 maps) of the stage-1 conv-MLP on the tcgen05 tensor cores
// (reference: Mlp.conv2, test_phase/models/visformer.py:146-148,157-159), forward and data-gradient.
//
// Formulation.  An image is laid out in shared memory as a zero-haloed 22 x 22 raster, one 16-byte (8-channel) chunk per
// pixel and K-chunk plane:  slab[plane c][haloed pixel p][8 channels].  That is exactly the NO-SWIZZLE K-major UMMA operand
// layout (core matrix = 8 consecutive rows of 16 B; SBO = 128 B between 8-row groups, LBO = plane pitch between K
// chunks), and in that layout a row shift is just +16 B on the descriptor start address.  With the output indexed on the
// same 22-wide raster (o = y*22 + x), the input row of filter tap (dy, dx) is o + dy*22 + dx for EVERY output row, so each
// tap is one MMA over the same resident slab with a shifted start address: the image is fetched from L2 once, not nine
// times, and the 32x32 block-diagonal weights of a group are a native N = 32 MMA (no zero blocks multiplied).
// Outputs that land on halo columns / rows (o % 22 >= 20 or o >= 438; 22 % of the 4 x 128 rows) are computed and dropped.
//
// Persistent, warp-specialised CTA (one per SM): a CTA owns one PAIR of groups (weights resident, 36 KB) and walks the
// images;  4 loader warps fill a 3-deep slab ring with cp.async (coalesced 128 B per pixel), one thread issues the
// 144 MMAs of an (image, group pair) item into double-buffered TMEM (2 x 256 columns), 8 epilogue warps drain it
// (GELU / pre-activation copy / chain-rule factor) with one 64-byte row segment per thread.
#include "common.cuh"

#include <cuda.h>

namespace {

constexpr int HW = 20, HP = 22, NPIX = 400, GC = 32;
constexpr int PLANE_ROWS = HP * HP;                 // 484 haloed pixels
constexpr int PLANE_BYTES = PLANE_ROWS * 16;        // 7744 (== 4 mod 8 in 16-B units: conflict-free 4-pixel x 8-chunk fills)
constexpr int GROUP_BYTES = 4 * PLANE_BYTES;        // 32 channels = 4 K chunks
constexpr int SLAB_BYTES = 2 * GROUP_BYTES;         // a pair of groups: 61,952 B
constexpr int STAGES = 3;
constexpr int W_TAP_BYTES = GC * GC * 2;            // 2 KB: [4 k chunks][32 n][8 k]
constexpr int W_BYTES = 2 * 9 * W_TAP_BYTES;        // 36,864 B
constexpr int LAST_ROW = (HW - 1) * HP + HW - 1;    // 437: last valid output raster position
constexpr int M_TILES = 4;                          // 4 x 128 raster rows cover 0..437
#ifndef SUNB_GCONV_EPI_WARPS
#define SUNB_GCONV_EPI_WARPS 8
#endif
constexpr int LOAD_WARPS = 4, EPI_WARPS = SUNB_GCONV_EPI_WARPS;   // 8 (16 = split tiles by parity: measured slower, 80-register cap spills)
constexpr int TSTEP = EPI_WARPS / 8;
constexpr int THREADS = 32 * (1 + LOAD_WARPS + EPI_WARPS);
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * SLAB_BYTES + W_BYTES + BAR_BYTES + 128;   // 223,104 B
constexpr int TMEM_COLS = 512;                      // 2 buffers x 2 groups x 4 tiles x 32 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// bounded wait: a broken pipeline traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;     // fast path: no clock read (CS2R costs ~100+ cycles on the issuer's serial path)
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 4096; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("sunb gconv_tc: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; `lbo` = bytes between K chunks, `sbo` = bytes between
// 8-row groups.  The start address (16-byte units) is added per MMA.
__device__ __forceinline__ uint64_t make_noswz_desc(uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell); swizzle field (61-63) = 0
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {   // D fp32, A/B bf16, both K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// 32 consecutive bf16 of one output row <-> floats: 256-bit accesses when the segment is 32-byte aligned, else 128-bit
__device__ __forceinline__ void store32_bf16(bf16* p, const float* v) {
    if ((((size_t)p) & 31) == 0) {
        store16_bf16(p, v);
        store16_bf16(p + 16, v + 16);
    } else {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
            *reinterpret_cast<uint4*>(p + i) = u;
        }
    }
}
__device__ __forceinline__ void load32_bf16(const bf16* p, float* v) {
    if ((((size_t)p) & 31) == 0) {
        load16_bf16(p, v);
        load16_bf16(p + 16, v + 16);
    } else {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(p + i);
            const bf16* h = reinterpret_cast<const bf16*>(&u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i + j] = __bfloat162float(h[j]);
        }
    }
}

// x, y, y2, aux: bf16 [B*400, ld] NHWC rows; wg: bf16 [8 groups][9 taps][32 n][32 k] (sunb_gconv_pack)
__global__ void __launch_bounds__(THREADS, 1)
gconv3x3_tc_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ wg, bf16* __restrict__ y, int ldy,
                   bf16* __restrict__ y2, int ldy2, const bf16* __restrict__ aux, int ldaux, int B, int act, int dact) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t slab0 = base;
    const uint32_t wsm = base + STAGES * SLAB_BYTES;
    const uint32_t bars = wsm + W_BYTES;
    // barrier map (8 bytes each): full[3] | empty[3] | acc_full[buf][g2] (4) | acc_empty[buf][g2] (4) | tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (3 + s); };
    auto acc_full = [&](int buf, int g2) { return bars + 8u * (6 + buf * 2 + g2); };
    auto acc_empty = [&](int buf, int g2) { return bars + 8u * (10 + buf * 2 + g2); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * SLAB_BYTES + W_BYTES + 8 * 14);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gp = blockIdx.x & 3;                       // group pair
    const int img0 = blockIdx.x >> 2, img_step = gridDim.x >> 2;
    const int n_items = img0 < B ? (B - img0 + img_step - 1) / img_step : 0;

    // ---- one-time setup: zero the slab ring (the halo rows are never written again), stage the pair's weights
    {
        uint4* z = reinterpret_cast<uint4*>(base_ptr);
        for (int i = tid; i < STAGES * SLAB_BYTES / 16; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
        pdl_wait();               // the packed weights (and everything after) come from earlier kernels of the stream
        uint8_t* wdst = base_ptr + STAGES * SLAB_BYTES;
        for (int i = tid; i < 2 * 9 * GC * 4; i += THREADS) {
            const int c = i & 3, n = (i >> 2) & 31, gt = i >> 7;          // gt = g2 * 9 + tap
            const int g2 = gt / 9, tap = gt - g2 * 9;
            const uint4 u = *reinterpret_cast<const uint4*>(wg + ((size_t)((gp * 2 + g2) * 9 + tap) * GC + n) * GC + c * 8);
            *reinterpret_cast<uint4*>(wdst + (gt * 4 + c) * 512 + n * 16) = u;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), LOAD_WARPS * 32);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b)
            for (int g2 = 0; g2 < 2; ++g2) {
                mbar_init(acc_full(b, g2), 1);
                mbar_init(acc_empty(b, g2), EPI_WARPS / 2);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_proxy();          // generic-proxy zero fill / weight stores -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                // TMEM is allocated: the next kernel may start its set-up

    if (warp == 0) {
        // ================================================================ MMA issuer
        // the whole warp walks the pipeline; one elected lane issues (keeps the MMA sequence in uniform registers)
        {
            constexpr uint32_t idesc = make_idesc(128, GC);
            const uint64_t a_desc0 = make_noswz_desc(PLANE_BYTES, 128);
            const uint64_t b_desc0 = make_noswz_desc(512, 128);
            for (int k = 0; k < n_items; ++k) {
                const int s = k % STAGES, ph = (k / STAGES) & 1, buf = k & 1, bph = (k >> 1) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t slab = slab0 + s * SLAB_BYTES;
#pragma unroll 1
                for (int g2 = 0; g2 < 2; ++g2) {
                    mbar_wait(acc_empty(buf, g2), bph ^ 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int tile = 0; tile < M_TILES; ++tile) {
                        const uint32_t d = tmem_base + buf * 256 + g2 * 128 + tile * GC;
                        const uint32_t a_tile = slab + g2 * GROUP_BYTES + tile * 128 * 16;
                        const uint32_t b_grp = wsm + g2 * 9 * W_TAP_BYTES;
                        if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint32_t a_addr = a_tile + ks * 2 * PLANE_BYTES + ((tap / 3) * HP + tap % 3) * 16;
                                const uint32_t b_addr = b_grp + tap * W_TAP_BYTES + ks * 2 * 512;
                                umma_bf16(d, a_desc0 | (uint64_t)((a_addr >> 4) & 0x3FFF), b_desc0 | (uint64_t)((b_addr >> 4) & 0x3FFF),
                                          idesc, (tap | ks) != 0);
                            }
                        }
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(acc_full(buf, g2));
                    __syncwarp();
                }
                if (elect_one()) umma_commit(empty_bar(s));
                __syncwarp();
            }
        }
    } else if (warp <= LOAD_WARPS) {
        // ================================================================ loaders: cp.async, one item of lag
        const int ltid = tid - 32;
        for (int k = 0; k < n_items; ++k) {
            const int s = k % STAGES, ph = (k / STAGES) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            const int img = img0 + k * img_step;
            const bf16* src = x + (size_t)img * NPIX * ldx + gp * 2 * GC;
            const uint32_t slab = slab0 + s * SLAB_BYTES;
#pragma unroll 5
            for (int i = ltid; i < NPIX * 8; i += LOAD_WARPS * 32) {
                const int p = i >> 3, c = i & 7;
                const int py = p / HW, px = p - py * HW;
                cp_async16(slab + (c >> 2) * GROUP_BYTES + (c & 3) * PLANE_BYTES + ((py + 1) * HP + px + 1) * 16,
                           src + (size_t)p * ldx + c * 8);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (k > 0) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                fence_async_proxy();
                mbar_arrive(full_bar((k - 1) % STAGES));
            }
        }
        if (n_items > 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            fence_async_proxy();
            mbar_arrive(full_bar((n_items - 1) % STAGES));
        }
    } else {
        // ================================================================ epilogue: 2 groups x 4 TMEM lane quadrants
        const int e = warp - 1 - LOAD_WARPS;
        const int g2 = (e >> 2) & 1, q = warp & 3;       // a warp may only touch TMEM lanes 32 * (warp % 4) ...
        const int t0 = e >> 3;                           // first tile of this warp (tiles t0, t0 + TSTEP, ...)
        const int ch = (gp * 2 + g2) * GC;
        // the last tile this quadrant has valid rows in, and the last one THIS warp drains
        const int last_tile = min(M_TILES - 1, (LAST_ROW - q * 32) / 128);
        const int my_last = last_tile - ((last_tile - t0) % TSTEP + TSTEP) % TSTEP;
        for (int k = 0; k < n_items; ++k) {
            const int buf = k & 1, bph = (k >> 1) & 1;
            const int img = img0 + k * img_step;
            mbar_wait(acc_full(buf, g2), bph);
            tc_fence_after();
            if (my_last < t0) {                          // no tile for this warp in this quadrant: just hand the buffer back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf, g2));
            }
#pragma unroll 1
            for (int tile = t0; tile <= last_tile; tile += TSTEP) {
                const int o = tile * 128 + q * 32 + lane;
                const int oy = o / HP, ox = o - oy * HP;
                const bool valid = (oy < HW) && (ox < HW);
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + g2 * 128 + tile * GC, v);
                if (tile == my_last) {                   // accumulators are in registers: hand the TMEM buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(buf, g2));
                }
                if (valid) {
                    const size_t row = (size_t)img * NPIX + oy * HW + ox;
                    if (y2) store32_bf16(y2 + row * ldy2 + ch, v);
                    if (act == ACT_GELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
                    }
                    if (aux) {
                        float a[32];
                        load32_bf16(aux + row * ldaux + ch, a);
                        if (dact == ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= gelu_grad(a[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= act_grad(a[i], dact);
                        }
                    }
                    store32_bf16(y + row * ldy + ch, v);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}


// fp32 grouped weight [256][32][3][3] -> bf16 [8][9][32 n][32 k].  transpose_flip = 1 gives the conv-transpose operand
// for the data gradient: n = input channel, k = output channel, taps mirrored.
__global__ void gconv_pack_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int transpose_flip) {
    const int total = 8 * 9 * GC * GC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i % GC, n = (i / GC) % GC, tap = (i / (GC * GC)) % 9, grp = i / (GC * GC * 9);
        const int co = transpose_flip ? k : n, ci = transpose_flip ? n : k, st = transpose_flip ? 8 - tap : tap;
        dst[i] = __float2bfloat16(w[((size_t)(grp * GC + co) * GC + ci) * 9 + st]);
    }
}

}  // namespace

int sunb_launch_gconv_tc(const bf16* x, int ldx, const bf16* wg, bf16* y, int ldy, bf16* y2, int ldy2, const bf16* aux,
                         int ldaux, int B, int act, int dact, cudaStream_t stream) {
    SUNB_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && (((size_t)x) & 15) == 0 && (((size_t)wg) & 15) == 0 && (((size_t)y) & 15) == 0,
                 "gconv3x3: operands must be 16-byte aligned with row strides that are multiples of 8 elements");
    SUNB_REQUIRE(!y2 || (ldy2 % 8 == 0 && (((size_t)y2) & 15) == 0), "gconv3x3: y2 must be 16-byte aligned");
    SUNB_REQUIRE(!aux || (ldaux % 8 == 0 && (((size_t)aux) & 15) == 0), "gconv3x3: aux must be 16-byte aligned");
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&gconv3x3_tc_kernel), SMEM_BYTES));
    const int per_pair = max(1, min(sunb_num_sms() / 4, B));        // CTAs per group pair; 148 SMs = 4 pairs x 37
    SUNB_CHECK_CUDA(sunb_launch(&gconv3x3_tc_kernel, dim3(4 * per_pair), dim3(THREADS), SMEM_BYTES, stream, x, ldx, wg, y, ldy, y2, ldy2,
                                aux, ldaux, B, act, dact));
    return SUNB_OK;
}

extern "C" {

int sunb_gconv3x3(const void* x, int ldx, const void* wg, void* y, int ldy, void* y2, int ldy2, const void* aux, int ldaux,
                  int B, int act, int dact, void* stream) {
    SUNB_REQUIRE(x && wg && y && B > 0, "gconv3x3: bad arguments");
    return sunb_launch_gconv_tc(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(wg), reinterpret_cast<bf16*>(y),
                                ldy, reinterpret_cast<bf16*>(y2), ldy2, reinterpret_cast<const bf16*>(aux), ldaux, B, act, dact,
                                reinterpret_cast<cudaStream_t>(stream));
}

int sunb_gconv_pack(const float* w, void* dst, int transpose_flip, void* stream) {
    SUNB_REQUIRE(w && dst, "gconv_pack: bad arguments");
    gconv_pack_kernel<<<72, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, reinterpret_cast<bf16*>(dst), transpose_flip);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

}  // extern "C"
