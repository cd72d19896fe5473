// Weight-gradient GEMM on tcgen05:  dW[tap][m, n] += sum_p dY[p, m] * X_tap[p, n]
// Both operands are activations stored pixel-major (NHWC rows), i.e. "MN-major" from the tensor core's point of
// view: a K block is 64 pixel rows of 128 bytes (64 channels), loaded by TMA with the 128-byte swizzle, and the UMMA
// shared-memory descriptors carry the MN-major canonical layout (LBO = stride between 64-channel atoms,
// SBO = stride between 8-row groups).  3x3 convolutions use the same shifted 4-D NHWC boxes (zero-filled padding) as
// the forward kernel, so no transposed copy of any activation or gradient is ever written.
// The reduction over pixels is split across CTAs (split-K); partial tiles are accumulated into the fp32 gradient
// buffer with red.global.add.f32, so the buffer must be zeroed (or hold the running gradient) beforehand.
// Reference semantics: autograd's conv2d / linear weight gradient for the layers in test_phase/models/visformer.py.
#include "common.cuh"

#include <mutex>
#include <stdlib.h>


namespace {

constexpr int BM = 128, BKP = 64;     // 64 pixel rows per K block
constexpr int NUM_THREADS = 192;
constexpr int ATOM_BYTES = BKP * 128; // one [64 pixels x 64 channels] box

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of a converged warp (see gemm_tc.cu: keeps TMA / MMA issue in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends in hardware; the watchdog (a broken pipeline traps instead of hanging the GPU) is only
    // consulted every 4096 failed polls so the spin loop stays two instructions long
    if (mbar_try_wait(bar, parity)) return;     // fast path: no clock read (CS2R costs ~100+ cycles on the issuer's serial path)
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 4096; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
            printf("sunb wgrad_tc: mbarrier timeout block (%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// MN-major, 128-byte swizzle: 64-channel atoms ATOM_BYTES apart (LBO), 8-pixel-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(ATOM_BYTES >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D fp32, A/B bf16, both MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
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

template <int BN>
struct WLayout {
    static constexpr int A_BYTES = 2 * ATOM_BYTES;               // 128 dY channels
    static constexpr int B_BYTES = (BN / 64) * ATOM_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = TILE_BYTES + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const WgradParams p, const int n_tiles) {
    using L = WLayout<BN>;
    constexpr int STAGES = L::STAGES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + L::TILE_BYTES;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + L::TILE_BYTES + 8 * (2 * STAGES + 1));
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x % n_tiles, m_tile = blockIdx.x / n_tiles;
    const int tap = blockIdx.y % p.taps, g = blockIdx.y / p.taps;
    // K blocks of this split.  conv mode: a K block = one (bw x bh) spatial block of 64/(bw*bh) consecutive images
    const int box = p.mode ? p.bw * p.bh : BKP;
    const int nimg = BKP / box;
    const int tiles_x = p.mode ? p.W / p.bw : 1;
    const int spi = p.mode ? tiles_x * (p.H / p.bh) : 1;
    const int kblocks = p.mode ? spi * ((p.P / (p.H * p.W) + nimg - 1) / nimg) : (p.P + BKP - 1) / BKP;
    const int per = (kblocks + p.ksplit - 1) / p.ksplit;
    const int kb0 = blockIdx.z * per;
    const int kb1 = min(kb0 + per, kblocks);
    const int nk = kb1 - kb0;                 // may be <= 0 for trailing splits: CTA-uniform, nothing to add

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start its set-up; nothing above touched global memory
    pdl_wait();         // the previous kernel has completed, its writes are visible

    if (nk > 0) {
        if (warp == 0) {
            {
                const int dy = p.mode ? tap / 3 - 1 : 0, dx = p.mode ? tap % 3 - 1 : 0;
                const int ca = g * p.a_goff + m_tile * BM;
                const int cb = g * p.b_goff + n_tile * BN;
                for (int i = 0; i < nk; ++i) {
                    const int kb = kb0 + i;
                    const int s = i % STAGES;
                    const uint32_t ph = (i / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    const uint32_t a_dst = smem_base + s * L::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + L::A_BYTES;
                    if (elect_one()) {
                    mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
                    if (p.mode == 0) {
#pragma unroll
                        for (int a = 0; a < 2; ++a) tma_load_2d(a_dst + a * ATOM_BYTES, &tmY, full_bar(s), ca + a * 64, kb * BKP);
#pragma unroll
                        for (int a = 0; a < BN / 64; ++a) tma_load_2d(b_dst + a * ATOM_BYTES, &tmX, full_bar(s), cb + a * 64, kb * BKP);
                    } else {
                        const int blk = kb % spi, img0 = (kb / spi) * nimg;
                        const int y0 = (blk / tiles_x) * p.bh, x0 = (blk % tiles_x) * p.bw;
#pragma unroll
                        for (int a = 0; a < 2; ++a)
                            tma_load_4d(a_dst + a * ATOM_BYTES, &tmY, full_bar(s), ca + a * 64, x0, y0, img0);
#pragma unroll
                        for (int a = 0; a < BN / 64; ++a)
                            tma_load_4d(b_dst + a * ATOM_BYTES, &tmX, full_bar(s), cb + a * 64, x0 + dx, y0 + dy, img0);
                    }
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            {
                constexpr uint32_t idesc = make_idesc_mn(BM, BN);
                for (int i = 0; i < nk; ++i) {
                    const int s = i % STAGES;
                    const uint32_t ph = (i / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + s * L::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + L::A_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BKP / 16; ++k)      // 16 pixel rows = 2048 bytes per UMMA_K step
                            umma_bf16(tmem_base, make_mn_sw128_desc(a_addr + k * 2048), make_mn_sw128_desc(b_addr + k * 2048),
                                      idesc, (i | k) ? 1u : 0u);
                        umma_commit(empty_bar(s));
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(accum_bar);
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            const int row = m_tile * BM + q * 32 + lane;
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            float* orow = p.out + ((size_t)(g * p.taps + tap) * p.Ma + row) * p.ldo;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n_tile * BN + c * 32;
                if (col0 >= p.Nb) break;
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                if (row < p.Ma) {
                    float* dst = orow + col0;
                    if (col0 + 32 <= p.Nb && ((((size_t)dst) & 15) == 0)) {      // 8 x red.global.add.v4.f32 instead of 32 scalar REDs
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.Nb) atomicAdd(dst + i, v[i]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

int make_map(CUtensorMap* map, const bf16* base, int C, int ld, const WgradParams& p, bool conv) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        sunb_set_error("cuTensorMapEncodeTiled entry point unavailable");
        return SUNB_ERR_DRIVER;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (!conv) {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)p.P};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
        cuuint32_t box[2] = {64, BKP};
        r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const int B = p.P / (p.H * p.W);
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * p.W, (cuuint64_t)ld * 2 * p.W * p.H};
        cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)(BKP / (p.bw * p.bh))};
        r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        sunb_set_error("wgrad: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return SUNB_ERR_DRIVER;
    }
    return SUNB_OK;
}

template <int BN>
int launch_w(const WgradParams& p, const CUtensorMap& tmY, const CUtensorMap& tmX, cudaStream_t stream) {
    using L = WLayout<BN>;
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&wgrad_tc_kernel<BN>), L::TOTAL));
    const int n_tiles = (p.Nb + BN - 1) / BN, m_tiles = (p.Ma + BM - 1) / BM;
    dim3 grid(n_tiles * m_tiles, p.taps * p.groups, p.ksplit);
    SUNB_CHECK_CUDA(sunb_launch(&wgrad_tc_kernel<BN>, grid, dim3(NUM_THREADS), L::TOTAL, stream, tmY, tmX, p, n_tiles));
    return SUNB_OK;
}

}  // namespace

int sunb_launch_wgrad_tc(WgradParams p, cudaStream_t stream) {
    SUNB_REQUIRE(p.P > 0 && p.Ma > 0 && p.Nb > 0 && p.taps >= 1 && p.groups >= 1, "wgrad: empty problem");
    SUNB_REQUIRE((p.ldy % 8) == 0 && (p.ldx % 8) == 0 && (((size_t)p.dY) & 15) == 0 && (((size_t)p.X) & 15) == 0,
                 "wgrad: operands must be 16-byte aligned with row strides that are multiples of 8 elements");
    const bool conv = p.mode == 1;
    if (conv) {
        SUNB_REQUIRE(p.taps == 9 && p.bw > 0 && p.bh > 0 && BKP % (p.bw * p.bh) == 0 && (p.bw * p.bh) % 8 == 0 &&
                         p.W % p.bw == 0 && p.H % p.bh == 0 && p.P % (p.H * p.W) == 0,
                     "wgrad: bad conv geometry");
    } else {
        SUNB_REQUIRE(p.taps == 1, "wgrad: taps must be 1 in 2-D mode");
    }
    const int BN = p.Nb <= 64 ? 64 : (p.Nb <= 128 ? 128 : 256);
    const int n_tiles = (p.Nb + BN - 1) / BN, m_tiles = (p.Ma + BM - 1) / BM;
    const int box = conv ? p.bw * p.bh : BKP;
    const int kblocks = conv ? (p.W / p.bw) * (p.H / p.bh) * ((p.P / (p.H * p.W) + BKP / box - 1) / (BKP / box))
                             : (p.P + BKP - 1) / BKP;
    if (p.ksplit <= 0) {       // ~1.5 waves of CTAs (fewer split-K atomics than 2 waves), at least 4 K blocks per CTA
        const int tiles = n_tiles * m_tiles * p.taps * p.groups;
        constexpr int waves_x2 = 3;                      // CTA waves x 2.  Measured on the train step: 2 -> 11.4 ms, 3 -> 11.15, 4 -> 11.4, 6 -> 11.5
        int ks = (waves_x2 * 74 + tiles - 1) / tiles;
        ks = ks < 1 ? 1 : ks;
        const int max_ks = kblocks / 4 > 0 ? kblocks / 4 : 1;
        p.ksplit = ks < max_ks ? ks : max_ks;
    }
    CUtensorMap tmY, tmX;
    SUNB_TRY(make_map(&tmY, p.dY, p.Ca, p.ldy, p, conv));
    SUNB_TRY(make_map(&tmX, p.X, p.Cb, p.ldx, p, conv));
    switch (BN) {
        case 64: return launch_w<64>(p, tmY, tmX, stream);
        case 128: return launch_w<128>(p, tmY, tmX, stream);
        default: return launch_w<256>(p, tmY, tmX, stream);
    }
}
