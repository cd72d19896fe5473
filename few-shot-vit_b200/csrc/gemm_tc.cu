// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
// Persistent kernel, one CTA per SM, 128 x BN output tiles.  Warp roles (576 threads):
//   warp 0    : TMA producer  (cp.async.bulk.tensor into a STAGES-deep ring of 128B-swizzled K-major tiles)
//   warp 1    : TMEM allocator + tcgen05.mma issuer (one elect.sync lane; two fp32 accumulators in TMEM, ping-pong)
//   warps 2-17: epilogue (tcgen05.ld -> bias / residual / activation -> bf16 or fp32 global stores), four warps per TMEM
//               lane quarter, overlapped with the MMAs of the next tile
// The K loop runs over taps x 64-wide K blocks: a 1x1 convolution / linear layer is the taps == 1 case; in conv mode a 3x3
// convolution is nine shifted 4-D TMA box loads of the NHWC activation (TMA zero-fills the padding) -- the grouped-pair and
// odd-shape fall-back; the dense stem 3x3 convolutions go to conv_slab.cu, the grouped 3x3 to gconv_tc.cu.
#include "common.cuh"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                      // 64 bf16 = one 128-byte swizzle row
#ifndef SUNB_EPI_WARPS1
#define SUNB_EPI_WARPS1 16
#endif
// 1-CTA kernel: the accumulator drain (tcgen05.ld -> bias/GELU/residual -> store) is latency bound with two warps per SM
// sub-partition; four per TMEM lane quarter measured +7 % on the whole eval forward (8: 11.15 ms, 12: 10.74, 16: 10.39).
// Prefetching the next chunk's TMEM load / residual inside a warp measured slower (register pressure) and was dropped.
constexpr int EPI_WARPS1 = SUNB_EPI_WARPS1;
constexpr int THREADS1 = 64 + 32 * EPI_WARPS1;
constexpr int CSTEP1 = EPI_WARPS1 / 4;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// Bounded wait: a broken pipeline traps (launch error) instead of hanging the GPU.
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
            printf("sunb gemm_tc: mbarrier timeout block (%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// One lane of a converged warp.  Issuing TMA / MMA under elect.sync (instead of `lane == 0`) lets ptxas keep the whole
// issue sequence in uniform registers; a plain divergent branch wraps every UTCHMMA / UBLKCP in a uniformisation loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128-byte swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);     // start address, 16-byte units
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
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

template <int BN>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;     // 16 KB
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);   // ~192 KB ring, one persistent CTA per SM
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;       // 192 KB for every BN
    static constexpr int MAX_RING = 12;                           // barrier slots (weight-stationary mode has up to 10 A stages)
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = TILE_BYTES + BAR_BYTES + 1024;   // + slack for 1024-byte alignment
};

// Persistent kernel: CTA b processes tiles b, b + gridDim.x, ...  (tile id = (m_tile * groups + g) * n_tiles + n_tile).
// The smem ring runs continuously across tiles; two TMEM accumulators let the epilogue of tile i overlap the
// MMAs of tile i + 1.
// (A weight-stationary schedule and a cta_group::2 variant were measured neutral / slower in round 1 and removed.)
template <int BN>
__global__ void __launch_bounds__(THREADS1, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const GemmParams p, const int n_tiles,
                                                                 const int total_tiles) {
    using L = SmemLayout<BN>;
    constexpr int MAXR = L::MAX_RING;
    constexpr int STAGES = L::STAGES;
    constexpr uint32_t TMEM_COLS = 2 * BN;           // two accumulators; 128 / 256 / 512 columns

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + L::TILE_BYTES;    // full[MAXR], empty[MAXR], acc_full[2], acc_empty[2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + L::TILE_BYTES + 8 * (2 * MAXR + 5));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpt = (p.K + BK - 1) / BK;
    const int nk = p.taps * kpt;

    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAXR + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * MAXR + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * MAXR + 2 + a); };
    const uint32_t ring_base = smem_base;               // ring of (A|B) stages
    constexpr uint32_t RING_STRIDE = L::STAGE_BYTES;
    const int tile0 = blockIdx.x, tstep = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full(a), 1);
            mbar_init(acc_empty(a), EPI_WARPS1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp allocates TMEM, base address lands in shared memory
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

    if (warp == 0) {
        // whole warp walks the schedule; one elected lane issues the TMA traffic
        {
            if (elect_one()) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            }
            __syncwarp();
            const int nimg = p.a_mode ? BM / (p.bw * p.bh) : 1;
            const int tiles_x = p.a_mode ? p.W / p.bw : 1;
            const int spi = p.a_mode ? tiles_x * (p.H / p.bh) : 1;
            uint32_t it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstep) {
                const int n0 = (tile % n_tiles) * BN;
                const int g = (tile / n_tiles) % p.groups;
                const int m_tile = tile / (n_tiles * p.groups);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    const uint32_t a_dst = ring_base + s * RING_STRIDE;
                    const uint32_t b_dst = a_dst + L::A_BYTES;
                    const int tap = kb / kpt, kc = kb % kpt;
                    if (elect_one()) {
                        mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
                        if (p.a_mode == 0) {
                            tma_load_2d(a_dst, &tmA, full_bar(s), g * p.a_goff + kc * BK, m_tile * BM);
                        } else {
                            // one box: (64 channels, bw, bh, nimg images) of spatial block `blk`, shifted by the tap
                            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                            const int blk = m_tile % spi, img0 = (m_tile / spi) * nimg;
                            tma_load_4d(a_dst, &tmA, full_bar(s), g * p.a_goff + kc * BK, (blk % tiles_x) * p.bw + dx,
                                        (blk / tiles_x) * p.bh + dy, img0);
                        }
                        tma_load_2d(b_dst, &tmB, full_bar(s), kc * BK, (g * p.taps + tap) * p.N + n0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0, lt = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstep, ++lt) {
                const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
                mbar_wait(acc_empty(acc), aph ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = ring_base + s * RING_STRIDE;
                    const uint64_t a_desc = make_sw128_desc(a_addr);
                    const uint64_t b_desc = make_sw128_desc(a_addr + L::A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)   // advance 32 bytes (2 x 16-byte units) per UMMA_K = 16
                            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        umma_commit(empty_bar(s));            // frees the smem slot when these MMAs retire
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(acc_full(acc));  // accumulator complete
                __syncwarp();
            }
        }
    } else {
        // epilogue warps 2..9: TMEM lanes [32*(warp%4), +32); the two warps of a lane quarter split the columns
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;      // which share of the column chunks (0 .. CSTEP1-1)
        const int r = q * 32 + lane;
        const bool bias_gelu = epilogue_is_bias_gelu(p);      // kernel-uniform: straight-line epilogue (common.cuh)
        const float* const fb_bias = p.bias;
        bf16* const fb_out = p.out;
        const int fb_ldc = p.ldc, fb_M = p.M;
        uint32_t lt = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstep, ++lt) {
            const int n0 = (tile % n_tiles) * BN;
            const int g = (tile / n_tiles) % p.groups;
            const int m_tile = tile / (n_tiles * p.groups);
            const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
            mbar_wait(acc_full(acc), aph);
            tc_fence_after();
            // sub-boxes past the last image map to m >= M and are dropped by the epilogue
            const int mm = p.a_mode ? conv_tile_row_to_pixel(p, m_tile, r) : m_tile * BM + r;
#pragma unroll 1
            for (int c = half; c < BN / 32; c += CSTEP1) {
                if (n0 + c * 32 >= p.N) break;            // warp-uniform
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                if (bias_gelu) epilogue_row_bias_gelu<32>(fb_bias, fb_out, fb_ldc, fb_M, mm, n0 + c * 32, v);
                else epilogue_row<32>(p, g, mm, n0 + c * 32, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        sunb_set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old or no GPU)");
        return SUNB_ERR_DRIVER;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sunb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                       (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return SUNB_ERR_DRIVER;
    }
    return SUNB_OK;
}

template <int BN>
int launch_bn(const GemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<BN>), L::TOTAL));
    const int n_tiles = (p.N + BN - 1) / BN;
    int m_tiles = (p.M + BM - 1) / BM;
    if (p.a_mode) {       // spatial blocks x groups of 128/(bw*bh) images
        const int nimg = BM / (p.bw * p.bh), B = p.M / (p.H * p.W);
        m_tiles = (p.W / p.bw) * (p.H / p.bh) * ((B + nimg - 1) / nimg);
    }
    const long total = (long)n_tiles * m_tiles * p.groups;
    SUNB_REQUIRE(total < (1L << 31), "gemm_tc: too many tiles");
    const int sms = sunb_num_sms();
    const int grid = (int)(total < sms ? total : sms);     // persistent: one CTA per SM
    SUNB_CHECK_CUDA(sunb_launch(&gemm_tc_kernel<BN>, dim3(grid), dim3(THREADS1), L::TOTAL, stream, tmA, tmB, p, n_tiles, (int)total));
    return SUNB_OK;
}

}  // namespace

int sunb_encode_tensor_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                           const cuuint32_t* box) {
    return encode_map(map, base, rank, dims, strides_bytes, box);
}

int sunb_gemm_tc_pick_bn(const GemmParams& p) {
    if (p.N <= 64) return 64;
    if (p.N <= 128) return 128;
    // 256-wide tiles halve the A re-reads.  The tile loop is bound by operand traffic (every operand byte is written to shared
    // memory by TMA and read once by the MMAs), not by the MMAs, so up to 15 % more column padding is a good trade: the
    // stage-2 qkv GEMM (N = 864: 84 % vs 96 % column utilisation) runs 190 -> 150 us with 256-wide tiles
    const long m_tiles = (p.M + BM - 1) / BM;
    const long n256 = (p.N + 255) / 256, n128 = (p.N + 127) / 128;
    const double util256 = (double)p.N / (n256 * 256), util128 = (double)p.N / (n128 * 128);
#ifndef SUNB_BN_SLACK
#define SUNB_BN_SLACK 0.15
#endif
    if (util256 >= util128 - SUNB_BN_SLACK && m_tiles * n256 * p.groups >= 148) return 256;
    return 128;
}

int sunb_launch_gemm_tc(const GemmParams& p, cudaStream_t stream) {
    SUNB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_tc: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
    SUNB_REQUIRE((p.lda % 8) == 0 && (p.ldw % 8) == 0, "gemm_tc: lda/ldw must be multiples of 8 elements (16 B)");
    SUNB_REQUIRE((((size_t)p.A) & 15) == 0 && (((size_t)p.Wt) & 15) == 0, "gemm_tc: operands must be 16-byte aligned");
    if (sunb_conv_slab_supported(p) > 0) return sunb_launch_conv_slab(p, stream);     // dense 3x3: resident haloed slab (conv_slab.cu)
    SUNB_REQUIRE(p.pool_out == nullptr, "gemm_tc: the fused max-pool epilogue exists in the slab convolution only");
    const int BN = sunb_gemm_tc_pick_bn(p);
    CUtensorMap tmA, tmB;
    if (p.a_mode == 0) {
        SUNB_REQUIRE(p.taps == 1, "gemm_tc: taps must be 1 in 2-D mode");
        cuuint64_t dims[2] = {(cuuint64_t)(p.groups > 1 ? p.a_goff * (p.groups - 1) + p.K : p.K), (cuuint64_t)p.M};
        cuuint64_t strides[1] = {(cuuint64_t)p.lda * 2};
        cuuint32_t box[2] = {BK, BM};
        SUNB_TRY(encode_map(&tmA, p.A, 2, dims, strides, box));
    } else {
        SUNB_REQUIRE(p.taps == 9, "gemm_tc: conv mode needs 9 taps");
        SUNB_REQUIRE(p.bw > 0 && p.bh > 0 && 128 % (p.bw * p.bh) == 0 && (p.bw * p.bh) % 8 == 0 && p.W % p.bw == 0 &&
                         p.H % p.bh == 0 && p.M % (p.H * p.W) == 0,
                     "gemm_tc: bad conv geometry H=%d W=%d box %dx%d M=%d", p.H, p.W, p.bw, p.bh, p.M);
        const int B = p.M / (p.H * p.W);
        const int C = p.groups > 1 ? p.a_goff * (p.groups - 1) + p.K : p.K;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.lda * 2 * p.W, (cuuint64_t)p.lda * 2 * p.W * p.H};
        cuuint32_t box[4] = {BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)(BM / (p.bw * p.bh))};
        SUNB_TRY(encode_map(&tmA, p.A, 4, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.groups * p.taps * p.N};
        cuuint64_t strides[1] = {(cuuint64_t)p.ldw * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)BN};
        SUNB_TRY(encode_map(&tmB, p.Wt, 2, dims, strides, box));
    }
    switch (BN) {
        case 64: return launch_bn<64>(p, tmA, tmB, stream);
        case 128: return launch_bn<128>(p, tmA, tmB, stream);
        default: return launch_bn<256>(p, tmA, tmB, stream);
    }
}
