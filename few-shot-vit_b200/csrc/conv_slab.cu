// Dense 3x3 convolution (pad 1, stride 1) of the stem (conv2 64->128, conv3 128->128 on 40x40 maps; reference
// test_phase/models/visformer.py:226-236) and its data-gradients, as a tcgen05 implicit GEMM over a RESIDENT haloed slab.
//
// Why.  The tap-per-K-block schedule of gemm_tc.cu re-fetches the 128-pixel activation tile once per filter tap (9x) and
// streams the whole 9-tap weight set per 128 output pixels: 576 KB cross L2 -> SMEM per 128x128 output tile, and the
// kernel sits on the chip-wide L2 -> SM fill rate (ncu: ~9 TB/s, tensor pipe 30 %).  Here a CTA owns a band of RB image
// rows of one image (RB * (W+2) <= 256 raster positions = two M=128 accumulators):
//   * ONE 4-D TMA box per 64-channel atom brings the band with its halo -- (64 ch) x (W+2) x (RB+2) rows, out-of-range
//     coordinates zero-filled by the TMA unit -- into shared memory as a K-major SWIZZLE_128B operand whose row index is
//     the raster position r = y * (W+2) + x;
//   * filter tap (dy, dx) is the same slab read at a start address shifted by (dy*(W+2) + dx) * 128 bytes: the swizzle
//     XOR is a function of the absolute shared-memory address, so a row-shifted descriptor still addresses what the TMA
//     wrote (rows are 128 B, the 8-row pattern repeats every 1024 B, slabs are 1024-byte aligned);
//   * each weight block (tap, 64 channels) is fetched once per 256 raster rows and used by both accumulators.
// L2 -> SMEM traffic drops to (slab 84 KB + weights 288 KB, 144 KB per SM in a pair) per 240 useful pixels of conv3: 3x less.
// Raster positions on halo columns / past the band are computed and dropped in the epilogue (useful: 240 of 256 rows).
//
// Roles (608 threads): warp 0 slab producer, warp 1 MMA issuer (+ TMEM owner), warp 2 weight producer, warps 3-18 epilogue
// (chunk parity x 2 accumulator halves x 4 TMEM lane quarters).  Accumulators are double-buffered in TMEM (2 x 2 x BN columns).
// Slab atoms (64 channels each) sit in a ring of three and are released as soon as their nine taps are issued; weight blocks
// stream through a 6-12 deep ring.  The kernel runs as clusters of two CTAs sharing one weight stream through
// tcgen05.mma.cta_group::2 (the single-CTA predecessor of round 1 measured 0.61 vs 0.70 of the bf16 peak and was removed).
#include "common.cuh"

#include <cuda.h>

int sunb_encode_tensor_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                           const cuuint32_t* box);     // gemm_tc.cu

namespace {

#ifndef SUNB_SLAB_EPI_WARPS
#define SUNB_SLAB_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = SUNB_SLAB_EPI_WARPS;          // 8 or 16: (chunk parity) x 2 accumulator halves x 4 TMEM lane quarters
constexpr int CSTEP = EPI_WARPS / 8;
constexpr int THREADS = 32 * (3 + EPI_WARPS);
constexpr int ATOM_SLOTS = 3;          // ring of 64-channel slab atoms (an atom is released as soon as its 9 taps are issued)
constexpr int SMEM_LIMIT = 232448;

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
// bounded wait: a broken pipeline traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;     // fast path: no clock read (CS2R costs ~100+ cycles on the issuer's serial path)
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 4096; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("sunb conv_slab: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
// K-major SWIZZLE_128B operand: 128-byte rows, 8-row groups 1024 B apart; start address in 16-byte units (any row)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
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

constexpr int POOL_PITCH = 80;                  // bytes per staged row: 32 bf16 + 16 B pad (conflict-free 16-byte row stores)
constexpr int POOL_STAGE = 256 * POOL_PITCH;    // one 32-column chunk of a 256-row tile
struct SlabGeom {
    int pool_off;     // byte offset of the pooling stage from the 1024-aligned base (0 = no fused pooling)
    int P;            // raster pitch W + 2
    int RB;           // output image rows per band
    int bands;        // ceil(H / RB)
    int atom_bytes;   // (RB + 2) * P * 128, multiple of 1024
    int b_stages;
    int tiles;        // images * bands
};

// ------------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one TPC) works on two bands with ONE weight stream.  Each CTA loads its own
// band's slab atoms but only HALF of every weight block (BN/2 rows); the leader issues tcgen05.mma.cta_group::2 (M = 256:
// 128 raster rows from each CTA's slab, written to that CTA's TMEM).  Per SM the weight bytes fetched from L2 and the
// weight bytes the tensor core re-reads from shared memory per MMA are halved (12 KB instead of 16 KB of operands per pair
// of MMAs).  Both CTAs run the same schedule in lock-step, so slot / stage indices -- and with them the shared-memory
// offsets inside the descriptors -- are identical in the pair.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> leader's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {     // arrive on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv_slab2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
                  const SlabGeom g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int KC = p.K / 64;
    constexpr uint32_t B_BYTES = (BN / 2) * 128;          // this CTA's half of a (tap, 64-channel) weight block
    const uint32_t bring = base + ATOM_SLOTS * g.atom_bytes;
    const uint32_t bars = bring + g.b_stages * B_BYTES;
    auto atom_full = [&](int s) { return bars + 8u * s; };
    auto atom_empty = [&](int s) { return bars + 8u * (3 + s); };
    auto acc_full = [&](int a) { return bars + 8u * (6 + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (8 + a); };
    auto b_full = [&](int s) { return bars + 8u * (10 + s); };
    auto b_empty = [&](int s) { return bars + 8u * (26 + s); };
    const uint32_t tmem_slot_addr = bars + 8u * 42;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot_addr - smem_u32(smem_raw)));
    constexpr int TMEM_COLS = 4 * BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair0 = blockIdx.x >> 1, pstep = gridDim.x >> 1;
    const int pairs = (g.tiles + 1) >> 1;
    if (tid == 0) {
        for (int s = 0; s < ATOM_SLOTS; ++s) { mbar_init(atom_full(s), 1); mbar_init(atom_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 2 * EPI_WARPS); }     // drain warps of both CTAs
        for (int s = 0; s < g.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start its set-up; nothing above touched global memory
    pdl_wait();         // the previous kernel has completed, its writes are visible

    if (warp == 0) {
        // ================================================================ slab producer (own band; bytes reported to the leader)
        if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        __syncwarp();
        uint32_t ai = 0;
        for (int t = pair0; t < pairs; t += pstep) {
            const int tile = 2 * t + (int)rank;            // tile == g.tiles (odd count): image index past the batch -> zero box
            const int img = tile / g.bands, y0 = (tile % g.bands) * g.RB;
            for (int kc = 0; kc < KC; ++kc, ++ai) {
                const int s = ai % ATOM_SLOTS, ph = (ai / ATOM_SLOTS) & 1;
                mbar_wait(atom_empty(s), ph ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(atom_full(s), 2 * g.atom_bytes);
                    tma2_load_4d(base + s * g.atom_bytes, &tmA, atom_full(s), kc * 64, -1, y0 - 1, img);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ================================================================ weight producer: this CTA's half of every block
        if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        __syncwarp();
        uint32_t it = 0;
        const int nblk = 9 * KC;
        for (int t = pair0; t < pairs; t += pstep) {
            for (int blk = 0; blk < nblk; ++blk, ++it) {
                const int s = it % g.b_stages, ph = (it / g.b_stages) & 1;
                mbar_wait(b_empty(s), ph ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(b_full(s), 2 * B_BYTES);
                    tma2_load_2d(bring + s * B_BYTES, &tmB, b_full(s), (blk / 9) * 64, (blk % 9) * p.N + (int)rank * (BN / 2));
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc(256, BN);
            uint32_t it = 0, ai = 0;
            int lt = 0;
            for (int t = pair0; t < pairs; t += pstep, ++lt) {
                const int acc = lt & 1, aph = (lt >> 1) & 1;
                mbar_wait(acc_empty(acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * 2 * BN;
                for (int kc = 0; kc < KC; ++kc, ++ai) {
                    const int s = ai % ATOM_SLOTS, ph = (ai / ATOM_SLOTS) & 1;
                    mbar_wait(atom_full(s), ph);
                    tc_fence_after();
                    const uint32_t atom = base + s * g.atom_bytes;
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int bs = it % g.b_stages, bph = (it / g.b_stages) & 1;
                        mbar_wait(b_full(bs), bph);
                        tc_fence_after();
                        const uint32_t a_addr = atom + ((tap / 3) * g.P + tap % 3) * 128;
                        const uint32_t b_addr = bring + bs * B_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
#pragma unroll
                                for (int half = 0; half < 2; ++half)
                                    umma2_bf16(d0 + half * BN, make_sw128_desc(a_addr + half * 128 * 128 + k * 32),
                                               make_sw128_desc(b_addr + k * 32), idesc, (kc | tap | k) ? 1u : 0u);
                            }
                            umma2_commit_both(b_empty(bs));
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma2_commit_both(atom_empty(s));
                    __syncwarp();
                }
                if (elect_one()) umma2_commit_both(acc_full(acc));
                __syncwarp();
            }
        }
    } else {
        // ================================================================ epilogue: warps 3..10 of each CTA drain that CTA's TMEM
        const int q = warp & 3;
        const int half = ((warp - 3) >> 2) & 1;
        const int c0 = (warp - 3) >> 3;              // first 32-column chunk of this warp
        const int r = half * 128 + q * 32 + lane;
        const int yy = r / g.P, xx = r - yy * g.P;
        // fused 2x2 max-pool + position table (eval stem tail): the 8 warps that drain the same 32-column chunk exchange the
        // activated tile through a padded shared-memory stage (raster rows r, r+1, r+P, r+P+1 live in different TMEM lane
        // quarters), then 240 of their 256 threads each pool one (pooled pixel, 8 columns) item
        const bool pool = g.pool_off != 0;
        GemmParams pl = p;
        pl.out = nullptr; pl.out_f32 = nullptr; pl.out2 = nullptr;
        uint8_t* stage = smem_raw + (base - smem_u32(smem_raw)) + g.pool_off + c0 * POOL_STAGE;
        const int wtid = ((warp - 3) & 7) * 32 + lane;                  // thread index inside the 8-warp set
        const int pp = wtid >> 2, cc = wtid & 3;                         // pooled pixel of the band, 8-column group
        const int pw = p.W >> 1, py = pp / pw, px = pp - py * pw;
        int lt = 0;
        for (int t = pair0; t < pairs; t += pstep, ++lt) {
            const int acc = lt & 1, aph = (lt >> 1) & 1;
            const int tile = 2 * t + (int)rank;
            const int img = tile / g.bands, y0 = (tile % g.bands) * g.RB;
            const bool valid = (tile < g.tiles) && (yy < g.RB) && (xx < p.W) && (y0 + yy < p.H);
            const int m = valid ? (img * p.H + y0 + yy) * p.W + xx : p.M;
            mbar_wait(acc_full(acc), aph);
            tc_fence_after();
#pragma unroll 1
            for (int c = c0; c < BN / 32; c += CSTEP) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + half * BN + c * 32), v);
                if (c + CSTEP >= BN / 32) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty(acc) & PEER_MASK);      // the leader's barrier
                }
                if (!pool) {
                    epilogue_row<32>(p, 0, m, c * 32, v);
                    continue;
                }
                epilogue_row<32>(pl, 0, m, c * 32, v);                    // bias, residual, activation -- in registers only
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(stage + r * POOL_PITCH);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                        dst[j] = u;
                    }
                }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + c0), "r"(256) : "memory");
                if (tile < g.tiles && py < (g.RB >> 1) && y0 + 2 * py + 1 < p.H) {
                    const int r0 = 2 * py * g.P + 2 * px;
                    float mx[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) mx[e] = -INFINITY;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = r0 + (k >> 1) * g.P + (k & 1);
                        const uint4 u = *reinterpret_cast<const uint4*>(stage + rr * POOL_PITCH + cc * 16);
                        const bf16* hb = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], __bfloat162float(hb[e]));
                    }
                    const int prow = (y0 >> 1) + py;
                    const float* pos = p.pool_pos + (size_t)(prow * pw + px) * p.N + c * 32 + cc * 8;
                    const float4 q0 = *reinterpret_cast<const float4*>(pos), q1 = *reinterpret_cast<const float4*>(pos + 4);
                    uint4 o;
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
                    oh[0] = __floats2bfloat162_rn(mx[0] + q0.x, mx[1] + q0.y);
                    oh[1] = __floats2bfloat162_rn(mx[2] + q0.z, mx[3] + q0.w);
                    oh[2] = __floats2bfloat162_rn(mx[4] + q1.x, mx[5] + q1.y);
                    oh[3] = __floats2bfloat162_rn(mx[6] + q1.z, mx[7] + q1.w);
                    *reinterpret_cast<uint4*>(p.pool_out + ((size_t)(img * (p.H >> 1) + prow) * pw + px) * p.N + c * 32 + cc * 8) = o;
                }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + c0), "r"(256) : "memory");       // stage free for the next chunk
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();        // nobody leaves (or frees TMEM) while the partner may still read its smem / signal its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int BN>
int launch2(const GemmParams& p, const SlabGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB2, int smem, cudaStream_t stream) {
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&conv_slab2_kernel<BN>), smem));
    const int sms = sunb_num_sms();
    const int pairs = (g.tiles + 1) / 2, cl = sms / 2;
    const int grid = 2 * (pairs < cl ? pairs : cl);
    SUNB_CHECK_CUDA(sunb_launch(&conv_slab2_kernel<BN>, dim3(grid), dim3(THREADS), smem, stream, tmA, tmB2, p, g));
    return SUNB_OK;
}

}  // namespace

// 1 when the problem fits the slab kernel (then sunb_launch_conv_slab runs it), 0 to keep the tap-per-K-block GEMM
int sunb_conv_slab_supported(const GemmParams& p) {
    if (p.a_mode != 1 || p.taps != 9 || p.groups != 1 || p.out_map != 0) return p.pool_out ? -1 : 0;
    if (!(p.K == 64 || p.K == 128) || !(p.N == 64 || p.N == 128)) return 0;
    if (p.W + 2 > 128 || p.H < 1 || p.M % (p.H * p.W) != 0) return 0;
    return 1;
}

int sunb_launch_conv_slab(const GemmParams& p, cudaStream_t stream) {
    SUNB_REQUIRE((p.lda % 8) == 0 && (p.ldw % 8) == 0 && (((size_t)p.A) & 15) == 0 && (((size_t)p.Wt) & 15) == 0,
                 "conv_slab: operands must be 16-byte aligned");
    SlabGeom g;
    g.P = p.W + 2;
    g.RB = 0;
    for (int rb = 256 / g.P; rb >= 1; --rb)                     // largest band whose slab is a whole number of 1 KB swizzle groups
        if (((rb + 2) * g.P) % 8 == 0) { g.RB = rb > p.H ? p.H : rb; break; }
    SUNB_REQUIRE(g.RB >= 1, "conv_slab: no band height fits W=%d", p.W);
    int sr = g.RB + 2;
    while ((sr * g.P) % 8) ++sr;                                 // (only when RB was clamped to H)
    g.atom_bytes = sr * g.P * 128;
    g.bands = (p.H + g.RB - 1) / g.RB;
    const int B = p.M / (p.H * p.W);
    g.tiles = B * g.bands;
    const int BN = p.N;
    // the junk rows of the second accumulator read up to 2P+2 rows past the 256-row window: keep that inside the allocation
    const int slab_total = ATOM_SLOTS * g.atom_bytes;
    const bool pair = true;      // cta_group::2 pairs share one weight stream (a lone last tile is padded with an empty partner)
    const int b_block = (pair ? BN / 2 : BN) * 128;
    const int pool_bytes = p.pool_out ? CSTEP * POOL_STAGE : 0;
    if (p.pool_out) {
        SUNB_REQUIRE(p.pool_pos && EPI_WARPS == 16 && (g.RB % 2) == 0 && (p.H % 2) == 0 && (p.W % 2) == 0 && (p.N % 32) == 0 &&
                         (g.RB / 2) * (p.W / 2) <= 64 && (((size_t)p.pool_out) & 15) == 0,
                     "conv_slab: fused pooling needs even band / image sizes and at most 64 pooled pixels per band");
    }
    int b_stages = (SMEM_LIMIT - 1024 - 512 - slab_total - pool_bytes) / b_block;
    if (b_stages > (pair ? 16 : 8)) b_stages = pair ? 16 : 8;
    SUNB_REQUIRE(b_stages >= 2, "conv_slab: slab of %d bytes leaves no room for the weight ring", slab_total);
    g.b_stages = b_stages;
    g.pool_off = p.pool_out ? slab_total + b_stages * b_block + 512 : 0;
    const int smem = 1024 + slab_total + b_stages * b_block + 512 + pool_bytes;
    SUNB_REQUIRE((256 + 2 * g.P + 2) * 128 <= g.atom_bytes + b_stages * b_block, "conv_slab: over-read guard");

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.lda * 2 * p.W, (cuuint64_t)p.lda * 2 * p.W * p.H};
        cuuint32_t box[4] = {64, (cuuint32_t)g.P, (cuuint32_t)sr, 1};
        SUNB_TRY(sunb_encode_tensor_map(&tmA, p.A, 4, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.taps * p.N};
        cuuint64_t strides[1] = {(cuuint64_t)p.ldw * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(pair ? BN / 2 : BN)};
        SUNB_TRY(sunb_encode_tensor_map(&tmB, p.Wt, 2, dims, strides, box));
    }
    return BN == 64 ? launch2<64>(p, g, tmA, tmB, smem, stream) : launch2<128>(p, g, tmA, tmB, smem, stream);
}
