// Fused tail of the stage-1 conv-MLP block (reference: Mlp with spatial_conv, test_phase/models/visformer.py:146-163, and the
// residual of Block.forward :259-263), eval mode:
//     out = x + conv3( gelu( gconv3x3( h1 ) ) )          h1 = gelu(conv1(bn(x)))  (written by the preceding tcgen05 GEMM)
// in ONE kernel: the grouped 3x3 convolution (8 groups x 32 channels), its GELU, the 1x1 convolution 256 -> 128 and the
// residual add.  The 256-channel hidden tensor h2 never leaves the SM: each (M tile, group) chunk goes TMEM -> registers
// (GELU) -> shared memory (bf16, K-major) and is consumed at once as one K = 32 slice of the conv3 accumulation in TMEM.
// HBM traffic of a block drops from 2.74 GB (three kernels, B = 2500) to conv1's 0.77 GB + 1.03 GB here.
// (conv1 cannot join: its 64 KB of weights + the 12-row haloed 256-channel h1 slab (129 KB) + the 213 KB of streamed
//  grouped / conv3 weights do not fit 227 KB of shared memory at any band height that keeps the M = 128 tiles full; DESIGN.md.)
//
// Work item = half an image (10 output rows).  Layout as in gconv_tc.cu, but fed by TMA with the 128-byte swizzle:
//   * h1 arrives as four 64-channel ATOMS, each one 4-D TMA box (64 ch, x = -1..19, 12 rows, 1 image) with hardware zero fill:
//     a K-major SWIZZLE_128B operand whose row index is the raster position p = row * 21 + (x + 1).  Column 0 of every raster
//     row is the (zero) left halo AND the right halo of the row above, so one zero column per row is enough (pitch 21).
//   * filter tap (dy, dx) of group g = the same atom read through a descriptor shifted by (dy*21 + dx) rows of 128 B, plus
//     64 B for the odd group of the atom, plus 32 B per K step: no im2col, the activation is fetched once.
//   * the output raster (10 x 21 = 210 positions, 2 M tiles) carries one junk column per row; junk rows are dropped.
//   * weights stream from L2 per (item, group) as two bulk copies: 9 taps x [32 n][32 k] of the grouped conv (18 KB) and the
//     [128 n][32 k] slice of conv3 (8 KB), both pre-packed (sunb200/packing.py) as no-swizzle K-major operands.
// Warp roles (896 threads, one persistent CTA per SM): 0 activation-atom producer, 1 weight producer, 2 grouped-conv issuer,
// 3 conv3 issuer, 4-19 four teams of GELU / h2 writers (TMEM lane quarters), 20-27 conv3 epilogue (residual add, bf16, identity
// or space-to-depth rows).  TMEM: conv3 accumulators 2 tiles x 128 columns, grouped-conv accumulators 2 pairs x 2 x 32 columns.
#include "tc_common.cuh"
#include "../../include/sunb200.h"

#ifdef SUNB_TAIL_TRACE
// developer build only (tools/build_variants.sh trace): clock64 stamps of CTA 0's third work item, read by tools/tail_trace.py
__device__ long long g_tail_trace[16][12];          // [group 0..7, 8 = epilogue][event]
#define TTRACE(g, ev) do { if (blockIdx.x == 0 && it == 2 && lane == 0) g_tail_trace[g][ev] = clock64(); } while (0)
#else
#define TTRACE(g, ev) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int HW = 20, PITCH = 21, NPIX = 400;
constexpr int BAND = 10;                               // output rows per work item
constexpr int SLAB_ROWS = (BAND + 2) * PITCH;          // 252 raster positions written by TMA
constexpr int ATOM_BYTES = 256 * 128;                  // 64 channels x 256 rows (252 loaded + the zero position 252 + pad)
constexpr int ATOM_TX = SLAB_ROWS * 128;               // bytes one TMA box delivers
// Pipeline structure.  The serial cost of the barrier operations (an mbarrier wait costs ~100 cycles even when it is already
// complete, a tcgen05.commit a few tens) is what bounded the first version of this kernel (one issuer warp doing every wait
// and commit: 650-700 us at B = 2500, 283 us with ALL math, loads and stores compiled out).  Hence:
//   * a work unit is a PAIR = both M tiles of one (item, group): 36 grouped MMAs share one barrier set, the two tiles are
//     issued interleaved (two independent accumulators, one weight operand);
//   * TWO issuer warps: warp 2 issues the grouped-conv MMAs, warp 3 the conv3 K-slices -- they wait on disjoint barriers and
//     feed the same tensor pipe (different accumulators), so neither sits behind the other's waits;
//   * two pair buffers everywhere (grouped accumulators 2 x 64 TMEM columns, h2 slots 2 x 2 x 8 KB), four GELU teams
//     (pair buffer x tile), eight epilogue warps.
constexpr int NA = 3, NW2 = 3, NW3 = 4;                // ring depths: atoms, grouped-conv weights, conv3 slices
constexpr int W2_BYTES = 9 * 2048, W3_BYTES = 4 * 2048;
constexpr int W_GROUP = W2_BYTES + W3_BYTES;           // 26,624 bytes per group in the operand blob
constexpr int H2_BYTES = 4 * 2048;                     // [k chunk of 8][128 rows][16 B]
constexpr int A_OFF = 0, W2_OFF = NA * ATOM_BYTES, W3_OFF = W2_OFF + NW2 * W2_BYTES, H_OFF = W3_OFF + NW3 * W3_BYTES,
              BAR_OFF = H_OFF + 4 * H2_BYTES;
constexpr int SMEM_BYTES = 1024 + BAR_OFF + 512;
static_assert(SMEM_BYTES <= 232448, "convmlp_tc: shared memory budget");
constexpr int THREADS = 32 * 28;                       // 896: 2 producers, 2 issuers, 16 GELU warps, 8 epilogue warps
constexpr int D3_COL = 0, D2_COL = 256;                // TMEM columns
constexpr int VALID_POS = BAND * PITCH;                // 210 output raster positions per item
#ifndef SUNB_TAIL_DBG
#define SUNB_TAIL_DBG 0       // compile-time timing experiments (never set in the product build): 1 no conv3 MMAs, 2 one tap only,
                              // 4 no GELU math, 8 no residual / store work, 16 no activation TMA, 32 no weight copies
#endif

__global__ void __launch_bounds__(THREADS, 1)
convmlp_tail_kernel(const __grid_constant__ CUtensorMap tmH, const uint8_t* __restrict__ wblob, const bf16* __restrict__ resid,
                    bf16* __restrict__ out, int B, int s2d) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + BAR_OFF;
    auto ATOM_FULL = [&](int i) { return bars + 8u * i; };
    auto ATOM_EMPTY = [&](int i) { return bars + 8u * (4 + i); };
    auto W2_FULL = [&](int i) { return bars + 8u * (8 + i); };
    auto W2_EMPTY = [&](int i) { return bars + 8u * (12 + i); };
    auto W3_FULL = [&](int i) { return bars + 8u * (16 + i); };
    auto W3_EMPTY = [&](int i) { return bars + 8u * (20 + i); };
    auto D2_FULL = [&](int i) { return bars + 8u * (24 + i); };
    auto D2_EMPTY = [&](int i) { return bars + 8u * (28 + i); };
    auto H2_FULL = [&](int i) { return bars + 8u * (32 + i); };
    auto H2_EMPTY = [&](int i) { return bars + 8u * (36 + i); };
    auto D3_FULL = [&](int i) { return bars + 8u * (40 + i); };
    auto D3_EMPTY = [&](int i) { return bars + 8u * (42 + i); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + BAR_OFF + 8 * 44);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items_total = 2 * B;                   // (image, half)
    const int n_items = blockIdx.x < n_items_total ? (n_items_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // raster position 252 (the right halo of the last slab row) and the pad rows are never written by TMA: zero them once
    for (int i = threadIdx.x; i < NA * (256 - SLAB_ROWS) * 8; i += THREADS) {
        const int slot = i / ((256 - SLAB_ROWS) * 8), rem = i % ((256 - SLAB_ROWS) * 8);
        *reinterpret_cast<uint4*>(base_ptr + A_OFF + slot * ATOM_BYTES + SLAB_ROWS * 128 + rem * 16) = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(ATOM_FULL(i), 1); mbar_init(ATOM_EMPTY(i), 1); }
        for (int i = 0; i < NW2; ++i) { mbar_init(W2_FULL(i), 1); mbar_init(W2_EMPTY(i), 1); }
        for (int i = 0; i < NW3; ++i) { mbar_init(W3_FULL(i), 1); mbar_init(W3_EMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(D2_FULL(i), 1);
            mbar_init(D2_EMPTY(i), 8);
            mbar_init(H2_FULL(i), 8);
            mbar_init(H2_EMPTY(i), 1);
        }
        mbar_init(D3_FULL(0), 1);
        mbar_init(D3_EMPTY(0), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start its set-up; nothing above touched global memory
    pdl_wait();         // the previous kernel has completed, its writes are visible

    if (warp == 0) {
        // ================================================================ activation atoms: one haloed TMA box per 64 channels
        if (elect_one()) prefetch_tensormap(&tmH);
        __syncwarp();
        for (int it = 0; it < n_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int img = item >> 1, y0 = (item & 1) * BAND;
            for (int a = 0; a < 4; ++a) {
                const int an = it * 4 + a, slot = an % NA;
                mbar_wait(ATOM_EMPTY(slot), ((an / NA) & 1) ^ 1);
                if (elect_one()) {
                    if ((SUNB_TAIL_DBG & 16) && an >= NA) {
                        mbar_arrive(ATOM_FULL(slot));
                    } else {
                        mbar_expect_tx(ATOM_FULL(slot), ATOM_TX);
                        tma_load_4d(base + A_OFF + slot * ATOM_BYTES, &tmH, ATOM_FULL(slot), a * 64, -1, y0 - 1, img);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================================================ weights: two bulk copies per (item, group)
        for (int it = 0; it < n_items; ++it) {
            for (int g = 0; g < 8; ++g) {
                const int wn = it * 8 + g, s2 = wn % NW2, s3 = wn % NW3;
                mbar_wait(W2_EMPTY(s2), ((wn / NW2) & 1) ^ 1);
                if (elect_one()) {
                    if ((SUNB_TAIL_DBG & 32) && wn >= NW2) {
                        mbar_arrive(W2_FULL(s2));
                    } else {
                        mbar_expect_tx(W2_FULL(s2), W2_BYTES);
                        bulk_load(base + W2_OFF + s2 * W2_BYTES, wblob + (size_t)g * W_GROUP, W2_BYTES, W2_FULL(s2));
                    }
                }
                __syncwarp();
                mbar_wait(W3_EMPTY(s3), ((wn / NW3) & 1) ^ 1);
                if (elect_one()) {
                    if ((SUNB_TAIL_DBG & 32) && wn >= NW3) {
                        mbar_arrive(W3_FULL(s3));
                    } else {
                        mbar_expect_tx(W3_FULL(s3), W3_BYTES);
                        bulk_load(base + W3_OFF + s3 * W3_BYTES, wblob + (size_t)g * W_GROUP + W2_BYTES, W3_BYTES, W3_FULL(s3));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ================================================================ issuer 1: grouped 3x3 MMAs, one PAIR (both tiles) at a time
        constexpr uint32_t idesc_g = make_idesc(128, 32);
        for (int it = 0; it < n_items; ++it) {
            for (int g = 0; g < 8; ++g) {
                const int pr = it * 8 + g, pb = pr & 1, k = pr >> 1, s2 = pr % NW2;
                const int an = it * 4 + (g >> 1), as = an % NA;
                TTRACE(g, 0);
                if ((g & 1) == 0) mbar_wait(ATOM_FULL(as), (an / NA) & 1);
                TTRACE(g, 1);
                mbar_wait(W2_FULL(s2), (pr / NW2) & 1);
                mbar_wait(D2_EMPTY(pb), (k & 1) ^ 1);
                tc_fence_after();
                TTRACE(g, 2);
                const uint32_t a_sm = base + A_OFF + as * ATOM_BYTES + (g & 1) * 64;
                const uint32_t b_sm = base + W2_OFF + s2 * W2_BYTES;
                const uint32_t d0 = tmem_base + D2_COL + pb * 64;
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < ((SUNB_TAIL_DBG & 2) ? 1 : 9); ++tap) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint64_t bd = desc_k_noswz(b_sm + tap * 2048 + ks * 2 * 512, 512, 128);
                            const uint32_t ao = a_sm + ((tap / 3) * PITCH + tap % 3) * 128 + ks * 32;
                            umma_bf16(d0, desc_k_sw128(ao), bd, idesc_g, (tap | ks) ? 1u : 0u);
                            umma_bf16(d0 + 32, desc_k_sw128(ao + 128 * 128), bd, idesc_g, (tap | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(D2_FULL(pb));
                    umma_commit(W2_EMPTY(s2));
                    if (g & 1) umma_commit(ATOM_EMPTY(as));
                }
                __syncwarp();
                TTRACE(g, 3);
            }
        }
    } else if (warp == 3) {
        // ================================================================ issuer 2: conv3 K-slices (K = 32 per group) of both tiles
        constexpr uint32_t idesc_c = make_idesc(128, 128);
        for (int it = 0; it < n_items; ++it) {
            for (int g = 0; g < 8; ++g) {
                const int pr = it * 8 + g, pb = pr & 1, k = pr >> 1, s3 = pr % NW3;
                TTRACE(g, 4);
                if (g == 0) mbar_wait(D3_EMPTY(0), (it & 1) ^ 1);
                mbar_wait(W3_FULL(s3), (pr / NW3) & 1);
                mbar_wait(H2_FULL(pb), k & 1);
                tc_fence_after();
                TTRACE(g, 5);
                const uint32_t b_sm = base + W3_OFF + s3 * W3_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint32_t a_sm = base + H_OFF + (pb * 2 + t) * H2_BYTES;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            if (!(SUNB_TAIL_DBG & 1))
                                umma_bf16(tmem_base + D3_COL + t * 128, desc_k_noswz(a_sm + ks * 2 * 2048, 2048, 128),
                                          desc_k_noswz(b_sm + ks * 2 * 2048, 2048, 128), idesc_c, (g | ks) ? 1u : 0u);
                    }
                    umma_commit(H2_EMPTY(pb));
                    umma_commit(W3_EMPTY(s3));
                    if (g == 7) umma_commit(D3_FULL(0));
                }
                __syncwarp();
                TTRACE(g, 6);
            }
        }
    } else if (warp >= 4 && warp < 20) {
        // ================================================================ GELU teams: team = (pair buffer, tile)
        const int team = (warp - 4) >> 2, q = warp & 3;
        const int pb = team >> 1, t = team & 1;
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const int pairs = n_items * 8;
        for (int pr = pb; pr < pairs; pr += 2) {
            const int k = pr >> 1;
#ifdef SUNB_TAIL_TRACE
            const int it = pr >> 3;
            const bool tr = (warp == 4 || warp == 12);
            if (tr) TTRACE(pr & 7, 7);
#endif
            mbar_wait(D2_FULL(pb), k & 1);
            tc_fence_after();
#ifdef SUNB_TAIL_TRACE
            if (tr) TTRACE(pr & 7, 8);
#endif
            float v[32];
            tmem_ld32(tmem_base + lane_sel + D2_COL + pb * 64 + t * 32, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(D2_EMPTY(pb));
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                pk[j] = (SUNB_TAIL_DBG & 4) ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : pack_bf16x2(gelu_fast(v[2 * j]), gelu_fast(v[2 * j + 1]));
            mbar_wait(H2_EMPTY(pb), (k & 1) ^ 1);
            uint8_t* dst = base_ptr + H_OFF + (pb * 2 + t) * H2_BYTES + r * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(dst + c * 2048) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(H2_FULL(pb));
#ifdef SUNB_TAIL_TRACE
            if (tr) TTRACE(pr & 7, 9);
#endif
        }
    } else if (warp >= 20) {
        // ================================================================ conv3 epilogue: + residual, bf16, store.  Warp (q, half)
        // owns columns [half*64, +64) of TMEM lane quarter q, 16 columns at a time; the residual segment of the next step is in
        // flight while the current one is added and stored (its address does not depend on the MMAs).
        const int q = warp & 3, half = (warp - 20) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        for (int it = 0; it < n_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int img = item >> 1, y0 = (item & 1) * BAND;
            int mrow[2], orow[2];
            bool valid[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int o = t * 128 + r;
                const int oy = o / PITCH, ox = o - oy * PITCH;
                valid[t] = (o < VALID_POS) && (ox < HW) && !(SUNB_TAIL_DBG & 8);
                mrow[t] = img * NPIX + (y0 + oy) * HW + ox;
                orow[t] = mrow[t];
                if (s2d) {                              // 2x2 space-to-depth rows for the following PatchEmbed GEMM (MAP_S2D)
                    const int y = y0 + oy;
                    orow[t] = ((img * (HW / 2) + y / 2) * (HW / 2) + ox / 2) * 4 + (y & 1) * 2 + (ox & 1);
                }
            }
            uint32_t rb[8];
            if (valid[0]) ld_global_256(resid + (size_t)mrow[0] * 128 + half * 64, rb);
            if (warp == 20) TTRACE(8, 0);
            mbar_wait(D3_FULL(0), it & 1);
            tc_fence_after();
            if (warp == 20) TTRACE(8, 1);
#pragma unroll
            for (int step = 0; step < 8; ++step) {
                const int t = step >> 2, col = half * 64 + (step & 3) * 16;
                float v[16];
                tmem_ld16(tmem_base + lane_sel + D3_COL + t * 128 + col, v);
                if (step == 7) {                        // accumulators are in registers: hand the conv3 tiles back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(D3_EMPTY(0));
                }
                uint32_t rn[8];
                if (step < 7) {
                    const int tn = (step + 1) >> 2, coln = half * 64 + ((step + 1) & 3) * 16;
                    if (valid[tn]) ld_global_256(resid + (size_t)mrow[tn] * 128 + coln, rn);
                }
                if (valid[t]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rb[j]);
                        v[2 * j] += __bfloat162float(h.x);
                        v[2 * j + 1] += __bfloat162float(h.y);
                    }
                    store16_bf16(out + (size_t)orow[t] * 128 + col, v);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) rb[j] = rn[j];
            }
            if (warp == 20) TTRACE(8, 2);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_free(tmem_base, 512);
    }
}

}  // namespace

#ifdef SUNB_TAIL_TRACE
extern "C" int sunb_tail_trace_read(long long* dst) {
    return cudaMemcpyFromSymbol(dst, g_tail_trace, sizeof(g_tail_trace)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int sunb_convmlp_tail(const void* h1, const void* wblob, const void* resid, void* out, int B, int s2d, void* stream) {
    SUNB_REQUIRE(h1 && wblob && resid && out && B > 0, "convmlp_tail: bad arguments");
    SUNB_REQUIRE(((((size_t)h1) | ((size_t)wblob)) & 15) == 0 && ((((size_t)resid) | ((size_t)out)) & 31) == 0,
                 "convmlp_tail: h1 / weights must be 16-byte aligned, resid / out 32-byte aligned");
    SUNB_REQUIRE(resid != out, "convmlp_tail: the residual stream cannot be updated in place (neighbouring bands read halo rows)");
    CUtensorMap tmH;
    cuuint64_t dims[4] = {256, HW, HW, (cuuint64_t)B};
    cuuint64_t strides[3] = {512, 512 * HW, 512 * NPIX};
    cuuint32_t box[4] = {64, PITCH, BAND + 2, 1};
    SUNB_TRY(sunb_encode_tensor_map(&tmH, h1, 4, dims, strides, box));
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&convmlp_tail_kernel), SMEM_BYTES));
    const int sms = sunb_num_sms();
    const int grid = 2 * B < sms ? 2 * B : sms;
    SUNB_CHECK_CUDA(sunb_launch(&convmlp_tail_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), tmH,
                                reinterpret_cast<const uint8_t*>(wblob), reinterpret_cast<const bf16*>(resid),
                                reinterpret_cast<bf16*>(out), B, s2d));
    return SUNB_OK;
}
