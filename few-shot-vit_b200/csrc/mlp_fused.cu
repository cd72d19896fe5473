// Fused MLP of a stage-2 attention block, eval mode (reference: Mlp of test_phase/models/visformer.py:127-163 inside
// Block :259-263, BatchNorm folded into conv1):
//     out = x + conv3( gelu( conv1(x) + b1 ) )        x, out: bf16 [M, 256];  hidden width 1024
// in ONE kernel: the 1024-wide hidden tensor (512 MB written + 512 MB re-read per block at 2500 images) never leaves the SM.
// Per 128-row tile the hidden dimension is walked in 4 chunks of 256:
//     G1(c): acc1 = X (128 x 256, resident, SW128 K-major)  x  W1[c*256 .. +256, :]^T                 (16 MMAs 128x256x16)
//     GELU : acc1 -> registers -> + b1, GELU -> bf16 -> shared memory as [k-chunk of 8][row][16 B] (no-swizzle K-major)
//     G2(c): acc2 += HID(c) (128 x 256)  x  W3[:, c*256 .. +256]^T                                     (16 MMAs 128x256x16)
// issued as G1(0) G1(1) G2(0) G1(2) G2(1) ...: G1(c+1) starts as soon as the GELU warps hold chunk c in registers, and runs
// under their arithmetic; G2(c) follows when they have written the hidden tile.
// TMEM: acc1 256 columns + acc2 256 columns = all 512.  Shared memory: X 64 KB + hidden 64 KB + a 3-stage ring of 32 KB weight
// K-blocks ([256 rows x 64] of W1 or W3) = 224 KB.
// Warps (576 threads, one persistent CTA per SM): 0 TMA producer, 1 MMA issuer, 2-17 GELU + output warps (4 per TMEM lane
// quarter, 64 columns each).
// Measured (2500 images, CUDA events): 279 us vs 363 us for the two gemm_tc launches; an event trace of the first version
// (128-wide chunks, two accumulators: 286 us) showed the MMA issuer busy 3000 of every 3700 cycles per chunk with the MMAs
// retiring at ~70 % of their nominal rate and the GELU stage at 2600 -- the kernel is bound by the tensor core's operand
// fetch, not by the GELU warps; wider chunks (this version) trade the second accumulator for fewer, larger MMAs.
// Stage 3 (C = 512) cannot use this scheme: its output accumulator alone needs all 512 TMEM columns.
#include "tc_common.cuh"
#include "../../include/sunb200.h"

namespace {

using namespace tc;

constexpr int C = 256, HID = 1024, HC = 256, NCHUNK = HID / HC;
constexpr int X_ATOM = 128 * 128;                   // 128 rows x 64 channels
constexpr int X_BYTES = 4 * X_ATOM;                 // 64 KB
constexpr int HID_BYTES = (HC / 8) * 2048;          // 64 KB: [32 k-chunks][128 rows][16 B]
constexpr int W_STAGE = 32768, W_STAGES = 3;        // one K-block [256 rows x 64] of W1 (rows = hidden) or W3 (rows = output)
constexpr int X_OFF = 0, HID_OFF = X_BYTES, W_OFF = HID_OFF + HID_BYTES, BAR_OFF = W_OFF + W_STAGES * W_STAGE;
constexpr int SMEM_BYTES = 1024 + BAR_OFF + 256;
static_assert(SMEM_BYTES <= 232448, "mlp_fused: shared memory budget");
constexpr int EPI_WARPS = 16, THREADS = 64 + 32 * EPI_WARPS;
constexpr int COL_ACC1 = 0, COL_ACC2 = HC;

__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW3, const float* __restrict__ b1, const bf16* __restrict__ resid,
                 bf16* __restrict__ out, int M, int s2d, int oH, int oW) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + BAR_OFF;
    const uint32_t X_FULL = bars, X_EMPTY = bars + 8;
    auto W_FULL = [&](int i) { return bars + 8u * (2 + i); };
    auto W_EMPTY = [&](int i) { return bars + 8u * (5 + i); };
    const uint32_t ACC1_FULL = bars + 8u * 8, ACC1_EMPTY = bars + 8u * 9, HID_FULL = bars + 8u * 10, HID_EMPTY = bars + 8u * 11;
    const uint32_t ACC2_FULL = bars + 8u * 12, ACC2_EMPTY = bars + 8u * 13;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + BAR_OFF + 8 * 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (M + 127) / 128;
    const int n_local = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        mbar_init(X_FULL, 1);
        mbar_init(X_EMPTY, 1);
        for (int i = 0; i < W_STAGES; ++i) { mbar_init(W_FULL(i), 1); mbar_init(W_EMPTY(i), 1); }
        mbar_init(ACC1_FULL, 1);
        mbar_init(ACC1_EMPTY, EPI_WARPS);
        mbar_init(HID_FULL, EPI_WARPS);
        mbar_init(HID_EMPTY, 1);
        mbar_init(ACC2_FULL, 1);
        mbar_init(ACC2_EMPTY, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start its set-up; nothing above touched global memory
    pdl_wait();         // the previous kernel has completed, its writes are visible

    const uint32_t x_sm = base + X_OFF, hid_sm = base + HID_OFF, w_sm = base + W_OFF;

    if (warp == 0) {
        // ================================================================ TMA producer (same stage order as the MMA issuer)
        if (elect_one()) {
            prefetch_tensormap(&tmX);
            prefetch_tensormap(&tmW1);
            prefetch_tensormap(&tmW3);
        }
        __syncwarp();
        uint32_t wit = 0;
        auto load_x = [&](int lt) {                                       // X tile of local tile lt (waits until G1 of lt-1 is done)
            const int tile = blockIdx.x + lt * gridDim.x;
            mbar_wait(X_EMPTY, (lt & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(X_FULL, X_BYTES);
#pragma unroll
                for (int a = 0; a < 4; ++a) tma_load_2d(x_sm + a * X_ATOM, &tmX, X_FULL, a * 64, tile * 128);
            }
            __syncwarp();
        };
        auto load_w = [&](const CUtensorMap* map, int kcol, int row) {
            const int s = wit % W_STAGES;
            mbar_wait(W_EMPTY(s), ((wit / W_STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(W_FULL(s), W_STAGE);
                tma_load_2d(w_sm + s * W_STAGE, map, W_FULL(s), kcol, row);
            }
            __syncwarp();
            ++wit;
        };
        if (n_local > 0) load_x(0);
        for (int lt = 0; lt < n_local; ++lt) {
            for (int step = 0; step <= NCHUNK; ++step) {
                if (step < NCHUNK)
                    for (int kb = 0; kb < 4; ++kb) load_w(&tmW1, kb * 64, step * HC);           // W1 rows [step*256, +256), K block kb
                if (step == NCHUNK && lt + 1 < n_local) load_x(lt + 1);       // next tile's X as soon as the last G1 has read this one
                if (step >= 1)
                    for (int kb = 0; kb < 4; ++kb) load_w(&tmW3, (step - 1) * HC + kb * 64, 0);  // W3 all rows, hidden K block
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        constexpr uint32_t idesc = make_idesc(128, 256);
        uint32_t wit = 0;
        for (int lt = 0; lt < n_local; ++lt) {
            const uint32_t g0 = (uint32_t)lt * NCHUNK;
            mbar_wait(X_FULL, lt & 1);
            for (int step = 0; step <= NCHUNK; ++step) {
                if (step < NCHUNK) {
                    const uint32_t g = g0 + step;
                    mbar_wait(ACC1_EMPTY, (g & 1) ^ 1);                   // the GELU warps hold the previous chunk in registers
                    tc_fence_after();
                    for (int kb = 0; kb < 4; ++kb, ++wit) {
                        const int s = wit % W_STAGES;
                        mbar_wait(W_FULL(s), (wit / W_STAGES) & 1);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_base + COL_ACC1, desc_k_sw128(x_sm + kb * X_ATOM + k * 32),
                                          desc_k_sw128(w_sm + s * W_STAGE + k * 32), idesc, (kb | k) ? 1u : 0u);
                            umma_commit(W_EMPTY(s));
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(ACC1_FULL);
                        if (step == NCHUNK - 1) umma_commit(X_EMPTY);          // X is only read by the G1 chunks
                    }
                    __syncwarp();
                }
                if (step >= 1) {
                    const int c = step - 1;
                    const uint32_t g = g0 + c;
                    mbar_wait(HID_FULL, g & 1);
                    if (c == 0) mbar_wait(ACC2_EMPTY, (lt & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < 4; ++kb, ++wit) {
                        const int s = wit % W_STAGES;
                        mbar_wait(W_FULL(s), (wit / W_STAGES) & 1);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_base + COL_ACC2, desc_k_noswz(hid_sm + 2 * (kb * 4 + k) * 2048, 2048, 128),
                                          desc_k_sw128(w_sm + s * W_STAGE + k * 32), idesc, (c | kb | k) ? 1u : 0u);
                            umma_commit(W_EMPTY(s));
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(HID_EMPTY);
                        if (c == NCHUNK - 1) umma_commit(ACC2_FULL);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================================================================ GELU + output warps: 64 of the 256 columns each
        const int q = warp & 3, part = (warp - 2) >> 2;                 // TMEM lane quarter, column share
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint8_t* hrow = base_ptr + HID_OFF + r * 16;
        for (int lt = 0; lt < n_local; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            const uint32_t g0 = (uint32_t)lt * NCHUNK;
            for (int c = 0; c < NCHUNK; ++c) {
                const uint32_t g = g0 + c;
                mbar_wait(ACC1_FULL, g & 1);
                tc_fence_after();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int col = part * 64 + hh * 32;
                    float v[32];
                    tmem_ld32(tmem_base + lane_sel + COL_ACC1 + col, v);
                    if (hh == 1) {                                       // the whole share is in registers: G1 of the next chunk may run
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(ACC1_EMPTY);
                    }
                    const float4* b4 = reinterpret_cast<const float4*>(b1 + c * HC + col);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bb = __ldg(b4 + i / 4);
                        pk[i / 2] = pack_bf16x2(gelu_fast(v[i] + bb.x), gelu_fast(v[i + 1] + bb.y));
                        pk[i / 2 + 1] = pack_bf16x2(gelu_fast(v[i + 2] + bb.z), gelu_fast(v[i + 3] + bb.w));
                    }
                    if (hh == 0) mbar_wait(HID_EMPTY, (g & 1) ^ 1);      // G2 of the previous chunk has finished reading the buffer
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<uint4*>(hrow + (col / 8 + jj) * 2048) =
                            make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
                }
                fence_async_proxy();
                __syncwarp();
                if (lane == 0) mbar_arrive(HID_FULL);
            }
            // ---- output: out = x + acc2, 64 columns per warp.  The residual row is fetched BEFORE the wait for the last G2: the
            //      global-load latency hides under it instead of sitting between two tiles on these warps' critical path.
            const int m = tile * 128 + r;
            int orow = m;
            if (s2d && m < M) {
                const int hw = oH * oW, img = m / hw, rem = m % hw, y = rem / oW, x = rem % oW;
                orow = ((img * (oH / 2) + y / 2) * (oW / 2) + x / 2) * 4 + (y & 1) * 2 + (x & 1);
            }
            uint32_t rr[4][8];
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ld_global_256(resid + (size_t)m * C + part * 64 + j * 16, rr[j]);
            }
            mbar_wait(ACC2_FULL, lt & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = part * 64 + cc * 32;
                float v[32];
                tmem_ld32(tmem_base + lane_sel + COL_ACC2 + col, v);
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rr[cc * 2 + j / 8][j % 8]);
                        v[2 * j] += __bfloat162float(h.x);
                        v[2 * j + 1] += __bfloat162float(h.y);
                    }
                    store16_bf16(out + (size_t)orow * C + col, v);
                    store16_bf16(out + (size_t)orow * C + col + 16, v + 16);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ACC2_EMPTY);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free(tmem_base, 512);
    }
}

}  // namespace

extern "C" int sunb_mlp_fused(const void* x, const void* w1, const float* b1, const void* w3, void* out, int M, int s2d, int oH,
                              int oW, void* stream) {
    SUNB_REQUIRE(x && w1 && b1 && w3 && out && M > 0, "mlp_fused: bad arguments");
    SUNB_REQUIRE(((((size_t)x) | ((size_t)out)) & 31) == 0 && ((((size_t)w1) | ((size_t)w3) | ((size_t)b1)) & 15) == 0,
                 "mlp_fused: x / out must be 32-byte aligned, weights and bias 16-byte aligned");
    SUNB_REQUIRE(!s2d || (oH > 0 && oW > 0 && oH % 2 == 0 && oW % 2 == 0 && M % (oH * oW) == 0),
                 "mlp_fused: space-to-depth output needs an even oH x oW raster that divides M");
    SUNB_REQUIRE(!s2d || out != x, "mlp_fused: the space-to-depth output cannot alias the input");
    CUtensorMap tmX, tmW1, tmW3;
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, 128};
        SUNB_TRY(sunb_encode_tensor_map(&tmX, x, 2, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)HID};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, 256};
        SUNB_TRY(sunb_encode_tensor_map(&tmW1, w1, 2, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)HID, (cuuint64_t)C};
        cuuint64_t strides[1] = {(cuuint64_t)HID * 2};
        cuuint32_t box[2] = {64, 256};
        SUNB_TRY(sunb_encode_tensor_map(&tmW3, w3, 2, dims, strides, box));
    }
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&mlp_fused_kernel), SMEM_BYTES));
    const int n_tiles = (M + 127) / 128;
    const int sms = sunb_num_sms();
    const int grid = n_tiles < sms ? n_tiles : sms;
    SUNB_CHECK_CUDA(sunb_launch(&mlp_fused_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), tmX,
                                tmW1, tmW3, b1, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out), M, s2d, oH, oW));
    return SUNB_OK;
}
