// Backward of the attention core on tcgen05 / TMEM / TMA (forward: attention_tc.cu; reference: autograd through
// test_phase/models/visformer.py:183-190) for the padded head layout (head stride ds = 48 for d = 42, ds = 96 for d = 85).
//   qkv  : bf16 [B*S, ld_qkv] saved forward input, channel (x*heads + y)*ds + z
//   dout : bf16 [B*S, ld_out] gradient of the head-concatenated output, channel y*ds + z
//   dqkv : bf16 [B*S, ld_qkv] same channel order as qkv (pad channels come out as exact zeros)
// Per 128-row tile (S = 100: one (image, head); S = 25: five images of one head, block-diagonal mask) five GEMMs run on
// the tensor core, every operand read in place from the four TMA boxes Q, K, V, dO (64 channels x rows, SWIZZLE_128B):
//   S  = Q K^T      A = Q  (K-major)            B = K  (K-major)        -> TMEM
//   dP = dO V^T     A = dO (K-major)            B = V  (K-major)        -> TMEM
//   dV = P^T dO     A = P  (MN-major, no swz)   B = dO (MN-major)       reduction over the 128 query rows
//   dK = dS^T Q     A = dS (MN-major, no swz)   B = Q  (MN-major)
//   dQ = dS K       A = dS (K-major, no swz)    B = K  (MN-major)       reduction over the padded keys
// P and dS are written once by the softmax threads as [key chunk of 8][query row][16 B]: that block of 8 x 8 core
// matrices is at the same time the K-major layout of the [query x key] matrix and the MN-major layout of its transpose,
// so P^T / dS^T need no second copy.
// Warp roles (192 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 one query row
// (= one key row in the epilogue) per thread: softmax statistics in registers, D = rowsum(P * dP), dS = P (dP - D) * scale.
// Rows / keys outside the tile's valid range get P = dS = 0, so whatever the boxes hold there never reaches a result.
#include "tc_common.cuh"

namespace {

using namespace tc;

template <int S_, int IPT_, int NPAD_, int KA_, int NST_>
struct BCfg {
    static constexpr int S = S_, IPT = IPT_, NPAD = NPAD_, KA = KA_, NST = NST_;
    static constexpr int DP = KA_ == 1 ? 48 : 96;             // padded head width
    static constexpr int ROWS = S_ * IPT_;                    // valid query rows per tile (<= 128)
    static constexpr int Q_ATOM = 128 * 128;                  // 128 rows x 64 channels
    static constexpr int KV_ATOM = NPAD_ * 128;
    static constexpr int Q_OFF = 0, DO_OFF = KA_ * Q_ATOM, K_OFF = 2 * KA_ * Q_ATOM, V_OFF = K_OFF + KA_ * KV_ATOM;
    static constexpr int STAGE = 2 * KA_ * (Q_ATOM + KV_ATOM);
    static constexpr int PBUF = 16 * 2048;                    // [16 key chunks][128 rows][16 B]
    static constexpr int P_OFF = NST_ * STAGE, DS_OFF = P_OFF + PBUF;
    static constexpr int BAR_OFF = DS_OFF + PBUF;
    static constexpr int SMEM = 1024 + BAR_OFF + 256;
    static_assert(SMEM <= 232448, "attention_bwd_tc: shared memory budget");
    static constexpr int THREADS = 192;
    static constexpr int NCHUNK = (NPAD_ + 31) / 32;
    static constexpr int OW = KA_ * 64;                       // TMEM columns of one output accumulator
    static constexpr bool ALIAS = 256 + 3 * OW > 512;         // dQ / dK reuse the S / dP columns when TMEM is short
    static constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256;
    static constexpr int COL_DK = ALIAS ? COL_DP : 256 + OW, COL_DQ = ALIAS ? COL_S : 256 + 2 * OW;
};

// MN-major operand without swizzle: core matrices of 8 K rows x 16 bytes (8 elements of M / N); `lbo` = bytes between
// consecutive 8-row K groups, `sbo` = bytes between consecutive 8-element M / N groups.  A K step of 16 advances by 2 * lbo.
__device__ __forceinline__ uint64_t desc_mn_noswz(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return desc_k_noswz(smem_addr, lbo, sbo);
}

template <typename C>
__global__ void __launch_bounds__(C::THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmDO, bf16* __restrict__ dqkv, int B, int heads, int ld_qkv,
                        float scale) {
    constexpr int S = C::S, IPT = C::IPT, NPAD = C::NPAD, KA = C::KA, DP = C::DP, NST = C::NST;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + C::BAR_OFF;
    auto LOAD_FULL = [&](int i) { return bars + 8u * i; };
    auto LOAD_EMPTY = [&](int i) { return bars + 8u * (2 + i); };
    const uint32_t SDP_FULL = bars + 8u * 4, SDP_EMPTY = bars + 8u * 5, PDS_FULL = bars + 8u * 6, OUT_FULL = bars + 8u * 7,
                   OUT_EMPTY = bars + 8u * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + C::BAR_OFF + 8 * 10);
    static_assert(NST <= 2, "barrier map");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img_tiles = (B + IPT - 1) / IPT;
    const int n_tiles = img_tiles * heads;
    const int n_local = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int inner = heads * DP;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(LOAD_FULL(i), 1); mbar_init(LOAD_EMPTY(i), 1); }
        mbar_init(SDP_FULL, 1);
        mbar_init(SDP_EMPTY, 4);
        mbar_init(PDS_FULL, 4);
        mbar_init(OUT_FULL, 1);
        mbar_init(OUT_EMPTY, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the P / dS buffers start as zeros: a thread only ever writes the key chunks its row attends to (the same ones for every
    // tile), the rest -- keys beyond NPAD, other images' keys with five images per tile -- stay zero for the whole kernel
    for (int i = threadIdx.x; i < 2 * C::PBUF / 16; i += C::THREADS)
        reinterpret_cast<uint4*>(base_ptr + C::P_OFF)[i] = make_uint4(0, 0, 0, 0);
    fence_async_proxy();
    if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start its set-up; nothing above touched global memory
    pdl_wait();         // the previous kernel has completed, its writes are visible

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one()) {
            prefetch_tensormap(&tmQ);
            prefetch_tensormap(&tmKV);
            prefetch_tensormap(&tmDO);
        }
        __syncwarp();
        for (int i = 0; i < n_local; ++i) {
            const int t = blockIdx.x + i * gridDim.x;
            const int head = t % heads, row0 = (t / heads) * IPT * S;
            const int st = i % NST;
            mbar_wait(LOAD_EMPTY(st), ((i / NST) & 1) ^ 1);
            if (elect_one()) {
                const uint32_t sb = base + st * C::STAGE;
                mbar_expect_tx(LOAD_FULL(st), C::STAGE);
#pragma unroll
                for (int a = 0; a < KA; ++a) {
                    tma_load_2d(sb + C::Q_OFF + a * C::Q_ATOM, &tmQ, LOAD_FULL(st), head * DP + a * 64, row0);
                    tma_load_2d(sb + C::K_OFF + a * C::KV_ATOM, &tmKV, LOAD_FULL(st), inner + head * DP + a * 64, row0);
                    tma_load_2d(sb + C::DO_OFF + a * C::Q_ATOM, &tmDO, LOAD_FULL(st), head * DP + a * 64, row0);
                    tma_load_2d(sb + C::V_OFF + a * C::KV_ATOM, &tmKV, LOAD_FULL(st), 2 * inner + head * DP + a * 64, row0);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        constexpr uint32_t idesc_qk = make_idesc(128, NPAD);                            // S, dP
        constexpr uint32_t idesc_t = make_idesc(128, C::OW, true, true);                // dV, dK: A^T from P / dS, B MN-major
        constexpr uint32_t idesc_dq = make_idesc(128, C::OW, false, true);              // dQ
        const uint32_t p_sm = base + C::P_OFF, ds_sm = base + C::DS_OFF;
        for (int i = 0; i < n_local; ++i) {
            const int st = i % NST, ph = i & 1;
            const uint32_t sb = base + st * C::STAGE;
            const uint32_t q_sm = sb + C::Q_OFF, do_sm = sb + C::DO_OFF, k_sm = sb + C::K_OFF, v_sm = sb + C::V_OFF;
            mbar_wait(LOAD_FULL(st), (i / NST) & 1);
            mbar_wait(SDP_EMPTY, ph ^ 1);
            if (C::ALIAS) mbar_wait(OUT_EMPTY, ph ^ 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int a = 0; a < KA; ++a) {
                    const int ksteps = (DP - a * 64) >= 64 ? 4 : (DP - a * 64) / 16;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < ksteps)
                            umma_bf16(tmem_base + C::COL_S, desc_k_sw128(q_sm + a * C::Q_ATOM + k * 32),
                                      desc_k_sw128(k_sm + a * C::KV_ATOM + k * 32), idesc_qk, (a | k) ? 1u : 0u);
                }
#pragma unroll
                for (int a = 0; a < KA; ++a) {
                    const int ksteps = (DP - a * 64) >= 64 ? 4 : (DP - a * 64) / 16;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < ksteps)
                            umma_bf16(tmem_base + C::COL_DP, desc_k_sw128(do_sm + a * C::Q_ATOM + k * 32),
                                      desc_k_sw128(v_sm + a * C::KV_ATOM + k * 32), idesc_qk, (a | k) ? 1u : 0u);
                }
                umma_commit(SDP_FULL);
            }
            __syncwarp();
            mbar_wait(PDS_FULL, ph);
            if (!C::ALIAS) mbar_wait(OUT_EMPTY, ph ^ 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k)                                             // dV = P^T dO over 128 query rows
                    umma_bf16(tmem_base + C::COL_DV, desc_mn_noswz(p_sm + k * 256, 128, 2048),
                              desc_mn_sw128(do_sm + k * 2048, C::Q_ATOM), idesc_t, k ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)                                             // dK = dS^T Q
                    umma_bf16(tmem_base + C::COL_DK, desc_mn_noswz(ds_sm + k * 256, 128, 2048),
                              desc_mn_sw128(q_sm + k * 2048, C::Q_ATOM), idesc_t, k ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < NPAD / 16; ++k)                                     // dQ = dS K over the padded keys
                    umma_bf16(tmem_base + C::COL_DQ, desc_k_noswz(ds_sm + 2 * k * 2048, 2048, 128),
                              desc_mn_sw128(k_sm + k * 2048, C::KV_ATOM), idesc_dq, k ? 1u : 0u);
                umma_commit(OUT_FULL);
                umma_commit(LOAD_EMPTY(st));
            }
            __syncwarp();
        }
    } else {
        // ================================================================ softmax / dS / epilogue: one row per thread
        const int q = warp & 3;                                        // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const int blk = r < C::ROWS ? r / S : 0;
        const int lo = blk * S, hi = lo + S;                           // valid key columns of this row
        const int wlo = min(q * 32, C::ROWS - 1) / S * S, whi = min(q * 32 + 31, C::ROWS - 1) / S * S + S;
        const int c_lo = IPT == 1 ? 0 : wlo / 32, c_hi = IPT == 1 ? (S - 1) / 32 : (whi - 1) / 32;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const float scale_log2e = scale * 1.4426950408889634f;
        constexpr int NL = (IPT == 1) ? C::NCHUNK : 3;
        uint8_t* prow = base_ptr + C::P_OFF + r * 16;
        uint8_t* dsrow = base_ptr + C::DS_OFF + r * 16;
        for (int i = 0; i < n_local; ++i) {
            const int t = blockIdx.x + i * gridDim.x;
            const int head = t % heads, img0 = (t / heads) * IPT;
            const int rows_valid = min(IPT, B - img0) * S;
            const bool row_ok = r < rows_valid;
            const int ph = i & 1;
            mbar_wait(SDP_FULL, ph);
            tc_fence_after();
            // ---- pass 1: score row -> e = exp2((s - max) * scale * log2 e), sum
            float e[NL][32];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;                                // warp-uniform
                if (IPT == 1 && NPAD - k * 32 < 32) tmem_ld16(tmem_base + lane_sel + C::COL_S + c * 32, e[k]);
                else tmem_ld32(tmem_base + lane_sel + C::COL_S + c * 32, e[k]);
            }
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    if (IPT == 1 && k * 32 + jj >= NPAD) continue;
                    const int col = c * 32 + jj;
                    const bool ok = (IPT == 1) ? (col < S) : (col >= lo && col < hi);
                    if (ok) mx = fmaxf(mx, e[k][jj]);
                }
            }
            const float mxs = mx * scale_log2e;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    if (IPT == 1 && k * 32 + jj >= NPAD) continue;
                    const int col = c * 32 + jj;
                    const bool ok = row_ok && ((IPT == 1) ? (col < S) : (col >= lo && col < hi));
                    const float x = ok ? ex2_approx(fmaf(e[k][jj], scale_log2e, -mxs)) : 0.f;
                    e[k][jj] = x;
                    sum += x;
                }
            }
            const float inv = row_ok ? 1.f / sum : 0.f;
            // ---- pass 2: D = rowsum(P * dP)
            float D = 0.f;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
                float dp[32];
                const bool half = IPT == 1 && NPAD - k * 32 < 32;
                if (half) tmem_ld16(tmem_base + lane_sel + C::COL_DP + c * 32, dp);
                else tmem_ld32(tmem_base + lane_sel + C::COL_DP + c * 32, dp);
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    if (half && jj >= 16) continue;
                    D = fmaf(e[k][jj], dp[jj], D);                     // e == 0 wherever the key is masked
                }
            }
            D *= inv;
            // ---- pass 3: P and dS = P (dP - D) * scale as bf16, [key chunk][row][16 B]
            //      (key chunks no row of this warp ever attends to keep the zeros written at kernel start)
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
                float dp[32];
                const bool half = IPT == 1 && NPAD - k * 32 < 32;
                if (half) tmem_ld16(tmem_base + lane_sel + C::COL_DP + c * 32, dp);
                else tmem_ld32(tmem_base + lane_sel + C::COL_DP + c * 32, dp);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (half && jj >= 2) continue;
                    uint32_t pp[4], dd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j0 = jj * 8 + 2 * u;
                        const float p0 = e[k][j0] * inv, p1 = e[k][j0 + 1] * inv;
                        pp[u] = pack_bf16x2(p0, p1);
                        dd[u] = pack_bf16x2(p0 * (dp[j0] - D) * scale, p1 * (dp[j0 + 1] - D) * scale);
                    }
                    *reinterpret_cast<uint4*>(prow + (c * 4 + jj) * 2048) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
                    *reinterpret_cast<uint4*>(dsrow + (c * 4 + jj) * 2048) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
                }
            }
            tc_fence_before();                                         // S / dP have been read: the accumulators may be reused
            fence_async_proxy();                                       // P / dS visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(SDP_EMPTY);
                mbar_arrive(PDS_FULL);
            }
            // ---- epilogue: row r of dQ (query r) and of dK, dV (key r)
            mbar_wait(OUT_FULL, ph);
            tc_fence_after();
            bf16* orow = dqkv + ((size_t)img0 * S + r) * ld_qkv + head * DP;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const uint32_t col = x == 0 ? C::COL_DQ : (x == 1 ? C::COL_DK : C::COL_DV);
                bf16* o = orow + x * inner;
#pragma unroll
                for (int c = 0; c < DP / 32; ++c) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_sel + col + c * 32, v);
                    if (row_ok) {
                        store16_bf16(o + c * 32, v);
                        store16_bf16(o + c * 32 + 16, v + 16);
                    }
                }
                if (DP % 32) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_sel + col + (DP / 32) * 32, v);
                    if (row_ok) store16_bf16(o + (DP / 32) * 32, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(OUT_EMPTY);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free(tmem_base, 512);
    }
}

template <typename C>
int launch_bwd_tc(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int heads, int ld_qkv, int ld_out, float scale,
                  cudaStream_t stream) {
    CUtensorMap tmQ, tmKV, tmDO;
    cuuint64_t dims[2] = {(cuuint64_t)(3 * heads * C::DP), (cuuint64_t)B * C::S};
    cuuint64_t strides[1] = {(cuuint64_t)ld_qkv * 2};
    cuuint64_t dims_o[2] = {(cuuint64_t)(heads * C::DP), (cuuint64_t)B * C::S};
    cuuint64_t strides_o[1] = {(cuuint64_t)ld_out * 2};
    cuuint32_t boxq[2] = {64, 128}, boxkv[2] = {64, (cuuint32_t)C::NPAD};
    SUNB_TRY(sunb_encode_tensor_map(&tmQ, qkv, 2, dims, strides, boxq));
    SUNB_TRY(sunb_encode_tensor_map(&tmKV, qkv, 2, dims, strides, boxkv));
    SUNB_TRY(sunb_encode_tensor_map(&tmDO, dout, 2, dims_o, strides_o, boxq));
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&attention_bwd_tc_kernel<C>), C::SMEM));
    const int n_tiles = ((B + C::IPT - 1) / C::IPT) * heads;
    const int sms = sunb_num_sms();
    const int grid = n_tiles < sms ? n_tiles : sms;
    SUNB_CHECK_CUDA(sunb_launch(&attention_bwd_tc_kernel<C>, dim3(grid), dim3(C::THREADS), C::SMEM, stream, tmQ, tmKV, tmDO, dqkv, B,
                                heads, ld_qkv, scale));
    return SUNB_OK;
}

}  // namespace

extern "C" int sunb_attention_backward(const void* qkv_, const void* dout_, void* dqkv_, int B, int S, int d, int d_stride,
                                       int heads, int ld_qkv, int ld_out, void* stream) {
    SUNB_REQUIRE(qkv_ && dout_ && dqkv_ && B > 0 && heads > 0 && d > 0 && d_stride >= d, "attention_backward: bad arguments");
    SUNB_REQUIRE(ld_qkv >= 3 * heads * d_stride && ld_out >= heads * d_stride, "attention_backward: row strides too small");
    const bf16* qkv = reinterpret_cast<const bf16*>(qkv_);
    const bf16* dout = reinterpret_cast<const bf16*>(dout_);
    bf16* dqkv = reinterpret_cast<bf16*>(dqkv_);
    const bool aligned = !((((size_t)qkv) & 15) || (((size_t)dout) & 15) || (((size_t)dqkv) & 31) || (ld_qkv % 16) || (ld_out % 8));
    const bool shape = (S == 100 && d_stride == 48) || (S == 25 && d_stride == 96 && d > 48);
    SUNB_REQUIRE(aligned && shape,
                 "attention_backward: unsupported problem S=%d d=%d d_stride=%d ld_qkv=%d ld_out=%d (supported: S=100 with d_stride 48, "
                 "S=25 with d_stride 96; 16-byte aligned inputs, 32-byte aligned dqkv, ld_qkv %% 16 == 0, ld_out %% 8 == 0)",
                 S, d, d_stride, ld_qkv, ld_out);
    const float scale = 1.0f / sqrtf((float)d);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (S == 100) return launch_bwd_tc<BCfg<100, 1, 112, 1, 2>>(qkv, dout, dqkv, B, heads, ld_qkv, ld_out, scale, st);
    return launch_bwd_tc<BCfg<25, 5, 128, 2, 1>>(qkv, dout, dqkv, B, heads, ld_qkv, ld_out, scale, st);
}
