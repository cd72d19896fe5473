// Multi-head self-attention core on tcgen05 / TMEM / TMA (reference: test_phase/models/visformer.py:183-190) for the
// engines' padded head layout (head stride ds = 48 for d = 42, ds = 96 for d = 85; pad channels are exact zeros).
//   qkv : bf16 [B*S, ld_qkv], channel c = x*(heads*ds) + y*ds + z   (x in {q,k,v}, y head, z < ds)
//   out : bf16 [B*S, ld_out], channel y*ds + z          P = softmax(q k^T * d^-0.5), O = P v
//
// One 128-row tile per work item:
//   stage 2 (S = 100): one (image, head) problem per tile -- 100 query rows, keys padded to 112;
//   stage 3 (S = 25) : FIVE images of one head per tile -- 125 query rows, 128 key rows, block-diagonal mask (a query only
//                      attends to the 25 keys of its own image); the off-diagonal 80 % of the score tile is computed and
//                      masked, which is cheaper than five 32-row problems on a 128-lane tensor core.
// Persistent, warp-specialised CTA (one per SM, 320 threads):
//   warp 0      TMA producer: Q, K, V boxes (64 channels x rows, SWIZZLE_128B) of the tile; (Q, K) and V travel through
//               separate rings -- (Q, K) is released as soon as QK^T has been issued, V lives until P V -- so the loads run
//               several tiles ahead of the softmax (the kernel is bound by HBM latency x bytes in flight otherwise);
//   warp 1      MMA issuer: S = Q K^T (K-major operands, N = keys) into TMEM, later O = P V with P read from shared memory
//               (K-major, no swizzle) and V used in place as an MN-major operand (no transposed copy); software-pipelined so
//               QK^T of tile i+1 is issued before P V of tile i;
//   warps 2-9   two softmax groups (even / odd tiles) of four warps, one query row per thread: tcgen05.ld of the whole score
//               row into registers (the S buffer is handed back at once), max and sum in registers, exp2 with the scale
//               folded in, bf16 probabilities to shared memory, then the O epilogue (1/sum, bf16, 32-byte global stores).
// Accumulators: S double-buffered (2 x 128 TMEM columns), O double-buffered (2 x 128).
#include "tc_common.cuh"

namespace {

using namespace tc;

template <int S_, int IPT_, int NPAD_, int KA_, int NQK_, int NV_, int PBUFS_>
struct ACfg {
    static constexpr int S = S_, IPT = IPT_, NPAD = NPAD_, KA = KA_;
    static constexpr int NQK = NQK_, NV = NV_, PBUFS = PBUFS_;      // ring depths: (Q, K) stages, V stages, P buffers
    static constexpr int DP = KA_ == 1 ? 48 : 96;             // padded head width
    static constexpr int ROWS = S_ * IPT_;                    // valid query rows per tile (<= 128)
    static constexpr int Q_ATOM = 128 * 128;                  // 128 rows x 64 channels
    static constexpr int KV_ATOM = NPAD_ * 128;
    static constexpr int QK_STAGE = KA_ * (Q_ATOM + KV_ATOM);
    static constexpr int V_STAGE = KA_ * KV_ATOM;
    static constexpr int P_BYTES = (NPAD_ / 8) * 2048;        // [key chunk of 8][128 rows][16 B]
    static constexpr int V_OFF = NQK_ * QK_STAGE;
    static constexpr int P_OFF = V_OFF + NV_ * V_STAGE;
    static constexpr int BAR_OFF = P_OFF + PBUFS_ * P_BYTES;
    static constexpr int SMEM = 1024 + BAR_OFF + 512;
    static_assert(SMEM <= 232448, "attention_tc: shared memory budget");
    static constexpr int THREADS = 320;
    static constexpr int NCHUNK = (NPAD_ + 31) / 32;          // 32-column chunks of the score row (the last one may be 16 wide)
};

template <typename C>
__global__ void __launch_bounds__(C::THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, bf16* __restrict__ out,
                    int B, int heads, int ld_out, float scale_log2e) {
    constexpr int S = C::S, IPT = C::IPT, NPAD = C::NPAD, KA = C::KA, DP = C::DP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t vbase = base + C::V_OFF, pbase = base + C::P_OFF, bars = base + C::BAR_OFF;
    // barriers (8 B each)
    auto QK_FULL = [&](int i) { return bars + 8u * i; };
    auto QK_EMPTY = [&](int i) { return bars + 8u * (8 + i); };
    auto V_FULL = [&](int i) { return bars + 8u * (16 + i); };
    auto V_EMPTY = [&](int i) { return bars + 8u * (24 + i); };
    auto S_FULL = [&](int i) { return bars + 8u * (32 + i); };
    auto S_EMPTY = [&](int i) { return bars + 8u * (34 + i); };
    auto P_FULL = [&](int i) { return bars + 8u * (36 + i); };
    auto P_EMPTY = [&](int i) { return bars + 8u * (38 + i); };
    auto O_FULL = [&](int i) { return bars + 8u * (40 + i); };
    auto O_EMPTY = [&](int i) { return bars + 8u * (42 + i); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + C::BAR_OFF + 8 * 44);
    static_assert(C::NQK <= 8 && C::NV <= 8 && C::PBUFS <= 2, "barrier map");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img_tiles = (B + IPT - 1) / IPT;
    const int n_tiles = img_tiles * heads;
    const int n_local = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int inner = heads * DP;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NQK; ++i) { mbar_init(QK_FULL(i), 1); mbar_init(QK_EMPTY(i), 1); }
        for (int i = 0; i < C::NV; ++i) { mbar_init(V_FULL(i), 1); mbar_init(V_EMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(S_FULL(i), 1);
            mbar_init(S_EMPTY(i), 4);
            mbar_init(P_FULL(i), 4);
            mbar_init(P_EMPTY(i), 1);
            mbar_init(O_FULL(i), 1);
            mbar_init(O_EMPTY(i), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        }
        __syncwarp();
        // Q / K loads run one tile ahead of the V loads: a full V ring (V lives until P V) must not hold back the next Q K^T
        for (int i = 0; i <= n_local; ++i) {
            if (i < n_local) {
                const int t = blockIdx.x + i * gridDim.x;
                const int head = t % heads, row0 = (t / heads) * IPT * S;
                const int sq = i % C::NQK;
                mbar_wait(QK_EMPTY(sq), ((i / C::NQK) & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t st = base + sq * C::QK_STAGE;
                    mbar_expect_tx(QK_FULL(sq), C::QK_STAGE);
#pragma unroll
                    for (int a = 0; a < KA; ++a) {
                        tma_load_2d(st + a * C::Q_ATOM, &tmQ, QK_FULL(sq), head * DP + a * 64, row0);
                        tma_load_2d(st + KA * C::Q_ATOM + a * C::KV_ATOM, &tmKV, QK_FULL(sq), inner + head * DP + a * 64, row0);
                    }
                }
                __syncwarp();
            }
            if (i >= 1) {
                const int j = i - 1;
                const int t = blockIdx.x + j * gridDim.x;
                const int head = t % heads, row0 = (t / heads) * IPT * S;
                const int sv = j % C::NV;
                mbar_wait(V_EMPTY(sv), ((j / C::NV) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(V_FULL(sv), C::V_STAGE);
#pragma unroll
                    for (int a = 0; a < KA; ++a)
                        tma_load_2d(vbase + sv * C::V_STAGE + a * C::KV_ATOM, &tmKV, V_FULL(sv), 2 * inner + head * DP + a * 64, row0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        constexpr uint32_t idesc_qk = make_idesc(128, NPAD);
        constexpr uint32_t idesc_pv = make_idesc(128, KA * 64, false, true);       // B = V, MN-major
        for (int i = 0; i <= n_local; ++i) {
            if (i < n_local) {
                const int s = i & 1, ph = (i >> 1) & 1, sq = i % C::NQK;
                mbar_wait(QK_FULL(sq), (i / C::NQK) & 1);
                mbar_wait(S_EMPTY(s), ph ^ 1);
                tc_fence_after();
                const uint32_t q_sm = base + sq * C::QK_STAGE, k_sm = q_sm + KA * C::Q_ATOM;
                const uint32_t d_s = tmem_base + s * 128;
                if (elect_one()) {
#pragma unroll
                    for (int a = 0; a < KA; ++a) {
                        const int ksteps = (DP - a * 64) >= 64 ? 4 : (DP - a * 64) / 16;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < ksteps)
                                umma_bf16(d_s, desc_k_sw128(q_sm + a * C::Q_ATOM + k * 32), desc_k_sw128(k_sm + a * C::KV_ATOM + k * 32),
                                          idesc_qk, (a | k) ? 1u : 0u);
                    }
                    umma_commit(S_FULL(s));
                    umma_commit(QK_EMPTY(sq));          // Q and K are dead once QK^T has run: refill while the softmax works
                }
                __syncwarp();
            }
            if (i >= 1) {
                const int j = i - 1, s = j & 1, ph = (j >> 1) & 1, sv = j % C::NV;
                const int pb = j % C::PBUFS, pph = (j / C::PBUFS) & 1;
                mbar_wait(V_FULL(sv), (j / C::NV) & 1);
                mbar_wait(P_FULL(pb), pph);
                mbar_wait(O_EMPTY(s), ph ^ 1);
                tc_fence_after();
                const uint32_t v_sm = vbase + sv * C::V_STAGE;
                const uint32_t p_sm = pbase + pb * C::P_BYTES;
                const uint32_t d_o = tmem_base + 256 + s * 128;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < NPAD / 16; ++k)
                        umma_bf16(d_o, desc_k_noswz(p_sm + 2 * k * 2048, 2048, 128), desc_mn_sw128(v_sm + k * 2048, C::KV_ATOM),
                                  idesc_pv, k ? 1u : 0u);
                    umma_commit(O_FULL(s));
                    umma_commit(V_EMPTY(sv));
                    umma_commit(P_EMPTY(pb));
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ softmax + output: group g owns tiles g, g+2, ...
        const int g = (warp - 2) >> 2, q = warp & 3;
        const int r = q * 32 + lane;                                   // query row of this thread = TMEM lane
        const int blk = r < C::ROWS ? r / S : 0;
        const int lo = blk * S, hi = lo + S;                           // valid key columns of this row
        // 32-column chunks any row of this warp needs (warp-uniform)
        const int wlo = min(q * 32, C::ROWS - 1) / S * S, whi = min(q * 32 + 31, C::ROWS - 1) / S * S + S;
        const int c_lo = IPT == 1 ? 0 : wlo / 32, c_hi = IPT == 1 ? (S - 1) / 32 : (whi - 1) / 32;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        for (int i = g; i < n_local; i += 2) {
            const int t = blockIdx.x + i * gridDim.x;
            const int head = t % heads, img0 = (t / heads) * IPT;
            const int rows_valid = min(IPT, B - img0) * S;
            const int ph = (i >> 1) & 1, pb = i % C::PBUFS, pph = (i / C::PBUFS) & 1;
            const uint32_t t_s = tmem_base + lane_sel + g * 128;
            mbar_wait(S_FULL(g), ph);
            tc_fence_after();
            // ---- the score row in registers (only the 32-column chunks a row of this warp attends to: all of them for one
            //      image per tile, at most NL = 3 of 4 with five images per tile); the S buffer goes straight back to the MMA warp
            constexpr int NL = (IPT == 1) ? C::NCHUNK : 3;
            float v[NL][32];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;                                // warp-uniform
                if (IPT == 1 && NPAD - k * 32 < 32) tmem_ld16(t_s + c * 32, v[k]);
                else tmem_ld32(t_s + c * 32, v[k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(S_EMPTY(g));
            // ---- row max over the valid keys, then exp2 with the scale folded in, sum, bf16 probabilities (packed in place)
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int col = c * 32 + jj;
                    if (IPT == 1 && k * 32 + jj >= NPAD) continue;
                    const bool ok = (IPT == 1) ? (k * 32 + jj < S) : (col >= lo && col < hi);
                    if (ok) mx = fmaxf(mx, v[k][jj]);
                }
            }
            const float mxs = mx * scale_log2e;
            float sum = 0.f;
            uint32_t pk[NL][16];                                       // bf16 pairs of this row's probabilities
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int col = c * 32 + jj;
                    const bool in = !(IPT == 1 && k * 32 + jj >= NPAD);
                    const bool ok = in && ((IPT == 1) ? (k * 32 + jj < S) : (col >= lo && col < hi));
                    const float e = ok ? ex2_approx(fmaf(v[k][jj], scale_log2e, -mxs)) : 0.f;
                    v[k][jj] = e;
                    sum += e;
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) pk[k][jj] = pack_bf16x2(v[k][2 * jj], v[k][2 * jj + 1]);
            }
            // the P buffer is free once P V of the tile that used it has completed; only the stores sit behind that wait
            mbar_wait(P_EMPTY(pb), pph ^ 1);
            uint8_t* prow = base_ptr + C::P_OFF + pb * C::P_BYTES + r * 16;
            if (IPT != 1) {                                            // key chunks no row of this warp attends to: zeros
#pragma unroll
                for (int c = 0; c < C::NCHUNK; ++c)
                    if (c < c_lo || c > c_hi) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) *reinterpret_cast<uint4*>(prow + (c * 4 + jj) * 2048) = make_uint4(0, 0, 0, 0);
                    }
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const int c = c_lo + k;
                if (c > c_hi) continue;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (IPT == 1 && k * 32 + jj * 8 >= NPAD) continue;
                    *reinterpret_cast<uint4*>(prow + (c * 4 + jj) * 2048) =
                        make_uint4(pk[k][jj * 4], pk[k][jj * 4 + 1], pk[k][jj * 4 + 2], pk[k][jj * 4 + 3]);
                }
            }
            // probabilities written -> visible to the tensor core (async proxy)
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(P_FULL(pb));
            const float inv = 1.f / sum;
            // ---- output: O / sum -> bf16, DP channels of this head
            mbar_wait(O_FULL(g), ph);
            tc_fence_after();
            const uint32_t t_o = tmem_base + lane_sel + 256 + g * 128;
            bf16* orow = out + ((size_t)img0 * S + r) * ld_out + head * DP;
#pragma unroll
            for (int c = 0; c < DP / 32; ++c) {
                float v[32];
                tmem_ld32(t_o + c * 32, v);
                if (r < rows_valid) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) v[jj] *= inv;
                    store16_bf16(orow + c * 32, v);
                    store16_bf16(orow + c * 32 + 16, v + 16);
                }
            }
            if (DP % 32) {
                float v[16];
                tmem_ld16(t_o + (DP / 32) * 32, v);
                if (r < rows_valid) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) v[jj] *= inv;
                    store16_bf16(orow + (DP / 32) * 32, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(O_EMPTY(g));
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
int launch_tc(const bf16* qkv, bf16* out, int B, int heads, int ld_qkv, int ld_out, float scale, cudaStream_t stream) {
    CUtensorMap tmQ, tmKV;
    cuuint64_t dims[2] = {(cuuint64_t)(3 * heads * C::DP), (cuuint64_t)B * C::S};
    cuuint64_t strides[1] = {(cuuint64_t)ld_qkv * 2};
    cuuint32_t boxq[2] = {64, 128}, boxkv[2] = {64, (cuuint32_t)C::NPAD};
    SUNB_TRY(sunb_encode_tensor_map(&tmQ, qkv, 2, dims, strides, boxq));
    SUNB_TRY(sunb_encode_tensor_map(&tmKV, qkv, 2, dims, strides, boxkv));
    SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&attention_tc_kernel<C>), C::SMEM));
    const int n_tiles = ((B + C::IPT - 1) / C::IPT) * heads;
    const int sms = sunb_num_sms();
    const int grid = n_tiles < sms ? n_tiles : sms;
    SUNB_CHECK_CUDA(sunb_launch(&attention_tc_kernel<C>, dim3(grid), dim3(C::THREADS), C::SMEM, stream, tmQ, tmKV, out, B, heads,
                                ld_out, scale * 1.4426950408889634f));
    return SUNB_OK;
}

}  // namespace

// The product serves the padded head layouts of the two engines (head stride 48 for d = 42 at S = 100, 96 for d = 85 at S = 25);
// the reference's packed layout (head stride == d: segments not 16-byte aligned, no TMA box) is only implemented by the
// warp-MMA cross-check kernel of the test library.
static int attention_tc_supported(const bf16* qkv, const bf16* out, int S, int d, int ds, int ld_qkv, int ld_out) {
    if ((((size_t)qkv) & 15) || (((size_t)out) & 31) || (ld_qkv % 8) || (ld_out % 16)) return 0;
    if (S == 100 && ds == 48 && d <= 48) return 1;
    if (S == 25 && ds == 96 && d <= 96 && d > 48) return 1;
    return 0;
}

int sunb_launch_attention(const bf16* qkv, bf16* out, int B, int S, int d, int ds, int heads, int ld_qkv, int ld_out,
                          cudaStream_t stream) {
    SUNB_REQUIRE(B > 0 && heads > 0, "attention: empty problem");
    SUNB_REQUIRE(ds >= d && ld_qkv >= 3 * heads * ds && ld_out >= heads * ds, "attention: head stride %d / row strides too small", ds);
    SUNB_REQUIRE(attention_tc_supported(qkv, out, S, d, ds, ld_qkv, ld_out),
                 "attention: unsupported problem S=%d d=%d d_stride=%d ld_qkv=%d ld_out=%d (supported: S=100 with d_stride 48, S=25 with "
                 "d_stride 96; qkv 16-byte aligned with ld_qkv %% 8 == 0, out 32-byte aligned with ld_out %% 16 == 0)",
                 S, d, ds, ld_qkv, ld_out);
    const float scale = 1.0f / sqrtf((float)d);
    if (S == 100) return launch_tc<ACfg<100, 1, 112, 1, 3, 5, 2>>(qkv, out, B, heads, ld_qkv, ld_out, scale, stream);
    return launch_tc<ACfg<25, 5, 128, 2, 2, 2, 1>>(qkv, out, B, heads, ld_qkv, ld_out, scale, stream);
}
