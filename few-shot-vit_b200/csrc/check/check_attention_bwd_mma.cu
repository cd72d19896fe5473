// TEST-ONLY warp-MMA cross-check of attention_bwd_tc.cu (tests/native/libsunb200_check.so): never linked into libsunb200.so.
// Backward of the attention core on the reference's PACKED head layout (forward: check_attention_mma.cu; reference: autograd
// through visformer.py:183-190).
//   inputs : qkv bf16 [B*S, ld_qkv] (saved forward input), dout bf16 [B*S, ld_out] (gradient of the head-concatenated output)
//   output : dqkv bf16 [B*S, ld_qkv], same channel order (qkv, head, d)
// Per (image, head) the whole problem lives in shared memory (zero padded to MMA shapes) and is processed in two phases,
// both on mma.sync.m16n8k16 (bf16 in, fp32 accumulate), FlashAttention-2 style:
//   A (warp owns 16 query rows): S = Q K^T -> row max / sum (kept in smem), dP = dO V^T, D = rowsum(P*dP),
//                                dS = P*(dP - D)*scale, dQ = dS K
//   B (warp owns 16 key rows)  : S^T = K Q^T -> P^T from the stored row statistics, dP^T = V dO^T,
//                                dV = P^T dO, dK = dS^T Q
// Probabilities / dS are rounded to bf16 only as MMA operands; statistics and accumulators stay fp32.
#include "../common.cuh"

#ifndef SUNB_ATTB_CTAS
#define SUNB_ATTB_CTAS 2
#endif

namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t ld32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

template <int S_PAD, int D_PAD, int PAIRS>
struct BwdCfg {
    static constexpr int RLD = D_PAD + 8;          // row-major tiles  [S_PAD][RLD]
    static constexpr int TLD = S_PAD + 8;          // transposed tiles [D_PAD][TLD]
    static constexpr int WARPS_PER_PAIR = S_PAD / 16;
    static constexpr int THREADS = 32 * WARPS_PER_PAIR * PAIRS;
    static constexpr int ROW_ELEMS = S_PAD * RLD, T_ELEMS = D_PAD * TLD;
    static constexpr int PAIR_BF16 = 4 * ROW_ELEMS + 3 * T_ELEMS;                 // Q K V dO | Qt Kt dOt
    static constexpr size_t PAIR_BYTES = (size_t)PAIR_BF16 * 2 + 3 * S_PAD * sizeof(float);   // + rowmax, 1/rowsum, D
    static constexpr size_t SMEM = PAIRS * PAIR_BYTES;
    // S = 100: shared memory allows two CTAs per SM, so cap the registers for two (175 -> <= 144) -- the kernel is latency bound
    static constexpr int MIN_CTAS = (S_PAD > 32 && D_PAD <= 48) ? SUNB_ATTB_CTAS : 1;
};

// C = A(16 rows starting at a_row0 of a row-major tile, K = KD) x B^T(rows n of a row-major tile), NT n-tiles of 8
template <int NT, int KD, int LD>
__device__ __forceinline__ void mma_rows_x_rows(float (&acc)[NT][4], const bf16* a_tile, int a_row0, const bf16* b_tile,
                                                int g, int t) {
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < KD; k0 += 16) {
        const uint32_t a0 = ld32(a_tile + (a_row0 + g) * LD + k0 + t * 2);
        const uint32_t a1 = ld32(a_tile + (a_row0 + g + 8) * LD + k0 + t * 2);
        const uint32_t a2 = ld32(a_tile + (a_row0 + g) * LD + k0 + 8 + t * 2);
        const uint32_t a3 = ld32(a_tile + (a_row0 + g + 8) * LD + k0 + 8 + t * 2);
#pragma unroll
        for (int j = 0; j < NT; ++j)
            mma16816(acc[j], a0, a1, a2, a3, ld32(b_tile + (j * 8 + g) * LD + k0 + t * 2),
                     ld32(b_tile + (j * 8 + g) * LD + k0 + 8 + t * 2));
    }
}

// out(16 x D_PAD) = P(16 x S_PAD, from accumulator registers) x M(S_PAD x D_PAD) with M given transposed [D_PAD][TLD]
template <int NT, int OT, int TLD>
__device__ __forceinline__ void mma_regs_x_t(float (&out)[OT][4], const float (&p)[NT][4], const bf16* mt, int g, int t) {
#pragma unroll
    for (int j = 0; j < OT; ++j) out[j][0] = out[j][1] = out[j][2] = out[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        const uint32_t a0 = pack2(p[2 * kk][0], p[2 * kk][1]);
        const uint32_t a1 = pack2(p[2 * kk][2], p[2 * kk][3]);
        const uint32_t a2 = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        const uint32_t a3 = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int j = 0; j < OT; ++j)
            mma16816(out[j], a0, a1, a2, a3, ld32(mt + (j * 8 + g) * TLD + kk * 16 + t * 2),
                     ld32(mt + (j * 8 + g) * TLD + kk * 16 + 8 + t * 2));
    }
}

template <int S_PAD, int D_PAD, int PAIRS>
__global__ void __launch_bounds__(BwdCfg<S_PAD, D_PAD, PAIRS>::THREADS, BwdCfg<S_PAD, D_PAD, PAIRS>::MIN_CTAS)
attention_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int n_pairs,
                         int S, int d, int heads, int ld_qkv, int ld_out, float scale) {
    using Cfg = BwdCfg<S_PAD, D_PAD, PAIRS>;
    constexpr int RLD = Cfg::RLD, TLD = Cfg::TLD, NT = S_PAD / 8, OT = D_PAD / 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int inner = heads * d;
    const bf16 zero = __float2bfloat16(0.f);

    for (int pl = 0; pl < PAIRS; ++pl) {
        const int pair = blockIdx.x * PAIRS + pl;
        bf16* base = reinterpret_cast<bf16*>(smem_raw + pl * Cfg::PAIR_BYTES);
        bf16 *sq = base, *sk = sq + Cfg::ROW_ELEMS, *sv = sk + Cfg::ROW_ELEMS, *sdo = sv + Cfg::ROW_ELEMS;
        bf16 *sqt = sdo + Cfg::ROW_ELEMS, *skt = sqt + Cfg::T_ELEMS, *sdot = skt + Cfg::T_ELEMS;
        const bool live = pair < n_pairs;
        const int img = live ? pair / heads : 0, head = live ? pair % heads : 0;
        for (int i = threadIdx.x; i < S_PAD * D_PAD; i += Cfg::THREADS) {
            const int tk = i / D_PAD, z = i % D_PAD;
            const bool in = live && tk < S && z < d;
            const bf16* row = qkv + (size_t)(img * S + tk) * ld_qkv + head * d + z;
            const bf16 q = in ? row[0] : zero, k = in ? row[inner] : zero, v = in ? row[2 * inner] : zero;
            const bf16 o = in ? dout[(size_t)(img * S + tk) * ld_out + head * d + z] : zero;
            sq[tk * RLD + z] = q; sk[tk * RLD + z] = k; sv[tk * RLD + z] = v; sdo[tk * RLD + z] = o;
            sqt[z * TLD + tk] = q; skt[z * TLD + tk] = k; sdot[z * TLD + tk] = o;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pl = warp / Cfg::WARPS_PER_PAIR, wb = warp % Cfg::WARPS_PER_PAIR;
    const int pair = blockIdx.x * PAIRS + pl;
    const bool live = pair < n_pairs;
    bf16* base = reinterpret_cast<bf16*>(smem_raw + pl * Cfg::PAIR_BYTES);
    const bf16 *sq = base, *sk = sq + Cfg::ROW_ELEMS, *sv = sk + Cfg::ROW_ELEMS, *sdo = sv + Cfg::ROW_ELEMS;
    const bf16 *sqt = sdo + Cfg::ROW_ELEMS, *skt = sqt + Cfg::T_ELEMS, *sdot = skt + Cfg::T_ELEMS;
    float* rmax = reinterpret_cast<float*>(base + Cfg::PAIR_BF16);
    float* rinv = rmax + S_PAD;
    float* rD = rinv + S_PAD;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = wb * 16;
    const float sl2 = scale * 1.4426950408889634f;
    const int img = live ? pair / heads : 0, head = live ? pair % heads : 0;

    // ---------------- phase A: query-row block r0 .. r0+15
    {
        float sc[NT][4], dp[NT][4];
        mma_rows_x_rows<NT, D_PAD, RLD>(sc, sq, r0, sk, g, t);        // S = Q K^T
        mma_rows_x_rows<NT, D_PAD, RLD>(dp, sdo, r0, sv, g, t);       // dP = dO V^T
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = j * 8 + t * 2;
            if (c < S) { mx0 = fmaxf(mx0, sc[j][0]); mx1 = fmaxf(mx1, sc[j][2]); }
            if (c + 1 < S) { mx0 = fmaxf(mx0, sc[j][1]); mx1 = fmaxf(mx1, sc[j][3]); }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = j * 8 + t * 2;
            sc[j][0] = (c < S) ? exp2f((sc[j][0] - mx0) * sl2) : 0.f;
            sc[j][1] = (c + 1 < S) ? exp2f((sc[j][1] - mx0) * sl2) : 0.f;
            sc[j][2] = (c < S) ? exp2f((sc[j][2] - mx1) * sl2) : 0.f;
            sc[j][3] = (c + 1 < S) ? exp2f((sc[j][3] - mx1) * sl2) : 0.f;
            sum0 += sc[j][0] + sc[j][1];
            sum1 += sc[j][2] + sc[j][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
        float D0 = 0.f, D1 = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sc[j][0] *= inv0; sc[j][1] *= inv0; sc[j][2] *= inv1; sc[j][3] *= inv1;
            D0 += sc[j][0] * dp[j][0] + sc[j][1] * dp[j][1];
            D1 += sc[j][2] * dp[j][2] + sc[j][3] * dp[j][3];
        }
        D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
        D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
        if (t == 0) {
            rmax[r0 + g] = mx0; rinv[r0 + g] = inv0; rD[r0 + g] = D0;
            rmax[r0 + g + 8] = mx1; rinv[r0 + g + 8] = inv1; rD[r0 + g + 8] = D1;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {          // dS (scale folded in), reuse sc
            sc[j][0] = sc[j][0] * (dp[j][0] - D0) * scale;
            sc[j][1] = sc[j][1] * (dp[j][1] - D0) * scale;
            sc[j][2] = sc[j][2] * (dp[j][2] - D1) * scale;
            sc[j][3] = sc[j][3] * (dp[j][3] - D1) * scale;
        }
        float dq[OT][4];
        mma_regs_x_t<NT, OT, TLD>(dq, sc, skt, g, t);                 // dQ = dS K
        if (live) {
            const int row0 = r0 + g, row1 = r0 + g + 8;
            bf16* o0 = dqkv + (size_t)(img * S + row0) * ld_qkv + head * d;
            bf16* o1 = dqkv + (size_t)(img * S + row1) * ld_qkv + head * d;
#pragma unroll
            for (int j = 0; j < OT; ++j) {
                const int c = j * 8 + t * 2;
                if (row0 < S) { if (c < d) o0[c] = __float2bfloat16(dq[j][0]); if (c + 1 < d) o0[c + 1] = __float2bfloat16(dq[j][1]); }
                if (row1 < S) { if (c < d) o1[c] = __float2bfloat16(dq[j][2]); if (c + 1 < d) o1[c + 1] = __float2bfloat16(dq[j][3]); }
            }
        }
    }
    __syncthreads();

    // ---------------- phase B: key-row block r0 .. r0+15 (columns = queries)
    {
        float pt[NT][4], dpt[NT][4];
        mma_rows_x_rows<NT, D_PAD, RLD>(pt, sk, r0, sq, g, t);        // S^T = K Q^T
        mma_rows_x_rows<NT, D_PAD, RLD>(dpt, sv, r0, sdo, g, t);      // dP^T = V dO^T
        const bool k0ok = r0 + g < S, k1ok = r0 + g + 8 < S;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = j * 8 + t * 2 + e;          // query index of this column
                const bool iok = i < S;
                const float m = iok ? rmax[i] : 0.f, inv = iok ? rinv[i] : 0.f, Dv = iok ? rD[i] : 0.f;
                const float p0 = (iok && k0ok) ? exp2f((pt[j][e] - m) * sl2) * inv : 0.f;
                const float p1 = (iok && k1ok) ? exp2f((pt[j][2 + e] - m) * sl2) * inv : 0.f;
                pt[j][e] = p0;
                pt[j][2 + e] = p1;
                dpt[j][e] = p0 * (dpt[j][e] - Dv) * scale;            // dS^T
                dpt[j][2 + e] = p1 * (dpt[j][2 + e] - Dv) * scale;
            }
        }
        float dv[OT][4], dk[OT][4];
        mma_regs_x_t<NT, OT, TLD>(dv, pt, sdot, g, t);                // dV = P^T dO
        mma_regs_x_t<NT, OT, TLD>(dk, dpt, sqt, g, t);                // dK = dS^T Q
        if (live) {
            const int row0 = r0 + g, row1 = r0 + g + 8;
            bf16* o0 = dqkv + (size_t)(img * S + row0) * ld_qkv + head * d;
            bf16* o1 = dqkv + (size_t)(img * S + row1) * ld_qkv + head * d;
#pragma unroll
            for (int j = 0; j < OT; ++j) {
                const int c = j * 8 + t * 2;
                if (row0 < S) {
                    if (c < d) { o0[inner + c] = __float2bfloat16(dk[j][0]); o0[2 * inner + c] = __float2bfloat16(dv[j][0]); }
                    if (c + 1 < d) { o0[inner + c + 1] = __float2bfloat16(dk[j][1]); o0[2 * inner + c + 1] = __float2bfloat16(dv[j][1]); }
                }
                if (row1 < S) {
                    if (c < d) { o1[inner + c] = __float2bfloat16(dk[j][2]); o1[2 * inner + c] = __float2bfloat16(dv[j][2]); }
                    if (c + 1 < d) { o1[inner + c + 1] = __float2bfloat16(dk[j][3]); o1[2 * inner + c + 1] = __float2bfloat16(dv[j][3]); }
                }
            }
        }
    }
}

template <int S_PAD, int D_PAD, int PAIRS>
int launch_bwd(const bf16* qkv, const bf16* dout, bf16* dqkv, int n_pairs, int S, int d, int heads, int ld_qkv, int ld_out,
               cudaStream_t stream) {
    using Cfg = BwdCfg<S_PAD, D_PAD, PAIRS>;
    SUNB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_mma_kernel<S_PAD, D_PAD, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int blocks = (n_pairs + PAIRS - 1) / PAIRS;
    attention_bwd_mma_kernel<S_PAD, D_PAD, PAIRS><<<blocks, Cfg::THREADS, Cfg::SMEM, stream>>>(
        qkv, dout, dqkv, n_pairs, S, d, heads, ld_qkv, ld_out, 1.0f / sqrtf((float)d));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

}  // namespace

// TEST-ONLY cross-check of the tcgen05 attention backward: the reference's PACKED head layout (head stride == d)
extern "C" int sunb_check_attention_backward(const void* qkv, const void* dout, void* dqkv, int B, int S, int d, int heads,
                                             int ld_qkv, int ld_out, void* stream) {
    SUNB_REQUIRE(qkv && dout && dqkv && B > 0 && heads > 0 && d > 0, "check_attention_backward: bad arguments");
    const bf16* q = reinterpret_cast<const bf16*>(qkv);
    const bf16* o = reinterpret_cast<const bf16*>(dout);
    bf16* dq = reinterpret_cast<bf16*>(dqkv);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_pairs = B * heads;
    if (S <= 32 && d <= 96 && d > 48) return launch_bwd<32, 96, 2>(q, o, dq, n_pairs, S, d, heads, ld_qkv, ld_out, st);
    if (S <= 32 && d <= 48) return launch_bwd<32, 48, 2>(q, o, dq, n_pairs, S, d, heads, ld_qkv, ld_out, st);
    if (S <= 112 && d <= 48) return launch_bwd<112, 48, 1>(q, o, dq, n_pairs, S, d, heads, ld_qkv, ld_out, st);
    if (S <= 112 && d <= 96) return launch_bwd<112, 96, 1>(q, o, dq, n_pairs, S, d, heads, ld_qkv, ld_out, st);
    sunb_set_error("check_attention_backward: unsupported shape S=%d d=%d (supported: S <= 112, d <= 96)", S, d);
    return SUNB_ERR_ARG;
}
