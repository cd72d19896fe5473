// Multi-head self-attention core (reference: test_phase/models/visformer.py:183-190).
//   qkv : bf16 [B*S, ld_qkv], channel c = x*(heads*ds) + y*ds + z   (x in {q,k,v}, y head, z in [0,d), ds = head stride >= d)
//   out : bf16 [B*S, ld_out], channel y*ds + z
// ds == d is the reference's packed layout (visformer.py:186).  The eval engine pads every head to ds = 48 / 96 channels
// (zero weight rows -> the pad channels of q, k, v are exact zeros): every (token, head) segment then starts on a 16-byte
// boundary, the tiles are staged with 16-byte cp.async and the output (pad channels = 0) is written as bf16 pairs.
//   P = softmax(q k^T * d^-0.5), O = P v.
// Warp-level tensor-core kernel: the whole sequence (S = 100 or 25 tokens) of one (image, head) problem sits in
// shared memory (zero-padded to MMA shapes: S -> 112 / 32 keys, d 42 -> 48, 85 -> 96); each warp owns 16 query rows,
// computes the full score row block with mma.sync.m16n8k16 (bf16 in, fp32 accumulate), does the softmax on the
// accumulator registers (row max / sum via quad shuffles, exp2 with the scale folded in) and feeds the probabilities
// straight back as the A operand of the P.V MMAs.
// TEST-ONLY warp-MMA cross-check of attention_tc.cu (tests/native/libsunb200_check.so): never linked into libsunb200.so.
// It was round 1's product kernel; both engines run the tcgen05 kernel on the padded layout now, and this one also covers
// the reference's packed layout (ds == d), which the product rejects.
#include "../common.cuh"

#ifndef SUNB_ATT_CTAS
#define SUNB_ATT_CTAS 3
#endif
namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int S_PAD, int D_PAD, int PAIRS>
struct AttnCfg {
    static constexpr int QK_LD = D_PAD + 8;        // bf16 elements per Q/K/V row (conflict-free fragment loads)
    static constexpr int WARPS_PER_PAIR = S_PAD / 16;
    static constexpr int THREADS = 32 * WARPS_PER_PAIR * PAIRS;
    static constexpr int PAIR_ELEMS = 3 * S_PAD * QK_LD;
    static constexpr size_t SMEM = (size_t)PAIRS * PAIR_ELEMS * sizeof(bf16);
    // the S = 100 kernel is latency bound at 2 CTAs (14 warps) per SM: cap the registers for 3
    static constexpr int MIN_CTAS = (S_PAD > 32 && D_PAD <= 48) ? SUNB_ATT_CTAS : 1;
};

template <int S_PAD, int D_PAD, int PAIRS>
__global__ void __launch_bounds__(AttnCfg<S_PAD, D_PAD, PAIRS>::THREADS, AttnCfg<S_PAD, D_PAD, PAIRS>::MIN_CTAS)
attention_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int n_pairs, int S, int d, int ds, int heads,
                     int ld_qkv, int ld_out, float scale_log2e) {
    using Cfg = AttnCfg<S_PAD, D_PAD, PAIRS>;
    constexpr int QK_LD = Cfg::QK_LD;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* smem = reinterpret_cast<bf16*>(smem_raw);
    const int inner = heads * ds;

    // ---- stage Q, K, V (all row-major [token][d]) for the PAIRS problems of this CTA.  Padding must read as zero:
    //      clear the tiles, then fill the valid part -- with 4-byte cp.async when every row segment is 4-byte aligned
    //      (d even: stage 2), else with scalar loads (d = 85: odd heads start on a 2-byte boundary).
    // fast path: head segments are 16-byte aligned and padded with zeros up to ds (a multiple of 8 channels)
    const bool vec16 = ((ds & 7) == 0) && ((ld_qkv & 7) == 0) && ((((size_t)qkv) & 15) == 0);
    const int ncopy = ds < D_PAD ? ds : D_PAD;        // channels taken from global memory per row (fast path)
    if (!vec16 || ncopy < D_PAD) {                      // something in the tiles is not overwritten: clear everything
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        for (int i = threadIdx.x; i < (int)(Cfg::SMEM / 16); i += Cfg::THREADS) z[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
    } else {                                            // only the key / value rows past the sequence (P = 0 there, V must be finite)
        constexpr int RCH = QK_LD / 8;                  // 16-byte chunks per row
        for (int i = threadIdx.x; i < PAIRS * 2 * (S_PAD - S) * RCH; i += Cfg::THREADS) {
            const int ch = i % RCH, row = (i / RCH) % (S_PAD - S), kv = (i / (RCH * (S_PAD - S))) % 2, pl = i / (2 * RCH * (S_PAD - S));
            *reinterpret_cast<uint4*>(smem + pl * Cfg::PAIR_ELEMS + (1 + kv) * S_PAD * QK_LD + (S + row) * QK_LD + ch * 8) =
                make_uint4(0, 0, 0, 0);
        }
    }
    const bool vec2 = ((d & 1) == 0) && ((ds & 1) == 0) && ((ld_qkv & 1) == 0);
    for (int pl = 0; pl < PAIRS; ++pl) {
        const int pair = blockIdx.x * PAIRS + pl;
        if (pair >= n_pairs) break;
        bf16* sq = smem + pl * Cfg::PAIR_ELEMS;
        bf16* sk = sq + S_PAD * QK_LD;
        bf16* sv = sk + S_PAD * QK_LD;
        const int img = pair / heads, head = pair % heads;
        const bf16* base = qkv + (size_t)img * S * ld_qkv + head * ds;
        if (vec16) {
            const int cpr = ncopy >> 3;
            for (int i = threadIdx.x; i < S * cpr; i += Cfg::THREADS) {
                const int t = i / cpr, z = (i % cpr) * 8;
                const bf16* row = base + (size_t)t * ld_qkv + z;
                const uint32_t dq = (uint32_t)__cvta_generic_to_shared(sq + t * QK_LD + z);
                const uint32_t dk = (uint32_t)__cvta_generic_to_shared(sk + t * QK_LD + z);
                const uint32_t dv = (uint32_t)__cvta_generic_to_shared(sv + t * QK_LD + z);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dq), "l"(row) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dk), "l"(row + inner) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dv), "l"(row + 2 * inner) : "memory");
            }
        } else if (vec2) {
            const int dh = d >> 1;
            for (int i = threadIdx.x; i < S * dh; i += Cfg::THREADS) {
                const int t = i / dh, z = (i % dh) * 2;
                const bf16* row = base + (size_t)t * ld_qkv + z;
                const uint32_t dq = (uint32_t)__cvta_generic_to_shared(sq + t * QK_LD + z);
                const uint32_t dk = (uint32_t)__cvta_generic_to_shared(sk + t * QK_LD + z);
                const uint32_t dv = (uint32_t)__cvta_generic_to_shared(sv + t * QK_LD + z);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dq), "l"(row) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dk), "l"(row + inner) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dv), "l"(row + 2 * inner) : "memory");
            }
        } else {
            for (int i = threadIdx.x; i < S * d; i += Cfg::THREADS) {
                const int t = i / d, z = i % d;
                const bf16* row = base + (size_t)t * ld_qkv + z;
                sq[t * QK_LD + z] = row[0];
                sk[t * QK_LD + z] = row[inner];
                sv[t * QK_LD + z] = row[2 * inner];
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pl = warp / Cfg::WARPS_PER_PAIR, wq = warp % Cfg::WARPS_PER_PAIR;
    const int pair = blockIdx.x * PAIRS + pl;
    if (pair >= n_pairs) return;
    const bf16* sq = smem + pl * Cfg::PAIR_ELEMS;
    const bf16* sk = sq + S_PAD * QK_LD;
    const bf16* sv = sk + S_PAD * QK_LD;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = wq * 16;

    // ---- scores: 16 rows x S_PAD keys
    constexpr int NT = S_PAD / 8;
    float sc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < D_PAD; k0 += 16) {
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(sq + (r0 + g) * QK_LD + k0 + t * 2);
        const uint32_t a1 = *reinterpret_cast<const uint32_t*>(sq + (r0 + g + 8) * QK_LD + k0 + t * 2);
        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(sq + (r0 + g) * QK_LD + k0 + 8 + t * 2);
        const uint32_t a3 = *reinterpret_cast<const uint32_t*>(sq + (r0 + g + 8) * QK_LD + k0 + 8 + t * 2);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sk + (j * 8 + g) * QK_LD + k0 + t * 2);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sk + (j * 8 + g) * QK_LD + k0 + 8 + t * 2);
            mma_bf16_16816(sc[j], a0, a1, a2, a3, b0, b1);
        }
    }

    // ---- softmax over the S valid keys (rows g and g+8 of this warp's block); thread owns cols j*8 + t*2 + {0,1}
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int c = j * 8 + t * 2;
        if (c < S) { mx0 = fmaxf(mx0, sc[j][0]); mx1 = fmaxf(mx1, sc[j][2]); }
        if (c + 1 < S) { mx0 = fmaxf(mx0, sc[j][1]); mx1 = fmaxf(mx1, sc[j][3]); }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int c = j * 8 + t * 2;
        const float e0 = (c < S) ? exp2f((sc[j][0] - mx0) * scale_log2e) : 0.f;
        const float e1 = (c + 1 < S) ? exp2f((sc[j][1] - mx0) * scale_log2e) : 0.f;
        const float e2 = (c < S) ? exp2f((sc[j][2] - mx1) * scale_log2e) : 0.f;
        const float e3 = (c + 1 < S) ? exp2f((sc[j][3] - mx1) * scale_log2e) : 0.f;
        sc[j][0] = e0; sc[j][1] = e1; sc[j][2] = e2; sc[j][3] = e3;
        sum0 += e0 + e1;
        sum1 += e2 + e3;
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V : P (bf16, un-normalised, <= 1) comes straight from the score accumulators
    constexpr int OT = D_PAD / 8;
    float oc[OT][4];
#pragma unroll
    for (int j = 0; j < OT; ++j) oc[j][0] = oc[j][1] = oc[j][2] = oc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < S_PAD / 16; ++kk) {
        const uint32_t a0 = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
        const uint32_t a1 = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
        const uint32_t a2 = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        const uint32_t a3 = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
        // B fragments of V (row-major [key][d]) through ldmatrix.trans: lanes 0-7 / 8-15 address keys kk*16 + 0..7 / 8..15,
        // lanes 16-31 the same keys of the next 8 head-dim columns -> (b0, b1) of two n-tiles per instruction
        const bf16* vrow = sv + (kk * 16 + (lane & 15)) * QK_LD + ((lane >> 4) << 3);
#pragma unroll
        for (int j = 0; j < OT; j += 2) {
            uint32_t b0, b1, b2, b3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                         : "r"((uint32_t)__cvta_generic_to_shared(vrow + j * 8)));
            mma_bf16_16816(oc[j], a0, a1, a2, a3, b0, b1);
            mma_bf16_16816(oc[j + 1], a0, a1, a2, a3, b2, b3);
        }
    }
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    const int img = pair / heads, head = pair % heads;
    const int row0 = r0 + g, row1 = r0 + g + 8;
    bf16* o0 = out + (size_t)(img * S + row0) * ld_out + head * ds;
    bf16* o1 = out + (size_t)(img * S + row1) * ld_out + head * ds;
    if (vec16 && ((ld_out & 1) == 0) && ((((size_t)out) & 3) == 0)) {
        // padded layout: write all min(ds, D_PAD) channels (the pad channels are exact zeros) as bf16 pairs
#pragma unroll
        for (int j = 0; j < OT; ++j) {
            const int c = j * 8 + t * 2;
            if (c < ncopy) {
                if (row0 < S) *reinterpret_cast<__nv_bfloat162*>(o0 + c) = __floats2bfloat162_rn(oc[j][0] * inv0, oc[j][1] * inv0);
                if (row1 < S) *reinterpret_cast<__nv_bfloat162*>(o1 + c) = __floats2bfloat162_rn(oc[j][2] * inv1, oc[j][3] * inv1);
            }
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < OT; ++j) {
        const int c = j * 8 + t * 2;
        if (row0 < S) {
            if (c < d) o0[c] = __float2bfloat16(oc[j][0] * inv0);
            if (c + 1 < d) o0[c + 1] = __float2bfloat16(oc[j][1] * inv0);
        }
        if (row1 < S) {
            if (c < d) o1[c] = __float2bfloat16(oc[j][2] * inv1);
            if (c + 1 < d) o1[c + 1] = __float2bfloat16(oc[j][3] * inv1);
        }
    }
}

template <int S_PAD, int D_PAD, int PAIRS>
int launch_cfg(const bf16* qkv, bf16* out, int n_pairs, int S, int d, int ds, int heads, int ld_qkv, int ld_out, float scale,
               cudaStream_t stream) {
    using Cfg = AttnCfg<S_PAD, D_PAD, PAIRS>;
    SUNB_CHECK_CUDA(cudaFuncSetAttribute(attention_mma_kernel<S_PAD, D_PAD, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int blocks = (n_pairs + PAIRS - 1) / PAIRS;
    attention_mma_kernel<S_PAD, D_PAD, PAIRS><<<blocks, Cfg::THREADS, Cfg::SMEM, stream>>>(
        qkv, out, n_pairs, S, d, ds, heads, ld_qkv, ld_out, scale * 1.4426950408889634f);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

}  // namespace

// TEST-ONLY cross-check of the tcgen05 attention core: any head stride ds >= d (the reference's packed layout ds == d included)
extern "C" int sunb_check_attention(const void* qkv_, void* out_, int B, int S, int d, int ds, int heads, int ld_qkv, int ld_out,
                                    void* stream_) {
    const bf16* qkv = reinterpret_cast<const bf16*>(qkv_);
    bf16* out = reinterpret_cast<bf16*>(out_);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SUNB_REQUIRE(qkv && out && B > 0 && heads > 0, "check_attention: empty problem");
    SUNB_REQUIRE(ds >= d && ld_qkv >= 3 * heads * ds && ld_out >= heads * ds, "check_attention: head stride %d / row strides too small", ds);
    const float scale = 1.0f / sqrtf((float)d);
    const int n_pairs = B * heads;
    if (S <= 32 && d <= 96 && d > 48) return launch_cfg<32, 96, 4>(qkv, out, n_pairs, S, d, ds, heads, ld_qkv, ld_out, scale, stream);
    if (S <= 32 && d <= 48) return launch_cfg<32, 48, 4>(qkv, out, n_pairs, S, d, ds, heads, ld_qkv, ld_out, scale, stream);
    if (S <= 112 && d <= 48) return launch_cfg<112, 48, 1>(qkv, out, n_pairs, S, d, ds, heads, ld_qkv, ld_out, scale, stream);
    if (S <= 112 && d <= 96) return launch_cfg<112, 96, 1>(qkv, out, n_pairs, S, d, ds, heads, ld_qkv, ld_out, scale, stream);
    sunb_set_error("check_attention: unsupported shape S=%d d=%d (supported: S <= 112, d <= 96)", S, d);
    return SUNB_ERR_ARG;
}
