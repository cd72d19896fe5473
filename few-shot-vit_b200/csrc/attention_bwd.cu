// Backward of the attention core (forward: attention.cu; reference: autograd through visformer.py:183-190).
//   inputs : qkv bf16 [B*S, ld_qkv] (saved forward input), dout bf16 [B*S, ld_out] (gradient of the head-concatenated output)
//   output : dqkv bf16 [B*S, ld_qkv], same channel order (qkv, head, d)
// Per (image, head): P = softmax(scale * q k^T) is recomputed, then
//   dV = P^T dO,  dP = dO V^T,  dS = P * (dP - rowsum(P * dP)),  dQ = scale * dS K,  dK = scale * dS^T Q.
// fp32 SIMT with the whole problem in shared memory (S <= 100, d <= 85); ~3 % of the backward FLOPs.
#include "common.cuh"

namespace {

constexpr int AB_WARPS = 8;

__global__ void __launch_bounds__(AB_WARPS * 32) attention_bwd_kernel(const bf16* __restrict__ qkv,
                                                                       const bf16* __restrict__ dout,
                                                                       bf16* __restrict__ dqkv, int S, int d, int heads,
                                                                       int ld_qkv, int ld_out, float scale) {
    extern __shared__ float sm[];
    const int dp = d | 1;
    float* q = sm;                       // [S][dp]
    float* k = q + S * dp;
    float* v = k + S * dp;
    float* dO = v + S * dp;
    float* P = dO + S * dp;              // [S][S]
    float* dS = P + S * S;               // [S][S]
    const int img = blockIdx.x / heads, head = blockIdx.x % heads;
    const int inner = heads * d;
    for (int i = threadIdx.x; i < S * d; i += blockDim.x) {
        const int t = i / d, z = i % d;
        const bf16* row = qkv + (size_t)(img * S + t) * ld_qkv + head * d + z;
        q[t * dp + z] = __bfloat162float(row[0]);
        k[t * dp + z] = __bfloat162float(row[inner]);
        v[t * dp + z] = __bfloat162float(row[2 * inner]);
        dO[t * dp + z] = __bfloat162float(dout[(size_t)(img * S + t) * ld_out + head * d + z]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float sl2 = scale * 1.4426950408889634f;
    // phase 1: rows of P and dS
    for (int i = warp; i < S; i += AB_WARPS) {
        float sc[4], dpv[4];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int key = lane + 32 * j;
            float s = -INFINITY, g = 0.f;
            if (key < S) {
                s = 0.f;
                for (int z = 0; z < d; ++z) {
                    s = fmaf(q[i * dp + z], k[key * dp + z], s);
                    g = fmaf(dO[i * dp + z], v[key * dp + z], g);
                }
                s *= sl2;
            }
            sc[j] = s;
            dpv[j] = g;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sc[j] = (lane + 32 * j < S) ? exp2f(sc[j] - mx) : 0.f;
            sum += sc[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sc[j] *= inv;
            dot = fmaf(sc[j], dpv[j], dot);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int key = lane + 32 * j;
            if (key < S) {
                P[i * S + key] = sc[j];
                dS[i * S + key] = sc[j] * (dpv[j] - dot) * scale;      // gradient w.r.t. q.k (scale folded in)
            }
        }
    }
    __syncthreads();
    // phase 2: dQ (row i), dK and dV (row j); lanes over the head dimension
    for (int i = warp; i < S; i += AB_WARPS) {
        bf16* oq = dqkv + (size_t)(img * S + i) * ld_qkv + head * d;
        for (int z = lane; z < d; z += 32) {
            float aq = 0.f, ak = 0.f, av = 0.f;
            for (int j = 0; j < S; ++j) {
                aq = fmaf(dS[i * S + j], k[j * dp + z], aq);
                ak = fmaf(dS[j * S + i], q[j * dp + z], ak);
                av = fmaf(P[j * S + i], dO[j * dp + z], av);
            }
            oq[z] = __float2bfloat16(aq);
            oq[inner + z] = __float2bfloat16(ak);
            oq[2 * inner + z] = __float2bfloat16(av);
        }
    }
}

}  // namespace

extern "C" int sunb_attention_backward(const void* qkv, const void* dout, void* dqkv, int B, int S, int d, int heads,
                                       int ld_qkv, int ld_out, void* stream) {
    SUNB_REQUIRE(qkv && dout && dqkv && B > 0, "attention_backward: bad arguments");
    SUNB_REQUIRE(S > 0 && S <= 128 && d > 0, "attention_backward: unsupported S=%d d=%d", S, d);
    const int dp = d | 1;
    const size_t smem = (size_t)(4 * S * dp + 2 * S * S) * sizeof(float);
    SUNB_REQUIRE(smem <= 220 * 1024, "attention_backward: problem does not fit shared memory");
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        SUNB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    attention_bwd_kernel<<<B * heads, AB_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dqkv), S, d, heads,
        ld_qkv, ld_out, 1.0f / sqrtf((float)d));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
