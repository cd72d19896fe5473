// Multi-head self-attention core (reference: test_phase/models/visformer.py:183-190).
//   qkv : bf16 [B*S, ld_qkv], channel c = x*(heads*d) + y*d + z   (x in {q,k,v}, y head, z in [0,d))
//   out : bf16 [B*S, ld_out], channel y*d + z
//   P = softmax(q k^T * scale) in fp32, O = P v in fp32.
// One CTA per (image, head); the whole sequence (S = 100 or 25) lives in shared memory.  fp32 SIMT math:
// QK^T and PV are 1.2 % of the encoder FLOPs.
#include "common.cuh"

namespace {

constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                    int S, int d, int heads, int ld_qkv, int ld_out,
                                                                    float scale_log2e) {
    extern __shared__ float sm[];
    const int dp = d | 1;                      // odd row stride -> conflict-free column walks
    float* q = sm;                             // [S][dp]
    float* k = q + S * dp;                     // [S][dp]
    float* v = k + S * dp;                     // [S][dp]
    float* prob = v + S * dp;                  // [ATT_WARPS][S]
    const int img = blockIdx.x / heads, head = blockIdx.x % heads;
    const int inner = heads * d;
    for (int i = threadIdx.x; i < S * d; i += blockDim.x) {
        const int t = i / d, z = i % d;
        const bf16* row = qkv + (size_t)(img * S + t) * ld_qkv + head * d + z;
        q[t * dp + z] = __bfloat162float(row[0]);
        k[t * dp + z] = __bfloat162float(row[inner]);
        v[t * dp + z] = __bfloat162float(row[2 * inner]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* pr = prob + warp * S;
    for (int t = warp; t < S; t += ATT_WARPS) {
        float sc[4];                            // S <= 128 keys: lane owns keys lane, lane+32, lane+64, lane+96
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int key = lane + 32 * j;
            float s = -INFINITY;
            if (key < S) {
                s = 0.f;
                for (int z = 0; z < d; ++z) s = fmaf(q[t * dp + z], k[key * dp + z], s);
                s *= scale_log2e;
            }
            sc[j] = s;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float e = (lane + 32 * j < S) ? exp2f(sc[j] - mx) : 0.f;
            sc[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < S) pr[lane + 32 * j] = sc[j] * inv;
        __syncwarp();
        for (int z = lane; z < d; z += 32) {
            float o = 0.f;
            for (int key = 0; key < S; ++key) o = fmaf(pr[key], v[key * dp + z], o);
            out[(size_t)(img * S + t) * ld_out + head * d + z] = __float2bfloat16(o);
        }
        __syncwarp();
    }
}

}  // namespace

int sunb_launch_attention(const bf16* qkv, bf16* out, int B, int S, int d, int heads, int ld_qkv, int ld_out,
                          cudaStream_t stream) {
    SUNB_REQUIRE(S > 0 && S <= 128 && d > 0 && heads > 0, "attention: unsupported S=%d d=%d", S, d);
    const int dp = d | 1;
    const size_t smem = (size_t)(3 * S * dp + ATT_WARPS * S) * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        SUNB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const float scale = 1.0f / sqrtf((float)d);
    attention_kernel<<<B * heads, ATT_WARPS * 32, smem, stream>>>(qkv, out, S, d, heads, ld_qkv, ld_out,
                                                                  scale * 1.4426950408889634f);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
