// Encoder tail and episode head (HBM-bound warp-level kernels).
//   final_norm_pool : last BatchNorm (folded to scale/shift) + global average pool
//                     (reference: test_phase/models/visformer.py:455-462)
//   episode_head    : prototype mean over shots, L2-normalise, temperature-scaled cosine / -sq-distance logits
//                     (reference: test_phase/models/meta_baseline.py:36-46, test_phase/utils/__init__.py:78-101)
//   logits_ce_acc   : mean cross-entropy + accuracy of [R, W] logits against int64 labels
//                     (reference: test_phase/test_few_shot.py:89-90, test_phase/utils/__init__.py:104-109)
#include "common.cuh"

namespace {

// thread = one (image, channel); tokens are walked with stride C so a warp reads 32 consecutive channels.
__global__ void final_norm_pool_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float* __restrict__ dense,
                                       bf16* __restrict__ dense_bf16, float* __restrict__ pooled,
                                       bf16* __restrict__ pooled_bf16, int B, int T, int C) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int img = i / C, c = i % C;
    const float s = scale[c], t = shift[c];
    float acc = 0.f;
    for (int tok = 0; tok < T; ++tok) {
        const size_t off = ((size_t)img * T + tok) * C + c;
        const float v = __bfloat162float(x[off]) * s + t;
        if (dense) dense[off] = v;
        if (dense_bf16) dense_bf16[off] = __float2bfloat16(v);
        acc += v;
    }
    acc /= (float)T;
    pooled[(size_t)img * C + c] = acc;
    if (pooled_bf16) pooled_bf16[(size_t)img * C + c] = __float2bfloat16(acc);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// metric: 0 dot, 1 cos (normalise both), 2 sqr.  proto_from_shots: average `shot` support features first.
__global__ void __launch_bounds__(256) episode_logits_kernel(const float* __restrict__ feat_shot,
                                                             const float* __restrict__ feat_query,
                                                             float* __restrict__ logits, int way, int shot, int Q, int D,
                                                             int metric, const float* __restrict__ temp_dev,
                                                             float temp_host) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float proto[];            // [way][D]
    const int e = blockIdx.x;
    const float temp = temp_dev ? *temp_dev : temp_host;
    const float* fs = feat_shot + (size_t)e * way * shot * D;
    for (int i = threadIdx.x; i < way * D; i += blockDim.x) {
        const int w = i / D, dd = i % D;
        float s = 0.f;
        for (int k = 0; k < shot; ++k) s += fs[((size_t)w * shot + k) * D + dd];
        proto[i] = s / (float)shot;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    if (metric == 1) {
        for (int w = warp; w < way; w += nwarp) {
            float ss = 0.f;
            for (int dd = lane; dd < D; dd += 32) ss += proto[w * D + dd] * proto[w * D + dd];
            ss = warp_sum(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            for (int dd = lane; dd < D; dd += 32) proto[w * D + dd] *= inv;
        }
        __syncthreads();
    }
    // blockIdx.y: query slice (an episode is split over gridDim.y blocks; each rebuilds the prototypes)
    for (int qi = blockIdx.y * nwarp + warp; qi < Q; qi += gridDim.y * nwarp) {
        const float* fq = feat_query + ((size_t)e * Q + qi) * D;
        float qn = 1.f;
        if (metric == 1) {
            float ss = 0.f;
            for (int dd = lane; dd < D; dd += 32) ss += fq[dd] * fq[dd];
            ss = warp_sum(ss);
            qn = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        }
        for (int w = 0; w < way; ++w) {
            float acc = 0.f;
            if (metric == 2) {
                for (int dd = lane; dd < D; dd += 32) {
                    const float df = fq[dd] - proto[w * D + dd];
                    acc = fmaf(df, df, acc);
                }
                acc = -warp_sum(acc);
            } else {
                for (int dd = lane; dd < D; dd += 32) acc = fmaf(fq[dd] * qn, proto[w * D + dd], acc);
                acc = warp_sum(acc);
            }
            if (lane == 0) logits[((size_t)e * Q + qi) * way + w] = acc * temp;
        }
    }
}

// one block; out[0] = mean CE, out[1] = accuracy (first-max argmax, as torch.argmax on distinct values)
__global__ void __launch_bounds__(256) logits_ce_acc_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ label, int R, int W,
                                                            float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_loss[256], s_hit[256];
    float loss = 0.f, hit = 0.f;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float* l = logits + (size_t)r * W;
        float mx = l[0];
        int am = 0;
        for (int w = 1; w < W; ++w)
            if (l[w] > mx) { mx = l[w]; am = w; }
        float se = 0.f;
        for (int w = 0; w < W; ++w) se += expf(l[w] - mx);
        const int y = (int)label[r];
        loss += logf(se) + mx - l[y];
        hit += (am == y) ? 1.f : 0.f;
    }
    s_loss[threadIdx.x] = loss;
    s_hit[threadIdx.x] = hit;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
            s_hit[threadIdx.x] += s_hit[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s_loss[0] / (float)R;
        out[1] = s_hit[0] / (float)R;
    }
}

}  // namespace

int sunb_launch_final_norm_pool(const bf16* x, const float* scale, const float* shift, float* dense, bf16* dense_bf16,
                                float* pooled, bf16* pooled_bf16, int B, int T, int C, cudaStream_t stream) {
    SUNB_REQUIRE(B > 0 && pooled, "final_norm_pool: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&final_norm_pool_kernel, dim3((B * C + 255) / 256), dim3(256), 0, stream, x, scale, shift, dense, dense_bf16, pooled,
                                                                    pooled_bf16, B, T, C));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_episode_logits(const float* feat_shot, const float* feat_query, float* logits, int E, int way, int shot,
                               int Q, int D, int metric, const float* temp_dev, float temp_host, cudaStream_t stream) {
    SUNB_REQUIRE(E > 0 && way > 0 && shot > 0 && Q > 0 && D > 0, "episode_logits: empty problem");
    SUNB_REQUIRE(metric >= 0 && metric <= 2, "episode_logits: metric must be 0 (dot), 1 (cos) or 2 (sqr)");
    const size_t smem = (size_t)way * D * sizeof(float);
    SUNB_REQUIRE(smem <= 200 * 1024, "episode_logits: way*D too large for shared memory");
    if (smem > 48 * 1024) SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&episode_logits_kernel), (int)smem));
    int qs = E >= 64 ? 1 : (Q + 7) / 8;           // few episodes: split the queries of an episode over several blocks
    if (qs > 16) qs = 16;
    SUNB_CHECK_CUDA(sunb_launch(&episode_logits_kernel, dim3(E, qs), dim3(256), smem, stream, feat_shot, feat_query, logits, way, shot, Q, D, metric, temp_dev,
                                                    temp_host));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_logits_ce_acc(const float* logits, const long long* label, int R, int W, float* out, cudaStream_t stream) {
    SUNB_REQUIRE(R > 0 && W > 0, "logits_ce_acc: empty problem");
    SUNB_CHECK_CUDA(sunb_launch(&logits_ce_acc_kernel, dim3(1), dim3(256), 0, stream, logits, label, R, W, out));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of episode_logits (autograd of meta_baseline.py:36-46 / utils.compute_logits): one block per episode.
//   dfeat_shot [E,way,shot,D], dfeat_query [E,Q,D], dtemp (scalar, atomically accumulated) from dlogits [E,Q,way].
// ------------------------------------------------------------------------------------------------
namespace {

// Kernel A, grid (E, QS): block (e, j) owns every QS-th group of 8 queries of episode e.  It rebuilds the (normalised)
// prototypes, writes dquery for its queries (accumulated per warp in shared memory, one global store per element) and adds
// its partial prototype gradient into dshot[e, w, 0, :] (zeroed by the launcher) with global atomics; dtemp likewise.
__global__ void __launch_bounds__(256) episode_logits_bwd_kernel(const float* __restrict__ feat_shot,
                                                                 const float* __restrict__ feat_query,
                                                                 const float* __restrict__ dlogits,
                                                                 float* __restrict__ dshot, float* __restrict__ dquery,
                                                                 float* __restrict__ dtemp, int way, int shot, int Q, int D,
                                                                 int metric, const float* __restrict__ temp_dev,
                                                                 float temp_host) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smh[];
    float* proto = smh;                    // [way][D]  (normalised for cos)
    float* dph = proto + way * D;          // [way][D]  partial gradient w.r.t. the (normalised) prototype
    float* sdq = dph + way * D;            // [8 warps][D] gradient w.r.t. the normalised query, one row per warp
    __shared__ float s_dtemp;
    const int e = blockIdx.x;
    const float temp = temp_dev ? *temp_dev : temp_host;
    const float* fs = feat_shot + (size_t)e * way * shot * D;
    if (threadIdx.x == 0) s_dtemp = 0.f;
    for (int i = threadIdx.x; i < way * D; i += blockDim.x) {
        const int w = i / D, dd = i % D;
        float s = 0.f;
        for (int k = 0; k < shot; ++k) s += fs[((size_t)w * shot + k) * D + dd];
        proto[i] = s / (float)shot;
        dph[i] = 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    if (metric == 1) {
        for (int w = warp; w < way; w += nwarp) {
            float ss = 0.f;
            for (int dd = lane; dd < D; dd += 32) ss += proto[w * D + dd] * proto[w * D + dd];
            ss = warp_sum(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            for (int dd = lane; dd < D; dd += 32) proto[w * D + dd] *= inv;
        }
        __syncthreads();
    }
    float dt_local = 0.f;
    float* mydq = sdq + warp * D;
    for (int qi = blockIdx.y * nwarp + warp; qi < Q; qi += gridDim.y * nwarp) {
        const float* fq = feat_query + ((size_t)e * Q + qi) * D;
        const float* dl = dlogits + ((size_t)e * Q + qi) * way;
        float qinv = 1.f;
        if (metric == 1) {
            float ss = 0.f;
            for (int dd = lane; dd < D; dd += 32) ss += fq[dd] * fq[dd];
            ss = warp_sum(ss);
            qinv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        }
        for (int dd = lane; dd < D; dd += 32) mydq[dd] = 0.f;
        for (int w = 0; w < way; ++w) {
            const float g = dl[w];
            float dot = 0.f;
            for (int dd = lane; dd < D; dd += 32) {
                const float qh = fq[dd] * qinv, ph = proto[w * D + dd];
                if (metric == 2) {
                    const float df = qh - ph;
                    dot = fmaf(df, df, dot);
                    mydq[dd] += -2.f * temp * g * df;
                    atomicAdd(&dph[w * D + dd], 2.f * temp * g * df);
                } else {
                    dot = fmaf(qh, ph, dot);
                    mydq[dd] += temp * g * ph;
                    atomicAdd(&dph[w * D + dd], temp * g * qh);
                }
            }
            dot = warp_sum(dot);
            dt_local += g * (metric == 2 ? -dot : dot);
        }
        float* dq = dquery + ((size_t)e * Q + qi) * D;
        if (metric == 1) {                 // back through the query normalisation
            float proj = 0.f;
            for (int dd = lane; dd < D; dd += 32) proj = fmaf(fq[dd] * qinv, mydq[dd], proj);
            proj = warp_sum(proj);
            for (int dd = lane; dd < D; dd += 32) dq[dd] = (mydq[dd] - fq[dd] * qinv * proj) * qinv;
        } else {
            for (int dd = lane; dd < D; dd += 32) dq[dd] = mydq[dd];
        }
    }
    if (lane == 0) atomicAdd(&s_dtemp, dt_local);
    __syncthreads();
    for (int i = threadIdx.x; i < way * D; i += blockDim.x) {
        const int w = i / D, dd = i % D;
        atomicAdd(dshot + ((size_t)(e * way + w) * shot) * D + dd, dph[i]);
    }
    if (threadIdx.x == 0 && dtemp) atomicAdd(dtemp, s_dtemp);
}

// Kernel B, grid E: dshot[e, w, 0, :] holds the summed gradient w.r.t. the (normalised) prototype; go back through the
// normalisation and the mean over shots, and write every shot's row.
__global__ void __launch_bounds__(256) episode_logits_bwd_finish_kernel(const float* __restrict__ feat_shot,
                                                                        float* __restrict__ dshot, int way, int shot, int D,
                                                                        int metric) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smh[];
    float* proto = smh;                    // [way][D]
    float* dph = proto + way * D;          // [way][D]
    const int e = blockIdx.x;
    const float* fs = feat_shot + (size_t)e * way * shot * D;
    for (int i = threadIdx.x; i < way * D; i += blockDim.x) {
        const int w = i / D, dd = i % D;
        float s = 0.f;
        for (int k = 0; k < shot; ++k) s += fs[((size_t)w * shot + k) * D + dd];
        proto[i] = s / (float)shot;
        dph[i] = dshot[((size_t)(e * way + w) * shot) * D + dd];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int w = warp; w < way; w += nwarp) {
        float inv = 1.f, proj = 0.f;
        if (metric == 1) {
            float ss = 0.f;
            for (int dd = lane; dd < D; dd += 32) ss += proto[w * D + dd] * proto[w * D + dd];
            ss = warp_sum(ss);
            inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            for (int dd = lane; dd < D; dd += 32) proj = fmaf(proto[w * D + dd] * inv, dph[w * D + dd], proj);
            proj = warp_sum(proj);
        }
        for (int dd = lane; dd < D; dd += 32) {
            float g = dph[w * D + dd];
            if (metric == 1) g = (g - proto[w * D + dd] * inv * proj) * inv;
            g /= (float)shot;
            for (int k = 0; k < shot; ++k) dshot[((size_t)(e * way + w) * shot + k) * D + dd] = g;
        }
    }
}

}  // namespace

int sunb_launch_episode_logits_bwd(const float* feat_shot, const float* feat_query, const float* dlogits, float* dshot,
                                   float* dquery, float* dtemp, int E, int way, int shot, int Q, int D, int metric,
                                   const float* temp_dev, float temp_host, cudaStream_t stream) {
    SUNB_REQUIRE(E > 0 && way > 0 && shot > 0 && Q > 0 && D > 0 && metric >= 0 && metric <= 2, "episode_logits_bwd: bad shape");
    const size_t smem = (size_t)(2 * way * D + 8 * D) * sizeof(float);
    SUNB_REQUIRE(smem <= 200 * 1024, "episode_logits_bwd: way*D too large for shared memory");
    if (smem > 48 * 1024) {
        SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&episode_logits_bwd_kernel), (int)smem));
        SUNB_TRY(sunb_opt_in_smem(reinterpret_cast<const void*>(&episode_logits_bwd_finish_kernel), (int)smem));
    }
    // an episode is split over QS blocks (8 queries per block and pass) so that a few episodes still fill the GPU
    int qs = (Q + 7) / 8;
    if (qs > 16) qs = 16;
    SUNB_CHECK_CUDA(cudaMemsetAsync(dshot, 0, (size_t)E * way * shot * D * sizeof(float), stream));
    SUNB_CHECK_CUDA(sunb_launch(&episode_logits_bwd_kernel, dim3(E, qs), dim3(256), smem, stream, feat_shot, feat_query, dlogits, dshot, dquery, dtemp, way, shot, Q,
                                                                  D, metric, temp_dev, temp_host));
    SUNB_CHECK_CUDA(cudaGetLastError());
    SUNB_CHECK_CUDA(sunb_launch(&episode_logits_bwd_finish_kernel, dim3(E), dim3(256), (size_t)(2 * way * D) * sizeof(float), stream, feat_shot, dshot, way, shot, D, metric));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

// ------------------------------------------------------------------------------------------------
// Channel LayerNorm over NHWC rows (reference: LayerNorm wrapper, test_phase/models/visformer.py:109-115 --
// not instantiated by visformer_micro_80; provided for API completeness).  One warp per row, fp32, two-pass variance.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y, long M,
                                                             int C, float eps) {
    pdl_trigger();
    pdl_wait();
    const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[row * C + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}
}  // namespace

int sunb_launch_layernorm_rows(const float* x, const float* gamma, const float* beta, float* y, long M, int C, float eps,
                               cudaStream_t stream) {
    SUNB_REQUIRE(x && gamma && beta && y && M > 0 && C > 0, "layernorm_rows: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&layernorm_rows_kernel, dim3((int)((M + 7) / 8)), dim3(256), 0, stream, x, gamma, beta, y, M, C, eps));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
