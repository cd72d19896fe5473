// Fused multi-tensor optimizers for the meta-tuning / meta-training steps: ONE launch updates every parameter tensor.
//   SGD   (momentum, weight decay)  : torch.optim.SGD semantics as utils.make_optimizer builds it
//                                     (reference meta_tuning_sun_m/utils/__init__.py:128-139; train_meta_warmup.py:140)
//   AdamW (decoupled weight decay)  : torch / timm AdamW semantics (reference sun_meta_training/offline.py:229)
// HBM-bound: 16 B (SGD: p, g, m read + p, m written = 20 B) / 28 B per parameter element, float4 accesses.
// The tensor table (pointers + sizes) travels BY VALUE in the kernel arguments (<= 96 tensors per launch, 4.6 KB), so no
// host-to-device copy happens at step time and a CUDA-graph capture of the training step records the table in its kernel
// node.  The hyper-parameters and the AdamW step counter live in device memory: a captured graph keeps working when the
// scheduler changes the learning rate or the step counter advances.
#include "common.cuh"
#include "../../include/sunb200.h"

namespace {

constexpr int CHUNK = 4096;                 // elements per block-iteration
constexpr int THREADS = 256;

// hyper-parameter block (fp32, device): [0] lr, [1] momentum | beta1, [2] weight_decay, [3] beta2, [4] eps, [5] step (as float)
__device__ __forceinline__ int find_tensor(const long long* __restrict__ chunk_prefix, int n, long long chunk) {
    int lo = 0, hi = n;                      // largest t with chunk_prefix[t] <= chunk
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_prefix[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int MAX_T = 96;                    // tensors per launch
struct OptArgs {
    SunbOptTensor t[MAX_T];
    long long prefix[MAX_T + 1];             // prefix sums of the per-tensor chunk counts
    int n;
};

template <bool ADAMW>
__global__ void __launch_bounds__(THREADS) fused_opt_kernel(const __grid_constant__ OptArgs args, const float* __restrict__ hp) {
    pdl_trigger();
    pdl_wait();
    const SunbOptTensor* tensors = args.t;
    const long long* chunk_prefix = args.prefix;
    const int n_tensors = args.n;
    const long long total = chunk_prefix[n_tensors];
    const float lr = hp[0], wd = hp[2];
    float mu = 0.f, b1 = 0.f, b2 = 0.f, eps = 0.f, step_size = 0.f, inv_sqrt_bc2 = 0.f;
    if (ADAMW) {
        b1 = hp[1]; b2 = hp[3]; eps = hp[4];
        const float t = hp[5];                                    // already incremented for this step
        step_size = lr / (1.f - powf(b1, t));
        inv_sqrt_bc2 = rsqrtf(1.f - powf(b2, t));
    } else {
        mu = hp[1];
    }
    for (long long c = blockIdx.x; c < total; c += gridDim.x) {
        const int ti = find_tensor(chunk_prefix, n_tensors, c);
        const SunbOptTensor T = tensors[ti];
        const long long base = (c - chunk_prefix[ti]) * CHUNK;
        float* p = reinterpret_cast<float*>(T.p);
        const float* g = reinterpret_cast<const float*>(T.g);
        float* m = reinterpret_cast<float*>(T.m);
        float* v = reinterpret_cast<float*>(T.v);
        if (g == nullptr) continue;                               // parameter without a gradient this step
        const long long end = base + CHUNK < T.n ? base + CHUNK : T.n;
        const bool vec = ((((size_t)p) | ((size_t)g) | ((size_t)m) | (ADAMW ? (size_t)v : 0)) & 15) == 0;
        if (vec) {
            for (long long i = base + threadIdx.x * 4; i + 3 < end; i += THREADS * 4) {
                float4 pp = *reinterpret_cast<float4*>(p + i);
                const float4 gg = *reinterpret_cast<const float4*>(g + i);
                float4 mm = *reinterpret_cast<float4*>(m + i);
                float* pa = reinterpret_cast<float*>(&pp);
                const float* ga = reinterpret_cast<const float*>(&gg);
                float* ma = reinterpret_cast<float*>(&mm);
                if (ADAMW) {
                    float4 vv = *reinterpret_cast<float4*>(v + i);
                    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pa[j] *= 1.f - lr * wd;
                        ma[j] = b1 * ma[j] + (1.f - b1) * ga[j];
                        va[j] = b2 * va[j] + (1.f - b2) * ga[j] * ga[j];
                        pa[j] -= step_size * ma[j] / (sqrtf(va[j]) * inv_sqrt_bc2 + eps);
                    }
                    *reinterpret_cast<float4*>(v + i) = vv;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = fmaf(wd, pa[j], ga[j]);
                        ma[j] = fmaf(mu, ma[j], d);
                        pa[j] -= lr * ma[j];
                    }
                }
                *reinterpret_cast<float4*>(p + i) = pp;
                *reinterpret_cast<float4*>(m + i) = mm;
            }
        }
        // scalar path: unaligned tensors, and the < 4-element tail of the last chunk
        const long long tail0 = vec ? base + ((end - base) & ~3LL) : base;
        for (long long i = tail0 + threadIdx.x; i < end; i += THREADS) {
            float pv = p[i];
            const float gv = g[i];
            if (ADAMW) {
                pv *= 1.f - lr * wd;
                const float mv = b1 * m[i] + (1.f - b1) * gv;
                const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
                pv -= step_size * mv / (sqrtf(vv) * inv_sqrt_bc2 + eps);
                m[i] = mv;
                v[i] = vv;
            } else {
                const float mv = fmaf(mu, m[i], fmaf(wd, pv, gv));
                m[i] = mv;
                pv -= lr * mv;
            }
            p[i] = pv;
        }
    }
}

__global__ void opt_step_inc_kernel(float* hp) {
    pdl_trigger();
    pdl_wait();
    hp[5] += 1.f;
}

template <bool ADAMW>
int launch_opt(const SunbOptTensor* tensors, int n_tensors, float* hp_dev, cudaStream_t st) {
    for (int t0 = 0; t0 < n_tensors; t0 += MAX_T) {
        OptArgs a;
        a.n = n_tensors - t0 < MAX_T ? n_tensors - t0 : MAX_T;
        a.prefix[0] = 0;
        for (int i = 0; i < a.n; ++i) {
            a.t[i] = tensors[t0 + i];
            SUNB_REQUIRE(a.t[i].p && a.t[i].m && a.t[i].n > 0 && (!ADAMW || a.t[i].v), "fused optimizer: bad tensor entry %d", t0 + i);
            a.prefix[i + 1] = a.prefix[i] + (a.t[i].n + CHUNK - 1) / CHUNK;
        }
        const long long total = a.prefix[a.n];
        const int grid = (int)(total < 148 * 8 ? total : 148 * 8);
        SUNB_CHECK_CUDA(sunb_launch(&fused_opt_kernel<ADAMW>, dim3(grid), dim3(THREADS), 0, st, a, hp_dev));
        SUNB_CHECK_CUDA(cudaGetLastError());
    }
    return SUNB_OK;
}

}  // namespace

extern "C" {

int sunb_fused_sgd(const SunbOptTensor* tensors, int n_tensors, float* hp_dev, void* stream) {
    SUNB_REQUIRE(tensors && hp_dev && n_tensors > 0, "fused_sgd: bad arguments");
    return launch_opt<false>(tensors, n_tensors, hp_dev, reinterpret_cast<cudaStream_t>(stream));
}

int sunb_fused_adamw(const SunbOptTensor* tensors, int n_tensors, float* hp_dev, void* stream) {
    SUNB_REQUIRE(tensors && hp_dev && n_tensors > 0, "fused_adamw: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SUNB_CHECK_CUDA(sunb_launch(&opt_step_inc_kernel, dim3(1), dim3(1), 0, st, hp_dev));    // the step counter lives on the device
    return launch_opt<true>(tensors, n_tensors, hp_dev, st);
}

}  // extern "C"
