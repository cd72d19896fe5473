// Train-mode support kernels: batch-statistics BatchNorm (forward and backward), max-pool tail of the stem,
// DropPath row scaling, layout helpers and weight preparation.  All HBM-bound, NHWC bf16 activations, fp32 statistics.
// Reference semantics: nn.BatchNorm2d in training mode (test_phase/models/visformer.py:118-124),
// ConvBlock tail (visformer.py:232-237), drop_path (visformer.py:89-97).
#include "common.cuh"
#include "../../include/sunb200.h"

#include <string.h>

int sunb_launch_wgrad_tc(WgradParams p, cudaStream_t stream);

namespace {

// ------------------------------------------------------------------------------------------------
// column statistics over rows:  sum[c] += sum_m x[m,c];  sq[c] += sum_m x[m,c]^2  (u == null)  or  x[m,c]*u[m,c]
// block = 256 threads = (256 / (C/8)) row lanes x (C/8) vectors of 8 channels; requires C % 8 == 0, C/8 <= 256
// ------------------------------------------------------------------------------------------------
// Optional fused finalize: the LAST block to finish (atomic ticket) turns the complete column sums into the BatchNorm
// coefficients, so a BatchNorm statistics pass is one launch instead of two (forward: scale / shift / running statistics;
// backward: the dx coefficients and dgamma / dbeta).
struct BnFin {
    int mode;                       // 0 none, 1 forward (bn_finalize), 2 backward (bn_bwd_finalize)
    unsigned int* ticket;           // zero before the launch
    float count;
    const float* gamma; const float* beta;
    float* rmean; float* rvar; long long* nbt; float momentum, eps;
    float* scale; float* shift; float* mean; float* rstd;          // forward: outputs; backward: mean / rstd are inputs
    int frozen; float* a; float* c1; float* c2; float* dgamma; float* dbeta;
};

__global__ void __launch_bounds__(256) colstats_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ u,
                                                       int ldu, int M, int C, int rows_per_block,
                                                       float* __restrict__ sum, float* __restrict__ sq, const BnFin fin) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float red[];          // [2][256][8]
    const int nvec = C / 8;
    const int lanes = 256 / nvec;
    const int vec = threadIdx.x % nvec, rl = threadIdx.x / nvec;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int m0 = blockIdx.x * rows_per_block;
    const int m1 = min(m0 + rows_per_block, M);
    if (rl < lanes) {
#pragma unroll 4
        for (int m = m0 + rl; m < m1; m += lanes) {
            const uint4 a = *reinterpret_cast<const uint4*>(x + (size_t)m * ldx + vec * 8);
            const bf16* ah = reinterpret_cast<const bf16*>(&a);
            if (u) {
                const uint4 b = *reinterpret_cast<const uint4*>(u + (size_t)m * ldu + vec * 8);
                const bf16* bh = reinterpret_cast<const bf16*>(&b);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = __bfloat162float(ah[j]);
                    s[j] += v;
                    q[j] = fmaf(v, __bfloat162float(bh[j]), q[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = __bfloat162float(ah[j]);
                    s[j] += v;
                    q[j] = fmaf(v, v, q[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x * 8 + j] = s[j];
        red[2048 + threadIdx.x * 8 + j] = q[j];
    }
    __syncthreads();
    // thread t < C reduces channel t over the row lanes
    for (int c = threadIdx.x; c < C; c += 256) {
        const int v = c / 8, j = c % 8;
        float a = 0.f, b = 0.f;
        for (int l = 0; l < lanes; ++l) {
            a += red[(l * nvec + v) * 8 + j];
            b += red[2048 + (l * nvec + v) * 8 + j];
        }
        atomicAdd(sum + c, a);
        atomicAdd(sq + c, b);
    }
    if (fin.mode == 0) return;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += 256) {
        const float s0 = __ldcg(sum + c), s1 = __ldcg(sq + c);
        if (fin.mode == 1) {
            const float mean = s0 / fin.count;
            const float var = fmaxf(s1 / fin.count - mean * mean, 0.f);
            const float rstd = rsqrtf(var + fin.eps);
            const float sc = fin.gamma[c] * rstd;
            fin.scale[c] = sc;
            fin.shift[c] = fin.beta[c] - mean * sc;
            fin.mean[c] = mean;
            fin.rstd[c] = rstd;
            if (fin.rmean) {
                fin.rmean[c] = (1.f - fin.momentum) * fin.rmean[c] + fin.momentum * mean;
                fin.rvar[c] = (1.f - fin.momentum) * fin.rvar[c] + fin.momentum * var * (fin.count / fmaxf(fin.count - 1.f, 1.f));
            }
        } else {
            const float mean = fin.mean[c], rstd = fin.rstd[c];
            const float cen = s1 - mean * s0;                    // sum dz*(x-mean)
            fin.a[c] = fin.gamma[c] * rstd;
            fin.c1[c] = fin.frozen ? 0.f : s0 / fin.count;
            fin.c2[c] = fin.frozen ? 0.f : rstd * rstd * cen / fin.count;
            fin.dgamma[c] += rstd * cen;
            fin.dbeta[c] += s0;
        }
    }
    if (threadIdx.x == 0 && fin.mode == 1 && fin.nbt) *fin.nbt += 1;
}

// one thread per channel: batch mean / biased var -> scale, shift; running stats with the unbiased var
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sq, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, long long* __restrict__ nbt,
                                   float momentum, float eps, int C, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    const float mean = sum[c] / count;
    const float var = fmaxf(sq[c] / count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float s = gamma[c] * rstd;
    scale[c] = s;
    shift[c] = beta[c] - mean * s;
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    if (rmean) {
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
    }
}

// frozen BatchNorm (module in eval() inside a training step, utils.freeze_bn): running statistics instead of batch statistics
__global__ void bn_frozen_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ rmean, const float* __restrict__ rvar, float eps, int C,
                                 float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                 float* __restrict__ rstd_out) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float rstd = rsqrtf(rvar[c] + eps);
    scale[c] = gamma[c] * rstd;
    shift[c] = beta[c] - rmean[c] * gamma[c] * rstd;
    mean_out[c] = rmean[c];
    rstd_out[c] = rstd;
}

// y = act(x*scale + shift) + tab[(m % tab_mod)][c]   (tab nullable); vector of 8 channels per thread
__global__ void bn_apply_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ scale,
                                const float* __restrict__ shift, int act, const float* __restrict__ tab, int tab_mod,
                                bf16* __restrict__ out, int ldo, long M, int C) {
    pdl_trigger();
    pdl_wait();
    const int nvec = C / 8;
    if (blockDim.x % nvec == 0) {
        // a thread keeps one 8-channel vector for its whole life (scale / shift in registers) and walks the rows
        const int c0 = (threadIdx.x % nvec) * 8, lanes = blockDim.x / nvec;
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
#pragma unroll 4
        for (long m = (long)blockIdx.x * lanes + threadIdx.x / nvec; m < M; m += (long)gridDim.x * lanes) {
            const uint4 a = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
            const bf16* ah = reinterpret_cast<const bf16*>(&a);
            float v[8];
            if (act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_fast(fmaf(__bfloat162float(ah[j]), sc[j], sh[j]));
            } else if (act == ACT_LRELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float t = fmaf(__bfloat162float(ah[j]), sc[j], sh[j]); v[j] = t > 0.f ? t : 0.1f * t; }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(__bfloat162float(ah[j]), sc[j], sh[j]);
            }
            if (tab) {
                const float4* t = reinterpret_cast<const float4*>(tab + (size_t)(m % tab_mod) * C + c0);
                const float4 t0 = t[0], t1 = t[1];
                v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
            }
            uint4 o;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(out + m * ldo + c0) = o;
        }
        return;
    }

    const long total = M * nvec;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = i / nvec;
        const int c0 = (int)(i % nvec) * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
        const bf16* ah = reinterpret_cast<const bf16*>(&a);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_apply(fmaf(__bfloat162float(ah[j]), scale[c0 + j], shift[c0 + j]), act);
        if (tab) {
            const float* t = tab + (size_t)(m % tab_mod) * C + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j];
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + m * ldo + c0) = o;
    }
}

// stem tail forward: out = maxpool2x2( lrelu( c3*s3+t3 + id*sd+td ) ) + pos
__global__ void stem_tail_fwd_kernel(const bf16* __restrict__ c3, const bf16* __restrict__ idn,
                                     const float* __restrict__ s3, const float* __restrict__ t3,
                                     const float* __restrict__ sd, const float* __restrict__ td,
                                     const float* __restrict__ pos, bf16* __restrict__ out, int B, int H, int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int oH = H / 2, oW = W / 2, cv = C / 8;
    const long total = (long)B * oH * oW * cv;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * 8;
        const long px = i / cv;
        const int ox = (int)(px % oW), oy = (int)((px / oW) % oH), img = (int)(px / ((long)oW * oH));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const size_t off = ((size_t)(img * H + 2 * oy + (d >> 1)) * W + 2 * ox + (d & 1)) * C + c0;
            const uint4 a = *reinterpret_cast<const uint4*>(c3 + off);
            const uint4 b = *reinterpret_cast<const uint4*>(idn + off);
            const bf16* ah = reinterpret_cast<const bf16*>(&a);
            const bf16* bh = reinterpret_cast<const bf16*>(&b);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float z = fmaf(__bfloat162float(ah[j]), s3[c0 + j], t3[c0 + j]) +
                          fmaf(__bfloat162float(bh[j]), sd[c0 + j], td[c0 + j]);
                z = z > 0.f ? z : 0.1f * z;
                m[j] = fmaxf(m[j], z);
            }
        }
        const float* pp = pos + (size_t)(oy * oW + ox) * C + c0;
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(m[2 * j] + pp[2 * j], m[2 * j + 1] + pp[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + (size_t)px * C + c0) = o;
    }
}

// stem tail backward: dz[full res] = g[pooled] * lrelu'(z) at the arg-max of each 2x2 window (first max in scan order), else 0
__global__ void stem_tail_bwd_kernel(const bf16* __restrict__ c3, const bf16* __restrict__ idn,
                                     const float* __restrict__ s3, const float* __restrict__ t3,
                                     const float* __restrict__ sd, const float* __restrict__ td,
                                     const bf16* __restrict__ g, bf16* __restrict__ dz, int B, int H, int W, int C) {
    pdl_trigger();
    pdl_wait();
    const int oH = H / 2, oW = W / 2, cv = C / 8;
    const long total = (long)B * oH * oW * cv;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * 8;
        const long px = i / cv;
        const int ox = (int)(px % oW), oy = (int)((px / oW) % oH), img = (int)(px / ((long)oW * oH));
        float best[8];
        int arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const size_t off = ((size_t)(img * H + 2 * oy + (d >> 1)) * W + 2 * ox + (d & 1)) * C + c0;
            const uint4 a = *reinterpret_cast<const uint4*>(c3 + off);
            const uint4 b = *reinterpret_cast<const uint4*>(idn + off);
            const bf16* ah = reinterpret_cast<const bf16*>(&a);
            const bf16* bh = reinterpret_cast<const bf16*>(&b);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float z = fmaf(__bfloat162float(ah[j]), s3[c0 + j], t3[c0 + j]) +
                          fmaf(__bfloat162float(bh[j]), sd[c0 + j], td[c0 + j]);
                z = z > 0.f ? z : 0.1f * z;                   // lrelu is monotonic: arg-max is unchanged
                if (z > best[j]) { best[j] = z; arg[j] = d; }
            }
        }
        const uint4 gg = *reinterpret_cast<const uint4*>(g + (size_t)px * C + c0);
        const bf16* gh = reinterpret_cast<const bf16*>(&gg);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            uint4 o;
            bf16* oh = reinterpret_cast<bf16*>(&o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float slope = best[j] > 0.f ? 1.f : 0.1f;
                oh[j] = __float2bfloat16(arg[j] == d ? __bfloat162float(gh[j]) * slope : 0.f);
            }
            const size_t off = ((size_t)(img * H + 2 * oy + (d >> 1)) * W + 2 * ox + (d & 1)) * C + c0;
            *reinterpret_cast<uint4*>(dz + off) = o;
        }
    }
}

// BN backward coefficients.  dx = a*(dz - c1 - (x - mean)*c2);  dgamma = rstd*(sum(dz*x) - mean*sum(dz));  dbeta = sum(dz)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sdz, const float* __restrict__ sdzx, float count,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       const float* __restrict__ gamma, int C, int frozen, float* __restrict__ a,
                                       float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float cen = sdzx[c] - mean[c] * sdz[c];        // sum dz*(x-mean)
    a[c] = gamma[c] * rstd[c];
    // frozen statistics do not depend on the batch: dx = a*dz, no mean / projection terms
    c1[c] = frozen ? 0.f : sdz[c] / count;
    c2[c] = frozen ? 0.f : rstd[c] * rstd[c] * cen / count;
    dgamma[c] += rstd[c] * cen;
    dbeta[c] += sdz[c];
}

// dx = a*(dz - c1 - (x - mean)*c2) + res     (res nullable)
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dz, int lddz, const bf16* __restrict__ x, int ldx,
                                    const float* __restrict__ a, const float* __restrict__ c1,
                                    const float* __restrict__ c2, const float* __restrict__ mean,
                                    const bf16* __restrict__ res, int ldr, bf16* __restrict__ out, int ldo, long M, int C) {
    pdl_trigger();
    pdl_wait();
    const int nvec = C / 8;
    if (blockDim.x % nvec == 0) {
        // fixed 8-channel vector per thread: dx = k1*dz + k2*x + k3 with k1 = a, k2 = -a*c2, k3 = a*(c2*mean - c1) in registers
        const int c0 = (threadIdx.x % nvec) * 8, lanes = blockDim.x / nvec;
        float k1[8], k2[8], k3[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float A = a[c0 + j], C2 = c2[c0 + j];
            k1[j] = A; k2[j] = -A * C2; k3[j] = A * (C2 * mean[c0 + j] - c1[c0 + j]);
        }
#pragma unroll 4
        for (long m = (long)blockIdx.x * lanes + threadIdx.x / nvec; m < M; m += (long)gridDim.x * lanes) {
            const uint4 d4 = *reinterpret_cast<const uint4*>(dz + m * lddz + c0);
            const uint4 x4 = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
            const bf16* dh = reinterpret_cast<const bf16*>(&d4);
            const bf16* xh = reinterpret_cast<const bf16*>(&x4);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(k1[j], __bfloat162float(dh[j]), fmaf(k2[j], __bfloat162float(xh[j]), k3[j]));
            if (res) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(res + m * ldr + c0);
                const bf16* rh = reinterpret_cast<const bf16*>(&r4);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += __bfloat162float(rh[j]);
            }
            uint4 o;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(out + m * ldo + c0) = o;
        }
        return;
    }

    const long total = M * nvec;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = i / nvec;
        const int c0 = (int)(i % nvec) * 8;
        const uint4 d4 = *reinterpret_cast<const uint4*>(dz + m * lddz + c0);
        const uint4 x4 = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
        const bf16* dh = reinterpret_cast<const bf16*>(&d4);
        const bf16* xh = reinterpret_cast<const bf16*>(&x4);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            v[j] = a[c0 + j] * (__bfloat162float(dh[j]) - c1[c0 + j] - (__bfloat162float(xh[j]) - mean[c0 + j]) * c2[c0 + j]);
        if (res) {
            const uint4 r4 = *reinterpret_cast<const uint4*>(res + m * ldr + c0);
            const bf16* rh = reinterpret_cast<const bf16*>(&r4);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __bfloat162float(rh[j]);
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + m * ldo + c0) = o;
    }
}

// out[r, c] = in[r, c] * rs[r / rows_per_img]
__global__ void scale_rows_kernel(const bf16* __restrict__ in, const float* __restrict__ rs, int rows_per_img,
                                  bf16* __restrict__ out, long M, int C) {
    pdl_trigger();
    pdl_wait();
    const int nvec = C / 8;
    const long total = M * nvec;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = i / nvec;
        const float s = rs[m / rows_per_img];
        const uint4 a = *reinterpret_cast<const uint4*>(in + i * 8);
        const bf16* ah = reinterpret_cast<const bf16*>(&a);
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            oh[j] = __floats2bfloat162_rn(__bfloat162float(ah[2 * j]) * s, __bfloat162float(ah[2 * j + 1]) * s);
        *reinterpret_cast<uint4*>(out + i * 8) = o;
    }
}

// out[s2d_row(m)] <- in[m] (dir 0: raster -> space-to-depth) or out[m] <- in[s2d_row(m)] (dir 1), rows of C channels
__global__ void s2d_reorder_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C, int dir) {
    pdl_trigger();
    pdl_wait();
    const int nvec = C / 8;
    const long total = (long)B * H * W * nvec;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = i / nvec;
        const int v = (int)(i % nvec);
        const int hw = H * W;
        const int img = (int)(m / hw), rem = (int)(m % hw), y = rem / W, x = rem % W;
        const long sm = ((long)(img * (H / 2) + y / 2) * (W / 2) + x / 2) * 4 + (y & 1) * 2 + (x & 1);
        const long src = dir ? sm : m, dst = dir ? m : sm;
        *reinterpret_cast<uint4*>(out + dst * C + v * 8) = *reinterpret_cast<const uint4*>(in + src * C + v * 8);
    }
}

// out[i] += sum_b g[b, i]   (i over S*C), fp32 accumulate -- gradient of a broadcast positional embedding
__global__ void batch_sum_kernel(const bf16* __restrict__ g, int B, long n, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += __bfloat162float(g[(size_t)b * n + i]);
        out[i] += a;
    }
}

// generic permute + cast: dst[a][b][c] (bf16, contiguous, last dim padded to ldd) = src[off + a*sa + b*sb + c*sc] (fp32)
__global__ void permute_cast_kernel(const float* __restrict__ src, long off, long sa, long sb, long sc, int A, int Bd,
                                    int Cd, int ldd, bf16* __restrict__ dst) {
    pdl_trigger();
    pdl_wait();
    const long total = (long)A * Bd * ldd;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ldd);
        const long ab = i / ldd;
        const int b = (int)(ab % Bd), a = (int)(ab / Bd);
        dst[i] = __float2bfloat16(c < Cd ? src[off + a * sa + b * sb + c * sc] : 0.f);
    }
}

// Every bf16 operand layout of the training step (forward + dgrad copies of all GEMM / conv weights) in ONE launch.
// Each entry is a 4-D permute + cast  dst[a][b][c][d] = src[off + a*sa + b*sb + c*sc + d*sd]  (last dim padded to ldd);
// the table travels by value in the kernel arguments (graph-capturable, nothing uploaded at step time).
constexpr int PK_MAX = 96, PK_T2 = 32, PK_T3 = 64;     // work unit: a 32 (dim 2) x 64 (dim 3) tile through shared memory
struct PackArgs {
    SunbPackDesc d[PK_MAX];
    int prefix[PK_MAX + 1];                             // prefix sums of the per-entry tile counts
    int n;
};
// Reads run along whichever of the two inner dimensions is denser in the SOURCE (a `.d` operand is the transpose of the
// master weight: its dim 2 is the contiguous one), writes always along dim 3 of the destination as bf16 pairs.
__global__ void __launch_bounds__(256) pack_multi_kernel(const __grid_constant__ PackArgs args) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[PK_T2][PK_T3 + 1];
    const int total = args.prefix[args.n];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x; c < total; c += gridDim.x) {
        int lo = 0, hi = args.n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (args.prefix[mid] <= c) lo = mid; else hi = mid;
        }
        const SunbPackDesc& D = args.d[lo];
        const float* __restrict__ src = reinterpret_cast<const float*>(D.src) + D.off;
        bf16* __restrict__ dst = reinterpret_cast<bf16*>(D.dst);
        const int D2 = D.dims[2], D3 = D.dims[3], ldd = D.ldd;
        const int v2 = D.valid2 ? D.valid2 : D2;
        const int t3n = (ldd + PK_T3 - 1) / PK_T3, t2n = (D2 + PK_T2 - 1) / PK_T2;
        int t = c - args.prefix[lo];
        const int t3 = t % t3n; t /= t3n;
        const int t2 = t % t2n; t /= t2n;
        const int d1 = t % D.dims[1], d0 = t / D.dims[1];
        const long sbase = (long)d0 * D.strides[0] + (long)d1 * D.strides[1];
        const long s2 = D.strides[2], s3 = D.strides[3];
        const int r0 = t2 * PK_T2, c0 = t3 * PK_T3;
        if ((s3 < 0 ? -s3 : s3) <= (s2 < 0 ? -s2 : s2)) {
            // source denser along dim 3: a warp reads rows of 64
            for (int r = warp; r < PK_T2; r += 8) {
                const int d2 = r0 + r;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int d3 = c0 + lane + 32 * h;
                    tile[r][lane + 32 * h] = (d2 < v2 && d3 < D3) ? src[sbase + d2 * s2 + d3 * s3] : 0.f;
                }
            }
        } else {
            // source denser along dim 2 (transposed operand): a warp reads columns of 32
            for (int cc = warp; cc < PK_T3; cc += 8) {
                const int d3 = c0 + cc, d2 = r0 + lane;
                tile[lane][cc] = (d2 < v2 && d3 < D3) ? src[sbase + d2 * s2 + d3 * s3] : 0.f;
            }
        }
        __syncthreads();
        const size_t obase = ((size_t)(d0 * D.dims[1] + d1) * D2) * ldd;
        for (int r = warp; r < PK_T2; r += 8) {
            const int d2 = r0 + r, d3 = c0 + 2 * lane;
            if (d2 < D2 && d3 < ldd) {
                bf16* o = dst + obase + (size_t)d2 * ldd + d3;
                if (d3 + 1 < ldd && ((reinterpret_cast<size_t>(o) & 3) == 0))
                    *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(tile[r][2 * lane], tile[r][2 * lane + 1]);
                else {
                    o[0] = __float2bfloat16(tile[r][2 * lane]);
                    if (d3 + 1 < ldd) o[1] = __float2bfloat16(tile[r][2 * lane + 1]);
                }
            }
        }
        __syncthreads();
    }
}

// grouped 3x3 weight [256,32,3,3] (8 groups) -> block-diagonal channel pairs [4][9][64][64] (bf16).
// transpose_flip = 0: forward operand  dst[p][tap][n][k] = W[p*64+n][k - 32*(n/32)][tap]   (zero off the diagonal)
// transpose_flip = 1: dgrad operand    dst[p][tap][k_in][n_out] with tap mirrored (8 - tap): conv-transpose weights
__global__ void grouped_pairs_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int transpose_flip) {
    pdl_trigger();
    pdl_wait();
    const int total = 4 * 9 * 64 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int col = i % 64, row = (i / 64) % 64, tap = (i / 4096) % 9, p = i / (4096 * 9);
        const int n = transpose_flip ? col : row;         // output channel within the pair
        const int k = transpose_flip ? row : col;         // input channel within the pair
        float v = 0.f;
        if (n / 32 == k / 32) {
            const int st = transpose_flip ? 8 - tap : tap;
            v = w[((size_t)(p * 64 + n) * 32 + (k % 32)) * 9 + st];
        }
        dst[i] = __float2bfloat16(v);
    }
}

// dW2[256][32][3][3] += diagonal 32x32 blocks of the pair-wise wgrad scratch [2 halves][9][128][128]
__global__ void grouped_wgrad_extract_kernel(const float* __restrict__ scratch, float* __restrict__ dw) {
    pdl_trigger();
    pdl_wait();
    const int total = 256 * 32 * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % 9, ci = (i / 9) % 32, co = i / (9 * 32);
        const int half = co / 128, r = co % 128, grp = r / 32;
        dw[i] += scratch[((size_t)(half * 9 + tap) * 128 + r) * 128 + grp * 32 + ci];
    }
}

// dy[b, t, c] = dpooled[b, c] / T (+ ddense[b, t, c])  -> bf16
__global__ void pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ ddense, bf16* __restrict__ dy,
                                int B, int T, int C) {
    pdl_trigger();
    pdl_wait();
    const long total = (long)B * T * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long b = i / ((long)T * C);
        float v = dpooled ? dpooled[b * C + c] / (float)T : 0.f;
        if (ddense) v += ddense[i];
        dy[i] = __float2bfloat16(v);
    }
}

// stem K=27 weight gradients: dW1[64][27] += sum_p da1[p, c] * patch(x)[p, :],  dWd[128][27] likewise from didn.
// The im2col matrix of the fp32 input image is materialised once as bf16 [B*1600, 32] (columns >= 27 zero; k = (ci*3+ky)*3+kx,
// the same K order as the forward kernel in stem_tc.cu) and the two gradients are plain split-K tcgen05 weight-gradient GEMMs
// over it (wgrad_tc.cu).  The first version kept 27 register accumulators per channel lane and was bound by its
// shared-memory broadcast loads (869 us at 480 images).
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ patches, int total) {
    pdl_trigger();
    pdl_wait();
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < total; m += gridDim.x * blockDim.x) {
        const int img = m / 1600, rem = m - img * 1600, oy = rem / 40, ox = rem - oy * 40;
        const float* xi = x + (size_t)img * 3 * 80 * 80;
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            v[k] = 0.f;
            if (k < 27) {
                const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
                const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                if (iy >= 0 && iy < 80 && ix >= 0 && ix < 80) v[k] = __ldg(xi + (ci * 80 + iy) * 80 + ix);
            }
        }
        store16_bf16(patches + (size_t)m * 32, v);
        store16_bf16(patches + (size_t)m * 32 + 16, v + 16);
    }
}

inline int grid_for(long total, int threads = 256) {
    long b = (total + threads - 1) / threads;
    const long cap = 148L * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// rows per block of the column-statistics pass: 16 rows per thread for large tensors (at most 8 blocks per SM), fewer for
// small ones so that a data-parallel shard's tensors (1.5 - 24 MB) still spread over ~2 blocks per SM with one or two
// load batches per thread instead of a 16-deep serial walk on a few dozen blocks
static void colstats_plan(long M, int C, int* rows_per_block, long* blocks) {
    const int lanes = 256 / (C / 8);
    const long target = 2L * sunb_num_sms();
    long rpt = (M + lanes * target - 1) / (lanes * target);
    rpt = rpt < 2 ? 2 : (rpt > 16 ? 16 : rpt);
    long rpb = lanes * rpt;
    long nb = (M + rpb - 1) / rpb;
    const long cap = 8L * sunb_num_sms();
    if (nb > cap) { rpb = (M + cap - 1) / cap; rpb = (rpb + lanes - 1) / lanes * lanes; nb = (M + rpb - 1) / rpb; }
    *rows_per_block = (int)rpb;
    *blocks = nb;
}

extern "C" {

int sunb_colstats(const void* x, int ldx, const void* u, int ldu, long M, int C, float* sum, float* sq, void* stream) {
    SUNB_REQUIRE(x && sum && sq && M > 0, "colstats: bad arguments");
    SUNB_REQUIRE(C % 8 == 0 && C / 8 <= 256 && ldx % 8 == 0, "colstats: C must be a multiple of 8 and <= 2048 (got %d)", C);
    int rpb;
    long blocks;
    colstats_plan(M, C, &rpb, &blocks);
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    SUNB_CHECK_CUDA(sunb_launch(&colstats_kernel, dim3((int)blocks), dim3(256), 2 * 2048 * sizeof(float), ST(stream), 
        reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(u), ldu, (int)M, C, rpb, sum, sq, fin));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

static int launch_colstats_fin(const void* x, int ldx, const void* u, int ldu, long M, int C, float* sum, float* sq,
                               const BnFin& fin, void* stream) {
    SUNB_REQUIRE(C % 8 == 0 && C / 8 <= 256 && ldx % 8 == 0, "bn_stats: C must be a multiple of 8 and <= 2048 (got %d)", C);
    int rpb;
    long blocks;
    colstats_plan(M, C, &rpb, &blocks);
    SUNB_CHECK_CUDA(sunb_launch(&colstats_kernel, dim3((int)blocks), dim3(256), 2 * 2048 * sizeof(float), ST(stream), 
        reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(u), ldu, (int)M, C, rpb, sum, sq, fin));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

// column statistics + bn_finalize in ONE launch.  sum, sq, ticket must be zero on entry.
int sunb_bn_stats_forward(const void* x, int ldx, long M, int C, float* sum, float* sq, void* ticket, const float* gamma,
                          const float* beta, float* rmean, float* rvar, int64_t* nbt, float momentum, float eps, float* scale,
                          float* shift, float* mean, float* rstd, void* stream) {
    SUNB_REQUIRE(x && sum && sq && ticket && gamma && beta && scale && shift && mean && rstd && M > 0, "bn_stats_forward: bad arguments");
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    fin.mode = 1; fin.ticket = reinterpret_cast<unsigned int*>(ticket); fin.count = (float)M;
    fin.gamma = gamma; fin.beta = beta; fin.rmean = rmean; fin.rvar = rvar; fin.nbt = reinterpret_cast<long long*>(nbt);
    fin.momentum = momentum; fin.eps = eps; fin.scale = scale; fin.shift = shift; fin.mean = mean; fin.rstd = rstd;
    return launch_colstats_fin(x, ldx, nullptr, 0, M, C, sum, sq, fin, stream);
}

// sum(dz), sum(dz*x) + bn_bwd_finalize in ONE launch.  sdz, sdzx, ticket must be zero on entry.
int sunb_bn_stats_backward(const void* dz, int lddz, const void* x, int ldx, long M, int C, float* sdz, float* sdzx, void* ticket,
                           float count, const float* mean, const float* rstd, const float* gamma, int frozen, float* a, float* c1,
                           float* c2, float* dgamma, float* dbeta, void* stream) {
    SUNB_REQUIRE(dz && x && sdz && sdzx && ticket && mean && rstd && gamma && a && c1 && c2 && dgamma && dbeta && M > 0,
                 "bn_stats_backward: bad arguments");
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    fin.mode = 2; fin.ticket = reinterpret_cast<unsigned int*>(ticket); fin.count = count;
    fin.gamma = gamma; fin.mean = const_cast<float*>(mean); fin.rstd = const_cast<float*>(rstd); fin.frozen = frozen;
    fin.a = a; fin.c1 = c1; fin.c2 = c2; fin.dgamma = dgamma; fin.dbeta = dbeta;
    return launch_colstats_fin(dz, lddz, x, ldx, M, C, sdz, sdzx, fin, stream);
}

int sunb_bn_finalize(const float* sum, const float* sq, float count, const float* gamma, const float* beta, float* rmean,
                     float* rvar, int64_t* nbt, float momentum, float eps, int C, float* scale, float* shift, float* mean,
                     float* rstd, void* stream) {
    SUNB_REQUIRE(sum && sq && gamma && beta && scale && shift && mean && rstd && C > 0, "bn_finalize: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, ST(stream), sum, sq, count, gamma, beta, rmean, rvar,
                                                                 reinterpret_cast<long long*>(nbt), momentum, eps, C,
                                                                 scale, shift, mean, rstd));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_bn_apply(const void* x, int ldx, const float* scale, const float* shift, int act, const float* tab, int tab_mod,
                  void* out, int ldo, long M, int C, void* stream) {
    SUNB_REQUIRE(x && out && scale && shift && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "bn_apply: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&bn_apply_kernel, dim3(grid_for(M * (C / 8))), dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(x), ldx, scale, shift, act,
                                                                    tab, tab_mod > 0 ? tab_mod : 1,
                                                                    reinterpret_cast<bf16*>(out), ldo, M, C));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_stem_tail_forward(const void* c3, const void* idn, const float* s3, const float* t3, const float* sd,
                           const float* td, const float* pos, void* out, int B, void* stream) {
    SUNB_REQUIRE(c3 && idn && out && pos, "stem_tail_forward: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&stem_tail_fwd_kernel, dim3(grid_for((long)B * 400 * 16)), dim3(256), 0, ST(stream), 
        reinterpret_cast<const bf16*>(c3), reinterpret_cast<const bf16*>(idn), s3, t3, sd, td, pos,
        reinterpret_cast<bf16*>(out), B, 40, 40, 128));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_stem_tail_backward(const void* c3, const void* idn, const float* s3, const float* t3, const float* sd,
                            const float* td, const void* g, void* dz, int B, void* stream) {
    SUNB_REQUIRE(c3 && idn && g && dz, "stem_tail_backward: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&stem_tail_bwd_kernel, dim3(grid_for((long)B * 400 * 16)), dim3(256), 0, ST(stream), 
        reinterpret_cast<const bf16*>(c3), reinterpret_cast<const bf16*>(idn), s3, t3, sd, td,
        reinterpret_cast<const bf16*>(g), reinterpret_cast<bf16*>(dz), B, 40, 40, 128));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_bn_frozen(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps, int C,
                   float* scale, float* shift, float* mean, float* rstd, void* stream) {
    SUNB_REQUIRE(gamma && beta && rmean && rvar && scale && shift && mean && rstd && C > 0, "bn_frozen: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&bn_frozen_kernel, dim3((C + 127) / 128), dim3(128), 0, ST(stream), gamma, beta, rmean, rvar, eps, C, scale, shift, mean, rstd));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_bn_bwd_finalize(const float* sdz, const float* sdzx, float count, const float* mean, const float* rstd,
                         const float* gamma, int C, int frozen, float* a, float* c1, float* c2, float* dgamma, float* dbeta,
                         void* stream) {
    SUNB_REQUIRE(sdz && sdzx && mean && rstd && gamma && a && c1 && c2 && dgamma && dbeta, "bn_bwd_finalize: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&bn_bwd_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, ST(stream), sdz, sdzx, count, mean, rstd, gamma, C, frozen, a, c1, c2,
                                                                     dgamma, dbeta));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_bn_bwd_apply(const void* dz, int lddz, const void* x, int ldx, const float* a, const float* c1, const float* c2,
                      const float* mean, const void* res, int ldr, void* out, int ldo, long M, int C, void* stream) {
    SUNB_REQUIRE(dz && x && out && C % 8 == 0, "bn_bwd_apply: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&bn_bwd_apply_kernel, dim3(grid_for(M * (C / 8))), dim3(256), 0, ST(stream), 
        reinterpret_cast<const bf16*>(dz), lddz, reinterpret_cast<const bf16*>(x), ldx, a, c1, c2, mean,
        reinterpret_cast<const bf16*>(res), ldr, reinterpret_cast<bf16*>(out), ldo, M, C));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_scale_rows(const void* in, const float* rs, int rows_per_img, void* out, long M, int C, void* stream) {
    SUNB_REQUIRE(in && rs && out && C % 8 == 0 && rows_per_img > 0, "scale_rows: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&scale_rows_kernel, dim3(grid_for(M * (C / 8))), dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(in), rs, rows_per_img,
                                                                      reinterpret_cast<bf16*>(out), M, C));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_s2d_reorder(const void* in, void* out, int B, int H, int W, int C, int dir, void* stream) {
    SUNB_REQUIRE(in && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "s2d_reorder: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&s2d_reorder_kernel, dim3(grid_for((long)B * H * W * (C / 8))), dim3(256), 0, ST(stream), 
        reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), B, H, W, C, dir));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_batch_sum(const void* g, int B, long n, float* out, void* stream) {
    SUNB_REQUIRE(g && out && B > 0 && n > 0, "batch_sum: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&batch_sum_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(g), B, n, out));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_permute_cast(const float* src, long off, long sa, long sb, long sc, int A, int B, int Cd, int ldd, void* dst,
                      void* stream) {
    SUNB_REQUIRE(src && dst && A > 0 && B > 0 && Cd > 0 && ldd >= Cd, "permute_cast: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&permute_cast_kernel, dim3(grid_for((long)A * B * ldd)), dim3(256), 0, ST(stream), src, off, sa, sb, sc, A, B, Cd, ldd,
                                                                              reinterpret_cast<bf16*>(dst)));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_pack_weights(const SunbPackDesc* descs, int n, void* stream) {
    SUNB_REQUIRE(descs && n > 0, "pack_weights: bad arguments");
    for (int t0 = 0; t0 < n; t0 += PK_MAX) {
        PackArgs a;
        a.n = n - t0 < PK_MAX ? n - t0 : PK_MAX;
        a.prefix[0] = 0;
        for (int i = 0; i < a.n; ++i) {
            a.d[i] = descs[t0 + i];
            const SunbPackDesc& D = a.d[i];
            SUNB_REQUIRE(D.src && D.dst && D.dims[0] > 0 && D.dims[1] > 0 && D.dims[2] > 0 && D.dims[3] > 0 && D.ldd >= D.dims[3],
                         "pack_weights: bad entry %d", t0 + i);
            const long tiles = (long)D.dims[0] * D.dims[1] * ((D.dims[2] + PK_T2 - 1) / PK_T2) * ((D.ldd + PK_T3 - 1) / PK_T3);
            SUNB_REQUIRE((long)D.dims[0] * D.dims[1] * D.dims[2] * D.ldd < (1L << 30) && a.prefix[i] + tiles < (1L << 30),
                         "pack_weights: entry %d too large", t0 + i);
            a.prefix[i + 1] = a.prefix[i] + (int)tiles;
        }
        const int total = a.prefix[a.n];
        SUNB_CHECK_CUDA(sunb_launch(&pack_multi_kernel, dim3(total < 148 * 16 ? total : 148 * 16), dim3(256), 0, ST(stream), a));
        SUNB_CHECK_CUDA(cudaGetLastError());
    }
    return SUNB_OK;
}

int sunb_grouped_pairs(const float* w, void* dst, int transpose_flip, void* stream) {
    SUNB_REQUIRE(w && dst, "grouped_pairs: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&grouped_pairs_kernel, dim3(grid_for(4 * 9 * 64 * 64)), dim3(256), 0, ST(stream), w, reinterpret_cast<bf16*>(dst), transpose_flip));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_grouped_wgrad_extract(const float* scratch, float* dw, void* stream) {
    SUNB_REQUIRE(scratch && dw, "grouped_wgrad_extract: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&grouped_wgrad_extract_kernel, dim3(grid_for(256 * 32 * 9)), dim3(256), 0, ST(stream), scratch, dw));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_pool_backward(const float* dpooled, const float* ddense, void* dy, int B, int T, int C, void* stream) {
    SUNB_REQUIRE(dy && (dpooled || ddense), "pool_backward: bad arguments");
    SUNB_CHECK_CUDA(sunb_launch(&pool_bwd_kernel, dim3(grid_for((long)B * T * C)), dim3(256), 0, ST(stream), dpooled, ddense, reinterpret_cast<bf16*>(dy), B, T, C));
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_stem_wgrad(const float* x, const void* da1, const void* didn, float* dw1, float* dwd, int B, void* scratch,
                    void* stream) {
    SUNB_REQUIRE(x && da1 && didn && dw1 && dwd && scratch && B > 0, "stem_wgrad: bad arguments");
    SUNB_REQUIRE((((size_t)scratch) & 31) == 0, "stem_wgrad: scratch must be 32-byte aligned");
    const int P = B * 1600;
    bf16* patches = reinterpret_cast<bf16*>(scratch);
    SUNB_CHECK_CUDA(sunb_launch(&stem_im2col_kernel, dim3(grid_for(P)), dim3(256), 0, ST(stream), x, patches, P));
    SUNB_CHECK_CUDA(cudaGetLastError());
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.P = P; p.Nb = 27; p.Cb = 32; p.groups = 1; p.taps = 1;
    p.X = patches; p.ldx = 32; p.ldo = 27;
    p.Ma = 64; p.Ca = 64; p.dY = reinterpret_cast<const bf16*>(da1); p.ldy = 64; p.out = dw1;
    SUNB_TRY(sunb_launch_wgrad_tc(p, ST(stream)));
    p.Ma = 128; p.Ca = 128; p.dY = reinterpret_cast<const bf16*>(didn); p.ldy = 128; p.out = dwd; p.ksplit = 0;
    SUNB_TRY(sunb_launch_wgrad_tc(p, ST(stream)));
    return SUNB_OK;
}

}  // extern "C"
