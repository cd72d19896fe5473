// SUN local-supervision head kernels (reference: sun_meta_training/offline.py:34-45, 57-76, 296-300).
//   softlabel        : background filtration + top-k pseudo labels from teacher patch logits (integer-exact)
//   soft_ce fwd/bwd  : mean_rows( sum_c -t_c * log_softmax(x)_c ) and its gradient
//   hard_ce bwd      : gradient of mean cross-entropy against int64 labels
// All HBM-bound: one warp per image / per row, coalesced over the class dimension.
#include "common.cuh"

namespace {

constexpr int MAX_CPL = 16;      // classes per lane -> n_cls <= 512 (miniImageNet 64, tieredImageNet 351 base classes)
constexpr int MAX_HW = 64;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per image.  logits element (b, c, p) at logits[b*sb + c*sc + p*sp].
// Ties (torch.topk leaves them unspecified): the lowest index wins, for patches and for classes.
__global__ void __launch_bounds__(128) softlabel_kernel(const float* __restrict__ logits, long sb, long sc, long sp,
                                                        int B, int n_cls, int hw, int k, int bp, float on, float off,
                                                        float* __restrict__ out) {
    __shared__ float s_pm[4][MAX_HW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= B) return;
    float* pm = s_pm[warp];
    const float* lb = logits + (size_t)b * sb;
    const int width = n_cls + 1;
    for (int p = 0; p < hw; ++p) {
        float m = -INFINITY;
        for (int c = lane; c < n_cls; c += 32) m = fmaxf(m, lb[c * sc + p * sp]);
        m = warp_max(m);
        if (lane == 0) pm[p] = m;
    }
    __syncwarp();
    const int n_fg = hw - bp;
    for (int p = 0; p < hw; ++p) {
        // rank of patch p among the per-patch maxima (descending, lowest index first on ties)
        int cnt = 0;
        const float mine = pm[p];
        for (int qi = lane; qi < hw; qi += 32) {
            const float o = pm[qi];
            cnt += (o > mine || (o == mine && qi < p)) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const bool fg = cnt < n_fg;
        float val[MAX_CPL];
        unsigned sel = 0;                      // bit j: class lane + 32*j selected
#pragma unroll
        for (int j = 0; j < MAX_CPL; ++j) {
            const int c = lane + 32 * j;
            val[j] = (c < n_cls) ? lb[c * sc + p * sp] : -INFINITY;
        }
        if (fg) {
            for (int it = 0; it < k; ++it) {
                float best = -INFINITY;
                int bi = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < MAX_CPL; ++j) {
                    const int c = lane + 32 * j;
                    if (c < n_cls && !((sel >> j) & 1u) && (val[j] > best || (val[j] == best && c < bi))) {
                        best = val[j];
                        bi = c;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if ((bi & 31) == lane && bi < n_cls) sel |= 1u << (bi >> 5);
            }
        }
        float* row = out + ((size_t)b * hw + p) * width;
        for (int c = lane; c < width; c += 32) {
            const bool hot = fg ? ((c < n_cls) && ((sel >> (c >> 5)) & 1u)) : (c == 1);   // bg column is 1 (offline.py:62,71)
            row[c] = hot ? on : off;
        }
    }
}

// warp per row: row_loss[r] = sum_c -t[r % Rt][c] * (x[r][c] - logsumexp(x[r]))
__global__ void __launch_bounds__(256) soft_ce_rows_kernel(const float* __restrict__ x, int ldx,
                                                           const float* __restrict__ t, int ldt, int R, int Rt, int C,
                                                           float* __restrict__ row_loss) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + (size_t)r * ldx;
    const float* tr = t + (size_t)(r % Rt) * ldt;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(xr[c] - mx);
    se = warp_sum(se);
    const float lse = logf(se) + mx;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += -tr[c] * (xr[c] - lse);
    acc = warp_sum(acc);
    if (lane == 0) row_loss[r] = acc;
}

// deterministic single-block mean of n floats
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float s[1024];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) a += v[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0] / (float)n;
}

// dx[r][c] = gscale/R * (softmax(x[r])[c] * sum(t[r]) - t[r][c])
__global__ void __launch_bounds__(256) soft_ce_bwd_kernel(const float* __restrict__ x, int ldx,
                                                          const float* __restrict__ t, int ldt, int R, int Rt, int C,
                                                          const float* __restrict__ gout, float gscale,
                                                          float* __restrict__ dx, int lddx) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + (size_t)r * ldx;
    const float* tr = t + (size_t)(r % Rt) * ldt;
    float mx = -INFINITY, ts = 0.f;
    for (int c = lane; c < C; c += 32) { mx = fmaxf(mx, xr[c]); ts += tr[c]; }
    mx = warp_max(mx);
    ts = warp_sum(ts);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(xr[c] - mx);
    se = warp_sum(se);
    const float g = (gout ? gout[0] : 1.f) * gscale / (float)R;
    for (int c = lane; c < C; c += 32) dx[(size_t)r * lddx + c] = g * (expf(xr[c] - mx) / se * ts - tr[c]);
}

// dlogits[r][c] = gscale/R * (softmax(l[r])[c] - [c == label[r]])
__global__ void __launch_bounds__(256) hard_ce_bwd_kernel(const float* __restrict__ l, const long long* __restrict__ label,
                                                          int R, int W, const float* __restrict__ gout, float gscale,
                                                          float* __restrict__ dl) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* lr = l + (size_t)r * W;
    float mx = -INFINITY;
    for (int c = lane; c < W; c += 32) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < W; c += 32) se += expf(lr[c] - mx);
    se = warp_sum(se);
    const float g = (gout ? gout[0] : 1.f) * gscale / (float)R;
    const int y = (int)label[r];
    for (int c = lane; c < W; c += 32) dl[(size_t)r * W + c] = g * (expf(lr[c] - mx) / se - (c == y ? 1.f : 0.f));
}

}  // namespace

int sunb_launch_softlabel(const float* logits, long sb, long sc, long sp, int B, int n_cls, int hw, int k, int bp,
                          double smoothing, float* out, cudaStream_t stream) {
    SUNB_REQUIRE(B > 0 && n_cls > 0 && n_cls <= 32 * MAX_CPL, "softlabel: n_cls must be in [1,512], got %d", n_cls);
    SUNB_REQUIRE(hw > 0 && hw <= MAX_HW && bp >= 0 && bp <= hw && k > 0 && k <= n_cls,
                 "softlabel: bad hw=%d bp=%d k=%d", hw, bp, k);
    const double off = smoothing / (double)n_cls;
    const double on = 1.0 - smoothing + off;
    softlabel_kernel<<<(B + 3) / 4, 128, 0, stream>>>(logits, sb, sc, sp, B, n_cls, hw, k, bp, (float)on, (float)off, out);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_soft_ce_forward(const float* x, int ldx, const float* t, int ldt, int R, int Rt, int C, float* row_loss,
                                float* loss, cudaStream_t stream) {
    SUNB_REQUIRE(R > 0 && Rt > 0 && R % Rt == 0 && C > 0, "soft_ce: rows %d must be a multiple of target rows %d", R, Rt);
    soft_ce_rows_kernel<<<(R + 7) / 8, 256, 0, stream>>>(x, ldx, t, ldt, R, Rt, C, row_loss);
    SUNB_CHECK_CUDA(cudaGetLastError());
    mean_kernel<<<1, 1024, 0, stream>>>(row_loss, R, loss);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_soft_ce_backward(const float* x, int ldx, const float* t, int ldt, int R, int Rt, int C,
                                 const float* gout, float gscale, float* dx, int lddx, cudaStream_t stream) {
    SUNB_REQUIRE(R > 0 && Rt > 0 && R % Rt == 0 && C > 0, "soft_ce_bwd: bad shape");
    soft_ce_bwd_kernel<<<(R + 7) / 8, 256, 0, stream>>>(x, ldx, t, ldt, R, Rt, C, gout, gscale, dx, lddx);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}

int sunb_launch_hard_ce_backward(const float* l, const long long* label, int R, int W, const float* gout, float gscale,
                                 float* dl, cudaStream_t stream) {
    SUNB_REQUIRE(R > 0 && W > 0, "hard_ce_bwd: bad shape");
    hard_ce_bwd_kernel<<<(R + 7) / 8, 256, 0, stream>>>(l, label, R, W, gout, gscale, dl);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
