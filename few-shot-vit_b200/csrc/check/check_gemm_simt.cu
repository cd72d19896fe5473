// Plain shared-memory-tiled SIMT GEMM / implicit GEMM with the same semantics and epilogue as gemm_tc.cu:
// an on-device cross-check for the tcgen05 kernel, called by the -m gpu tests through sunb_check_gemm.
// TEST-ONLY (tests/native/libsunb200_check.so): never linked into libsunb200.so.
#include "../gemm_desc.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    const int g = blockIdx.z;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4] = {};
    const int Ktot = p.taps * p.K;
    for (int k0 = 0; k0 < Ktot; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            const int r = i / TK, kk = i % TK;
            const int m = m0 + r, k = k0 + kk;
            float a = 0.f;
            if (m < p.M && k < Ktot) {
                const int tap = k / p.K, kc = k % p.K;
                if (p.a_mode == 0) {
                    a = __bfloat162float(p.A[(size_t)m * p.lda + g * p.a_goff + kc]);
                } else {
                    const int hw = p.H * p.W;
                    const int img = m / hw, rem = m % hw;
                    const int y = rem / p.W + tap / 3 - 1, x = rem % p.W + tap % 3 - 1;
                    if (y >= 0 && y < p.H && x >= 0 && x < p.W)
                        a = __bfloat162float(p.A[((size_t)(img * p.H + y) * p.W + x) * p.lda + g * p.a_goff + kc]);
                }
            }
            As[kk][r] = a;
        }
        for (int i = threadIdx.x; i < TN * TK; i += 256) {
            const int c = i / TK, kk = i % TK;
            const int n = n0 + c, k = k0 + kk;
            float b = 0.f;
            if (n < p.N && k < Ktot) {
                const int tap = k / p.K, kc = k % p.K;
                b = __bfloat162float(p.Wt[(size_t)((g * p.taps + tap) * p.N + n) * p.ldw + kc]);
            }
            Bs[kk][c] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) epilogue_row<4>(p, g, m0 + ty * 4 + i, n0 + tx * 4, acc[i]);
}

}  // namespace

extern "C" int sunb_check_gemm(const SunbGemmDesc* d, void* stream) {
    SUNB_REQUIRE(d != nullptr, "check_gemm: null descriptor");
    const GemmParams p = sunb_desc_to_params(d);
    SUNB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_simt: empty problem");
    dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, p.groups);
    gemm_simt_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
