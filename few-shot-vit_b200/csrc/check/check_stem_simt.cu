// Stem entry kernels (reference: test_phase/models/visformer.py:209-210,216,221-223,232 and :234-237,431).
//   stem_in  : fp32 NCHW image -> conv1 3x3 s2 (3->64) + BN1 + LeakyReLU  -> a1  [B,40,40,64]  bf16 NHWC
//                              -> downsample 3x3 s2 (3->128) + BN          -> idn [B,40,40,128] bf16 NHWC
//              K = 27 is too short for a tensor-core tile: CUDA-core kernel, both convolutions share the
//              staged input rows; BN is pre-folded into the fp32 weights/biases by the host packer.
//   pool_pos : 2x2 max-pool of the [B,40,40,128] map + pos_embed1 -> [B,20,20,128]
// TEST-ONLY CUDA-core cross-check of stem_tc.cu (tests/native/libsunb200_check.so): never linked into libsunb200.so.
#include "../common.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

constexpr int IMG = 80, OUT = 40, C1 = 64, CD = 128;

// block = one output row (40 pixels) of one image; 128 threads = 2 pixel halves x 64 channel lanes.
// lane c computes conv1 channel c and downsample channels c, c+64 for 20 pixels.
__global__ void __launch_bounds__(128) stem_in_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ wd,
                                                      const float* __restrict__ bd, bf16* __restrict__ a1,
                                                      bf16* __restrict__ idn, int B, int lrelu) {
    __shared__ float in[3][3][IMG + 2];     // [channel][input row ky][x + 1], zero padded
    const int img = blockIdx.x / OUT, oy = blockIdx.x % OUT;
    for (int i = threadIdx.x; i < 3 * 3 * (IMG + 2); i += blockDim.x) {
        const int c = i / (3 * (IMG + 2)), r = (i / (IMG + 2)) % 3, xx = i % (IMG + 2);
        const int iy = 2 * oy - 1 + r, ix = xx - 1;
        float v = 0.f;
        if (iy >= 0 && iy < IMG && ix >= 0 && ix < IMG) v = x[((size_t)(img * 3 + c) * IMG + iy) * IMG + ix];
        in[c][r][xx] = v;
    }
    const int c = threadIdx.x & 63, half = threadIdx.x >> 6;
    float k1[27], kd0[27], kd1[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) {
        k1[i] = w1[c * 27 + i];
        kd0[i] = wd[c * 27 + i];
        kd1[i] = wd[(c + 64) * 27 + i];
    }
    const float bias1 = b1[c], biasd0 = bd[c], biasd1 = bd[c + 64];
    __syncthreads();
    for (int ox = half * 20; ox < half * 20 + 20; ++ox) {
        float s1 = bias1, s2 = biasd0, s3 = biasd1;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = in[ci][ky][2 * ox + kx];      // (2*ox - 1 + kx) + 1
                    const int wi = (ci * 3 + ky) * 3 + kx;
                    s1 = fmaf(v, k1[wi], s1);
                    s2 = fmaf(v, kd0[wi], s2);
                    s3 = fmaf(v, kd1[wi], s3);
                }
        if (lrelu) s1 = s1 > 0.f ? s1 : 0.1f * s1;
        const size_t px = (size_t)(img * OUT + oy) * OUT + ox;
        a1[px * C1 + c] = __float2bfloat16(s1);
        idn[px * CD + c] = __float2bfloat16(s2);
        idn[px * CD + c + 64] = __float2bfloat16(s3);
    }
}

}  // namespace

extern "C" int sunb_check_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, void* a1,
                                  void* idn, int B, int lrelu, void* stream) {
    SUNB_REQUIRE(x && w1 && b1 && wd && bd && a1 && idn && B > 0, "check_stem_in: bad arguments");
    stem_in_kernel<<<B * OUT, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w1, b1, wd, bd, reinterpret_cast<bf16*>(a1),
                                                                                  reinterpret_cast<bf16*>(idn), B, lrelu);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
