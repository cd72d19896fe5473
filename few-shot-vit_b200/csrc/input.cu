// On-device input path of the episodic loaders (reference: test_phase/datasets/mini_imagenet.py:50-56, default_transform):
//     PIL image (uint8 84 x 84 x 3)  ->  Resize((88, 88), bilinear)  ->  CenterCrop(80)  ->  ToTensor  ->  Normalize(mean, std)
// as ONE kernel over a uint8 image store resident in HBM, with an optional gather index (the flat batch a CategoriesSampler
// yields), so an episode costs 8 bytes of H2D per image instead of 76.8 KB of fp32 pixels.
// Bit-exact with PIL + torchvision: PIL resamples 8-bit images in fixed point (22 fractional bits), horizontally first, and
// rounds / clips to uint8 after EACH pass (Resample.c); ToTensor divides by 255 and Normalize subtracts / divides in fp32
// (IEEE round-to-nearest ops here, no fast-math).  The coefficient tables are built on the host exactly as PIL's
// precompute_coeffs + normalize_coeffs_8bpc do (sunb200/input.py).
// HBM-bound: 21 KB read + 76.8 KB written per image.
#include "common.cuh"
#include "../../include/sunb200.h"

namespace {

constexpr int PRECISION_BITS = 22;     // 32 - 8 - 2, PIL Resample.c
constexpr int KMAX = 3;                // taps per output position (bilinear up-sampling touches at most 3 source pixels)

__device__ __forceinline__ int clip8(int v) {
    v >>= PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// thread = one output pixel (all three channels).  tab_min[o] = first source index, tab_k[o][KMAX] = fixed-point weights
// (zero padded) of output position o of the CROPPED axis; the same tables serve rows and columns (square images).
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ data, const long long* __restrict__ idx,
                                                            int n, int in_size, int out_size, const int* __restrict__ tab_min,
                                                            const int* __restrict__ tab_k, const float* __restrict__ mean_std,
                                                            float* __restrict__ out) {
    const long long total = (long long)n * out_size * out_size;
    const float m0 = mean_std[0], m1 = mean_std[1], m2 = mean_std[2], s0 = mean_std[3], s1 = mean_std[4], s2 = mean_std[5];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % out_size), y = (int)((i / out_size) % out_size);
        const long long img = i / ((long long)out_size * out_size);
        const long long src_img = idx ? idx[img] : img;
        const uint8_t* src = data + (size_t)src_img * in_size * in_size * 3;
        const int xmin = tab_min[x], ymin = tab_min[y];
        int kx[KMAX], ky[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) { kx[j] = tab_k[x * KMAX + j]; ky[j] = tab_k[y * KMAX + j]; }
        int acc[3] = {1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1)};
#pragma unroll
        for (int r = 0; r < KMAX; ++r) {
            if (ky[r] == 0) continue;                                  // zero-padded tap (also keeps the row index in range)
            const uint8_t* row = src + (size_t)(ymin + r) * in_size * 3;
            int h[3] = {1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1)};
#pragma unroll
            for (int c = 0; c < KMAX; ++c) {
                if (kx[c] == 0) continue;
                const uint8_t* px = row + (xmin + c) * 3;
                h[0] += px[0] * kx[c];
                h[1] += px[1] * kx[c];
                h[2] += px[2] * kx[c];
            }
            // horizontal pass result, rounded and clipped to 8 bits as PIL stores it before the vertical pass
            acc[0] += clip8(h[0]) * ky[r];
            acc[1] += clip8(h[1]) * ky[r];
            acc[2] += clip8(h[2]) * ky[r];
        }
        const size_t plane = (size_t)out_size * out_size;
        float* o = out + (size_t)img * 3 * plane + (size_t)y * out_size + x;
        o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc[0]), 255.f), m0), s0);
        o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc[1]), 255.f), m1), s1);
        o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(acc[2]), 255.f), m2), s2);
    }
}

}  // namespace

extern "C" int sunb_preprocess_u8(const void* data, const int64_t* idx, int n, int in_size, int out_size, const int32_t* tab_min,
                                  const int32_t* tab_k, const float* mean_std, float* out, void* stream) {
    SUNB_REQUIRE(data && tab_min && tab_k && mean_std && out && n > 0 && in_size > 0 && out_size > 0, "preprocess_u8: bad arguments");
    const long long total = (long long)n * out_size * out_size;
    const int blocks = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    preprocess_u8_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint8_t*>(data), reinterpret_cast<const long long*>(idx), n, in_size, out_size, tab_min, tab_k,
        mean_std, out);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
