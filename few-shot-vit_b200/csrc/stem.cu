// Stem tail kernel (reference: test_phase/models/visformer.py:234-237,431).
//   pool_pos : 2x2 max-pool of the [B,40,40,128] map + pos_embed1 -> [B,20,20,128]
// The stem entry convolutions live in stem_tc.cu (tcgen05); their CUDA-core cross-check kernel is test-only (check/).
#include "common.cuh"


namespace {

// thread = 8 channels of one pooled pixel
__global__ void pool_pos_kernel(const bf16* __restrict__ in, const float* __restrict__ pos, bf16* __restrict__ out,
                                int B, int H, int W, int C) {
    const int oH = H / 2, oW = W / 2, cv = C / 8;
    const long total = (long)B * oH * oW * cv;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        const long px = i / cv;
        const int ox = (int)(px % oW), oy = (int)((px / oW) % oH), img = (int)(px / ((long)oW * oH));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const uint4 u = *reinterpret_cast<const uint4*>(
                    in + ((size_t)(img * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c8 * 8);
                const bf16* h = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __bfloat162float(h[j]));
            }
        const float* pp = pos + (size_t)(oy * oW + ox) * C + c8 * 8;
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(m[2 * j] + pp[2 * j], m[2 * j + 1] + pp[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + (size_t)px * C + c8 * 8) = o;
    }
}

}  // namespace

int sunb_launch_stem_in_tc(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, bf16* a1,
                           bf16* idn, int B, int lrelu, cudaStream_t stream);

int sunb_launch_stem_in(const float* x, const float* w1, const float* b1, const float* wd, const float* bd, bf16* a1,
                        bf16* idn, int B, int lrelu, cudaStream_t stream) {
    return sunb_launch_stem_in_tc(x, w1, b1, wd, bd, a1, idn, B, lrelu, stream);
}

int sunb_launch_pool_pos(const bf16* in, const float* pos, bf16* out, int B, int H, int W, int C, cudaStream_t stream) {
    SUNB_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "pool_pos: bad shape");
    const long total = (long)B * (H / 2) * (W / 2) * (C / 8);
    const int blocks = (int)min((total + 255) / 256, (long)148 * 16);
    pool_pos_kernel<<<blocks, 256, 0, stream>>>(in, pos, out, B, H, W, C);
    SUNB_CHECK_CUDA(cudaGetLastError());
    return SUNB_OK;
}
