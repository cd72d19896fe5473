// How fast can 16 warps per SM run the bias + GELU + bf16-pack epilogue when nothing else is in the way?
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I few-shot-vit_b200/csrc -o tools/micro/gelu_bench tools/micro/gelu_bench.cu
#include <cstdio>
#include "common.cuh"
int sunb_opt_in_smem(const void*, int) { return 0; }
int sunb_num_sms() { return 148; }
bool sunb_pdl_allowed() { return false; }
void sunb_pdl_allow(bool) {}
void sunb_set_error(const char*, ...) {}
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(const float* __restrict__ bias, uint4* out, int iters, long long* cyc) {
    __shared__ uint4 sink[WARPS * 32];
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 1e-3f + i * 0.03f - 0.5f;
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + (it & 7) * 32);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 bb = __ldg(b4 + i / 4);
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(gelu_fast(v[i] + bb.x), gelu_fast(v[i + 1] + bb.y));
            const __nv_bfloat162 h1 = __floats2bfloat162_rn(gelu_fast(v[i + 2] + bb.z), gelu_fast(v[i + 3] + bb.w));
            pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h0);
            pk[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc.x ^= pk[4 * j]; acc.y ^= pk[4 * j + 1]; acc.z ^= pk[4 * j + 2]; acc.w ^= pk[4 * j + 3]; }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += 1e-4f * (float)(acc.x & 1);      // keep the inputs loop-carried
    }
    long long t1 = clock64();
    sink[threadIdx.x] = acc;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int WARPS> void run() {
    const int sms = 148, iters = 2000;
    float* bias; uint4* out; long long* cyc;
    cudaMalloc(&bias, 4096); cudaMemset(bias, 0, 4096); cudaMalloc(&out, sms * WARPS * 32 * 16); cudaMalloc(&cyc, sms * 8);
    k<WARPS><<<sms, WARPS * 32>>>(bias, out, iters, cyc); cudaDeviceSynchronize();
    k<WARPS><<<sms, WARPS * 32>>>(bias, out, iters, cyc); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%2d warps/SM: %.2f GELU elements per clock per SM (+1 FFMA per element of loop overhead)\n", WARPS, (double)WARPS * 32 * 32 * iters / avg);
}
int main() { run<4>(); run<8>(); run<16>(); run<32>(); return 0; }
