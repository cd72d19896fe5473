// Throughput of the special-function unit on this GPU: tanh.approx, ex2.approx, rcp.approx vs FFMA, per SM per clock.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/mufu_bench tools/micro/mufu_bench.cu && tools/micro/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i * 0.01f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
            if (OP == 4) asm volatile("{.reg .b32 t; tanh.approx.f16x2 t, %0; mov.b32 %0, t;}" : "+r"(*reinterpret_cast<unsigned*>(&v[i])));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, iters = 4096;
    float* out; long long* cyc; cudaMalloc(&out, sms * threads * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<OP><<<sms, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%-22s %.2f thread-ops per clock per SM\n", name, (double)threads * iters * 8 / avg);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("tanh.approx.f32"); run<1>("ex2.approx.ftz.f32"); run<2>("rcp.approx.ftz.f32"); run<3>("fma.rn.f32"); run<4>("tanh.approx.f16x2 (x2)");
    return 0;
}
