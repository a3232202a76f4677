#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma2(float* out, int iters) {
    unsigned long long c[16];
    for (int i = 0; i < 16; ++i) c[i] = (unsigned long long)threadIdx.x;
    unsigned long long a = 0x3f8000013f800001ull, b = 0x3f7fff003f7fff00ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(a), "l"(b));
    }
    unsigned long long s = 0; for (int i = 0; i < 16; ++i) s ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000; float ms;
        k_ffma2<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_ffma2<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double f2 = 148.0 * warps * 32 * iters * 32.0;
        printf("ffma2 warps/SM %2d: %.2f ms -> %.1f TFLOP/s (%.0f FMA/clk/SM)\n", warps, ms, 2 * f2 / ms / 1e9, f2 / 148 / (ms * 1e-3) / 1.9e9);
    }
    return 0;
}
