#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mma(float* out, int iters) {
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
    float c[32]; for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 1e-9f;
    float a = 1.0001f, b = 0.9999f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0; for (int i = 0; i < 32; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000; float ms;
        k_mma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_mma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * warps * iters * 8.0 * 1024.0;   // m16n8k8 = 1024 MACs
        printf("mma.sync tf32 warps/SM %2d: %.2f ms -> %.1f TFLOP/s (%.0f MAC/clk/SM @1.9GHz)\n", warps, ms, 2 * fma / ms / 1e9, fma / 148 / (ms * 1e-3) / 1.9e9);
        k_ffma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_ffma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double f2 = 148.0 * warps * 32 * iters * 32.0;
        printf("ffma          warps/SM %2d: %.2f ms -> %.1f TFLOP/s (%.0f FMA/clk/SM)\n", warps, ms, 2 * f2 / ms / 1e9, f2 / 148 / (ms * 1e-3) / 1.9e9);
    }
    return 0;
}
