// microbenchmark: issue/throughput of scalar vs packed (f32x2) FP32 add / fma on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0, float b0) {
    float x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a0 + threadIdx.x * 1e-3f + i; y[i] = b0 + i; }
    const float c = b0 * 0.5f, d = a0 * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i] = x[i] + c; y[i] = y[i] + d; }                           // 2 FADD
            if (MODE == 1) { x[i] = fmaf(x[i], c, d); y[i] = fmaf(y[i], d, c); }           // 2 FFMA
            if (MODE == 2) asm volatile("{ .reg .b64 p, q; mov.b64 p, {%0, %1}; mov.b64 q, {%2, %3}; add.rn.f32x2 p, p, q; mov.b64 {%0, %1}, p; }"
                                        : "+f"(x[i]), "+f"(y[i]) : "f"(c), "f"(d));          // 1 FADD2
            if (MODE == 3) asm volatile("{ .reg .b64 p, q, r; mov.b64 p, {%0, %1}; mov.b64 q, {%2, %3}; mov.b64 r, {%3, %2}; fma.rn.f32x2 p, p, q, r; mov.b64 {%0, %1}, p; }"
                                        : "+f"(x[i]), "+f"(y[i]) : "f"(c), "f"(d));          // 1 FFMA2
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out) {
    const int iters = 4096, grid = 148 * 8;
    k<MODE><<<grid, 256>>>(out, 16, 1.f, 2.f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<grid, 256>>>(out, iters, 1.f, 2.f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lane_ops = (double)grid * 256 * iters * 16;   // scalar-equivalent fp ops (add or fma) per lane
    printf("%-8s %.3f ms  %.2f T scalar-equivalent ops/s  (%.1f per clk per SM at 1.9 GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FADD", out); run<1>("FFMA", out); run<2>("FADD2", out); run<3>("FFMA2", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
