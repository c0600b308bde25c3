// microbenchmark: bandwidth of column-owner streaming over a row-major (rows x cols) fp32 matrix
// for different CTA widths / unrolls / per-element work.  nvcc -O3 -arch=sm_100a exp_colstream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

template <int BX, int U, int WORK>
__global__ void __launch_bounds__(BX) k_stream(const float* __restrict__ p, long long cols, long long rows,
                                               long long rps, unsigned* out) {
    const long long col = (long long)blockIdx.x * BX + threadIdx.x;
    if (col >= cols) return;
    const long long i0 = (long long)blockIdx.y * rps;
    const long long i1 = min(rows, i0 + rps);
    const float* src = p + i0 * cols + col;
    unsigned acc[4] = {0, 0, 0, 0};
    const unsigned b0 = 0x3a000000u + threadIdx.x, b1 = 0x3b000000u, b2 = 0x3c000000u, b3 = 0x3d000000u;
    long long i = i0;
#pragma unroll 1
    for (; i + U <= i1; i += U) {
        float f[U];
#pragma unroll
        for (int u = 0; u < U; ++u, src += cols) f[u] = __ldcs(src);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = __float_as_uint(f[u]);
            acc[0] += v >= b0;
            if (WORK >= 2) acc[1] += v >= b1;
            if (WORK >= 4) { acc[2] += v >= b2; acc[3] += v >= b3; }
            if (WORK >= 8) { acc[0] += v <= b1; acc[1] += v <= b2; acc[2] += v <= b3; acc[3] += v <= b0; }
        }
    }
    atomicAdd(out + (col & 1023), acc[0] + acc[1] + acc[2] + acc[3]);
}

// variant: private candidate list appends for a fraction of the elements (values in [0, 0.01): v < thr)
template <int BX, int U, int MODE>
__global__ void __launch_bounds__(BX) k_store(const float* __restrict__ p, long long cols, long long rows,
                                              long long rps, unsigned* out, unsigned* lists, unsigned cap, float thr) {
    const long long col = (long long)blockIdx.x * BX + threadIdx.x;
    if (col >= cols) return;
    const long long i0 = (long long)blockIdx.y * rps;
    const long long i1 = min(rows, i0 + rps);
    const float* src = p + i0 * cols + col;
    unsigned* list = lists + ((long long)blockIdx.y * cols + col) * cap;
    unsigned nl = 0, acc = 0;
    const unsigned tb = __float_as_uint(thr);
    long long i = i0;
#pragma unroll 1
    for (; i + U <= i1; i += U) {
        float f[U];
#pragma unroll
        for (int u = 0; u < U; ++u, src += cols) f[u] = __ldcs(src);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = __float_as_uint(f[u]);
            const bool in = v < tb;
            acc += v >= 0x3a000000u;
            if (MODE == 0) { if (in) { if (nl < cap) list[nl] = v; ++nl; } }
            if (MODE == 1) { if (in) { if (nl < cap) __stcs(list + nl, v); ++nl; } }
            if (MODE == 2) { if (in) ++nl; }      // count only
            if (MODE == 3) { if (in) ++nl; }
        }
    }
    atomicAdd(out + (col & 1023), acc + nl);
}

// variant: count + the warp writes one full 128-byte line every `period` rows (what drained staging buffers cost)
template <int BX, int U>
__global__ void __launch_bounds__(BX) k_line(const float* __restrict__ p, long long cols, long long rows,
                                             long long rps, unsigned* out, unsigned* lists, int period) {
    const long long col = (long long)blockIdx.x * BX + threadIdx.x;
    if (col >= cols) return;
    const long long i0 = (long long)blockIdx.y * rps;
    const long long i1 = min(rows, i0 + rps);
    const float* src = p + i0 * cols + col;
    unsigned* line = lists + ((long long)blockIdx.y * cols + (col & ~31ll)) * 64 + (threadIdx.x & 31);
    unsigned acc = 0; int k = 0;
    long long i = i0;
#pragma unroll 1
    for (; i + U <= i1; i += U) {
        float f[U];
#pragma unroll
        for (int u = 0; u < U; ++u, src += cols) f[u] = __ldcs(src);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __float_as_uint(f[u]) >= 0x3a000000u;
        k += U;
        if (k >= period) { k -= period; *line = acc; line += 32; }
    }
    atomicAdd(out + (col & 1023), acc);
}

// variant: each thread owns 4 adjacent columns (128-bit loads)
template <int BX, int U>
__global__ void __launch_bounds__(BX) k_stream4(const float4* __restrict__ p, long long cols4, long long rows,
                                                long long rps, unsigned* out) {
    const long long col = (long long)blockIdx.x * BX + threadIdx.x;
    if (col >= cols4) return;
    const long long i0 = (long long)blockIdx.y * rps;
    const long long i1 = min(rows, i0 + rps);
    const float4* src = p + i0 * cols4 + col;
    unsigned acc[4] = {0, 0, 0, 0};
    const unsigned b0 = 0x3a000000u + threadIdx.x;
    long long i = i0;
#pragma unroll 1
    for (; i + U <= i1; i += U) {
        float4 f[U];
#pragma unroll
        for (int u = 0; u < U; ++u, src += cols4) f[u] = __ldcs(src);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc[0] += __float_as_uint(f[u].x) >= b0; acc[1] += __float_as_uint(f[u].y) >= b0;
            acc[2] += __float_as_uint(f[u].z) >= b0; acc[3] += __float_as_uint(f[u].w) >= b0;
        }
    }
    atomicAdd(out + (col & 1023), acc[0] + acc[1] + acc[2] + acc[3]);
}

__global__ void fill(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = (h >> 8) * (1.0f / 16777216.0f) * 0.01f;
    }
}

template <typename F>
float time_it(F&& f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

template <int BX, int U, int WORK>
void run(const float* p, long long cols, long long rows, long long ctas_target, unsigned* out) {
    const long long tiles = (cols + BX - 1) / BX;
    long long splits = ctas_target / tiles; if (splits < 1) splits = 1;
    long long rps = (rows + splits - 1) / splits; splits = (rows + rps - 1) / rps;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    float ms = time_it([&] { k_stream<BX, U, WORK><<<grid, BX>>>(p, cols, rows, rps, out); });
    printf("BX=%4d U=%2d WORK=%d ctas=%6lld rps=%6lld : %.3f ms  %.0f GB/s\n", BX, U, WORK, tiles * splits, rps, ms,
           rows * cols * 4.0 / ms / 1e6);
}
template <int BX, int U, int MODE>
void runs(const float* p, long long cols, long long rows, long long ctas_target, unsigned* out, unsigned* lists, float frac) {
    const long long tiles = (cols + BX - 1) / BX;
    long long splits = ctas_target / tiles; if (splits < 1) splits = 1;
    long long rps = (rows + splits - 1) / splits; splits = (rows + rps - 1) / rps;
    unsigned cap = (unsigned)(rps * frac * 1.5 + 32);
    dim3 grid((unsigned)tiles, (unsigned)splits);
    float ms = time_it([&] { k_store<BX, U, MODE><<<grid, BX>>>(p, cols, rows, rps, out, lists, cap, 0.01f * frac); });
    printf("STORE mode=%d frac=%.3f BX=%4d U=%2d ctas=%6lld rps=%6lld cap=%u: %.3f ms  %.0f GB/s\n", MODE, frac, BX, U, tiles * splits, rps, cap, ms,
           rows * cols * 4.0 / ms / 1e6);
}
template <int BX, int U>
void run4(const float* p, long long cols, long long rows, long long ctas_target, unsigned* out) {
    const long long cols4 = cols / 4, tiles = (cols4 + BX - 1) / BX;
    long long splits = ctas_target / tiles; if (splits < 1) splits = 1;
    long long rps = (rows + splits - 1) / splits; splits = (rows + rps - 1) / rps;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    float ms = time_it([&] { k_stream4<BX, U><<<grid, BX>>>((const float4*)p, cols4, rows, rps, out); });
    printf("V4 BX=%4d U=%2d ctas=%6lld rps=%6lld : %.3f ms  %.0f GB/s\n", BX, U, tiles * splits, rps, ms,
           rows * cols * 4.0 / ms / 1e6);
}

int main() {
    const long long cols = 4096, rows = 488280;   // 8 GB
    float* p; unsigned* out;
    cudaMalloc(&p, rows * cols * 4); cudaMalloc(&out, 4096);
    fill<<<2048, 256>>>(p, rows * cols); cudaDeviceSynchronize();
    for (long long ctas : {148ll * 64}) {
        run<128, 8, 1>(p, cols, rows, ctas, out);
        run<128, 16, 1>(p, cols, rows, ctas, out);
        run<256, 8, 1>(p, cols, rows, ctas, out);
        run<256, 16, 1>(p, cols, rows, ctas, out);
        run<512, 8, 1>(p, cols, rows, ctas, out);
        run<1024, 8, 1>(p, cols, rows, ctas, out);
        run<1024, 16, 1>(p, cols, rows, ctas, out);
        run4<128, 4>(p, cols, rows, ctas, out);
        run4<128, 8>(p, cols, rows, ctas, out);
        run4<256, 8>(p, cols, rows, ctas, out);
        run4<1024, 4>(p, cols, rows, ctas, out);
    }
    unsigned* lists; cudaMalloc(&lists, (size_t)rows * cols * 4);
    for (float frac : {0.067f}) {
        runs<128, 8, 0>(p, cols, rows, 148 * 64, out, lists, frac);
        runs<128, 8, 1>(p, cols, rows, 148 * 64, out, lists, frac);
        runs<128, 8, 2>(p, cols, rows, 148 * 64, out, lists, frac);
        runs<128, 8, 0>(p, cols, rows, 148 * 16, out, lists, frac);
        runs<1024, 8, 0>(p, cols, rows, 148 * 64, out, lists, frac);
    }
    for (int period : {1 << 30, 64, 32, 16, 8}) {
        const long long tiles = cols / 128; long long splits = 148 * 64 / tiles; long long rps = (rows + splits - 1) / splits; splits = (rows + rps - 1) / rps;
        dim3 grid((unsigned)tiles, (unsigned)splits);
        float ms = time_it([&] { k_line<128, 8><<<grid, 128>>>(p, cols, rows, rps, out, lists, period); });
        printf("LINE period=%d rps=%lld: %.3f ms  %.0f GB/s read, write %.2f GB\n", period, rps, ms, rows * cols * 4.0 / ms / 1e6, rows * cols * 4.0 / period / 1e9);
    }
    for (int w : {2, 4, 8}) {
        if (w == 2) { run<128, 8, 2>(p, cols, rows, 148 * 64, out); run<1024, 8, 2>(p, cols, rows, 148 * 64, out); }
        if (w == 4) { run<128, 8, 4>(p, cols, rows, 148 * 64, out); run<1024, 8, 4>(p, cols, rows, 148 * 64, out); }
        if (w == 8) { run<128, 8, 8>(p, cols, rows, 148 * 64, out); run<1024, 8, 8>(p, cols, rows, 148 * 64, out); }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
