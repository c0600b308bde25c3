// accuracy of lg2.approx.ftz.f32 over the whole normal range against float64 log2 (is the exponent / mantissa
// split of power_to_dB needed?).  build: nvcc -arch=sm_100a -o /tmp/lg2_accuracy tools/exp/lg2_accuracy.cu
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cstring>
__global__ void k(const float* x, float* y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float l; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x[i])); y[i] = l; }
}
int main() {
    const int n = 1 << 24;
    std::vector<float> x(n), y(n);
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint32_t b = (uint32_t)(s >> 32);
        b = (b & 0x007FFFFFu) | ((1u + (b >> 23) % 253u) << 23);      // positive normal floats, every exponent
        float f; memcpy(&f, &b, 4); x[i] = f;
    }
    float *dx, *dy;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dy, n);
    cudaMemcpy(y.data(), dy, n * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst_near1 = 0, sum = 0; float wx = 0;
    for (int i = 0; i < n; ++i) {
        const double t = std::log2((double)x[i]);
        const double e = std::fabs((double)y[i] - t);
        // the float result itself cannot be closer than half an ulp of |t|
        const double e_ulp = e / (std::fabs(t) * 5.96e-8 + 1e-300);
        if (e > worst) { worst = e; wx = x[i]; }
        if (std::fabs(t) < 1.0 && e > worst_near1) worst_near1 = e;
        sum += (double)y[i] - t;
        (void)e_ulp;
    }
    printf("max abs error of lg2.approx.ftz %.3e (at x = %g, log2 = %g); for |log2| < 1: %.3e; mean signed error %.3e\n",
           worst, wx, std::log2((double)wx), worst_near1, sum / n);
    printf("in dB (x 3.0103): %.3e\n", worst * 3.0103);
    return 0;
}
