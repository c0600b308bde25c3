#include <cuda_runtime.h>
__device__ __forceinline__ float2 cmul_packed(float2 a, float2 w) {
    float2 d;
    asm("{ .reg .b64 pa, pas, cc, sm, p; mov.b64 pa, {%2, %3}; mov.b64 pas, {%3, %2}; mov.b64 cc, {%4, %4}; "
        "neg.f32 %0, %5; mov.b64 sm, {%0, %5}; mul.rn.f32x2 p, pa, cc; fma.rn.f32x2 p, pas, sm, p; mov.b64 {%0, %1}, p; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(w.x), "f"(w.y));
    return d;
}
__global__ void k(const float2* a, const float2* w, float2* o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float2 x = a[i], y = a[i + n], t = w[i], u = w[i + n];
        float2 r = cmul_packed(x, t), s = cmul_packed(y, u);
        // keep them in packed flow
        float2 z;
        asm("{ .reg .b64 p, q; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %5}; add.rn.f32x2 p, p, q; mov.b64 {%0, %1}, p; }"
            : "=f"(z.x), "=f"(z.y) : "f"(r.x), "f"(r.y), "f"(s.x), "f"(s.y));
        o[i] = z;
    }
}
