// iqw_elementwise.cu -- the public elementwise power transforms of the reference
// (/root/reference/src/iqwaveform/power_analysis.py:168-206 powtodB, 209-231 dBtopow, 234-257
// envtopow, 260-298 envtodB) as one streaming kernel family: float32 or complex64 in, float32 out,
// 128-bit loads and stores, one pass.  Bound: HBM, 8 B (real) / 12 B (complex) per element.
// The same expressions are the epilogues of kernel 1 (|X|^2, dB); here they stand alone so that the
// `power_analysis` surface of the drop-in is complete.
#include "iqw_common.cuh"

namespace iqw {

enum EwOp : int { EW_POWTODB = 0, EW_DBTOPOW = 1, EW_ENVTOPOW = 2, EW_ENVTODB = 3 };

// 10*log10(v) with the argument already formed (v = |x| + eps or x + eps); negative -> NaN,
// zero -> -inf, exactly like log10
__device__ __forceinline__ float dB_of(float v) {
    const uint32_t b = __float_as_uint(v);
    if (b - 0x00800000u < 0x7F000000u) {
        const float e = __uint_as_float(0x4B400000u | (b >> 23)) - 12583039.0f;
        const float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(m));
        return (e + l) * 3.01029995663981195f;
    }
    return power_to_dB_slow(v);
}

template <int OP, bool ABS>
__device__ __forceinline__ float ew_real(float x, float eps) {
    if (OP == EW_POWTODB) return dB_of((ABS ? fabsf(x) : x) + eps);
    if (OP == EW_ENVTODB) return 2.0f * dB_of((ABS ? fabsf(x) : x) + eps);
    if (OP == EW_DBTOPOW) return exp10f(x / 10.0f);
    return x * x;   // EW_ENVTOPOW: abs(x)**2
}

template <int OP, bool ABS>
__global__ void __launch_bounds__(256) ew_real_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       long long n, float eps, int vec_ok) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    long long done = 0;
    if (vec_ok) {
        const long long n4 = n / 4;
        const float4* in4 = reinterpret_cast<const float4*>(in);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (long long i = tid; i < n4; i += nth) {
            const float4 v = __ldcs(in4 + i);
            float4 r;
            r.x = ew_real<OP, ABS>(v.x, eps); r.y = ew_real<OP, ABS>(v.y, eps);
            r.z = ew_real<OP, ABS>(v.z, eps); r.w = ew_real<OP, ABS>(v.w, eps);
            __stcs(out4 + i, r);
        }
        done = n4 * 4;
    }
    for (long long i = done + tid; i < n; i += nth) out[i] = ew_real<OP, ABS>(in[i], eps);
}

template <int OP>
__global__ void __launch_bounds__(256) ew_complex_kernel(const float2* __restrict__ in, float* __restrict__ out,
                                                          long long n, float eps, int vec_ok) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    auto f = [&](float2 z) {
        const float p = z.x * z.x + z.y * z.y;
        if (OP == EW_ENVTOPOW) return p;
        // envtodB: 20*log10(|z| + eps); with eps == 0 this is 10*log10(|z|^2)
        return eps == 0.0f ? dB_of(p) : 2.0f * dB_of(sqrtf(p) + eps);
    };
    long long done = 0;
    if (vec_ok) {
        const long long n2 = n / 2;
        const float4* in4 = reinterpret_cast<const float4*>(in);
        float2* out2 = reinterpret_cast<float2*>(out);
        for (long long i = tid; i < n2; i += nth) {
            const float4 v = __ldcs(in4 + i);
            __stcs(out2 + i, make_float2(f(make_float2(v.x, v.y)), f(make_float2(v.z, v.w))));
        }
        done = n2 * 2;
    }
    for (long long i = done + tid; i < n; i += nth) out[i] = f(in[i]);
}

static int ew_grid(long long n, int* grid) {
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    long long g = (n / 4 + 255) / 256;
    if (g > (long long)sms * 16) g = (long long)sms * 16;
    if (g < 1) g = 1;
    *grid = (int)g;
    return IQW_OK;
}

}  // namespace iqw

using namespace iqw;

extern "C" int iqw_elementwise_f32(int32_t op, const float* d_in, float* d_out, int64_t n, int32_t use_abs,
                                   float eps, void* stream) {
    iqw::DeviceGuard _dev_guard(d_in);
    if (n < 0) return fail(IQW_ERR_INVALID, "negative size");
    if (n == 0) return IQW_OK;
    if (!d_in || !d_out) return fail(IQW_ERR_INVALID, "null pointer argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int grid = 1;
    if (int rc = ew_grid(n, &grid)) return rc;
    const int vec_ok = (((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0;
    IQW_PROFILE("elementwise", s);
    switch (op) {
        case EW_POWTODB:
            if (use_abs) ew_real_kernel<EW_POWTODB, true><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok);
            else ew_real_kernel<EW_POWTODB, false><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok);
            break;
        case EW_ENVTODB:
            if (use_abs) ew_real_kernel<EW_ENVTODB, true><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok);
            else ew_real_kernel<EW_ENVTODB, false><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok);
            break;
        case EW_DBTOPOW: ew_real_kernel<EW_DBTOPOW, true><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok); break;
        case EW_ENVTOPOW: ew_real_kernel<EW_ENVTOPOW, true><<<grid, 256, 0, s>>>(d_in, d_out, n, eps, vec_ok); break;
        default: return fail(IQW_ERR_INVALID, "unknown elementwise op %d", op);
    }
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_elementwise_c64(int32_t op, const void* d_in, float* d_out, int64_t n, float eps, void* stream) {
    iqw::DeviceGuard _dev_guard(d_in);
    if (n < 0) return fail(IQW_ERR_INVALID, "negative size");
    if (n == 0) return IQW_OK;
    if (!d_in || !d_out) return fail(IQW_ERR_INVALID, "null pointer argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int grid = 1;
    if (int rc = ew_grid(n, &grid)) return rc;
    const int vec_ok = ((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 7) == 0;
    const float2* in = static_cast<const float2*>(d_in);
    IQW_PROFILE("elementwise", s);
    switch (op) {
        case EW_ENVTOPOW: ew_complex_kernel<EW_ENVTOPOW><<<grid, 256, 0, s>>>(in, d_out, n, eps, vec_ok); break;
        case EW_ENVTODB: ew_complex_kernel<EW_ENVTODB><<<grid, 256, 0, s>>>(in, d_out, n, eps, vec_ok); break;
        default: return fail(IQW_ERR_INVALID, "op %d is not defined for complex input", op);
    }
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

// ---------------------------------------------------------------------------------------------
// layout: captures whose time axis is not the last one -- (N, C) with axis = 0, the layout the
// reference's noverlap = 0 path accepts (util.py:400-442 to_blocks) -- are brought to the (C, N)
// channel layout of the kernels by a tiled transpose through shared memory (8-byte elements,
// 32 x 32 tiles, both sides coalesced) instead of a framework copy kernel on the data path.
//   in  (batch, rows, cols) complex64 contiguous  ->  out (batch, cols, rows)
// Bound: HBM, 16 B per sample.
// ---------------------------------------------------------------------------------------------
namespace iqw {
__global__ void __launch_bounds__(256) transpose_c64_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                             long long rows, long long cols) {
    __shared__ float2 tile[32][33];
    const long long b = blockIdx.z;
    const float2* src = in + b * rows * cols;
    float2* dst = out + b * rows * cols;
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;   // the long axis on grid.x
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long long r = r0 + i, c = c0 + tx;
        if (r < rows && c < cols) tile[i][tx] = __ldcs(src + r * cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long long c = c0 + i, r = r0 + tx;
        if (r < rows && c < cols) __stcs(dst + c * rows + r, tile[tx][i]);
    }
}
}  // namespace iqw

extern "C" int iqw_transpose_c64(const void* d_in, int64_t batch, int64_t rows, int64_t cols, void* d_out, void* stream) {
    iqw::DeviceGuard _dev_guard(d_in);
    if (!d_in || !d_out) return fail(IQW_ERR_INVALID, "null pointer argument");
    if (batch < 0 || rows < 0 || cols < 0) return fail(IQW_ERR_INVALID, "negative size");
    if (batch == 0 || rows == 0 || cols == 0) return IQW_OK;
    const long long gx = (rows + 31) / 32, gy = (cols + 31) / 32;
    if (gx > 0x7FFFFFFFll || gy > 65535 || batch > 65535)
        return fail(IQW_ERR_UNSUPPORTED, "iqw_transpose_c64: more than 65535 column tiles or batches");
    transpose_c64_kernel<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(d_in), static_cast<float2*>(d_out), rows, cols);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}
