// iqw_binpower.cu -- kernel 3: envelope power statistics of contiguous time bins.
//
// Replaces /root/reference/src/iqwaveform/power_analysis.py:380-385: to_blocks (reshape into
// (n_bins, bin_len)), envtopow (|x|^2 materialised for the whole capture) and the mean / max / min
// detector over each bin, in one streaming pass that reads every sample once (8 B/sample) and
// writes 4 B per bin and statistic.
//
// Decomposition: a bin is cut into `splits` equal segments; one CTA reduces one segment with
// 128-bit loads (two complex64 per load, 4 loads in flight per thread), a shuffle tree and one
// shared-memory hop.  With splits == 1 the CTA writes the final value; otherwise segment partials
// are combined by the last CTA to finish the bin (threadfence + counter), in fixed segment order,
// so results do not depend on scheduling.  Bins shorter than 1024 samples use one warp per bin.
#include "iqw_common.cuh"

namespace iqw {

constexpr int kBpThreads = 256;
constexpr int kBpMaxSplits = 64;

struct Acc {
    float sum, mx, mn;
};

__device__ __forceinline__ void acc_add(Acc& a, float re, float im) {
    const float p = fmaf(re, re, im * im);
    a.sum += p;
    a.mx = fmaxf(a.mx, p);
    a.mn = fminf(a.mn, p);
}
__device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) {
    a.sum += b.sum;
    a.mx = fmaxf(a.mx, b.mx);
    a.mn = fminf(a.mn, b.mn);
}
__device__ __forceinline__ Acc acc_warp_reduce(Acc a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Acc b;
        b.sum = __shfl_xor_sync(0xffffffffu, a.sum, o);
        b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
        b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
        acc_merge(a, b);
    }
    return a;
}

// reduce samples [begin, end) of x with `nthreads` cooperating threads, this thread = `tid`
__device__ __forceinline__ Acc reduce_span(const float2* __restrict__ x, long long begin,
                                           long long end, int tid, int nthreads) {
    Acc a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i].sum = 0.f; a[i].mx = -INFINITY; a[i].mn = INFINITY; }
    long long i = begin;
    // peel to 16-byte alignment
    if ((reinterpret_cast<uintptr_t>(x + i) & 15) && i < end) {
        if (tid == 0) { const float2 v = __ldcs(x + i); acc_add(a[0], v.x, v.y); }
        ++i;
    }
    const float4* x4 = reinterpret_cast<const float4*>(x + i);
    const long long n4 = (end - i) >> 1;
    long long j = tid;
    for (; j + 3ll * nthreads < n4; j += 4ll * nthreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(x4 + j + (long long)u * nthreads);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc_add(a[u], v[u].x, v[u].y); acc_add(a[u], v[u].z, v[u].w); }
    }
    for (; j < n4; j += nthreads) {
        const float4 v = __ldcs(x4 + j);
        acc_add(a[0], v.x, v.y);
        acc_add(a[1], v.z, v.w);
    }
    if (((end - i) & 1) && tid == 0) { const float2 v = __ldcs(x + end - 1); acc_add(a[2], v.x, v.y); }
    acc_merge(a[0], a[1]);
    acc_merge(a[2], a[3]);
    acc_merge(a[0], a[2]);
    return a[0];
}

struct BpArgs {
    const float2* x;
    long long x_ch_stride, bin_len, n_bins, n_items;   // n_items = channels * bins * splits
    int splits;
    float *mean, *mx, *mn;
    float* partial;          // [items][3] when splits > 1
    unsigned int* counters;  // [channels*bins] when splits > 1, zeroed
};

__device__ __forceinline__ void bp_store(const BpArgs& a, long long bin_flat, const Acc& r) {
    if (a.mean) a.mean[bin_flat] = r.sum / (float)a.bin_len;
    if (a.mx) a.mx[bin_flat] = r.mx;
    if (a.mn) a.mn[bin_flat] = r.mn;
}

__global__ void __launch_bounds__(kBpThreads) bin_power_cta_kernel(const BpArgs a) {
    __shared__ Acc warp_acc[kBpThreads / 32];
    __shared__ bool is_last;
    for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const long long bin_flat = item / a.splits;       // channel * n_bins + bin
        const int split = (int)(item - bin_flat * a.splits);
        const long long c = bin_flat / a.n_bins, b = bin_flat - c * a.n_bins;
        const long long seg = (a.bin_len + a.splits - 1) / a.splits;
        const long long s0 = split * seg;
        const long long s1 = min(a.bin_len, s0 + seg);
        const float2* x = a.x + c * a.x_ch_stride + b * a.bin_len;

        Acc r = acc_warp_reduce(reduce_span(x, s0, s1, threadIdx.x, kBpThreads));
        if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = r;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int wi = 1; wi < kBpThreads / 32; ++wi) acc_merge(r, warp_acc[wi]);
            if (a.splits == 1) {
                bp_store(a, bin_flat, r);
            } else {
                float* p = a.partial + item * 3;
                p[0] = r.sum; p[1] = r.mx; p[2] = r.mn;
                __threadfence();
                is_last = atomicAdd(a.counters + bin_flat, 1u) == (unsigned)a.splits - 1;
                if (is_last) {
                    __threadfence();
                    const volatile float* q = a.partial + bin_flat * a.splits * 3;
                    Acc t{q[0], q[1], q[2]};
                    for (int s = 1; s < a.splits; ++s) {
                        Acc u{q[s * 3], q[s * 3 + 1], q[s * 3 + 2]};
                        acc_merge(t, u);
                    }
                    bp_store(a, bin_flat, t);
                }
            }
        }
        __syncthreads();
    }
}

// short bins: one warp per bin
__global__ void __launch_bounds__(kBpThreads) bin_power_warp_kernel(const BpArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kBpThreads / 32);
    for (long long bin_flat = (long long)blockIdx.x * (kBpThreads / 32) + (threadIdx.x >> 5);
         bin_flat < a.n_items; bin_flat += warps) {
        const long long c = bin_flat / a.n_bins, b = bin_flat - c * a.n_bins;
        const float2* x = a.x + c * a.x_ch_stride + b * a.bin_len;
        const Acc r = acc_warp_reduce(reduce_span(x, 0, a.bin_len, lane, 32));
        if (lane == 0) bp_store(a, bin_flat, r);
    }
}

// |x|^2 of (bins, bin_len) written transposed (bin_len, bins) through a 32x33 tile
__global__ void __launch_bounds__(256)
envtopow_transposed_kernel(const float2* __restrict__ x, long long x_ch_stride, long long bin_len,
                           long long n_bins, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const long long c = blockIdx.z;
    const long long b0 = (long long)blockIdx.y * 32, s0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const long long b = b0 + r, s = s0 + tx;
        float p = 0.f;
        if (b < n_bins && s < bin_len) {
            const float2 v = __ldcs(x + c * x_ch_stride + b * bin_len + s);
            p = fmaf(v.x, v.x, v.y * v.y);
        }
        tile[r][tx] = p;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long s = s0 + r, b = b0 + tx;
        if (b < n_bins && s < bin_len) out[(c * bin_len + s) * n_bins + b] = tile[tx][r];
    }
}

}  // namespace iqw

using namespace iqw;

extern "C" size_t iqw_bin_power_workspace_bytes(int64_t n_channels, int64_t bin_len, int64_t n_bins) {
    (void)bin_len;
    if (n_channels <= 0 || n_bins <= 0) return 256;
    const size_t bins = (size_t)n_channels * (size_t)n_bins;
    return 256 + ((bins * 4 + 255) / 256) * 256 + bins * kBpMaxSplits * 3 * sizeof(float);
}

extern "C" int iqw_bin_power_c64(const void* d_x, int64_t n_channels, int64_t x_channel_stride,
                                 int64_t bin_len, int64_t n_bins, float* d_mean, float* d_max,
                                 float* d_min, void* d_workspace, size_t workspace_bytes,
                                 void* stream) {
    iqw::DeviceGuard _dev_guard(d_x);
    if (!d_x) return fail(IQW_ERR_INVALID, "null input");
    if (!d_mean && !d_max && !d_min) return fail(IQW_ERR_INVALID, "no output requested");
    if (bin_len < 1 || n_bins < 0 || n_channels < 0) return fail(IQW_ERR_INVALID, "bad sizes");
    if (n_bins == 0 || n_channels == 0) return IQW_OK;
    if (n_channels > 1 && x_channel_stride < bin_len * n_bins)
        return fail(IQW_ERR_INVALID, "x_channel_stride smaller than n_bins*bin_len");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;

    BpArgs a{};
    a.x = static_cast<const float2*>(d_x);
    a.x_ch_stride = x_channel_stride;
    a.bin_len = bin_len;
    a.n_bins = n_bins;
    a.mean = d_mean; a.mx = d_max; a.mn = d_min;
    const long long bins = n_channels * n_bins;

    if (bin_len < 1024) {
        a.splits = 1;
        a.n_items = bins;
        long long blocks = (bins + kBpThreads / 32 - 1) / (kBpThreads / 32);
        const long long cap = (long long)sms * 16;
        if (blocks > cap) blocks = cap;
        { IQW_PROFILE("bin_power_warp", s); bin_power_warp_kernel<<<(unsigned)blocks, kBpThreads, 0, s>>>(a); }
        IQW_CUDA_OK(cudaGetLastError());
        return IQW_OK;
    }

    // enough CTAs for ~4 waves of 8 CTAs/SM, each segment at least 4096 samples
    long long splits = 1;
    const long long want = (long long)sms * 8 * 4;
    if (bins < want) {
        splits = (want + bins - 1) / bins;
        const long long max_by_len = bin_len / 4096;
        if (splits > max_by_len) splits = max_by_len;
        if (splits > kBpMaxSplits) splits = kBpMaxSplits;
        if (splits < 1) splits = 1;
    }
    a.splits = (int)splits;
    a.n_items = bins * splits;
    if (splits > 1) {
        const size_t counters_bytes = (((size_t)bins * 4 + 255) / 256) * 256;
        const size_t need = counters_bytes + (size_t)a.n_items * 3 * sizeof(float);
        if (!d_workspace || workspace_bytes < need)
            return fail(IQW_ERR_WORKSPACE, "bin power workspace %zu bytes < required %zu",
                        workspace_bytes, need);
        a.counters = static_cast<unsigned int*>(d_workspace);
        a.partial = reinterpret_cast<float*>(static_cast<char*>(d_workspace) + counters_bytes);
        IQW_CUDA_OK(cudaMemsetAsync(a.counters, 0, counters_bytes, s));
    }
    long long blocks = a.n_items;
    const long long cap = (long long)sms * 32;
    if (blocks > cap) blocks = cap;
    { IQW_PROFILE("bin_power_cta", s); bin_power_cta_kernel<<<(unsigned)blocks, kBpThreads, 0, s>>>(a); }
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_envtopow_transposed_c64(const void* d_x, int64_t n_channels,
                                           int64_t x_channel_stride, int64_t bin_len,
                                           int64_t n_bins, float* d_out, void* stream) {
    iqw::DeviceGuard _dev_guard(d_x);
    if (!d_x || !d_out) return fail(IQW_ERR_INVALID, "null pointer argument");
    if (bin_len < 1 || n_bins < 0 || n_channels < 0) return fail(IQW_ERR_INVALID, "bad sizes");
    if (n_bins == 0 || n_channels == 0) return IQW_OK;
    const long long gx = (bin_len + 31) / 32, gy = (n_bins + 31) / 32;
    if (gy > 65535 || n_channels > 65535) return fail(IQW_ERR_UNSUPPORTED, "too many bins/channels for the transposing grid");
    envtopow_transposed_kernel<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)n_channels), 256, 0,
                                 static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(d_x), x_channel_stride, bin_len, n_bins, d_out);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}
