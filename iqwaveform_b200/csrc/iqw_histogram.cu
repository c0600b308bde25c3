// iqw_histogram.cu -- kernel 5: per-row counts of where float32 samples fall among sorted edges.
// The array branches of two thin consumers of the power path:
//   power_analysis.py:552-583  sample_ccdf      : searchsorted(edges, a, 'left')  -> bincount -> cumsum
//   util.py:497-543            histogram_last_axis: searchsorted(bins, x, 'right') - 1 -> bincount per row
// (/root/reference/src/iqwaveform/).  Both reduce to counts[row][i] = #{ x in row : searchsorted(edges,
// x, side) == i }, i = 0..n_edges; the cumulative sum / column slicing on that tiny result stays on
// the host side of the call.
//
// One CTA handles a slice of one row; samples are read with 128-bit loads.  numpy compares the
// float32 samples with float64 edges in float64; the same decisions are taken here in float32
// against edges rounded once in the direction that keeps every comparison's outcome
// (x > e  <=>  x > round_down(e);  x >= e  <=>  x >= round_up(e) for float x).
// The class of a sample is GUESSED from the straight line through the first and last edge and
// then corrected against the neighbouring edges (exact for any ascending edges; one or two shared
// loads for evenly spaced ones, which is what linspace callers pass); after a few correction
// steps it falls back to a binary search.  Counters are private per thread while they fit in
// shared memory (no same-address collisions), per group of threads beyond that.
// Bound: HBM, 4 B per sample.
#include "iqw_common.cuh"

namespace iqw {

constexpr int kHiThreads = 256;
constexpr int kHiMaxEdges = 4096;
constexpr int kHiCounterBytes = 64 * 1024;

template <bool RIGHT>
__device__ __forceinline__ float edge_as_float(double e) { return RIGHT ? __double2float_ru(e) : __double2float_rd(e); }
template <bool RIGHT>
__device__ __forceinline__ bool beyond(float v, float e) { return RIGHT ? (v >= e) : (v > e); }

// number of edges the sample is beyond, in [0, n]; NaN counts as beyond all, as numpy sorts it last
template <bool RIGHT>
__device__ __forceinline__ int classify(const float* __restrict__ e, int n, float e0, float inv, float v) {
    if (v != v) return n;
    // guess: edges passed if they were evenly spaced (clamped; inf and huge values saturate)
    const float t = (v - e0) * inv;
    int k = t < 0.0f ? 0 : (t >= (float)n ? n : (int)t + 1);
    k = k > n ? n : k;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const bool down = k > 0 && !beyond<RIGHT>(v, e[k - 1]);
        const bool up = k < n && beyond<RIGHT>(v, e[k]);
        if (!down && !up) return k;
        k += up ? 1 : -1;
    }
    int lo = 0, len = n;        // uneven edges: plain binary search
    while (len > 0) {
        const int half = len >> 1;
        const bool go = beyond<RIGHT>(v, e[lo + half]);
        lo = go ? lo + half + 1 : lo;
        len = go ? len - half - 1 : half;
    }
    return lo;
}

template <bool RIGHT>
__global__ void __launch_bounds__(kHiThreads)
edge_count_kernel(const float* __restrict__ a, long long n_cols, const double* __restrict__ edges, int n_edges,
                  int copies, unsigned long long* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* e = reinterpret_cast<float*>(smem_raw);
    unsigned int* h = reinterpret_cast<unsigned int*>(e + ((n_edges + 3) & ~3));     // [copies][n_edges + 1]
    const int nb = n_edges + 1;
    for (int i = threadIdx.x; i < n_edges; i += kHiThreads) e[i] = edge_as_float<RIGHT>(edges[i]);
    for (int i = threadIdx.x; i < copies * nb; i += kHiThreads) h[i] = 0;
    __syncthreads();
    unsigned int* mine = h + (threadIdx.x % copies) * nb;
    const float e0 = e[0];
    const float span = e[n_edges - 1] - e0;
    const float inv = (n_edges > 1 && span > 0.0f && span < 3.0e38f) ? (float)(n_edges - 1) / span : 0.0f;

    const long long row = blockIdx.x;
    const float* src = a + row * n_cols;
    const long long per = ((n_cols + gridDim.y - 1) / gridDim.y + 3) & ~3LL;
    const long long c0 = min(n_cols, per * blockIdx.y), c1 = min(n_cols, c0 + per);
    const bool vec = ((reinterpret_cast<uintptr_t>(src + c0) & 15) == 0);
    const long long nvec = vec ? (c1 - c0) >> 2 : 0;
    const float4* src4 = reinterpret_cast<const float4*>(src + c0);
    auto count4 = [&](const float4& v) {
        atomicAdd(&mine[classify<RIGHT>(e, n_edges, e0, inv, v.x)], 1u);
        atomicAdd(&mine[classify<RIGHT>(e, n_edges, e0, inv, v.y)], 1u);
        atomicAdd(&mine[classify<RIGHT>(e, n_edges, e0, inv, v.z)], 1u);
        atomicAdd(&mine[classify<RIGHT>(e, n_edges, e0, inv, v.w)], 1u);
    };
    long long i = threadIdx.x;
    for (; i + 3 * kHiThreads < nvec; i += 4 * kHiThreads) {      // four 16-byte loads in flight per thread
        const float4 v0 = __ldcs(src4 + i), v1 = __ldcs(src4 + i + kHiThreads);
        const float4 v2 = __ldcs(src4 + i + 2 * kHiThreads), v3 = __ldcs(src4 + i + 3 * kHiThreads);
        count4(v0); count4(v1); count4(v2); count4(v3);
    }
    for (; i < nvec; i += kHiThreads) count4(__ldcs(src4 + i));
    for (long long k = c0 + nvec * 4 + threadIdx.x; k < c1; k += kHiThreads)
        atomicAdd(&mine[classify<RIGHT>(e, n_edges, e0, inv, src[k])], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += kHiThreads) {
        unsigned long long s = 0;
        for (int c = 0; c < copies; ++c) s += h[c * nb + b];
        if (s) atomicAdd(&counts[row * nb + b], s);
    }
}

}  // namespace iqw

using namespace iqw;

extern "C" int iqw_edge_counts_f32(const float* d_a, int64_t n_rows, int64_t n_cols, const double* d_edges,
                                   int32_t n_edges, int32_t side_right, int64_t* d_counts, void* stream) {
    iqw::DeviceGuard _dev_guard(d_a);
    if (!d_a || !d_edges || !d_counts || n_rows < 1 || n_cols < 1 || n_edges < 1)
        return fail(IQW_ERR_INVALID, "iqw_edge_counts_f32: bad argument");
    if (n_edges > kHiMaxEdges)
        return fail(IQW_ERR_UNSUPPORTED, "iqw_edge_counts_f32: %d edges, at most %d are built", n_edges, kHiMaxEdges);
    if (n_rows >= (1LL << 31)) return fail(IQW_ERR_UNSUPPORTED, "iqw_edge_counts_f32: more than 2^31-1 rows");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nb = (size_t)n_edges + 1;
    IQW_CUDA_OK(cudaMemsetAsync(d_counts, 0, (size_t)n_rows * nb * sizeof(int64_t), s));
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    long long slices = ((long long)sms * 8 + n_rows - 1) / n_rows;
    const long long most = (n_cols + kHiThreads * 16 - 1) / (kHiThreads * 16);
    if (slices > most) slices = most;
    if (slices < 1) slices = 1;
    // a slice must stay below 2^32 samples (32-bit shared counters)
    while ((n_cols + slices - 1) / slices >= (1LL << 32)) slices *= 2;
    if (slices > 65535) return fail(IQW_ERR_UNSUPPORTED, "iqw_edge_counts_f32: rows longer than 2^47 samples");
    dim3 grid((unsigned)n_rows, (unsigned)slices);
    unsigned long long* out = (unsigned long long*)d_counts;
    IQW_PROFILE("edge_count", s);
    // one private copy of the counters per thread while they fit, then per 2, 4 ... threads
    int copies = kHiThreads;
    while (copies > 1 && copies * nb * sizeof(unsigned int) > (size_t)kHiCounterBytes) copies >>= 1;
    const size_t smem = (size_t)((n_edges + 3) & ~3) * sizeof(float) + copies * nb * sizeof(unsigned int);
    auto kr = edge_count_kernel<true>;
    auto kl = edge_count_kernel<false>;
    IQW_CUDA_OK(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, kHiCounterBytes + 32 * 1024));
    IQW_CUDA_OK(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, kHiCounterBytes + 32 * 1024));
    if (side_right) kr<<<grid, kHiThreads, smem, s>>>(d_a, n_cols, d_edges, n_edges, copies, out);
    else kl<<<grid, kHiThreads, smem, s>>>(d_a, n_cols, d_edges, n_edges, copies, out);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}
