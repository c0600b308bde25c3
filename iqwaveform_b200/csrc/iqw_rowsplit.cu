// iqw_rowsplit.cu -- exact per-column order statistics of a matrix whose ROWS live on several
// GPUs (time-sharded persistence spectrum: SURVEY.md section 8e, "3-4 exchanges for exact
// quantiles").  The reference is single-process and calls np.quantile on the whole spectrogram
// (/root/reference/src/iqwaveform/fourier.py:1317-1320); here no rank ever holds all frames, so the
// order statistics are found by a most-significant-digit radix select on the order-preserving
// uint32 keys of the float32 values, 8 bits per level:
//
//   bracket   every rank finds, with kernel 2 on its own rows, the local order statistic at the
//             proportional rank floor(j*n_r/T); the minimum and maximum of those over the ranks
//             (two small all_reduce) bracket the global order statistic j -- a guaranteed bound,
//             not a statistical one -- and hold only ~sqrt(T) rows per column
//   collect   ONE pass over the local rows: per statistic the rows below its bracket are counted
//             and the rows inside any bracket (~0.5-5 % of them) are copied to per-thread
//             candidate segments                                       (bracket_collect_kernel)
//   level L   every rank counts, per column and per requested statistic, the candidates inside
//             the statistic's current key interval by their next 8 key bits
//             (candidate_count_kernel; radix_count_kernel does the same over the whole matrix
//             when a candidate segment overflowed -- heavy ties -- or no bracket is used)
//             -> all_reduce(SUM) of the (n_sel, n_cols, 256) int32 counts   (NCCL, by the host)
//             -> every rank walks the summed counts to the digit that holds the rank, extends the
//                prefix and narrows the interval; identical on all ranks  (radix_descend_kernel)
//
// After four levels the prefix IS the key of the order statistic.  order_stats_finish_kernel then
// applies dB and numpy's 'linear' interpolation exactly as kernel 2's finalize step does.
//
// Bound: HBM, 4 B per element per level (the local shard is re-read at every level; counts and
// intervals are KBs per column).  Data layout: column = bin = lane and a CTA covers 128 adjacent
// columns, so every row contributes 512 contiguous bytes per CTA; interval membership is two
// integer compares per statistic in registers.  Only rows inside an interval (a few hundred of
// ~5e5 per column with the bracket) reach the counters, which therefore live in global memory;
// runs of equal digits are merged in registers first, so tied columns (digital silence) cost one
// atomic per run instead of one per row.
#include "iqw_common.cuh"

namespace iqw {

constexpr int kRsCols = 128;       // columns per CTA
constexpr int kRsPhases = 2;       // row phases per CTA
constexpr int kRsThreads = kRsCols * kRsPhases;
constexpr int kRsUnroll = 8;       // rows in flight per thread
constexpr int kRsMaxSel = 8;       // statistics per launch
constexpr int kRsMaxStats = 16;

template <int NS, bool BELOW>
__global__ void __launch_bounds__(kRsThreads, 3)
radix_count_kernel(const float* __restrict__ p, long long rows, long long cols, const uint32_t* __restrict__ lo_g,
                   const uint32_t* __restrict__ hi_g, int n_sel, int level, int* __restrict__ counts,
                   int* __restrict__ below_g) {
    const int phase = threadIdx.x / kRsCols;
    const long long col = (long long)blockIdx.x * kRsCols + (threadIdx.x % kRsCols);
    if (col >= cols) return;
    uint32_t lo[NS], hi[NS], run_digit[NS];
    int below[NS], run[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {      // unused slots get an empty interval
        const bool use = s < n_sel;
        lo[s] = use ? lo_g[s * cols + col] : 0xFFFFFFFFu;
        hi[s] = use ? hi_g[s * cols + col] : 0u;
        below[s] = 0; run[s] = 0; run_digit[s] = 0;
    }
    const long long per = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = per * blockIdx.y, r1 = min(rows, r0 + per);
    const int shift = 24 - 8 * level;
#pragma unroll 1
    for (long long r = r0 + phase; r < r1; r += kRsPhases * kRsUnroll) {
        float v[kRsUnroll];
#pragma unroll
        for (int u = 0; u < kRsUnroll; ++u) {
            const long long rr = r + (long long)u * kRsPhases;
            v[u] = rr < r1 ? __ldcs(p + rr * cols + col) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kRsUnroll; ++u) {
            if (r + (long long)u * kRsPhases >= r1) break;
            const uint32_t key = float_to_key(v[u]);
            const uint32_t digit = (key >> shift) & 255u;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const bool ge = key >= lo[s];
                if (BELOW) below[s] += ge ? 0 : 1;
                if (ge && key <= hi[s]) {
                    if (run[s] && digit != run_digit[s]) {
                        atomicAdd(&counts[((long long)s * cols + col) * 256 + run_digit[s]], run[s]);
                        run[s] = 0;
                    }
                    run_digit[s] = digit;
                    ++run[s];
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (run[s]) atomicAdd(&counts[((long long)s * cols + col) * 256 + run_digit[s]], run[s]);
        if (BELOW && s < n_sel && below[s]) atomicAdd(&below_g[s * cols + col], below[s]);
    }
}

// ---- bracketed path: one pass that keeps the rows inside any bracket ------------------------
// A thread owns one column and kSegRows consecutive rows; what it keeps goes to ITS segment of
// the candidate store (kSegCap slots), so appending needs no atomics and no ordering; the store is
// [segment][slot][column], coalesced across a warp for equal slots.  cnt[segment][column] may
// exceed kSegCap: the surplus was dropped and the host falls back to the matrix passes.
constexpr int kSegRows = 512;
constexpr int kSegCap = 64;
constexpr int kBcCols = 128;

template <int NS>
__global__ void __launch_bounds__(kBcCols, 6)
bracket_collect_kernel(const float* __restrict__ p, long long rows, long long cols, const uint32_t* __restrict__ lo_g,
                       const uint32_t* __restrict__ hi_g, int n_sel, int* __restrict__ below_g,
                       int* __restrict__ cnt, float* __restrict__ cand) {
    const long long col = (long long)blockIdx.x * kBcCols + threadIdx.x;
    if (col >= cols) return;
    const long long seg = blockIdx.y;
    uint32_t lo[NS], hi[NS];
    int below[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {      // unused slots get an empty interval
        const bool use = s < n_sel;
        lo[s] = use ? lo_g[s * cols + col] : 0xFFFFFFFFu;
        hi[s] = use ? hi_g[s * cols + col] : 0u;
        below[s] = 0;
    }
    const long long r0 = seg * kSegRows, r1 = min(rows, r0 + kSegRows);
    float* mine = cand + seg * kSegCap * cols + col;
    int kept = 0;
    const float* src = p + r0 * cols + col;
    float va[kRsUnroll], vb[kRsUnroll];
    auto load = [&](float (&v)[kRsUnroll], long long r) {
#pragma unroll
        for (int u = 0; u < kRsUnroll; ++u) v[u] = r + u < r1 ? __ldcs(src + (r + u - r0) * cols) : 0.0f;
    };
    auto visit = [&](const float (&v)[kRsUnroll], long long r) {
#pragma unroll
        for (int u = 0; u < kRsUnroll; ++u) {
            if (r + u >= r1) break;
            const uint32_t key = float_to_key(v[u]);
            bool any = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const bool ge = key >= lo[s];
                below[s] += ge ? 0 : 1;
                any |= ge && key <= hi[s];
            }
            if (any) {
                if (kept < kSegCap) mine[(long long)kept * cols] = v[u];
                ++kept;
            }
        }
    };
    load(va, r0);
#pragma unroll 1
    for (long long r = r0; r < r1; r += 2 * kRsUnroll) {
        load(vb, r + kRsUnroll);
        visit(va, r);
        load(va, r + 2 * kRsUnroll);
        visit(vb, r + kRsUnroll);
    }
    cnt[seg * cols + col] = kept;
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (s < n_sel && below[s]) atomicAdd(&below_g[s * cols + col], below[s]);
    if (kept > kSegCap) atomicAdd(&below_g[(long long)n_sel * cols + col], 1);      // overflow row
}

// digit counts over the candidate store instead of the matrix (same contract as radix_count_kernel)
template <int NS>
__global__ void __launch_bounds__(kBcCols)
candidate_count_kernel(const int* __restrict__ cnt, const float* __restrict__ cand, long long segs, long long cols,
                       const uint32_t* __restrict__ lo_g, const uint32_t* __restrict__ hi_g, int n_sel, int level,
                       int* __restrict__ counts) {
    const long long col = (long long)blockIdx.x * kBcCols + threadIdx.x;
    if (col >= cols) return;
    uint32_t lo[NS], hi[NS], run_digit[NS];
    int run[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const bool use = s < n_sel;
        lo[s] = use ? lo_g[s * cols + col] : 0xFFFFFFFFu;
        hi[s] = use ? hi_g[s * cols + col] : 0u;
        run[s] = 0; run_digit[s] = 0;
    }
    const long long per = (segs + gridDim.y - 1) / gridDim.y;
    const long long g0 = per * blockIdx.y, g1 = min(segs, g0 + per);
    const int shift = 24 - 8 * level;
#pragma unroll 1
    for (long long g = g0; g < g1; ++g) {
        const int n = min(cnt[g * cols + col], kSegCap);
        const float* mine = cand + g * kSegCap * cols + col;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const uint32_t key = float_to_key(mine[(long long)i * cols]);
            const uint32_t digit = (key >> shift) & 255u;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (key >= lo[s] && key <= hi[s]) {
                    if (run[s] && digit != run_digit[s]) {
                        atomicAdd(&counts[((long long)s * cols + col) * 256 + run_digit[s]], run[s]);
                        run[s] = 0;
                    }
                    run_digit[s] = digit;
                    ++run[s];
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (run[s]) atomicAdd(&counts[((long long)s * cols + col) * 256 + run_digit[s]], run[s]);
}

// one thread per (statistic, column): find the digit whose cumulative count covers the residual
// rank, append it to the prefix, make the rank relative to that digit's bucket and narrow the key
// interval to the bucket.  A rank outside the counted rows (a bracket that does not hold it, ranks
// computed for another row count) poisons that column with the key of NaN instead of returning a
// plausible wrong value.
__global__ void radix_descend_kernel(const int* __restrict__ counts, int n_sel, long long cols, int level,
                                     long long* __restrict__ rank, uint32_t* __restrict__ prefix,
                                     uint32_t* __restrict__ lo, uint32_t* __restrict__ hi) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_sel * cols) return;
    const int4* hv = reinterpret_cast<const int4*>(counts + idx * 256);
    const long long r = rank[idx];
    long long cum = 0;
    int digit = -1;
    if (r >= 0) {
        for (int q = 0; q < 64 && digit < 0; ++q) {
            const int4 n4 = hv[q];
            const int n[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (digit < 0) {
                    if (r < cum + n[k]) digit = q * 4 + k;
                    else cum += n[k];
                }
            }
        }
    }
    const int shift = 24 - 8 * level;
    if (digit < 0) {
        rank[idx] = -1;
        prefix[idx] = 0xFFFFFFFFu >> shift;      // ends as key 0xFFFFFFFF = NaN
        lo[idx] = 0xFFFFFFFFu;                    // empty interval: nothing more is counted
        hi[idx] = 0u;
        return;
    }
    const uint32_t pre = level == 0 ? (uint32_t)digit : ((prefix[idx] << 8) | (uint32_t)digit);
    prefix[idx] = pre;
    rank[idx] = r - cum;
    const uint32_t b0 = pre << shift, b1 = b0 | (shift ? (0xFFFFFFFFu >> (32 - shift)) : 0u);
    lo[idx] = max(lo[idx], b0);
    hi[idx] = min(hi[idx], b1);
}

struct FinishPlan {
    int n_stats;
    int kind[kRsMaxStats], ia[kRsMaxStats], ib[kRsMaxStats];
    float gamma[kRsMaxStats];
};

// dB of the selected order statistics (monotone, so equal to selecting among dB values) and the
// separate-rounding float32 lerp of numpy's 'linear' method (_function_base_impl.py _lerp)
__global__ void order_stats_finish_kernel(const uint32_t* __restrict__ keys, long long cols, FinishPlan fp, int to_dB,
                                          float eps, float* __restrict__ out) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    auto xf = [&](uint32_t key) {
        const float v = key_to_float(key);
        return to_dB ? power_to_dB(v, eps) : v;
    };
    for (int t = 0; t < fp.n_stats; ++t) {
        const float a = xf(keys[fp.ia[t] * cols + col]);
        const float b = xf(keys[fp.ib[t] * cols + col]);
        float r;
        if (fp.kind[t] == IQW_STAT_MEDIAN) {
            r = __fmul_rn(__fadd_rn(a, b), 0.5f);
        } else {
            const float g = fp.gamma[t];
            const float d = __fsub_rn(b, a);
            r = g >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g))) : __fadd_rn(a, __fmul_rn(d, g));
        }
        out[(long long)t * cols + col] = r;
    }
}

}  // namespace iqw

using namespace iqw;

template <int NS>
static int launch_count(const float* d_p, int64_t n_rows, int64_t n_cols, int n_sel, const uint32_t* lo,
                        const uint32_t* hi, int level, int32_t* counts, int32_t* below, dim3 grid, cudaStream_t s) {
    if (below) radix_count_kernel<NS, true><<<grid, kRsThreads, 0, s>>>(d_p, n_rows, n_cols, lo, hi, n_sel, level, counts, below);
    else radix_count_kernel<NS, false><<<grid, kRsThreads, 0, s>>>(d_p, n_rows, n_cols, lo, hi, n_sel, level, counts, nullptr);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_radix_count_f32(const float* d_p, int64_t n_rows, int64_t n_cols, int32_t n_sel,
                                   const uint32_t* d_lo, const uint32_t* d_hi, int32_t level,
                                   int32_t* d_counts, int32_t* d_below, void* stream) {
    iqw::DeviceGuard _dev_guard(d_lo);
    if (n_rows < 0 || n_cols < 1 || n_sel < 1 || level < 0 || level > 3 || !d_counts || !d_lo || !d_hi ||
        (n_rows > 0 && !d_p))
        return fail(IQW_ERR_INVALID, "iqw_radix_count_f32: bad argument");
    if (n_rows >= (int64_t)1 << 31) return fail(IQW_ERR_UNSUPPORTED, "iqw_radix_count_f32: more than 2^31-1 rows");
    cudaStream_t s = (cudaStream_t)stream;
    IQW_CUDA_OK(cudaMemsetAsync(d_counts, 0, (size_t)n_sel * n_cols * 256 * sizeof(int32_t), s));
    if (d_below) IQW_CUDA_OK(cudaMemsetAsync(d_below, 0, (size_t)n_sel * n_cols * sizeof(int32_t), s));
    if (n_rows == 0) return IQW_OK;
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    const long long tiles = (n_cols + kRsCols - 1) / kRsCols;
    long long splits = (16LL * sms + tiles - 1) / tiles;       // ~4 resident CTAs per SM, 4 waves
    const long long most = (n_rows + kRsPhases * kRsUnroll - 1) / (kRsPhases * kRsUnroll);
    if (splits > most) splits = most;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    for (int s0 = 0; s0 < n_sel; s0 += kRsMaxSel) {
        const int ns = n_sel - s0 < kRsMaxSel ? n_sel - s0 : kRsMaxSel;
        const uint32_t *lo = d_lo + (size_t)s0 * n_cols, *hi = d_hi + (size_t)s0 * n_cols;
        int32_t* cn = d_counts + (size_t)s0 * n_cols * 256;
        int32_t* bl = d_below ? d_below + (size_t)s0 * n_cols : nullptr;
        IQW_PROFILE("radix_count", s);
        int rc;
        if (ns <= 1) rc = launch_count<1>(d_p, n_rows, n_cols, ns, lo, hi, level, cn, bl, grid, s);
        else if (ns <= 2) rc = launch_count<2>(d_p, n_rows, n_cols, ns, lo, hi, level, cn, bl, grid, s);
        else if (ns <= 4) rc = launch_count<4>(d_p, n_rows, n_cols, ns, lo, hi, level, cn, bl, grid, s);
        else rc = launch_count<8>(d_p, n_rows, n_cols, ns, lo, hi, level, cn, bl, grid, s);
        if (rc) return rc;
    }
    return IQW_OK;
}

static inline long long collect_segments(int64_t n_rows) { return (n_rows + kSegRows - 1) / kSegRows; }

extern "C" size_t iqw_bracket_collect_workspace_bytes(int64_t n_rows, int64_t n_cols) {
    if (n_rows < 0 || n_cols < 1) return 0;
    const size_t segs = (size_t)collect_segments(n_rows);
    return (segs * n_cols * sizeof(int32_t) + 255) / 256 * 256 + segs * kSegCap * n_cols * sizeof(float) + 256;
}

struct CollectView { int* cnt; float* cand; long long segs; };
static CollectView collect_view(void* ws, int64_t n_rows, int64_t n_cols) {
    CollectView v;
    v.segs = collect_segments(n_rows);
    v.cnt = (int*)ws;
    v.cand = (float*)((char*)ws + ((size_t)v.segs * n_cols * sizeof(int32_t) + 255) / 256 * 256);
    return v;
}

extern "C" int iqw_bracket_collect_f32(const float* d_p, int64_t n_rows, int64_t n_cols, int32_t n_sel,
                                       const uint32_t* d_lo, const uint32_t* d_hi, int32_t* d_below,
                                       void* d_workspace, size_t workspace_bytes, void* stream) {
    iqw::DeviceGuard _dev_guard(d_lo);
    if (n_rows < 0 || n_cols < 1 || n_sel < 1 || n_sel > kRsMaxSel || !d_lo || !d_hi || !d_below || (n_rows > 0 && !d_p))
        return fail(IQW_ERR_INVALID, "iqw_bracket_collect_f32: bad argument (1 <= n_sel <= %d)", kRsMaxSel);
    if (n_rows >= (int64_t)1 << 31) return fail(IQW_ERR_UNSUPPORTED, "iqw_bracket_collect_f32: more than 2^31-1 rows");
    if (workspace_bytes < iqw_bracket_collect_workspace_bytes(n_rows, n_cols) || (n_rows > 0 && !d_workspace))
        return fail(IQW_ERR_WORKSPACE, "iqw_bracket_collect_f32: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    IQW_CUDA_OK(cudaMemsetAsync(d_below, 0, (size_t)(n_sel + 1) * n_cols * sizeof(int32_t), s));
    if (n_rows == 0) return IQW_OK;
    const CollectView v = collect_view(d_workspace, n_rows, n_cols);
    if (v.segs > 65535) return fail(IQW_ERR_UNSUPPORTED, "iqw_bracket_collect_f32: more than %d rows", 65535 * kSegRows);
    dim3 grid((unsigned)((n_cols + kBcCols - 1) / kBcCols), (unsigned)v.segs);
    IQW_PROFILE("bracket_collect", s);
#define IQW_BC(NS) bracket_collect_kernel<NS><<<grid, kBcCols, 0, s>>>(d_p, n_rows, n_cols, d_lo, d_hi, n_sel, d_below, v.cnt, v.cand)
    if (n_sel <= 1) IQW_BC(1); else if (n_sel <= 2) IQW_BC(2); else if (n_sel <= 4) IQW_BC(4); else IQW_BC(8);
#undef IQW_BC
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_candidate_count_f32(const void* d_workspace, int64_t n_rows, int64_t n_cols, int32_t n_sel,
                                       const uint32_t* d_lo, const uint32_t* d_hi, int32_t level,
                                       int32_t* d_counts, void* stream) {
    iqw::DeviceGuard _dev_guard(d_workspace);
    if (n_rows < 0 || n_cols < 1 || n_sel < 1 || n_sel > kRsMaxSel || level < 0 || level > 3 || !d_counts || !d_lo ||
        !d_hi || (n_rows > 0 && !d_workspace))
        return fail(IQW_ERR_INVALID, "iqw_candidate_count_f32: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    IQW_CUDA_OK(cudaMemsetAsync(d_counts, 0, (size_t)n_sel * n_cols * 256 * sizeof(int32_t), s));
    if (n_rows == 0) return IQW_OK;
    const CollectView v = collect_view(const_cast<void*>(d_workspace), n_rows, n_cols);
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    const long long tiles = (n_cols + kBcCols - 1) / kBcCols;
    long long splits = (8LL * sms + tiles - 1) / tiles;
    if (splits > v.segs) splits = v.segs;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    IQW_PROFILE("candidate_count", s);
#define IQW_CC(NS) candidate_count_kernel<NS><<<grid, kBcCols, 0, s>>>(v.cnt, v.cand, v.segs, n_cols, d_lo, d_hi, n_sel, level, d_counts)
    if (n_sel <= 1) IQW_CC(1); else if (n_sel <= 2) IQW_CC(2); else if (n_sel <= 4) IQW_CC(4); else IQW_CC(8);
#undef IQW_CC
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_radix_descend(const int32_t* d_counts, int32_t n_sel, int64_t n_cols, int32_t level,
                                 int64_t* d_rank, uint32_t* d_prefix, uint32_t* d_lo, uint32_t* d_hi,
                                 void* stream) {
    iqw::DeviceGuard _dev_guard(d_counts);
    if (!d_counts || !d_rank || !d_prefix || !d_lo || !d_hi || n_sel < 1 || n_cols < 1 || level < 0 || level > 3)
        return fail(IQW_ERR_INVALID, "iqw_radix_descend: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = (long long)n_sel * n_cols;
    IQW_PROFILE("radix_descend", s);
    radix_descend_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_counts, n_sel, n_cols, level,
                                                                      (long long*)d_rank, d_prefix, d_lo, d_hi);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

extern "C" int iqw_order_stats_finish_f32(const uint32_t* d_keys, int32_t n_sel, const int64_t* sel_rank,
                                          int64_t n_rows_total, int64_t n_cols, const iqw_stat* stats,
                                          int32_t n_stats, int32_t to_dB, float eps, float* d_out, void* stream) {
    iqw::DeviceGuard _dev_guard(d_keys);
    if (!d_keys || !sel_rank || !stats || !d_out || n_sel < 1 || n_cols < 1 || n_stats < 1 || n_stats > kRsMaxStats)
        return fail(IQW_ERR_INVALID, "iqw_order_stats_finish_f32: bad argument");
    FinishPlan fp;
    fp.n_stats = n_stats;
    auto find = [&](int64_t r) {
        for (int i = 0; i < n_sel; ++i)
            if (sel_rank[i] == r) return i;
        return -1;
    };
    for (int t = 0; t < n_stats; ++t) {
        const iqw_stat& st = stats[t];
        if (st.kind != IQW_STAT_QUANTILE && st.kind != IQW_STAT_MEDIAN)
            return fail(IQW_ERR_INVALID, "iqw_order_stats_finish_f32: statistic %d is not an order statistic", t);
        const int64_t lo = st.kind == IQW_STAT_MEDIAN ? (n_rows_total - 1) / 2 : st.rank_lo;
        const int64_t hi = st.kind == IQW_STAT_MEDIAN ? n_rows_total / 2 : st.rank_hi;
        fp.kind[t] = st.kind;
        fp.gamma[t] = st.gamma;
        fp.ia[t] = find(lo);
        fp.ib[t] = find(hi);
        if (fp.ia[t] < 0 || fp.ib[t] < 0)
            return fail(IQW_ERR_INVALID, "iqw_order_stats_finish_f32: ranks (%lld, %lld) of statistic %d were not selected",
                        (long long)lo, (long long)hi, t);
    }
    cudaStream_t s = (cudaStream_t)stream;
    IQW_PROFILE("order_stats_finish", s);
    order_stats_finish_kernel<<<(unsigned)((n_cols + 127) / 128), 128, 0, s>>>(d_keys, n_cols, fp, to_dB, eps, d_out);
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}
