// iqw_stats.cu -- kernel 2: statistics over the time axis of a (rows, cols) float32 matrix:
// exact order statistics (+ numpy 'linear' lerp), mean, max, min -- per column.
//
// Replaces /root/reference/src/iqwaveform/fourier.py:1311-1325 (np.quantile with float32 q and
// the named ufuncs over the time axis; 73 % of the reference's CPU time is ndarray.partition)
// and the in-place powtodB at fourier.py:1298-1299.
//
// Exact selection without sorting and without atomics in the hot loop.  The matrix is frame-major
// (a row = one frame, written coalesced by kernel 1), so a THREAD OWNS A COLUMN: lanes of a warp
// read 32 adjacent columns of one row (128 B, coalesced) and every thread streams down the time
// axis keeping PRIVATE counters in shared memory, laid out [bucket][thread] so that bank == lane
// (conflict free, no atomics).  The m <= 8 distinct target ranks of a call are located level by
// level; per column the state is a short ordered list of disjoint key INTERVALS, each known to
// contain a contiguous group of the target ranks, with the exact number of keys below it:
//
//   range   : min/max key of ~2k sampled rows per column  -> bucket map of that column
//   L0      : 256-bucket histogram of every element (bucket 0 / 255 catch keys outside the sampled
//             range, so counts are exact whatever the sample missed); also exact min / max and the
//             float64 sum (of dB values when to_dB) for the named statistics
//   scan0   : per column, group the target ranks by bucket -> intervals
//   R (x k) : 32 sub-buckets per pending interval, all intervals in one pass; the scan that follows
//             splits each interval by sub-bucket.  Repeated (kernels exit at once when nothing is
//             pending) until an interval holds <= CAP keys or is a single key
//   collect : keys inside the final intervals -> candidate lists (<= CAP each)
//   resolve : bitonic sort of the candidates in shared memory, pick the ranks, dB, lerp, store
//
// Keys are the order-preserving uint32 image of the float (iqw_common.cuh float_to_key), so the
// selection is exact for any input, ties and signed zeros included.  NaNs sort above +inf.
#include "iqw_common.cuh"

namespace iqw {

constexpr int kMaxRanks = 8;       // distinct target ranks (and therefore intervals) per call
constexpr int kMaxStats = 32;      // output rows per call
constexpr int kNB0 = 256;          // level-0 buckets
constexpr int kNSub = 32;          // sub-buckets per interval and refinement level
constexpr int kCap = 1024;         // candidates kept per (column, interval)
constexpr int kRefineLevels = 6;   // 2^32 / 32^6 < 254: always enough
constexpr int kBX = 128;           // columns (= threads) per CTA in the streaming passes
constexpr int kUnroll = 16;        // rows in flight per thread

enum IvStatus : uint32_t { IV_REFINE = 0, IV_COLLECT = 1, IV_RESOLVED = 2 };

struct RankPlan {                  // same for every column: depends only on n_rows
    int n_ranks;
    unsigned int rank[kMaxRanks];  // ascending, distinct
};

struct StatPlan {
    int n_stats;
    int kind[kMaxStats];
    int ia[kMaxStats], ib[kMaxStats];   // indices into RankPlan::rank
    float gamma[kMaxStats];
};

// per-channel workspace, carved by carve_workspace()
struct Work {
    uint32_t* range_lo;   // [cols]
    uint32_t* range_hi;   // [cols]
    uint32_t* kmin;       // [cols]
    uint32_t* kmax;       // [cols]
    double* dsum;         // [cols]
    uint32_t* hist0;      // [cols][256]
    uint32_t* n_iv;       // [cols]
    uint32_t* iv_klo;     // [cols][8]  first key of the interval
    uint32_t* iv_aux;     // [cols][8]  REFINE: sub-bucket shift; COLLECT: last key (inclusive)
    uint32_t* iv_below;   // [cols][8]  number of keys < klo in the column
    uint32_t* iv_status;  // [cols][8]
    uint32_t* iv_first;   // [cols][8]  first target-rank index inside
    uint32_t* iv_nr;      // [cols][8]  number of target ranks inside
    uint32_t* r_key;      // [cols][8]  key of each target rank once known
    uint32_t* hist1;      // [cols][8][32]
    uint32_t* cursor;     // [cols][8]
    uint32_t* cand;       // [cols][8][kCap]
    uint32_t* pending;    // [1] number of intervals in IV_REFINE
    size_t zero_bytes;    // leading bytes that must be zero before a channel starts
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve_workspace(void* base, int64_t cols, Work* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off = align_up(off + bytes, 256);
        return p;
    };
    const size_t c = (size_t)cols, m = kMaxRanks;
    Work t{};
    // --- zero-initialised region first ---
    t.range_hi = (uint32_t*)take(4 * c);
    t.kmax = (uint32_t*)take(4 * c);
    t.dsum = (double*)take(8 * c);
    t.hist0 = (uint32_t*)take(4 * c * kNB0);
    t.hist1 = (uint32_t*)take(4 * c * m * kNSub);
    t.cursor = (uint32_t*)take(4 * c * m);
    t.n_iv = (uint32_t*)take(4 * c);
    t.pending = (uint32_t*)take(256);
    t.zero_bytes = off;
    // --- 0xFF-initialised ---
    t.range_lo = (uint32_t*)take(4 * c);
    t.kmin = (uint32_t*)take(4 * c);
    // --- written before read ---
    t.iv_klo = (uint32_t*)take(4 * c * m);
    t.iv_aux = (uint32_t*)take(4 * c * m);
    t.iv_below = (uint32_t*)take(4 * c * m);
    t.iv_status = (uint32_t*)take(4 * c * m);
    t.iv_first = (uint32_t*)take(4 * c * m);
    t.iv_nr = (uint32_t*)take(4 * c * m);
    t.r_key = (uint32_t*)take(4 * c * m);
    t.cand = (uint32_t*)take(4 * c * m * kCap);
    if (w) *w = t;
    return off;
}

__device__ __forceinline__ float load_stream(const float* p) { return __ldcs(p); }

// ---------------------------------------------------------------------------------------------
// range: min / max key over sampled rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBX)
range_kernel(const float* __restrict__ p, long long rows, long long cols, long long row_step,
             Work w) {
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    if (col >= cols) return;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (long long r = (long long)blockIdx.y * row_step; r < rows; r += row_step * gridDim.y) {
        const uint32_t k = float_to_key(load_stream(p + r * cols + col));
        lo = min(lo, k);
        hi = max(hi, k);
    }
    atomicMin(w.range_lo + col, lo);
    atomicMax(w.range_hi + col, hi);
}

// bucket map of level 0: 0 = below the sampled range, 255 = above, 1..254 inside
__device__ __forceinline__ uint32_t l0_shift(uint32_t lo, uint32_t hi) {
    const uint32_t span = hi - lo;
    uint32_t s = 0;
    while ((span >> s) >= (uint32_t)(kNB0 - 2)) ++s;
    return s;
}
__device__ __forceinline__ uint32_t l0_bucket(uint32_t k, uint32_t lo, uint32_t hi, uint32_t s) {
    if (k < lo) return 0;
    if (k > hi) return kNB0 - 1;
    return 1 + ((k - lo) >> s);
}

// ---------------------------------------------------------------------------------------------
// L0: private 256-bucket histogram per column + min / max / sum
// ---------------------------------------------------------------------------------------------
template <bool WANT_SUM, bool TO_DB>
__global__ void __launch_bounds__(kBX)
l0_kernel(const float* __restrict__ p, long long rows, long long cols, long long rows_per_split,
          float eps, Work w) {
    extern __shared__ uint16_t hist[];   // [kNB0][kBX]
    for (int i = threadIdx.x; i < kNB0 * kBX; i += kBX) hist[i] = 0;
    __syncthreads();

    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    if (col >= cols) return;
    const uint32_t lo = w.range_lo[col], hi = w.range_hi[col];
    const uint32_t s = l0_shift(lo, hi);
    uint16_t* h = hist + threadIdx.x;

    const long long r0 = (long long)blockIdx.y * rows_per_split;
    const long long r1 = min(rows, r0 + rows_per_split);
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    double dsum = 0.0;
    const float* src = p + r0 * cols + col;

    long long r = r0;
    for (; r + kUnroll <= r1; r += kUnroll, src += (long long)kUnroll * cols) {
        float v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = load_stream(src + (long long)u * cols);
        float part = 0.f;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t k = float_to_key(v[u]);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
            h[l0_bucket(k, lo, hi, s) * kBX] += 1;
            if (WANT_SUM) part += TO_DB ? power_to_dB(v[u], eps) : v[u];
        }
        if (WANT_SUM) dsum += (double)part;
    }
    for (; r < r1; ++r, src += cols) {
        const float v = load_stream(src);
        const uint32_t k = float_to_key(v);
        kmin = min(kmin, k);
        kmax = max(kmax, k);
        h[l0_bucket(k, lo, hi, s) * kBX] += 1;
        if (WANT_SUM) dsum += (double)(TO_DB ? power_to_dB(v, eps) : v);
    }

    uint32_t* g = w.hist0 + col * kNB0;
    for (int b = 0; b < kNB0; ++b) {
        const uint32_t c = h[b * kBX];
        if (c) atomicAdd(g + b, c);
    }
    atomicMin(w.kmin + col, kmin);
    atomicMax(w.kmax + col, kmax);
    if (WANT_SUM) atomicAdd(w.dsum + col, dsum);
}

// ---------------------------------------------------------------------------------------------
// interval bookkeeping (one thread per column, sequential)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ceil_log2_u64(unsigned long long v) {
    uint32_t l = 0;
    while ((1ull << l) < v) ++l;
    return l;
}

struct IvList {
    uint32_t n;
    uint32_t klo[kMaxRanks], aux[kMaxRanks], below[kMaxRanks], status[kMaxRanks],
        first[kMaxRanks], nr[kMaxRanks];
};

// append the interval [klo, klo + span) holding `cnt` keys, `below` keys under it, and target
// ranks [first, first + nr).  `single` = every key in it is the same key `klo_exact`.
__device__ void iv_append(IvList& L, const Work& w, long long col, uint32_t klo,
                          unsigned long long span, uint32_t below, uint32_t cnt, uint32_t first,
                          uint32_t nr, bool single, uint32_t key_exact) {
    const uint32_t i = L.n++;
    L.klo[i] = klo;
    L.below[i] = below;
    L.first[i] = first;
    L.nr[i] = nr;
    L.aux[i] = 0;
    if (single || span == 1ull) {
        L.status[i] = IV_RESOLVED;
        for (uint32_t q = 0; q < nr; ++q) w.r_key[col * kMaxRanks + first + q] = single ? key_exact : klo;
    } else if (cnt <= (uint32_t)kCap) {
        L.status[i] = IV_COLLECT;
        const unsigned long long last = (unsigned long long)klo + span - 1ull;
        L.aux[i] = last > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)last;
    } else {
        L.status[i] = IV_REFINE;
        const uint32_t l = ceil_log2_u64(span);
        L.aux[i] = l > 5 ? l - 5 : 0;
        atomicAdd(w.pending, 1u);
    }
}

__device__ void iv_store(const IvList& L, const Work& w, long long col) {
    w.n_iv[col] = L.n;
    for (uint32_t i = 0; i < L.n; ++i) {
        const long long x = col * kMaxRanks + i;
        w.iv_klo[x] = L.klo[i];
        w.iv_aux[x] = L.aux[i];
        w.iv_below[x] = L.below[i];
        w.iv_status[x] = L.status[i];
        w.iv_first[x] = L.first[i];
        w.iv_nr[x] = L.nr[i];
    }
}

__global__ void scan0_kernel(long long cols, RankPlan rp, Work w) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const uint32_t lo = w.range_lo[col], hi = w.range_hi[col];
    const uint32_t s0 = l0_shift(lo, hi);
    const uint32_t* h = w.hist0 + col * kNB0;

    IvList L;
    L.n = 0;
    uint32_t cum = 0;
    int i = 0;
    for (int b = 0; b < kNB0 && i < rp.n_ranks; ++b) {
        const uint32_t c = h[b];
        const uint32_t next = cum + c;
        if (rp.rank[i] < next) {
            const int first = i;
            while (i < rp.n_ranks && rp.rank[i] < next) ++i;
            const unsigned long long start =
                b == 0 ? 0ull
                       : b == kNB0 - 1 ? (unsigned long long)hi + 1ull
                                       : (unsigned long long)lo + ((unsigned long long)(b - 1) << s0);
            const unsigned long long end =
                b == 0 ? (unsigned long long)lo
                       : b == kNB0 - 1 ? 0x100000000ull
                                       : (unsigned long long)lo + ((unsigned long long)b << s0);
            const bool interior = b >= 1 && b <= kNB0 - 2;
            iv_append(L, w, col, (uint32_t)start, end - start, cum, c, (uint32_t)first,
                      (uint32_t)(i - first), interior && s0 == 0, lo + (uint32_t)(b - 1));
        }
        cum = next;
    }
    iv_store(L, w, col);
}

__global__ void scan_refine_kernel(long long cols, RankPlan rp, Work w) {
    if (*w.pending == 0) return;
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const uint32_t n_old = w.n_iv[col];
    bool any = false;
    for (uint32_t v = 0; v < n_old; ++v) any |= w.iv_status[col * kMaxRanks + v] == IV_REFINE;
    if (!any) return;

    IvList O, L;
    O.n = n_old;
    for (uint32_t v = 0; v < n_old; ++v) {
        const long long x = col * kMaxRanks + v;
        O.klo[v] = w.iv_klo[x]; O.aux[v] = w.iv_aux[x]; O.below[v] = w.iv_below[x];
        O.status[v] = w.iv_status[x]; O.first[v] = w.iv_first[x]; O.nr[v] = w.iv_nr[x];
    }
    L.n = 0;
    for (uint32_t v = 0; v < n_old; ++v) {
        if (O.status[v] != IV_REFINE) {
            const uint32_t i = L.n++;
            L.klo[i] = O.klo[v]; L.aux[i] = O.aux[v]; L.below[i] = O.below[v];
            L.status[i] = O.status[v]; L.first[i] = O.first[v]; L.nr[i] = O.nr[v];
            continue;
        }
        atomicSub(w.pending, 1u);
        uint32_t* h = w.hist1 + (col * kMaxRanks + v) * kNSub;
        const uint32_t sh = O.aux[v];
        uint32_t cum = O.below[v];
        uint32_t i = O.first[v];
        const uint32_t i_end = O.first[v] + O.nr[v];
        for (int b = 0; b < kNSub; ++b) {
            const uint32_t c = h[b];
            h[b] = 0;                                   // ready for the next level
            const uint32_t next = cum + c;
            if (i < i_end && rp.rank[i] < next) {
                const uint32_t first = i;
                while (i < i_end && rp.rank[i] < next) ++i;
                const uint32_t klo = O.klo[v] + ((uint32_t)b << sh);
                iv_append(L, w, col, klo, 1ull << sh, cum, c, first, i - first, sh == 0, klo);
            }
            cum = next;
        }
    }
    iv_store(L, w, col);
}

// ---------------------------------------------------------------------------------------------
// refine / collect share the interval lookup: intervals are disjoint and ordered by key, so the
// candidate interval of key k is the LAST one with klo <= k; unused entries hold klo = 0xFFFFFFFF
// and an aux value that rejects everything.
// ---------------------------------------------------------------------------------------------
struct IvTable {                 // shared memory, [interval][thread]
    uint32_t klo[kMaxRanks][kBX];
    uint32_t aux[kMaxRanks][kBX];
};

template <int M>   // M = number of target ranks of the call = max intervals per column
__global__ void __launch_bounds__(kBX)
refine_kernel(const float* __restrict__ p, long long rows, long long cols,
              long long rows_per_split, Work w) {
    if (*w.pending == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IvTable* tab = reinterpret_cast<IvTable*>(smem_raw);
    uint16_t* hist = reinterpret_cast<uint16_t*>(smem_raw + sizeof(IvTable));   // [M*32][kBX]

    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    const int t = threadIdx.x;
    const bool live = col < cols;
    const uint32_t n = live ? w.n_iv[col] : 0;
    int active = 0;
#pragma unroll
    for (int v = 0; v < M; ++v) {
        uint32_t klo = 0xFFFFFFFFu, sh = 0xFFFFFFFFu;
        if ((uint32_t)v < n) {
            const long long x = col * kMaxRanks + v;
            klo = w.iv_klo[x];
            if (w.iv_status[x] == IV_REFINE) { sh = w.iv_aux[x]; active = 1; }
        }
        tab->klo[v][t] = klo;
        tab->aux[v][t] = sh;
    }
    for (int i = t; i < M * kNSub * kBX; i += kBX) hist[i] = 0;
    if (!__syncthreads_or(active)) return;
    if (!active) return;

    const long long r0 = (long long)blockIdx.y * rows_per_split;
    const long long r1 = min(rows, r0 + rows_per_split);
    const float* src = p + r0 * cols + col;

    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        int v = -1;
#pragma unroll
        for (int q = 0; q < M; ++q) v += (k >= tab->klo[q][t]) ? 1 : 0;
        if (v < 0) return;
        const uint32_t sh = tab->aux[v][t];
        if (sh > 31u) return;
        const uint32_t sub = (k - tab->klo[v][t]) >> sh;
        if (sub < (uint32_t)kNSub) hist[(v * kNSub + sub) * kBX + t] += 1;
    };

    long long r = r0;
    for (; r + kUnroll <= r1; r += kUnroll, src += (long long)kUnroll * cols) {
        float f[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) f[u] = load_stream(src + (long long)u * cols);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) visit(f[u]);
    }
    for (; r < r1; ++r, src += cols) visit(load_stream(src));

    uint32_t* g = w.hist1 + col * (kMaxRanks * kNSub);
    for (uint32_t v = 0; v < n; ++v) {
        if (tab->aux[v][t] > 31u) continue;
        for (int b = 0; b < kNSub; ++b) {
            const uint32_t c = hist[(v * kNSub + b) * kBX + t];
            if (c) atomicAdd(g + v * kNSub + b, c);
        }
    }
}

template <int M>
__global__ void __launch_bounds__(kBX)
collect_kernel(const float* __restrict__ p, long long rows, long long cols,
               long long rows_per_split, Work w) {
    __shared__ IvTable tab;                          // aux = last key of a COLLECT interval
    __shared__ uint8_t is_collect[kMaxRanks][kBX];   // other intervals reject every key
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    const int t = threadIdx.x;
    const bool live = col < cols;
    const uint32_t n = live ? w.n_iv[col] : 0;
    int active = 0;
#pragma unroll
    for (int v = 0; v < M; ++v) {
        uint32_t klo = 0xFFFFFFFFu, khi = 0u;
        uint8_t flag = 0;
        if ((uint32_t)v < n) {
            const long long x = col * kMaxRanks + v;
            klo = w.iv_klo[x];
            if (w.iv_status[x] == IV_COLLECT) { khi = w.iv_aux[x]; flag = 1; active = 1; }
        }
        tab.klo[v][t] = klo;
        tab.aux[v][t] = khi;
        is_collect[v][t] = flag;
    }
    if (!__syncthreads_or(active)) return;
    if (!active) return;

    const long long r0 = (long long)blockIdx.y * rows_per_split;
    const long long r1 = min(rows, r0 + rows_per_split);
    const float* src = p + r0 * cols + col;

    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        int v = -1;
#pragma unroll
        for (int q = 0; q < M; ++q) v += (k >= tab.klo[q][t]) ? 1 : 0;
        if (v < 0) return;
        if (!is_collect[v][t] || k > tab.aux[v][t]) return;
        const long long x = col * kMaxRanks + v;
        const uint32_t pos = atomicAdd(w.cursor + x, 1u);
        if (pos < (uint32_t)kCap) w.cand[x * kCap + pos] = k;
    };

    long long r = r0;
    for (; r + kUnroll <= r1; r += kUnroll, src += (long long)kUnroll * cols) {
        float f[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) f[u] = load_stream(src + (long long)u * cols);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) visit(f[u]);
    }
    for (; r < r1; ++r, src += cols) visit(load_stream(src));
}

// ---------------------------------------------------------------------------------------------
// resolve: one CTA per column; sort each COLLECT interval's candidates, pick ranks, write rows
// ---------------------------------------------------------------------------------------------
constexpr int kResolveThreads = 256;

__global__ void __launch_bounds__(kResolveThreads)
resolve_kernel(long long rows, long long cols, RankPlan rp, StatPlan st, int to_dB, float eps,
               Work w, float* __restrict__ out /* [n_stats][cols] */) {
    __shared__ uint32_t keys[kCap];
    __shared__ uint32_t rkey[kMaxRanks];
    const long long col = blockIdx.x;
    const int t = threadIdx.x;

    if (t < rp.n_ranks) rkey[t] = w.r_key[col * kMaxRanks + t];
    const uint32_t n_iv = rp.n_ranks ? w.n_iv[col] : 0;
    for (uint32_t v = 0; v < n_iv; ++v) {
        const long long x = col * kMaxRanks + v;
        if (w.iv_status[x] != IV_COLLECT) continue;     // RESOLVED ranks are already in r_key
        const uint32_t n = min(w.cursor[x], (uint32_t)kCap);
        int m = 1;
        while (m < (int)n) m <<= 1;
        __syncthreads();
        for (int i = t; i < m; i += kResolveThreads)
            keys[i] = i < (int)n ? w.cand[x * kCap + i] : 0xFFFFFFFFu;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = t; i < m; i += kResolveThreads) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = keys[i], b = keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        const uint32_t first = w.iv_first[x], nr = w.iv_nr[x], below = w.iv_below[x];
        if ((uint32_t)t < nr) {
            const uint32_t pos = rp.rank[first + t] - below;
            rkey[first + t] = pos < n ? keys[pos] : 0xFFFFFFFFu;
        }
    }
    __syncthreads();

    if (t < st.n_stats) {
        float r;
        const int kind = st.kind[t];
        auto xf = [&](uint32_t key) {
            const float v = key_to_float(key);
            return to_dB ? power_to_dB(v, eps) : v;
        };
        if (kind == IQW_STAT_MEAN) {
            r = (float)(w.dsum[col] / (double)rows);
        } else if (kind == IQW_STAT_MAX) {
            r = xf(w.kmax[col]);
        } else if (kind == IQW_STAT_MIN) {
            r = xf(w.kmin[col]);
        } else {
            const float a = xf(rkey[st.ia[t]]), b = xf(rkey[st.ib[t]]);
            if (kind == IQW_STAT_MEDIAN) {
                r = __fmul_rn(__fadd_rn(a, b), 0.5f);
            } else {
                // numpy _lerp: a + (b-a)*g, or b - (b-a)*(1-g) where g >= 0.5; no fused ops
                const float g = st.gamma[t];
                const float d = __fsub_rn(b, a);
                r = g >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)))
                              : __fadd_rn(a, __fmul_rn(d, g));
            }
        }
        out[(long long)t * cols + col] = r;
    }
}

__global__ void init_ff_kernel(uint32_t* a, uint32_t* b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = 0xFFFFFFFFu; b[i] = 0xFFFFFFFFu; }
}

// ---------------------------------------------------------------------------------------------
static int build_plans(const iqw_stat* stats, int n_stats, int64_t rows, RankPlan* rp, StatPlan* st,
                       bool* want_sum) {
    st->n_stats = n_stats;
    *want_sum = false;
    int64_t ranks[2 * kMaxStats];
    int nr = 0;
    for (int i = 0; i < n_stats; ++i) {
        const iqw_stat& s = stats[i];
        st->kind[i] = s.kind;
        st->gamma[i] = s.gamma;
        st->ia[i] = st->ib[i] = 0;
        if (s.kind < 0 || s.kind > IQW_STAT_MEDIAN)
            return fail(IQW_ERR_INVALID, "statistic %d: unknown kind %d", i, s.kind);
        if (s.kind == IQW_STAT_MEAN) *want_sum = true;
        if (s.kind == IQW_STAT_QUANTILE || s.kind == IQW_STAT_MEDIAN) {
            const int64_t lo = s.kind == IQW_STAT_MEDIAN ? (rows - 1) / 2 : s.rank_lo;
            const int64_t hi = s.kind == IQW_STAT_MEDIAN ? rows / 2 : s.rank_hi;
            if (lo < 0 || hi < lo || hi >= rows)
                return fail(IQW_ERR_INVALID, "statistic %d: ranks (%lld, %lld) outside [0, %lld)", i,
                            (long long)lo, (long long)hi, (long long)rows);
            ranks[nr++] = lo;
            ranks[nr++] = hi;
        }
    }
    for (int i = 1; i < nr; ++i)
        for (int j = i; j > 0 && ranks[j] < ranks[j - 1]; --j) {
            int64_t tmp = ranks[j]; ranks[j] = ranks[j - 1]; ranks[j - 1] = tmp;
        }
    int nu = 0;
    for (int i = 0; i < nr; ++i)
        if (nu == 0 || ranks[i] != ranks[nu - 1]) ranks[nu++] = ranks[i];
    if (nu > kMaxRanks)
        return fail(IQW_ERR_UNSUPPORTED,
                    "%d distinct order statistics requested; at most %d per call (split the request)",
                    nu, kMaxRanks);
    rp->n_ranks = nu;
    for (int i = 0; i < nu; ++i) rp->rank[i] = (unsigned)ranks[i];
    for (int i = 0; i < n_stats; ++i) {
        const iqw_stat& s = stats[i];
        if (s.kind != IQW_STAT_QUANTILE && s.kind != IQW_STAT_MEDIAN) continue;
        const int64_t lo = s.kind == IQW_STAT_MEDIAN ? (rows - 1) / 2 : s.rank_lo;
        const int64_t hi = s.kind == IQW_STAT_MEDIAN ? rows / 2 : s.rank_hi;
        for (int k = 0; k < nu; ++k) {
            if (ranks[k] == lo) st->ia[i] = k;
            if (ranks[k] == hi) st->ib[i] = k;
        }
    }
    return IQW_OK;
}

template <int M>
static void launch_refine_collect(dim3 grid, size_t rf_smem, cudaStream_t s, const float* p,
                                  long long rows, long long cols, long long rps, RankPlan rp,
                                  Work w, unsigned cblocks, unsigned cthreads) {
    cudaFuncSetAttribute(refine_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_smem);
    for (int level = 0; level < kRefineLevels; ++level) {
        static const char* const names[kRefineLevels] = {"stats_refine_1", "stats_refine_2", "stats_refine_3",
                                                        "stats_refine_4", "stats_refine_5", "stats_refine_6"};
        { IQW_PROFILE(names[level], s); refine_kernel<M><<<grid, kBX, rf_smem, s>>>(p, rows, cols, rps, w); }
        { IQW_PROFILE("stats_scan", s); scan_refine_kernel<<<cblocks, cthreads, 0, s>>>(cols, rp, w); }
    }
    { IQW_PROFILE("stats_collect", s); collect_kernel<M><<<grid, kBX, 0, s>>>(p, rows, cols, rps, w); }
}

}  // namespace iqw

using namespace iqw;

extern "C" size_t iqw_time_stats_workspace_bytes(int64_t n_channels, int64_t n_rows, int64_t n_cols,
                                                 int32_t n_stats) {
    (void)n_channels; (void)n_rows; (void)n_stats;
    if (n_cols <= 0) return 256;
    return carve_workspace(nullptr, n_cols, nullptr);
}

extern "C" int iqw_time_stats_f32(const float* d_p, int64_t n_channels, int64_t n_rows,
                                  int64_t n_cols, int64_t p_channel_stride, const iqw_stat* stats,
                                  int32_t n_stats, int32_t to_dB, float eps, float* d_out,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_p || !stats || !d_out || !d_workspace) return fail(IQW_ERR_INVALID, "null pointer argument");
    if (n_stats < 1 || n_stats > kMaxStats)
        return fail(IQW_ERR_INVALID, "n_stats=%d outside 1..%d", n_stats, kMaxStats);
    if (n_channels < 0 || n_rows < 1 || n_cols < 1)
        return fail(IQW_ERR_INVALID, "empty matrix (rows=%lld cols=%lld)", (long long)n_rows, (long long)n_cols);
    if (n_rows >= 0xFFFFFFFFll) return fail(IQW_ERR_UNSUPPORTED, "n_rows >= 2^32");
    if (((uintptr_t)d_workspace & 255) != 0) return fail(IQW_ERR_INVALID, "workspace not 256-byte aligned");
    Work w{};
    const size_t need = carve_workspace(d_workspace, n_cols, &w);
    if (workspace_bytes < need)
        return fail(IQW_ERR_WORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need);

    RankPlan rp{};
    StatPlan st{};
    bool want_sum = false;
    if (int rc = build_plans(stats, n_stats, n_rows, &rp, &st, &want_sum)) return rc;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;

    const long long col_tiles = (n_cols + kBX - 1) / kBX;
    // time splits: enough CTAs to fill the machine ~4x, rows per split < 65536 (uint16 counters)
    long long splits = (4ll * sms * 3 + col_tiles - 1) / col_tiles;
    if (splits < 1) splits = 1;
    long long rows_per_split = (n_rows + splits - 1) / splits;
    if (rows_per_split > 65535) rows_per_split = 65535;
    if (rows_per_split < kUnroll) rows_per_split = kUnroll;
    splits = (n_rows + rows_per_split - 1) / rows_per_split;
    if (splits > 65535) return fail(IQW_ERR_UNSUPPORTED, "n_rows too large for the split grid");
    const dim3 grid((unsigned)col_tiles, (unsigned)splits);

    long long row_step = n_rows / 2048;
    if (row_step < 1) row_step = 1;
    const long long sample_rows = (n_rows + row_step - 1) / row_step;
    long long rsplits = sample_rows / 64;
    if (rsplits < 1) rsplits = 1;
    if (rsplits > 64) rsplits = 64;

    const size_t l0_smem = sizeof(uint16_t) * kNB0 * kBX;
    const int m_pad = rp.n_ranks <= 2 ? 2 : rp.n_ranks <= 4 ? 4 : 8;
    const size_t rf_smem = sizeof(IvTable) + sizeof(uint16_t) * m_pad * kNSub * kBX;
    IQW_CUDA_OK(cudaFuncSetAttribute(l0_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l0_smem));
    IQW_CUDA_OK(cudaFuncSetAttribute(l0_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l0_smem));
    IQW_CUDA_OK(cudaFuncSetAttribute(l0_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l0_smem));

    const unsigned cthreads = 128;
    const unsigned cblocks = (unsigned)((n_cols + cthreads - 1) / cthreads);

    for (int64_t c = 0; c < n_channels; ++c) {
        const float* p = d_p + c * p_channel_stride;
        float* out = d_out + c * (int64_t)n_stats * n_cols;

        IQW_CUDA_OK(cudaMemsetAsync(d_workspace, 0, w.zero_bytes, s));
        { IQW_PROFILE("stats_init", s); init_ff_kernel<<<cblocks, cthreads, 0, s>>>(w.range_lo, w.kmin, n_cols); }

        { IQW_PROFILE("stats_range", s); range_kernel<<<dim3((unsigned)col_tiles, (unsigned)rsplits), kBX, 0, s>>>(p, n_rows, n_cols, row_step, w); }
        {
        IQW_PROFILE("stats_l0", s);
        if (want_sum && to_dB)
            l0_kernel<true, true><<<grid, kBX, l0_smem, s>>>(p, n_rows, n_cols, rows_per_split, eps, w);
        else if (want_sum)
            l0_kernel<true, false><<<grid, kBX, l0_smem, s>>>(p, n_rows, n_cols, rows_per_split, eps, w);
        else
            l0_kernel<false, false><<<grid, kBX, l0_smem, s>>>(p, n_rows, n_cols, rows_per_split, eps, w);
        }

        if (rp.n_ranks > 0) {
            { IQW_PROFILE("stats_scan", s); scan0_kernel<<<cblocks, cthreads, 0, s>>>(n_cols, rp, w); }
            if (m_pad == 2)
                launch_refine_collect<2>(grid, rf_smem, s, p, n_rows, n_cols, rows_per_split, rp, w, cblocks, cthreads);
            else if (m_pad == 4)
                launch_refine_collect<4>(grid, rf_smem, s, p, n_rows, n_cols, rows_per_split, rp, w, cblocks, cthreads);
            else
                launch_refine_collect<8>(grid, rf_smem, s, p, n_rows, n_cols, rows_per_split, rp, w, cblocks, cthreads);
        }
        { IQW_PROFILE("stats_resolve", s); resolve_kernel<<<(unsigned)n_cols, kResolveThreads, 0, s>>>(n_rows, n_cols, rp, st, to_dB, eps, w, out); }
        IQW_CUDA_OK(cudaGetLastError());
    }
    return IQW_OK;
}
