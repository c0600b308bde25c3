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
// EXACT PIPELINE (any row set):
//   range   : min/max key of ~2k rows per column  -> bucket map of that column
//   L0      : 256-bucket histogram of every element (bucket 0 / 255 catch keys outside the range
//             seen by `range`, so counts are exact whatever it missed)
//   scan0   : per column, group the target ranks by bucket -> intervals
//   R (x 7) : 32 sub-buckets per pending interval, all intervals in one pass; the scan that follows
//             splits each interval by sub-bucket.  Kernels exit at once when nothing is pending;
//             7 levels always suffice (32 key bits / 5 bits per level)
//   collect : keys inside the final intervals -> candidate lists (<= CAP each)
//   resolve : bitonic sort of the candidates in shared memory -> key of every target rank
//   finalize: dB, numpy lerp, mean/max/min, store
//
// LONG COLUMNS (rows >= 32768): two full reads instead of four.
//   1. the exact pipeline above runs on a pseudo-random 1-in-k ROW SAMPLE (~16k rows) and returns
//      sample order statistics that bracket every target rank with a 6-sigma margin
//   2. `bracket` pass over ALL rows: per column, exact count of keys below / between / above the
//      brackets and a 32-bucket histogram inside each bracket (+ exact min, max, sum)
//   3. scan: the counts say exactly which sub-bucket (or, if a bracket missed, which gap) holds
//      each rank -> intervals of a few hundred keys; a missed bracket simply becomes a wide
//      interval that the R levels refine, so the result is exact whatever the sample looked like
//   4. collect + resolve + finalize as above
//
// Keys are the order-preserving uint32 image of the float (iqw_common.cuh float_to_key), so the
// selection is exact for any input, ties and signed zeros included.  NaNs sort above +inf.
#include <cmath>
#include "iqw_common.cuh"

namespace iqw {

constexpr int kMaxRanks = 8;       // distinct target ranks (and therefore intervals) per call
constexpr int kMaxStats = 32;      // output rows per call
constexpr int kNB0 = 256;          // level-0 buckets
constexpr int kNSub = 32;          // sub-buckets per interval and refinement level
constexpr int kCap = 2048;         // candidates kept per (column, interval)
constexpr int kRefineLevels = 7;   // 32 key bits / 5 bits per level
constexpr int kBX = 128;           // columns (= threads) per CTA in the streaming passes
constexpr int kUnroll = 8;         // rows in flight per thread
constexpr long long kSampleMinRows = 32768;   // below this the exact pipeline reads all rows
constexpr long long kSampleRows = 16384;      // target size of the row sample
constexpr int kMaxGroups = 4;      // rank groups (brackets) per call on the sampled path

enum IvStatus : uint32_t { IV_REFINE = 0, IV_COLLECT = 1, IV_RESOLVED = 2, IV_BRACKET = 3 };

struct RankPlan {                  // same for every column: depends only on the row count
    int n_ranks;
    unsigned int rank[kMaxRanks];  // ascending, distinct
};

struct StatPlan {
    int n_stats;
    int kind[kMaxStats];
    int ia[kMaxStats], ib[kMaxStats];   // indices into RankPlan::rank
    float gamma[kMaxStats];
};

// which sample order statistics bracket which group of full-matrix ranks
struct BracketPlan {
    int n_groups;
    int first[kMaxGroups], nr[kMaxGroups];       // target ranks [first, first+nr) of the full plan
    int s_lo[kMaxGroups], s_hi[kMaxGroups];      // indices into the SAMPLE RankPlan
    int open_lo[kMaxGroups], open_hi[kMaxGroups];  // bracket extends to the end of the key space
};

// rows visited by a streaming pass: all of them (step == 1) or one pseudo-random row out of every
// `step` consecutive rows (a stratified sample; the jitter defeats periodic captures)
struct RowMap {
    long long n;       // rows visited
    long long step;
    uint32_t seed;
};
__device__ __forceinline__ long long map_row(const RowMap& m, long long i) {
    uint32_t h = (uint32_t)i * 2654435761u + m.seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return i * m.step + (long long)(h % (uint32_t)m.step);
}

// per-pipeline state, carved from the caller's workspace
struct Work {
    uint32_t* range_lo;   // [cols]
    uint32_t* range_hi;   // [cols]
    uint32_t* kmin;       // [cols]
    uint32_t* kmax;       // [cols]
    double* dsum;         // [cols]
    uint32_t* hist0;      // [cols][256]
    uint32_t* n_iv;       // [cols]
    uint32_t* iv_klo;     // [cols][8]  first key of the interval
    uint32_t* iv_khi;     // [cols][8]  last key (COLLECT, BRACKET)
    uint32_t* iv_shift;   // [cols][8]  sub-bucket shift (REFINE, BRACKET)
    uint32_t* iv_below;   // [cols][8]  number of keys < klo in the column
    uint32_t* iv_status;  // [cols][8]
    uint32_t* iv_first;   // [cols][8]  first target-rank index inside
    uint32_t* iv_nr;      // [cols][8]  number of target ranks inside
    uint32_t* iv_cnt;     // [cols][8]  number of keys inside (COLLECT: checked against the cursor)
    uint32_t* r_key;      // [cols][8]  key of each target rank once known
    uint32_t* hist1;      // [cols][8][32]
    uint32_t* gap;        // [cols][9]   bracket pass: keys below / between / above the brackets
    uint32_t* cursor;     // [cols][8]
    uint32_t* cand;       // [cols][8][kCap]
    uint32_t* pending;    // [1] number of intervals in IV_REFINE
    size_t zero_bytes;    // leading bytes that must be zero before a pipeline starts
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve_work(void* base, int64_t cols, Work* w) {
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
    t.gap = (uint32_t*)take(4 * c * (m + 1));
    t.cursor = (uint32_t*)take(4 * c * m);
    t.n_iv = (uint32_t*)take(4 * c);
    t.pending = (uint32_t*)take(256);
    t.zero_bytes = off;
    // --- 0xFF-initialised ---
    t.range_lo = (uint32_t*)take(4 * c);
    t.kmin = (uint32_t*)take(4 * c);
    // --- written before read ---
    t.iv_klo = (uint32_t*)take(4 * c * m);
    t.iv_khi = (uint32_t*)take(4 * c * m);
    t.iv_shift = (uint32_t*)take(4 * c * m);
    t.iv_below = (uint32_t*)take(4 * c * m);
    t.iv_status = (uint32_t*)take(4 * c * m);
    t.iv_first = (uint32_t*)take(4 * c * m);
    t.iv_nr = (uint32_t*)take(4 * c * m);
    t.iv_cnt = (uint32_t*)take(4 * c * m);
    t.r_key = (uint32_t*)take(4 * c * m);
    t.cand = (uint32_t*)take(4 * c * m * kCap);
    if (w) *w = t;
    return off;
}

// ---------------------------------------------------------------------------------------------
// streaming skeleton: thread `col` visits rows i0..i1 of the row map, kUnroll loads in flight
// ---------------------------------------------------------------------------------------------
template <bool SAMPLED, typename F>
__device__ __forceinline__ void stream_column(const float* __restrict__ p, long long cols,
                                              long long col, const RowMap& rm, long long i0,
                                              long long i1, F&& visit) {
    long long i = i0;
    if (!SAMPLED) {
        const float* src = p + i0 * cols + col;
        for (; i + kUnroll <= i1; i += kUnroll, src += (long long)kUnroll * cols) {
            float f[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) f[u] = __ldcs(src + (long long)u * cols);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) visit(f[u]);
        }
        for (; i < i1; ++i, src += cols) visit(__ldcs(src));
    } else {
        for (; i + kUnroll <= i1; i += kUnroll) {
            float f[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) f[u] = __ldg(p + map_row(rm, i + u) * cols + col);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) visit(f[u]);
        }
        for (; i < i1; ++i) visit(__ldg(p + map_row(rm, i) * cols + col));
    }
}

#define IQW_STREAM(SAMPLED_FLAG, ...)                                                      \
    do {                                                                                   \
        if (SAMPLED_FLAG) stream_column<true>(__VA_ARGS__);                                \
        else stream_column<false>(__VA_ARGS__);                                            \
    } while (0)

// ---------------------------------------------------------------------------------------------
// range: min / max key over ~2k of the visited rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBX)
range_kernel(const float* __restrict__ p, long long cols, RowMap rm, long long stride, Work w) {
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    if (col >= cols) return;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (long long i = (long long)blockIdx.y * stride; i < rm.n; i += stride * gridDim.y) {
        const long long r = rm.step == 1 ? i : map_row(rm, i);
        const uint32_t k = float_to_key(__ldg(p + r * cols + col));
        lo = min(lo, k);
        hi = max(hi, k);
    }
    atomicMin(w.range_lo + col, lo);
    atomicMax(w.range_hi + col, hi);
}

// bucket map of level 0: 0 = below the range, 255 = above, 1..254 inside
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

// exact min / max / sum of a column (named statistics), shared by L0 and the bracket pass
template <bool WANT_SUM, bool TO_DB>
struct Named {
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    double dsum = 0.0;
    float part = 0.f;
    int n_part = 0;
    __device__ __forceinline__ void add(float f, uint32_t k, float eps) {
        kmin = min(kmin, k);
        kmax = max(kmax, k);
        if (WANT_SUM) {
            part += TO_DB ? power_to_dB(f, eps) : f;
            if (++n_part == 8) { dsum += (double)part; part = 0.f; n_part = 0; }
        }
    }
    __device__ __forceinline__ void flush(const Work& w, long long col) {
        atomicMin(w.kmin + col, kmin);
        atomicMax(w.kmax + col, kmax);
        if (WANT_SUM) atomicAdd(w.dsum + col, dsum + (double)part);
    }
};

// ---------------------------------------------------------------------------------------------
// L0: private 256-bucket histogram per column (+ named statistics when NAMED)
// ---------------------------------------------------------------------------------------------
template <bool NAMED, bool WANT_SUM, bool TO_DB>
__global__ void __launch_bounds__(kBX)
l0_kernel(const float* __restrict__ p, long long cols, RowMap rm, long long rows_per_split,
          float eps, Work w) {
    extern __shared__ uint16_t hist[];   // [kNB0][kBX]
    for (int i = threadIdx.x; i < kNB0 * kBX; i += kBX) hist[i] = 0;
    __syncthreads();

    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    if (col >= cols) return;
    const uint32_t lo = w.range_lo[col], hi = w.range_hi[col];
    const uint32_t s = l0_shift(lo, hi);
    uint16_t* h = hist + threadIdx.x;
    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(rm.n, i0 + rows_per_split);
    Named<WANT_SUM, TO_DB> named;

    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        if (NAMED) named.add(f, k, eps);
        h[l0_bucket(k, lo, hi, s) * kBX] += 1;
    };
    IQW_STREAM(rm.step != 1, p, cols, col, rm, i0, i1, visit);

    uint32_t* g = w.hist0 + col * kNB0;
    for (int b = 0; b < kNB0; ++b) {
        const uint32_t c = h[b * kBX];
        if (c) atomicAdd(g + b, c);
    }
    if (NAMED) named.flush(w, col);
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
    uint32_t klo[kMaxRanks], khi[kMaxRanks], shift[kMaxRanks], below[kMaxRanks],
        status[kMaxRanks], first[kMaxRanks], nr[kMaxRanks], cnt[kMaxRanks];
};

// append the interval [klo, klo + span) holding `cnt` keys, `below` keys under it, and target
// ranks [first, first + nr).  `single` = every key in it equals `key_exact`.
__device__ void iv_append(IvList& L, const Work& w, long long col, uint32_t klo,
                          unsigned long long span, uint32_t below, uint32_t cnt, uint32_t first,
                          uint32_t nr, bool single, uint32_t key_exact) {
    const uint32_t i = L.n++;
    const unsigned long long last = (unsigned long long)klo + span - 1ull;
    L.klo[i] = klo;
    L.khi[i] = last > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)last;
    L.below[i] = below;
    L.first[i] = first;
    L.nr[i] = nr;
    L.cnt[i] = cnt;
    L.shift[i] = 0;
    if (single || span == 1ull) {
        L.status[i] = IV_RESOLVED;
        for (uint32_t q = 0; q < nr; ++q) w.r_key[col * kMaxRanks + first + q] = single ? key_exact : klo;
    } else if (cnt <= (uint32_t)kCap) {
        L.status[i] = IV_COLLECT;
    } else {
        L.status[i] = IV_REFINE;
        const uint32_t l = ceil_log2_u64(span);
        L.shift[i] = l > 5 ? l - 5 : 0;
        atomicAdd(w.pending, 1u);
    }
}

__device__ void iv_store(const IvList& L, const Work& w, long long col) {
    w.n_iv[col] = L.n;
    for (uint32_t i = 0; i < L.n; ++i) {
        const long long x = col * kMaxRanks + i;
        w.iv_klo[x] = L.klo[i];
        w.iv_khi[x] = L.khi[i];
        w.iv_shift[x] = L.shift[i];
        w.iv_below[x] = L.below[i];
        w.iv_status[x] = L.status[i];
        w.iv_first[x] = L.first[i];
        w.iv_nr[x] = L.nr[i];
        w.iv_cnt[x] = L.cnt[i];
    }
}

__device__ void iv_load(IvList& L, const Work& w, long long col) {
    L.n = w.n_iv[col];
    for (uint32_t i = 0; i < L.n; ++i) {
        const long long x = col * kMaxRanks + i;
        L.klo[i] = w.iv_klo[x]; L.khi[i] = w.iv_khi[x]; L.shift[i] = w.iv_shift[x];
        L.below[i] = w.iv_below[x]; L.status[i] = w.iv_status[x];
        L.first[i] = w.iv_first[x]; L.nr[i] = w.iv_nr[x]; L.cnt[i] = w.iv_cnt[x];
    }
}

// split the interval O[v] by its sub-bucket counts h[0..kNSub) (zeroed on the way), `cum` = number
// of keys below the first sub-bucket on entry, advanced past the interval on exit
__device__ void split_by_subbuckets(IvList& L, const IvList& O, uint32_t v, uint32_t* h,
                                    const RankPlan& rp, uint32_t& i, uint32_t& cum, const Work& w,
                                    long long col) {
    const uint32_t sh = O.shift[v];
    const uint32_t i_end = O.first[v] + O.nr[v];
    for (int b = 0; b < kNSub; ++b) {
        const uint32_t c = h[b];
        h[b] = 0;
        const uint32_t next = cum + c;
        if (i < i_end && rp.rank[i] < next) {
            const uint32_t first = i;
            while (i < i_end && rp.rank[i] < next) ++i;
            // the sub-bucket is clamped to its parent: keys above the parent's last key were NOT
            // counted in c (they belong to the gap / interval that follows)
            const uint32_t klo = O.klo[v] + ((uint32_t)b << sh);
            unsigned long long span = 1ull << sh;
            const unsigned long long room = (unsigned long long)O.khi[v] - klo + 1ull;
            if (span > room) span = room;
            iv_append(L, w, col, klo, span, cum, c, first, i - first, sh == 0, klo);
        }
        cum = next;
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
            unsigned long long end =
                b == 0 ? (unsigned long long)lo
                       : b == kNB0 - 1 ? 0x100000000ull
                                       : (unsigned long long)lo + ((unsigned long long)b << s0);
            // keys above `hi` were counted in the last bucket, not in the interior bucket that
            // straddles hi
            if (b >= 1 && b <= kNB0 - 2 && end > (unsigned long long)hi + 1ull) end = (unsigned long long)hi + 1ull;
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
    IvList O, L;
    iv_load(O, w, col);
    bool any = false;
    for (uint32_t v = 0; v < O.n; ++v) any |= O.status[v] == IV_REFINE;
    if (!any) return;
    L.n = 0;
    for (uint32_t v = 0; v < O.n; ++v) {
        if (O.status[v] != IV_REFINE) {
            const uint32_t i = L.n++;
            L.klo[i] = O.klo[v]; L.khi[i] = O.khi[v]; L.shift[i] = O.shift[v]; L.below[i] = O.below[v];
            L.status[i] = O.status[v]; L.first[i] = O.first[v]; L.nr[i] = O.nr[v]; L.cnt[i] = O.cnt[v];
            continue;
        }
        atomicSub(w.pending, 1u);
        uint32_t cum = O.below[v], i = O.first[v];
        split_by_subbuckets(L, O, v, w.hist1 + (col * kMaxRanks + v) * kNSub, rp, i, cum, w, col);
    }
    iv_store(L, w, col);
}

// after the bracket pass: regions in key order are gap0, bracket0, gap1, bracket1, ..., gapM with
// exact counts; locate every target rank in its region
__global__ void scan_bracket_kernel(long long cols, RankPlan rp, Work w) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    IvList O, L;
    iv_load(O, w, col);
    L.n = 0;
    uint32_t cum = 0, i = 0;
    uint32_t* gap = w.gap + col * (kMaxRanks + 1);
    for (uint32_t g = 0; g <= O.n; ++g) {
        const uint32_t c = gap[g];
        gap[g] = 0;
        const uint32_t next = cum + c;
        if (i < (uint32_t)rp.n_ranks && rp.rank[i] < next) {      // the bracket missed: rank is in a gap
            const uint32_t first = i;
            while (i < (uint32_t)rp.n_ranks && rp.rank[i] < next) ++i;
            const unsigned long long start = g == 0 ? 0ull : (unsigned long long)O.khi[g - 1] + 1ull;
            const unsigned long long end = g == O.n ? 0x100000000ull : (unsigned long long)O.klo[g];
            iv_append(L, w, col, (uint32_t)start, end - start, cum, c, first, i - first, false, 0u);
        }
        cum = next;
        if (g == O.n) break;
        // ranks inside bracket g; its [first, nr) covers the ranks the plan aimed at it, but after
        // a miss any rank may land here, so walk with the global rank cursor
        IvList T = O;
        T.first[g] = i;
        T.nr[g] = (uint32_t)rp.n_ranks - i;
        split_by_subbuckets(L, T, g, w.hist1 + (col * kMaxRanks + g) * kNSub, rp, i, cum, w, col);
    }
    iv_store(L, w, col);
}

// ---------------------------------------------------------------------------------------------
// interval lookup shared by refine / bracket / collect: intervals are disjoint and ordered by key;
// the candidate of key k is the LAST interval with klo <= k.  Bounds live in registers.
// ---------------------------------------------------------------------------------------------
template <int M>
struct IvRegs {
    uint32_t klo[M], khi[M], shift[M];
    uint32_t n;
    __device__ __forceinline__ void load(const Work& w, long long col, bool live, uint32_t want_status,
                                         int& active) {
        n = live ? w.n_iv[col] : 0;
#pragma unroll
        for (int v = 0; v < M; ++v) {
            // unused / not-wanted entries keep their klo (the lookup needs the ordering) but carry
            // shift = 0x80000000, which every consumer treats as "reject"
            klo[v] = 0xFFFFFFFFu; khi[v] = 0u; shift[v] = 0x80000000u;
            if ((uint32_t)v < n) {
                const long long x = col * kMaxRanks + v;
                klo[v] = w.iv_klo[x];
                if (w.iv_status[x] == want_status) {
                    khi[v] = w.iv_khi[x];
                    shift[v] = w.iv_shift[x];
                    active = 1;
                }
            }
        }
    }
    // returns interval index or -1; sets (kl, kh, sh) of that interval
    __device__ __forceinline__ int find(uint32_t k, uint32_t& kl, uint32_t& kh, uint32_t& sh) const {
        int v = -1;
        kl = 0; kh = 0; sh = 0x80000000u;
#pragma unroll
        for (int q = 0; q < M; ++q) {
            const bool ge = k >= klo[q];
            v = ge ? q : v;
            kl = ge ? klo[q] : kl;
            kh = ge ? khi[q] : kh;
            sh = ge ? shift[q] : sh;
        }
        return v;
    }
};

// refine: 32 sub-buckets per pending interval, private counters [interval*32+sub][thread]
template <int M>
__global__ void __launch_bounds__(kBX)
refine_kernel(const float* __restrict__ p, long long cols, RowMap rm, long long rows_per_split,
              Work w) {
    if (*w.pending == 0) return;
    extern __shared__ uint16_t hist[];   // [M*32][kBX]
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    const int t = threadIdx.x;
    int active = 0;
    IvRegs<M> iv;
    iv.load(w, col, col < cols, IV_REFINE, active);
    for (int i = t; i < M * kNSub * kBX; i += kBX) hist[i] = 0;
    if (!__syncthreads_or(active)) return;
    if (!active) return;

    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(rm.n, i0 + rows_per_split);
    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        uint32_t kl, kh, sh;
        const int v = iv.find(k, kl, kh, sh);
        if (v >= 0 && k <= kh && sh < 32u) hist[(v * kNSub + ((k - kl) >> sh)) * kBX + t] += 1;
    };
    IQW_STREAM(rm.step != 1, p, cols, col, rm, i0, i1, visit);

    uint32_t* g = w.hist1 + col * (kMaxRanks * kNSub);
#pragma unroll
    for (int v = 0; v < M; ++v) {
        if (iv.shift[v] >= 32u) continue;
        for (int b = 0; b < kNSub; ++b) {
            const uint32_t c = hist[(v * kNSub + b) * kBX + t];
            if (c) atomicAdd(g + v * kNSub + b, c);
        }
    }
}

// bracket pass over all rows: histogram inside each bracket, counts of the gaps, named statistics
template <int M, bool WANT_SUM, bool TO_DB>
__global__ void __launch_bounds__(kBX)
bracket_kernel(const float* __restrict__ p, long long cols, RowMap rm, long long rows_per_split,
               float eps, Work w) {
    extern __shared__ uint16_t hist[];   // [M*32 + M + 1][kBX]
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    const int t = threadIdx.x;
    int active = 0;
    IvRegs<M> iv;
    iv.load(w, col, col < cols, IV_BRACKET, active);
    for (int i = t; i < (M * kNSub + M + 1) * kBX; i += kBX) hist[i] = 0;
    __syncthreads();
    if (col >= cols) return;

    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(rm.n, i0 + rows_per_split);
    Named<WANT_SUM, TO_DB> named;
    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        named.add(f, k, eps);
        uint32_t kl, kh, sh;
        const int v = iv.find(k, kl, kh, sh);
        // inside bracket v -> its sub-bucket; otherwise the gap that follows bracket v (gap v+1)
        const int slot = (v >= 0 && k <= kh) ? v * kNSub + (int)((k - kl) >> sh) : M * kNSub + v + 1;
        hist[slot * kBX + t] += 1;
    };
    IQW_STREAM(false, p, cols, col, rm, i0, i1, visit);

    uint32_t* g = w.hist1 + col * (kMaxRanks * kNSub);
#pragma unroll
    for (int v = 0; v < M; ++v) {
        if ((uint32_t)v >= iv.n) break;
        for (int b = 0; b < kNSub; ++b) {
            const uint32_t c = hist[(v * kNSub + b) * kBX + t];
            if (c) atomicAdd(g + v * kNSub + b, c);
        }
    }
    uint32_t* gg = w.gap + col * (kMaxRanks + 1);
    for (uint32_t v = 0; v <= iv.n; ++v) {
        const uint32_t c = hist[(M * kNSub + v) * kBX + t];
        if (c) atomicAdd(gg + v, c);
    }
    named.flush(w, col);
}

template <int M>
__global__ void __launch_bounds__(kBX)
collect_kernel(const float* __restrict__ p, long long cols, RowMap rm, long long rows_per_split,
               Work w) {
    const long long col = (long long)blockIdx.x * kBX + threadIdx.x;
    int active = 0;
    IvRegs<M> iv;
    iv.load(w, col, col < cols, IV_COLLECT, active);
    if (!__syncthreads_or(active)) return;
    if (!active) return;

    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(rm.n, i0 + rows_per_split);
    auto visit = [&](float f) {
        const uint32_t k = float_to_key(f);
        uint32_t kl, kh, sh;
        const int v = iv.find(k, kl, kh, sh);
        if (v >= 0 && k <= kh && sh < 32u) {
            const long long x = col * kMaxRanks + v;
            const uint32_t pos = atomicAdd(w.cursor + x, 1u);
            if (pos < (uint32_t)kCap) w.cand[x * kCap + pos] = k;
        }
    };
    IQW_STREAM(rm.step != 1, p, cols, col, rm, i0, i1, visit);
}

// ---------------------------------------------------------------------------------------------
// resolve: one CTA per column; sort each COLLECT interval's candidates -> key of every rank
// ---------------------------------------------------------------------------------------------
constexpr int kResolveThreads = 256;

__global__ void __launch_bounds__(kResolveThreads)
resolve_kernel(long long cols, RankPlan rp, Work w) {
    __shared__ uint32_t keys[kCap];
    const long long col = blockIdx.x;
    const int t = threadIdx.x;
    const uint32_t n_iv = w.n_iv[col];
    for (uint32_t v = 0; v < n_iv; ++v) {
        const long long x = col * kMaxRanks + v;
        if (w.iv_status[x] != IV_COLLECT) continue;     // RESOLVED ranks are already in r_key
        // invariant: the collect pass found exactly the keys the histograms counted; if not, the
        // ranks of this interval are poisoned (NaN) instead of silently wrong
        const bool sane = w.cursor[x] == w.iv_cnt[x] && w.cursor[x] <= (uint32_t)kCap;
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
            w.r_key[col * kMaxRanks + first + t] = (sane && pos < n) ? keys[pos] : 0xFFFFFFFFu;
        }
    }
}

// sample order statistics -> brackets of the full pass (one thread per column)
__global__ void make_brackets_kernel(long long cols, BracketPlan bp, Work ws, Work wf) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    IvList L;
    L.n = 0;
    for (int g = 0; g < bp.n_groups; ++g) {
        const uint32_t lo = bp.open_lo[g] ? 0u : ws.r_key[col * kMaxRanks + bp.s_lo[g]];
        const uint32_t hi = bp.open_hi[g] ? 0xFFFFFFFFu : ws.r_key[col * kMaxRanks + bp.s_hi[g]];
        if (L.n > 0 && lo <= L.khi[L.n - 1]) {          // overlaps the previous bracket: merge
            const uint32_t i = L.n - 1;
            L.khi[i] = max(L.khi[i], hi);
            L.nr[i] += bp.nr[g];
        } else {
            const uint32_t i = L.n++;
            L.klo[i] = lo; L.khi[i] = max(hi, lo);
            L.first[i] = bp.first[g]; L.nr[i] = bp.nr[g];
            L.below[i] = 0; L.cnt[i] = 0; L.status[i] = IV_BRACKET;
        }
    }
    for (uint32_t i = 0; i < L.n; ++i) {
        const unsigned long long span = (unsigned long long)L.khi[i] - L.klo[i] + 1ull;
        const uint32_t l = ceil_log2_u64(span);
        L.shift[i] = l > 5 ? l - 5 : 0;
    }
    iv_store(L, wf, col);
}

// final rows: dB, numpy lerp, named statistics
__global__ void finalize_kernel(long long rows, long long cols, StatPlan st, int to_dB, float eps,
                                Work w, float* __restrict__ out /* [n_stats][cols] */) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    auto xf = [&](uint32_t key) {
        const float v = key_to_float(key);
        return to_dB ? power_to_dB(v, eps) : v;
    };
    for (int t = 0; t < st.n_stats; ++t) {
        float r;
        const int kind = st.kind[t];
        if (kind == IQW_STAT_MEAN) {
            r = (float)(w.dsum[col] / (double)rows);
        } else if (kind == IQW_STAT_MAX) {
            r = xf(w.kmax[col]);
        } else if (kind == IQW_STAT_MIN) {
            r = xf(w.kmin[col]);
        } else {
            const float a = xf(w.r_key[col * kMaxRanks + st.ia[t]]);
            const float b = xf(w.r_key[col * kMaxRanks + st.ib[t]]);
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
// host side
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

// sample plan: group the target ranks, bracket each group by two sample order statistics a
// 6-sigma (+2) margin away.  Returns false when the sampled path does not apply.
static double g_margin_sigmas = 6.0;   // test aid: iqw_debug_set_sample_margin
static int g_margin_extra = 2;

static bool build_sample_plan(const RankPlan& rp, int64_t rows, int64_t srows, RankPlan* srp,
                              BracketPlan* bp) {
    const double f = (double)srows / (double)rows;
    auto margin = [&](double r) {
        const double q = r / (double)rows;
        return (int64_t)std::ceil(g_margin_sigmas * std::sqrt((double)srows * q * (1.0 - q))) + g_margin_extra;
    };
    int64_t lo[kMaxRanks], hi[kMaxRanks];
    int first[kMaxRanks], nr[kMaxRanks], ng = 0;
    for (int i = 0; i < rp.n_ranks; ++i) {
        const int64_t a = (int64_t)std::floor(rp.rank[i] * f) - margin(rp.rank[i]);
        const int64_t b = (int64_t)std::ceil(rp.rank[i] * f) + margin(rp.rank[i]);
        if (ng > 0 && a <= hi[ng - 1]) {               // overlaps the previous group: same bracket
            hi[ng - 1] = b > hi[ng - 1] ? b : hi[ng - 1];
            nr[ng - 1] += 1;
        } else {
            lo[ng] = a; hi[ng] = b; first[ng] = i; nr[ng] = 1; ++ng;
        }
    }
    if (ng > kMaxGroups) return false;
    bp->n_groups = ng;
    int64_t sr[2 * kMaxGroups];
    int ns = 0;
    for (int g = 0; g < ng; ++g) {
        bp->first[g] = first[g]; bp->nr[g] = nr[g];
        bp->open_lo[g] = lo[g] <= 0;
        bp->open_hi[g] = hi[g] >= srows - 1;
        if (!bp->open_lo[g]) sr[ns++] = lo[g];
        if (!bp->open_hi[g]) sr[ns++] = hi[g];
    }
    // sample ranks are ascending by construction (groups do not overlap); dedupe defensively
    int nu = 0;
    for (int i = 0; i < ns; ++i)
        if (nu == 0 || sr[i] != sr[nu - 1]) sr[nu++] = sr[i];
    srp->n_ranks = nu;
    for (int i = 0; i < nu; ++i) srp->rank[i] = (unsigned)sr[i];
    for (int g = 0; g < ng; ++g) {
        bp->s_lo[g] = bp->s_hi[g] = 0;
        for (int k = 0; k < nu; ++k) {
            if (!bp->open_lo[g] && sr[k] == lo[g]) bp->s_lo[g] = k;
            if (!bp->open_hi[g] && sr[k] == hi[g]) bp->s_hi[g] = k;
        }
    }
    return true;
}

struct Grid {
    dim3 grid;
    long long rows_per_split;
};

static int plan_grid(long long n_rows, long long col_tiles, int sms, Grid* g) {
    // time splits: enough CTAs to fill the machine a few times, < 65536 rows each (uint16 counters)
    long long splits = (4ll * sms * 4 + col_tiles - 1) / col_tiles;
    if (splits < 1) splits = 1;
    long long rps = (n_rows + splits - 1) / splits;
    if (rps > 65535) rps = 65535;
    if (rps < 4 * kUnroll) rps = 4 * kUnroll;
    splits = (n_rows + rps - 1) / rps;
    if (splits > 65535) return fail(IQW_ERR_UNSUPPORTED, "too many rows for the split grid");
    g->grid = dim3((unsigned)col_tiles, (unsigned)splits);
    g->rows_per_split = rps;
    return IQW_OK;
}

template <int M>
static void launch_refine_levels(const Grid& g, cudaStream_t s, const float* p, long long cols,
                                 RowMap rm, const RankPlan& rp, const Work& w, unsigned cblocks,
                                 unsigned cthreads, const char* tag) {
    static const char* const names[2][kRefineLevels] = {
        {"stats_refine_1", "stats_refine_2", "stats_refine_3", "stats_refine_4", "stats_refine_5",
         "stats_refine_6", "stats_refine_7"},
        {"sample_refine_1", "sample_refine_2", "sample_refine_3", "sample_refine_4",
         "sample_refine_5", "sample_refine_6", "sample_refine_7"}};
    const int which = tag[1] == 'a' ? 1 : 0;   // "sample" vs "stats"
    const size_t smem = sizeof(uint16_t) * M * kNSub * kBX;
    cudaFuncSetAttribute(refine_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int level = 0; level < kRefineLevels; ++level) {
        { IQW_PROFILE(names[which][level], s);
          refine_kernel<M><<<g.grid, kBX, smem, s>>>(p, cols, rm, g.rows_per_split, w); }
        { IQW_PROFILE(which ? "sample_scan" : "stats_scan", s);
          scan_refine_kernel<<<cblocks, cthreads, 0, s>>>(cols, rp, w); }
    }
}

template <int M>
static void launch_collect(const Grid& g, cudaStream_t s, const float* p, long long cols, RowMap rm,
                           const Work& w, const char* name) {
    IQW_PROFILE(name, s);
    collect_kernel<M><<<g.grid, kBX, 0, s>>>(p, cols, rm, g.rows_per_split, w);
}

template <int M>
static void launch_bracket(const Grid& g, cudaStream_t s, const float* p, long long cols, RowMap rm,
                           bool want_sum, bool to_dB, float eps, const Work& w) {
    const size_t smem = sizeof(uint16_t) * (M * kNSub + M + 1) * kBX;
    IQW_PROFILE("stats_bracket", s);
    if (want_sum && to_dB) {
        cudaFuncSetAttribute(bracket_kernel<M, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bracket_kernel<M, true, true><<<g.grid, kBX, smem, s>>>(p, cols, rm, g.rows_per_split, eps, w);
    } else if (want_sum) {
        cudaFuncSetAttribute(bracket_kernel<M, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bracket_kernel<M, true, false><<<g.grid, kBX, smem, s>>>(p, cols, rm, g.rows_per_split, eps, w);
    } else {
        cudaFuncSetAttribute(bracket_kernel<M, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bracket_kernel<M, false, false><<<g.grid, kBX, smem, s>>>(p, cols, rm, g.rows_per_split, eps, w);
    }
}

#define IQW_DISPATCH_M(m, CALL)                  \
    do {                                         \
        if ((m) <= 2) { constexpr int M = 2; CALL; }        \
        else if ((m) <= 4) { constexpr int M = 4; CALL; }   \
        else { constexpr int M = 8; CALL; }                 \
    } while (0)

// exact pipeline on the rows of `rm`: leaves the key of every target rank in w.r_key
static int run_exact(const float* p, long long cols, RowMap rm, const RankPlan& rp, bool named,
                     bool want_sum, bool to_dB, float eps, const Work& w, int sms, cudaStream_t s,
                     bool sample) {
    const long long col_tiles = (cols + kBX - 1) / kBX;
    Grid g;
    if (int rc = plan_grid(rm.n, col_tiles, sms, &g)) return rc;
    const unsigned cthreads = 128, cblocks = (unsigned)((cols + cthreads - 1) / cthreads);

    long long stride = rm.n / 2048;
    if (stride < 1) stride = 1;
    long long rsplits = (rm.n + stride - 1) / stride / 64;
    if (rsplits < 1) rsplits = 1;
    if (rsplits > 64) rsplits = 64;
    { IQW_PROFILE(sample ? "sample_range" : "stats_range", s);
      range_kernel<<<dim3((unsigned)col_tiles, (unsigned)rsplits), kBX, 0, s>>>(p, cols, rm, stride, w); }

    const size_t l0_smem = sizeof(uint16_t) * kNB0 * kBX;
    {
        IQW_PROFILE(sample ? "sample_l0" : "stats_l0", s);
#define IQW_L0(A, B, C)                                                                              \
    do {                                                                                             \
        cudaFuncSetAttribute(l0_kernel<A, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l0_smem); \
        l0_kernel<A, B, C><<<g.grid, kBX, l0_smem, s>>>(p, cols, rm, g.rows_per_split, eps, w);       \
    } while (0)
        if (!named) IQW_L0(false, false, false);
        else if (want_sum && to_dB) IQW_L0(true, true, true);
        else if (want_sum) IQW_L0(true, true, false);
        else IQW_L0(true, false, false);
#undef IQW_L0
    }
    if (rp.n_ranks > 0) {
        { IQW_PROFILE(sample ? "sample_scan" : "stats_scan", s);
          scan0_kernel<<<cblocks, cthreads, 0, s>>>(cols, rp, w); }
        IQW_DISPATCH_M(rp.n_ranks, launch_refine_levels<M>(g, s, p, cols, rm, rp, w, cblocks, cthreads,
                                                           sample ? "sample" : "stats"));
        IQW_DISPATCH_M(rp.n_ranks, launch_collect<M>(g, s, p, cols, rm, w,
                                                     sample ? "sample_collect" : "stats_collect"));
        { IQW_PROFILE(sample ? "sample_resolve" : "stats_resolve", s);
          resolve_kernel<<<(unsigned)cols, kResolveThreads, 0, s>>>(cols, rp, w); }
    }
    IQW_CUDA_OK(cudaGetLastError());
    return IQW_OK;
}

static int reset_work(const Work& w, long long cols, cudaStream_t s) {
    IQW_CUDA_OK(cudaMemsetAsync(w.range_hi, 0, w.zero_bytes, s));   // range_hi is the first field
    const unsigned cthreads = 128, cblocks = (unsigned)((cols + cthreads - 1) / cthreads);
    init_ff_kernel<<<cblocks, cthreads, 0, s>>>(w.range_lo, w.kmin, cols);
    return IQW_OK;
}

}  // namespace iqw

using namespace iqw;

extern "C" int iqw_debug_set_sample_margin(double sigmas, int extra) {
    g_margin_sigmas = sigmas;
    g_margin_extra = extra;
    return IQW_OK;
}

extern "C" size_t iqw_time_stats_workspace_bytes(int64_t n_channels, int64_t n_rows, int64_t n_cols,
                                                 int32_t n_stats) {
    (void)n_channels; (void)n_stats;
    if (n_cols <= 0) return 256;
    const size_t one = carve_work(nullptr, n_cols, nullptr);
    return n_rows >= kSampleMinRows ? 2 * one : one;
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

    RankPlan rp{};
    StatPlan st{};
    bool want_sum = false;
    if (int rc = build_plans(stats, n_stats, n_rows, &rp, &st, &want_sum)) return rc;

    // sampled path?
    RankPlan srp{};
    BracketPlan bp{};
    RowMap full{n_rows, 1, 0u};
    RowMap samp = full;
    bool sampled = false;
    if (n_rows >= kSampleMinRows && rp.n_ranks > 0) {
        samp.step = n_rows / kSampleRows;
        if (samp.step < 2) samp.step = 2;
        samp.n = n_rows / samp.step;
        samp.seed = 0x9E3779B9u;
        sampled = build_sample_plan(rp, n_rows, samp.n, &srp, &bp);
    }

    Work wf{}, ws{};
    const size_t one = carve_work(d_workspace, n_cols, &wf);
    size_t need = one;
    if (sampled) {
        carve_work(static_cast<char*>(d_workspace) + one, n_cols, &ws);
        need = 2 * one;
    }
    if (workspace_bytes < need)
        return fail(IQW_ERR_WORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need);

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    const long long col_tiles = (n_cols + kBX - 1) / kBX;
    const unsigned cthreads = 128, cblocks = (unsigned)((n_cols + cthreads - 1) / cthreads);

    for (int64_t c = 0; c < n_channels; ++c) {
        const float* p = d_p + c * p_channel_stride;
        float* out = d_out + c * (int64_t)n_stats * n_cols;
        { IQW_PROFILE("stats_init", s); if (int rc = reset_work(wf, n_cols, s)) return rc; }

        if (!sampled) {
            if (int rc = run_exact(p, n_cols, full, rp, true, want_sum, to_dB != 0, eps, wf, sms, s, false))
                return rc;
        } else {
            { IQW_PROFILE("stats_init", s); if (int rc = reset_work(ws, n_cols, s)) return rc; }
            if (srp.n_ranks > 0)
                if (int rc = run_exact(p, n_cols, samp, srp, false, false, false, eps, ws, sms, s, true))
                    return rc;
            { IQW_PROFILE("stats_scan", s);
              make_brackets_kernel<<<cblocks, cthreads, 0, s>>>(n_cols, bp, ws, wf); }
            Grid g;
            if (int rc = plan_grid(n_rows, col_tiles, sms, &g)) return rc;
            IQW_DISPATCH_M(bp.n_groups, launch_bracket<M>(g, s, p, n_cols, full, want_sum, to_dB != 0, eps, wf));
            { IQW_PROFILE("stats_scan", s);
              scan_bracket_kernel<<<cblocks, cthreads, 0, s>>>(n_cols, rp, wf); }
            IQW_DISPATCH_M(rp.n_ranks, launch_refine_levels<M>(g, s, p, n_cols, full, rp, wf, cblocks, cthreads, "stats"));
            IQW_DISPATCH_M(rp.n_ranks, launch_collect<M>(g, s, p, n_cols, full, wf, "stats_collect"));
            { IQW_PROFILE("stats_resolve", s);
              resolve_kernel<<<(unsigned)n_cols, kResolveThreads, 0, s>>>(n_cols, rp, wf); }
        }
        { IQW_PROFILE("stats_finalize", s);
          finalize_kernel<<<cblocks, cthreads, 0, s>>>(n_rows, n_cols, st, to_dB, eps, wf, out); }
        IQW_CUDA_OK(cudaGetLastError());
    }
    return IQW_OK;
}
