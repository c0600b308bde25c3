// fft_core.cuh -- register-resident mixed-radix (16/8/4/2) Stockham FFT building blocks.
//
// Replaces, inside the fused STFT kernel, the reference's batched FFT call
// (/root/reference/src/iqwaveform/fourier.py:200-218, called from fourier.py:1044): forward,
// unnormalised, complex fp32, along the frame axis.
//
// Everything here is __host__ __device__ so that tests/host/fft_emulate.cpp can run the exact same
// butterfly / twiddle / index code thread-by-thread on the CPU (there is no GPU in the build
// container) and compare it with a float64 DFT.
//
// Algorithm (autosort Stockham, decimation in time).  N = R_0 * R_1 * ... * R_{P-1}.  Pass p has
// radix R = R_p and Ns = R_0*...*R_{p-1}.  Butterfly j (0 <= j < N/R) of pass p
//     reads    in[j + r*N/R]                          r = 0..R-1
//     scales   by W_{Ns*R}^{r*(j mod Ns)}             (nothing in pass 0, Ns = 1)
//     does     an R-point DFT
//     writes   out[(j / Ns)*Ns*R + (j mod Ns) + r*Ns]
// Input and output are both in natural order; pass 0 reads global memory (stride N/R across r,
// unit stride across j => coalesced) and the last pass writes bins j + r*N/R (coalesced).
#pragma once

#if defined(__CUDACC__)
#define IQW_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define IQW_HD inline
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace iqw {

// ------------------------------------------------------------------------------------------
// compile-time plan: radices per log2(N).  E = complex elements held per thread.
// ------------------------------------------------------------------------------------------
constexpr int kMaxPasses = 4;

IQW_HD constexpr int plan_radix(int log2n, int p) {
    // clang-format off
    switch (log2n) {
        case 4:  return p == 0 ? 16 : 1;
        case 5:  return p == 0 ? 8  : p == 1 ? 4 : 1;
        case 6:  return p <= 1 ? 8 : 1;
        case 7:  return p == 0 ? 16 : p == 1 ? 8 : 1;
        case 8:  return p <= 1 ? 16 : 1;
        case 9:  return p <= 2 ? 8 : 1;
        case 10: return p <= 1 ? 16 : p == 2 ? 4 : 1;
        case 11: return p <= 1 ? 16 : p == 2 ? 8 : 1;
        case 12: return p <= 2 ? 16 : 1;
        case 13: return p <= 1 ? 16 : p == 2 ? 8 : p == 3 ? 4 : 1;
        default: return 1;
    }
    // clang-format on
}

IQW_HD constexpr int plan_passes(int log2n) {
    int n = 0;
    for (int p = 0; p < kMaxPasses; ++p) n += plan_radix(log2n, p) > 1 ? 1 : 0;
    return n;
}

IQW_HD constexpr int plan_elems(int log2n) {   // elements per thread
    return log2n == 5 || log2n == 6 ? 8 : 16;
}

IQW_HD constexpr int plan_ns(int log2n, int p) {   // product of the radices before pass p
    int ns = 1;
    for (int q = 0; q < p; ++q) ns *= plan_radix(log2n, q);
    return ns;
}

// twiddle table: pass p >= 1 owns (R_p - 1) * Ns_p entries, entry (r-1)*Ns + i = W_{Ns*R}^{r*i}
IQW_HD constexpr int plan_tw_offset(int log2n, int p) {
    int off = 0;
    for (int q = 1; q < p; ++q) off += (plan_radix(log2n, q) - 1) * plan_ns(log2n, q);
    return off;
}
IQW_HD constexpr int plan_tw_size(int log2n) { return plan_tw_offset(log2n, plan_passes(log2n)); }

// shared-memory exchange index with one pad element every 16 (kills the 16-way conflict of the
// stride-R stores after pass 0; see DESIGN.md "bank conflicts")
IQW_HD constexpr int pad_index(int i) { return i + (i >> 4); }
IQW_HD constexpr int padded_size(int n) { return n + (n >> 4); }

// ------------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------------
// complex add / subtract: on the device one packed f32x2 instruction (sm_100a FADD2 / FFMA2) on the
// (re, im) register pair -- same IEEE results as the scalar pair, half the issue slots
#if defined(__CUDA_ARCH__) && defined(IQW_PACKED_F32X2)
IQW_HD float2 cadd(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 p, q; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %5}; add.rn.f32x2 p, p, q; mov.b64 {%0, %1}, p; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
IQW_HD float2 csub(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 p, q; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %5}; sub.rn.f32x2 p, p, q; mov.b64 {%0, %1}, p; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
#else
IQW_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
IQW_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
// complex multiply and real scaling.  Packed form: (a.x, a.y)*(b.x, b.x) then fma((a.y, a.x),
// (-b.y, b.y), .) -- ptxas folds the half swap, the single negation and the scalar broadcasts into
// operand modifiers of FMUL2 / FFMA2, so a complex multiply is 2 instructions instead of 4.
#if defined(__CUDA_ARCH__) && defined(IQW_PACKED_F32X2)
IQW_HD float2 cmul(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 pa, pas, cc, sm, p; .reg .f32 ns; mov.b64 pa, {%2, %3}; mov.b64 pas, {%3, %2}; "
        "mov.b64 cc, {%4, %4}; neg.f32 ns, %5; mov.b64 sm, {ns, %5}; mul.rn.f32x2 p, pa, cc; "
        "fma.rn.f32x2 p, pas, sm, p; mov.b64 {%0, %1}, p; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
IQW_HD float2 cscale(float2 a, float s) {      // a * s, s real
    float2 d;
    asm("{ .reg .b64 pa, ss; mov.b64 pa, {%2, %3}; mov.b64 ss, {%4, %4}; mul.rn.f32x2 pa, pa, ss; mov.b64 {%0, %1}, pa; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(s));
    return d;
}
#else
IQW_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
IQW_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#endif
// fused forms used where a scaling meets the first radix-2 stage of a butterfly: one rounding less per value
#if defined(__CUDA_ARCH__) && defined(IQW_PACKED_F32X2)
IQW_HD float2 cfma_rs(float2 a, float s, float2 c) {     // a * s + c, s real: one FFMA2
    float2 d;
    asm("{ .reg .b64 pa, ss, pc; mov.b64 pa, {%2, %3}; mov.b64 ss, {%4, %4}; mov.b64 pc, {%5, %6}; "
        "fma.rn.f32x2 pa, pa, ss, pc; mov.b64 {%0, %1}, pa; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
    return d;
}
IQW_HD float2 cfma(float2 a, float2 b, float2 c) {       // a * b + c, all complex: two FFMA2
    float2 d;
    asm("{ .reg .b64 pa, pas, cc, sm, p; .reg .f32 ns; mov.b64 pa, {%2, %3}; mov.b64 pas, {%3, %2}; "
        "mov.b64 cc, {%4, %4}; neg.f32 ns, %5; mov.b64 sm, {ns, %5}; mov.b64 p, {%6, %7}; "
        "fma.rn.f32x2 p, pa, cc, p; fma.rn.f32x2 p, pas, sm, p; mov.b64 {%0, %1}, p; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
IQW_HD float2 ctwice_minus(float2 u, float2 s) {         // 2u - s: one FFMA2
    float2 d;
    asm("{ .reg .b64 pu, two, ps; .reg .f32 nx, ny; mov.b64 pu, {%2, %3}; mov.b64 two, {0f40000000, 0f40000000}; "
        "neg.f32 nx, %4; neg.f32 ny, %5; mov.b64 ps, {nx, ny}; fma.rn.f32x2 pu, pu, two, ps; mov.b64 {%0, %1}, pu; }"
        : "=f"(d.x), "=f"(d.y) : "f"(u.x), "f"(u.y), "f"(s.x), "f"(s.y));
    return d;
}
#else
IQW_HD float2 cfma_rs(float2 a, float s, float2 c) { return make_float2(a.x * s + c.x, a.y * s + c.y); }
IQW_HD float2 cfma(float2 a, float2 b, float2 c) {
    return make_float2(a.x * b.x - a.y * b.y + c.x, a.x * b.y + a.y * b.x + c.y);
}
IQW_HD float2 ctwice_minus(float2 u, float2 s) { return make_float2(2.f * u.x - s.x, 2.f * u.y - s.y); }
#endif
IQW_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }          // a * (-i)
// a * exp(-i*pi/4)  and  a * exp(-3i*pi/4)
#if defined(__CUDA_ARCH__) && defined(IQW_PACKED_F32X2)
IQW_HD float2 mul_w8_1(float2 a) {
    const float h = 0.70710678118654752440f;
    return cmul(a, make_float2(h, -h));
}
IQW_HD float2 mul_w8_3(float2 a) {
    const float h = 0.70710678118654752440f;
    return cmul(a, make_float2(-h, -h));
}
#else
IQW_HD float2 mul_w8_1(float2 a) {
    const float h = 0.70710678118654752440f;
    return make_float2((a.x + a.y) * h, (a.y - a.x) * h);
}
IQW_HD float2 mul_w8_3(float2 a) {
    const float h = 0.70710678118654752440f;
    return make_float2((a.y - a.x) * h, -(a.x + a.y) * h);
}
#endif

// ------------------------------------------------------------------------------------------
// in-register forward DFTs; element n lives at a[n*S]; outputs in natural order
// ------------------------------------------------------------------------------------------
template <int S>
IQW_HD void bfly2(float2* a) {
    float2 t = a[S];
    a[S] = csub(a[0], t);
    a[0] = cadd(a[0], t);
}

template <int S>
IQW_HD void bfly4(float2* a) {
    float2 t0 = cadd(a[0], a[2 * S]);
    float2 t1 = csub(a[0], a[2 * S]);
    float2 t2 = cadd(a[S], a[3 * S]);
    float2 t3 = mul_mi(csub(a[S], a[3 * S]));
    a[0] = cadd(t0, t2);
    a[S] = cadd(t1, t3);
    a[2 * S] = csub(t0, t2);
    a[3 * S] = csub(t1, t3);
}

// second half of the radix-8 butterfly: sums s_n = a_n + a_{n+4} and differences d_n = a_n - a_{n+4} in,
// X[2k] = DFT4(s), X[2k+1] = DFT4(d_n W8^n) out (decimation in frequency)
template <int S>
IQW_HD void bfly8_tail(float2* a, float2 s0, float2 s1, float2 s2, float2 s3, float2 d0, float2 d1, float2 d2, float2 d3) {
    float2 e[4] = {s0, s1, s2, s3};
    float2 o[4] = {d0, mul_w8_1(d1), mul_mi(d2), mul_w8_3(d3)};
    bfly4<1>(e);
    bfly4<1>(o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[(2 * k) * S] = e[k];
        a[(2 * k + 1) * S] = o[k];
    }
}

template <int S>
IQW_HD void bfly8(float2* a) {
    bfly8_tail<S>(a, cadd(a[0], a[4 * S]), cadd(a[S], a[5 * S]), cadd(a[2 * S], a[6 * S]), cadd(a[3 * S], a[7 * S]),
                  csub(a[0], a[4 * S]), csub(a[S], a[5 * S]), csub(a[2 * S], a[6 * S]), csub(a[3 * S], a[7 * S]));
}

// radix-8 butterfly of a[n*S] * w[n*S] (w real: the window): the scaling rides on the first radix-2 stage,
// u = a_n w_n, s = a_{n+4} w_{n+4} + u, d = -a_{n+4} w_{n+4} + u: 12 packed instructions instead of 16
template <int S>
IQW_HD void bfly8_scaled(float2* a, const float* w) {
    float2 sv[4], dv[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float2 u = cscale(a[n * S], w[n * S]);
        sv[n] = cfma_rs(a[(n + 4) * S], w[(n + 4) * S], u);
        dv[n] = cfma_rs(a[(n + 4) * S], -w[(n + 4) * S], u);
    }
    bfly8_tail<S>(a, sv[0], sv[1], sv[2], sv[3], dv[0], dv[1], dv[2], dv[3]);
}
// radix-8 butterfly of a[n*S] * t[n] (t complex twiddles; t[0] is 1 when FIRST_ONE): u = a_n t_n,
// s = a_{n+4} t_{n+4} + u (two FFMA2), d = 2u - s (one): 5 packed instructions per pair instead of 6
template <int S, bool FIRST_ONE>
IQW_HD void bfly8_twiddled(float2* a, const float2* t) {
    float2 sv[4], dv[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float2 u = (FIRST_ONE && n == 0) ? a[0] : cmul(a[n * S], t[n]);
        sv[n] = cfma(a[(n + 4) * S], t[n + 4], u);
        dv[n] = ctwice_minus(u, sv[n]);
    }
    bfly8_tail<S>(a, sv[0], sv[1], sv[2], sv[3], dv[0], dv[1], dv[2], dv[3]);
}
// the same for the radix-4 first stage of the 32-point transform (pairs n, n+2)
template <int S>
IQW_HD void bfly4_tail(float2* a, float2 t0, float2 t2, float2 t1, float2 d13) {
    // t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, d13 = a1 - a3
    const float2 t3 = mul_mi(d13);
    a[0] = cadd(t0, t2);
    a[S] = cadd(t1, t3);
    a[2 * S] = csub(t0, t2);
    a[3 * S] = csub(t1, t3);
}
template <int S>
IQW_HD void bfly4_scaled(float2* a, const float* w) {
    const float2 u0 = cscale(a[0], w[0]), u1 = cscale(a[S], w[S]);
    const float2 t0 = cfma_rs(a[2 * S], w[2 * S], u0), t1 = cfma_rs(a[2 * S], -w[2 * S], u0);
    const float2 t2 = cfma_rs(a[3 * S], w[3 * S], u1), d13 = cfma_rs(a[3 * S], -w[3 * S], u1);
    bfly4_tail<S>(a, t0, t2, t1, d13);
}
template <int S, bool FIRST_ONE>
IQW_HD void bfly4_twiddled(float2* a, const float2* t) {
    const float2 u0 = FIRST_ONE ? a[0] : cmul(a[0], t[0]);
    const float2 u1 = cmul(a[S], t[1]);
    const float2 t0 = cfma(a[2 * S], t[2], u0), t1 = ctwice_minus(u0, t0);
    const float2 t2 = cfma(a[3 * S], t[3], u1), d13 = ctwice_minus(u1, t2);
    bfly4_tail<S>(a, t0, t2, t1, d13);
}

template <int S>
IQW_HD void bfly16(float2* a) {
    // n = 4*n1 + n0, k = k0 + 4*k1:
    //   y[n0][k0] = sum_n1 a[4 n1 + n0] W4^{n1 k0};  y *= W16^{n0 k0};  X[k0 + 4 k1] = sum_n0 y W4^{n0 k1}
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin(pi/8)
    float2 y[16];
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) {
        float2 t[4] = {a[n0 * S], a[(n0 + 4) * S], a[(n0 + 8) * S], a[(n0 + 12) * S]};
        bfly4<1>(t);
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) y[n0 * 4 + k0] = t[k0];
    }
    // W16^m = exp(-2 pi i m / 16)
    y[1 * 4 + 1] = cmul(y[1 * 4 + 1], make_float2(c1, -s1));     // m = 1
    y[1 * 4 + 2] = mul_w8_1(y[1 * 4 + 2]);                       // m = 2
    y[1 * 4 + 3] = cmul(y[1 * 4 + 3], make_float2(s1, -c1));     // m = 3
    y[2 * 4 + 1] = mul_w8_1(y[2 * 4 + 1]);                       // m = 2
    y[2 * 4 + 2] = mul_mi(y[2 * 4 + 2]);                         // m = 4
    y[2 * 4 + 3] = mul_w8_3(y[2 * 4 + 3]);                       // m = 6
    y[3 * 4 + 1] = cmul(y[3 * 4 + 1], make_float2(s1, -c1));     // m = 3
    y[3 * 4 + 2] = mul_w8_3(y[3 * 4 + 2]);                       // m = 6
    y[3 * 4 + 3] = cmul(y[3 * 4 + 3], make_float2(-c1, s1));     // m = 9
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) {
        float2 t[4] = {y[0 * 4 + k0], y[1 * 4 + k0], y[2 * 4 + k0], y[3 * 4 + k0]};
        bfly4<1>(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) a[(k0 + 4 * k1) * S] = t[k1];
    }
}

template <int R, int S>
IQW_HD void bfly(float2* a) {
    if constexpr (R == 2) bfly2<S>(a);
    else if constexpr (R == 4) bfly4<S>(a);
    else if constexpr (R == 8) bfly8<S>(a);
    else bfly16<S>(a);
}

// ------------------------------------------------------------------------------------------
// one pass for one thread.  The thread owns butterflies j = ltid + q*TPF, q = 0..E/R-1, and keeps
// element r of butterfly q in v[q*R + r].
//   FIRST: v was filled by the caller from global memory (windowed samples), no twiddle
//   LAST : v is left in registers for the caller's epilogue; bin of v[q*R + r] is j + r*N/R
// `src`/`dst` are this frame's ping-pong exchange buffers (padded), `t` the thread's twiddles of
// this pass (load_twiddles).
// ------------------------------------------------------------------------------------------
// twiddles of pass P (>= 1) for this thread: t[q*(R-1) + r-1] = W_{Ns*R}^{r*(j mod Ns)}.  Loaded
// separately from the pass so that the caller can issue the loads BEFORE the barrier that
// precedes the pass (they do not depend on the exchanged data).
template <int LOG2N, int P>
IQW_HD void load_twiddles(float2* t, const float2* tw, int ltid) {
    constexpr int N = 1 << LOG2N;
    constexpr int E = plan_elems(LOG2N);
    constexpr int TPF = N / E;
    constexpr int R = plan_radix(LOG2N, P);
    constexpr int Ns = plan_ns(LOG2N, P);
    constexpr int NB = E / R;
    constexpr int TWO = plan_tw_offset(LOG2N, P);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int i = (ltid + q * TPF) & (Ns - 1);
#pragma unroll
        for (int r = 1; r < R; ++r) t[q * (R - 1) + r - 1] = tw[TWO + (r - 1) * Ns + i];
    }
}

// STRIDE == 0: one frame per exchange buffer, padded index (pad_index).  STRIDE > 0: STRIDE
// independent transforms interleaved element by element (element i of transform c lives at
// [i*STRIDE + c]; the caller passes src/dst already offset by c) -- used by the large-nfft
// column pass, where the lanes of a warp run different transforms.
template <int LOG2N, int P, int STRIDE = 0>
IQW_HD void fft_pass(float2* v, const float2* src, float2* dst, const float2* t, int ltid) {
    constexpr int N = 1 << LOG2N;
    constexpr int E = plan_elems(LOG2N);
    constexpr int TPF = N / E;
    constexpr int R = plan_radix(LOG2N, P);
    constexpr int Ns = plan_ns(LOG2N, P);
    constexpr int NB = E / R;
    constexpr bool FIRST = (P == 0);
    constexpr bool LAST = (P == plan_passes(LOG2N) - 1);

#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int j = ltid + q * TPF;
        float2* a = v + q * R;
        if constexpr (!FIRST) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if constexpr (STRIDE > 0)
                    a[r] = src[(j + r * (N / R)) * STRIDE];
                else if constexpr ((N / R) % 16 == 0)    // keep the r-offset a compile-time immediate
                    a[r] = src[pad_index(j) + pad_index(r * (N / R))];
                else
                    a[r] = src[pad_index(j + r * (N / R))];
            }
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], t[q * (R - 1) + r - 1]);
        }
        bfly<R, 1>(a);
        if constexpr (!LAST) {
            const int base = (j / Ns) * (Ns * R) + (j & (Ns - 1));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if constexpr (STRIDE > 0)
                    dst[(base + r * Ns) * STRIDE] = a[r];
                else if constexpr (Ns % 16 == 0)
                    dst[pad_index(base) + pad_index(r * Ns)] = a[r];
                else
                    dst[pad_index(base + r * Ns)] = a[r];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// radix-32 / radix-64 in-register DFTs for the two-pass kernels (iqw_stft2p.cu): 8 x 4 and 8 x 8
// decompositions over bfly4 / bfly8 with compile-time W64 constants.  All indices are compile-time
// after unrolling, so the 64 values stay in registers.
// ------------------------------------------------------------------------------------------
IQW_HD constexpr float cos64_q(int m) {      // cos(2 pi m / 64), 0 <= m <= 16
    switch (m) {
        case 0: return 1.0f; case 1: return 9.951847267e-01f; case 2: return 9.807852804e-01f; case 3: return 9.569403357e-01f; case 4: return 9.238795325e-01f; case 5: return 8.819212643e-01f; case 6: return 8.314696123e-01f; case 7: return 7.730104534e-01f; case 8: return 7.071067812e-01f; case 9: return 6.343932842e-01f; case 10: return 5.555702330e-01f; case 11: return 4.713967368e-01f; case 12: return 3.826834324e-01f; case 13: return 2.902846773e-01f; case 14: return 1.950903220e-01f; case 15: return 9.801714033e-02f; case 16: return 0.0f;
    }
    return 0.0f;
}
IQW_HD constexpr float cos64(int m) {
    m &= 63;
    if (m > 32) m = 64 - m;
    return m > 16 ? -cos64_q(32 - m) : cos64_q(m);
}
IQW_HD constexpr float sin64(int m) { return cos64((m + 48) & 63); }

// a * W64^m = a * exp(-2 pi i m / 64), m a compile-time constant after unrolling
IQW_HD float2 mul_w64(float2 a, int m) {
    m &= 63;
    if (m == 0) return a;
    if (m == 16) return mul_mi(a);
    if (m == 32) return make_float2(-a.x, -a.y);
    if (m == 48) return make_float2(-a.y, a.x);
    if (m == 8) return mul_w8_1(a);
    if (m == 24) return mul_w8_3(a);
    return cmul(a, make_float2(cos64(m), -sin64(m)));
}

// 64-point forward DFT in place, natural order in and out.  r = 8a + b, k = c + 8d:
// W64^(rk) = W8^(ac) W64^(bc) W8^(bd)
// everything after the first stage (a[8c + b] = y[b][c]): internal twiddles, second stage, natural order
IQW_HD void bfly64_tail(float2* a);
IQW_HD void bfly64(float2* a) {
#pragma unroll
    for (int b = 0; b < 8; ++b) bfly8<8>(a + b);                 // a[8c + b] = y[b][c]
    bfly64_tail(a);
}
IQW_HD void bfly64_tail(float2* a) {
#pragma unroll
    for (int b = 1; b < 8; ++b)
#pragma unroll
        for (int c = 1; c < 8; ++c) a[8 * c + b] = mul_w64(a[8 * c + b], b * c);
#pragma unroll
    for (int c = 0; c < 8; ++c) bfly8<1>(a + 8 * c);             // a[8c + d] = X[c + 8d]
    float2 t[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) t[k] = a[8 * (k & 7) + (k >> 3)];
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] = t[k];
}

// 32-point forward DFT in place.  r = 8a + b (a < 4), k = c + 4d (c < 4): W32^(rk) = W4^(ac) W32^(bc) W8^(bd)
IQW_HD void bfly32_tail(float2* a);
IQW_HD void bfly32(float2* a) {
#pragma unroll
    for (int b = 0; b < 8; ++b) bfly4<8>(a + b);                 // a[8c + b] = y[b][c], c < 4
    bfly32_tail(a);
}
IQW_HD void bfly32_tail(float2* a) {
#pragma unroll
    for (int b = 1; b < 8; ++b)
#pragma unroll
        for (int c = 1; c < 4; ++c) a[8 * c + b] = mul_w64(a[8 * c + b], 2 * b * c);
#pragma unroll
    for (int c = 0; c < 4; ++c) bfly8<1>(a + 8 * c);             // a[8c + d] = X[c + 4d]
    float2 t[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) t[k] = a[8 * (k & 3) + (k >> 2)];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = t[k];
}

template <int R>
IQW_HD void bfly_big(float2* a) {
    if constexpr (R == 64) bfly64(a);
    else bfly32(a);
}
// a[r] * w[r] (real scales, the window) -> DFT, the scaling fused into the first stage
template <int R>
IQW_HD void bfly_big_scaled(float2* a, const float* w) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if constexpr (R == 64) bfly8_scaled<8>(a + b, w + b);
        else bfly4_scaled<8>(a + b, w + b);
    }
    if constexpr (R == 64) bfly64_tail(a);
    else bfly32_tail(a);
}
// a[8k + b] * A[k] * B[b] (A[0] = B[0] = 1 implied) -> DFT: the pass-B twiddles of the two-pass kernels,
// products formed per first-stage butterfly and fused into it
template <int R>
IQW_HD void bfly_big_twiddled(float2* a, const float2* A, const float2* B) {
    constexpr int NK = R / 8;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float2 t[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) t[k] = k == 0 ? B[b] : (b == 0 ? A[k] : cmul(A[k], B[b]));
        if constexpr (R == 64) {
            if (b == 0) bfly8_twiddled<8, true>(a + b, t);
            else bfly8_twiddled<8, false>(a + b, t);
        } else {
            if (b == 0) bfly4_twiddled<8, true>(a + b, t);
            else bfly4_twiddled<8, false>(a + b, t);
        }
    }
    if constexpr (R == 64) bfly64_tail(a);
    else bfly32_tail(a);
}

}  // namespace iqw
