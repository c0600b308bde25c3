// Host emulation of the device FFT building blocks (iqwaveform_b200/csrc/fft_core.cuh).
// Runs every "thread" of one frame sequentially, pass by pass, with the same ping-pong exchange
// buffers and twiddle-table layout as the CUDA kernel, and prints the max error against a float64
// DFT for every supported size.  Built and run by tests/test_host_fft.py (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "../../iqwaveform_b200/csrc/fft_core.cuh"

using namespace iqw;

template <int LOG2N, int P>
struct Runner {
    static void run(std::vector<float2>& regs, std::vector<float2>& a, std::vector<float2>& b,
                    const std::vector<float2>& tw) {
        constexpr int N = 1 << LOG2N, E = plan_elems(LOG2N), TPF = N / E;
        for (int t = 0; t < TPF; ++t) {
            float2 tn[E];
            if constexpr (P > 0) load_twiddles<LOG2N, P>(tn, tw.data(), t);
            fft_pass<LOG2N, P>(&regs[t * E], a.data(), b.data(), tn, t);
        }
        if constexpr (P + 1 < plan_passes(LOG2N)) Runner<LOG2N, P + 1>::run(regs, b, a, tw);
    }
};

template <int LOG2N>
double check(unsigned seed) {
    constexpr int N = 1 << LOG2N, E = plan_elems(LOG2N), TPF = N / E, NP = plan_passes(LOG2N);
    constexpr int R0 = plan_radix(LOG2N, 0), RL = plan_radix(LOG2N, NP - 1);
    srand(seed);
    std::vector<std::complex<double>> x(N);
    for (auto& v : x) v = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};

    // twiddle table, same layout as the device init kernel
    std::vector<float2> tw(plan_tw_size(LOG2N) + 1);
    for (int p = 1; p < NP; ++p) {
        const int R = plan_radix(LOG2N, p), Ns = plan_ns(LOG2N, p), off = plan_tw_offset(LOG2N, p);
        for (int r = 1; r < R; ++r)
            for (int i = 0; i < Ns; ++i) {
                double ang = -2.0 * M_PI * (double)(r * i) / (double)(Ns * R);
                tw[off + (r - 1) * Ns + i] = make_float2((float)cos(ang), (float)sin(ang));
            }
    }
    std::vector<float2> regs(N), a(padded_size(N)), b(padded_size(N));
    for (int t = 0; t < TPF; ++t)
        for (int q = 0; q < E / R0; ++q)
            for (int r = 0; r < R0; ++r) {
                int n = (t + q * TPF) + r * (N / R0);
                regs[t * E + q * R0 + r] = make_float2((float)x[n].real(), (float)x[n].imag());
            }
    Runner<LOG2N, 0>::run(regs, a, b, tw);

    std::vector<std::complex<double>> X(N);
    for (int t = 0; t < TPF; ++t)
        for (int q = 0; q < E / RL; ++q)
            for (int r = 0; r < RL; ++r) {
                int k = (t + q * TPF) + r * (N / RL);
                X[k] = {regs[t * E + q * RL + r].x, regs[t * E + q * RL + r].y};
            }
    // float64 DFT of the float32-rounded input
    double maxerr = 0, maxmag = 0;
    std::vector<std::complex<double>> w(N);
    for (int m = 0; m < N; ++m) w[m] = std::polar(1.0, -2.0 * M_PI * m / N);
    for (int k = 0; k < N; ++k) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n)
            acc += std::complex<double>((float)x[n].real(), (float)x[n].imag()) * w[(size_t)k * n % N];
        maxerr = std::fmax(maxerr, std::abs(acc - X[k]));
        maxmag = std::fmax(maxmag, std::abs(acc));
    }
    return maxerr / maxmag;
}

int main() {
    double e[14] = {0};
    e[4] = check<4>(1); e[5] = check<5>(2); e[6] = check<6>(3); e[7] = check<7>(4);
    e[8] = check<8>(5); e[9] = check<9>(6); e[10] = check<10>(7); e[11] = check<11>(8);
    e[12] = check<12>(9); e[13] = check<13>(10);
    int bad = 0;
    for (int l = 4; l <= 13; ++l) {
        printf("N=%d rel_err=%.3e\n", 1 << l, e[l]);
        if (!(e[l] < 2e-6)) bad = 1;
    }
    return bad;
}
