// iqw_stft.cuh -- pieces of the STFT kernels shared by iqw_stft.cu (nfft <= 8192, one CTA per frame
// group) and iqw_stft_large.cu (nfft 16384..65536, two-kernel four-step).
#pragma once
#include "iqw_common.cuh"
#include "fft_core.cuh"

namespace iqw {

struct StftArgs {
    const float2* x;
    long long n_samples, x_ch_stride;
    int n_channels;
    const float* window;
    const float2* twiddle;
    long long hop, n_frames;
    float eps;
    int bin_lo, bin_hi;
    void* out;
    long long out_ch_stride;
    long long n_groups;        // n_channels * groups_per_channel
    long long groups_per_ch;   // ceil(n_frames / FPC)
};

// immutable per-(device, log2 n) twiddle table of the in-CTA FFT (built on first use)
int get_twiddles(int log2n, cudaStream_t stream, const float2** out);

// nfft = 2^14 .. 2^16 (iqw_stft_large.cu)
size_t stft_large_workspace_bytes(int log2n, long long n_channels, long long n_frames);
int launch_stft_large(const StftArgs& a, int log2n, int mode, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream);

}  // namespace iqw
