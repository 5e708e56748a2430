// fused.h -- single-kernel (shared-memory resident) modulator / receiver for the
// shapes that have a hand-written fused path.  See fused_modem.cu.
#pragma once
#include "common.cuh"

#include <complex>
#include <vector>

namespace gfdm {

struct FusedImpl;

class FusedModem {
public:
    // taps: already normalised, L*M entries, FFT order
    void init_tx(int M, int K, int L, const std::vector<std::complex<float>>& taps);
    void init_rx(int M, int K, int L, const std::vector<std::complex<float>>& taps,
                 const std::vector<std::complex<float>>& ic_taps);
    bool available() const { return impl_ != nullptr; }
    // in/out: [frames][N] device pointers; returns the number of kernel launches
    int modulate(cpx* out, const cpx* in, size_t frames, cudaStream_t s);
    // out_td (soft symbols) and/or out_fd (fft_filter_downsample result) may be null; eq may be null
    int demodulate(cpx* out_td, cpx* out_fd, const cpx* in, const cpx* eq, size_t frames, cudaStream_t s);
    const char* mod_name() const;
    const char* rx_name() const;
    void destroy();

private:
    FusedImpl* impl_ = nullptr;
};

} // namespace gfdm
