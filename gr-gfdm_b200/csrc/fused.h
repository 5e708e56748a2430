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
    // advanced receiver: successive interference cancellation resident in the receiver kernel.
    // Call after init_rx; returns false when the shape / constellation has no fused path.
    bool init_sic(const std::vector<std::complex<float>>& ic_taps, const std::vector<std::complex<float>>& points,
                  int rule, const std::vector<int>& subcarrier_map);
    bool sic_available() const;
    int demodulate_sic(cpx* out, const cpx* in, const cpx* eq, size_t frames, int ic_iter, int phase_comp,
                       cudaStream_t s);
    const char* sic_name() const;
    const char* mod_name() const;
    const char* rx_name() const;
    void destroy();

private:
    FusedImpl* impl_ = nullptr;
};

} // namespace gfdm
