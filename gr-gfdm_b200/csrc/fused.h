// fused.h -- single-kernel (shared-memory resident) modulator / receiver for the
// shapes that have a hand-written fused path.  See fused_modem.cu.
#pragma once
#include "common.cuh"

#include <complex>
#include <vector>

namespace gfdm {

struct FusedImpl;

// transmitter chain fused into the modulator kernel (mapper gather in front, preamble + cyclic prefix/suffix +
// window behind); all pointers are device pointers owned by the transmitter handle
constexpr int GFDM_TX_MAX_ANT = 4;
struct TxArgs {
    const int* inv_map = nullptr;   // [K] position of subcarrier k in the sorted subcarrier map, -1 = unused
    const cpx* front = nullptr;     // [ramp] leading window ramp
    const cpx* back = nullptr;      // [ramp] trailing window ramp
    const cpx* preambles = nullptr; // [n_shifts][P]
    size_t ant_stride = 0;          // elements between the outputs of two antennas (cyclic shifts)
    int A = 0, per_timeslot = 1, n_in = 0; // active subcarriers, symbol order, symbols per frame (<= A*M)
    int cp = 0, cs = 0, ramp = 0, P = 0;
    int n_ant = 1;
    int shift[GFDM_TX_MAX_ANT] = { 0, 0, 0, 0 };
    int pre_idx[GFDM_TX_MAX_ANT] = { 0, 0, 0, 0 };
    // chunk input (one byte per symbol): constellation points, device pointer
    const cpx* points = nullptr;
    int n_points = 0;
    // short_burst_shaper epilogue (lib/short_burst_shaper_impl.cc:161-182): every output row becomes
    // [pre_pad zeros | scale * (preamble | frame) | post_pad zeros]
    int shaped = 0, pre_pad = 0, post_pad = 0;
    float sc_re = 1.f, sc_im = 0.f;
};
constexpr int GFDM_FUSED_MAX_POINTS = 64; // constellation size the fused kernels keep in shared memory

// host helpers shared by fused_modem.cu and fused_twopass.cu
int fused_grid_cap(const void* fn, int threads, size_t smem); // persistent grid = SMs x resident CTAs
std::vector<cpx> make_row_twiddles(int R1, int R2);           // W_K^{n0*k1} as [k1][n0], K = R1*R2
// folded filter/twiddle table C[m][n1]; sign +1: modulator, -1: receiver
std::vector<std::complex<double>> make_fold_table_d(int M, int K, int L, const std::vector<std::complex<float>>& taps,
                                                    int sign, bool with_taps);
std::vector<cpx> make_fold_table(int M, int K, int L, const std::vector<std::complex<float>>& taps, int sign,
                                 bool with_taps);

// two-pass kernels for frames larger than shared memory (fused_twopass.cu)
struct TwoPass;
bool twopass_supported(int M, int K);
TwoPass* twopass_create_tx(int M, int K, int L, const std::vector<std::complex<float>>& taps);
TwoPass* twopass_create_rx(int M, int K, int L, const std::vector<std::complex<float>>& taps);
int twopass_modulate(TwoPass* t, cpx* out, const cpx* in, size_t frames, cudaStream_t s);
int twopass_demodulate(TwoPass* t, cpx* out, const cpx* in, const cpx* eq, int mode, size_t frames, cudaStream_t s); // mode 0: y, 1: R
bool twopass_supports_eq(const TwoPass* t);
const char* twopass_name(const TwoPass* t);
void twopass_destroy(TwoPass* t);

class FusedModem {
public:
    // taps: already normalised, L*M entries, FFT order
    void init_tx(int M, int K, int L, const std::vector<std::complex<float>>& taps);
    void init_rx(int M, int K, int L, const std::vector<std::complex<float>>& taps,
                 const std::vector<std::complex<float>>& ic_taps);
    bool available() const { return impl_ != nullptr; }
    // the one-tap equaliser is fused for single-pass shapes only
    bool supports_eq() const;
    // in/out: [frames][N] device pointers; returns the number of kernel launches
    int modulate(cpx* out, const cpx* in, size_t frames, cudaStream_t s);
    // whole transmitter chain in one kernel (single-pass shapes, even n_in, <= GFDM_TX_MAX_ANT antennas):
    // in [frames][n_in] compact symbols -> out [n_ant][...][P+cp+N+cs]
    bool supports_tx_chain(const TxArgs& tx) const;
    int transmit(cpx* out, const cpx* in, const TxArgs& tx, size_t frames, cudaStream_t s);
    const char* tx_name() const;
    // chunk entries: the symbol side of a frame is one byte per symbol.
    // modulate_chunks: chunks [frames][N] -> out [frames][N]; transmit_chunks: chunks [frames][tx.n_in];
    // demodulate_decide: in [frames][N] samples -> chunks_out [frames][N] hard decisions (full grid).
    bool supports_chunks(int n_points) const;
    int modulate_chunks(cpx* out, const unsigned char* chunks, const cpx* d_points, int n_points, size_t frames,
                        cudaStream_t s);
    bool supports_tx_chain_chunks(const TxArgs& tx) const;
    int transmit_chunks(cpx* out, const unsigned char* chunks, const TxArgs& tx, size_t frames, cudaStream_t s);
    int demodulate_decide(unsigned char* chunks_out, const cpx* in, const cpx* eq, const cpx* d_points, int n_points,
                          int rule, const DecideGrid& grid, size_t frames, cudaStream_t s);
    const char* modc_name() const;
    const char* txc_name() const;
    const char* rxd_name() const;
    // out_td (soft symbols) and/or out_fd (fft_filter_downsample result) may be null; eq may be null
    // in_stride != 0: frame f is read from in + f*in_stride (single-pass shapes; even strides and 16-byte aligned starts)
    bool supports_stride() const;
    int demodulate(cpx* out_td, cpx* out_fd, const cpx* in, const cpx* eq, size_t frames, cudaStream_t s, size_t in_stride = 0);
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
