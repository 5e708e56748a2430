// engine.h -- host-callable launchers of the CUDA kernels (one translation unit each).
#pragma once
#include "common.cuh"

namespace gfdm {

// ---------------------------------------------------------------- fft_engine.cu
// Any-length batched c2c DFT (unnormalised), Stockham autosort passes with
// radix 4/2/3/5/... from the factorisation of n.  Used directly for shapes that
// have no fused kernel and by the estimator.
struct FftPlan {
    int n = 0;
    std::vector<int> radices;
    cpx* d_tw = nullptr; // W_n^j = exp(-2 pi i j / n), generated in double
    void init(int n);
    void destroy();
};
// `batch` contiguous transforms; out != in; scratch holds batch*n elements
// (only touched when the plan has more than one pass).  Returns #launches.
int fft_exec(const FftPlan& p, cpx* out, const cpx* in, cpx* scratch, size_t batch, bool inverse,
             float scale, cudaStream_t s);

// ---------------------------------------------------------------- generic_smem.cu
// Any-shape modulator / receiver with the frame resident in shared memory (N <= 12288): one kernel per batch, 16 N bytes
// of HBM traffic per frame (24 N with a per-frame channel).  For the shapes that have no fused kernel.
bool generic_smem_supported(int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n);
int launch_generic_smem_mod(cpx* out, const cpx* in, int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n,
                            const cpx* d_taps, size_t frames, cudaStream_t s);
// mode 0: soft symbols (generic_work[_equalize]), mode 1: R (fft_[equalize_]filter_downsample); eq may be null
int launch_generic_smem_rx(cpx* out, const cpx* in, const cpx* eq, int mode, int M, int K, int L, const FftPlan& fft_m,
                           const FftPlan& fft_n, const cpx* d_taps, size_t frames, cudaStream_t s);

// ---------------------------------------------------------------- stage_kernels.cu
// modulator: X[b*M+m] = sum_i T[((i+h)%L)*M+m] * D[((b-i+h) mod K)*M+m], m < part_len
void launch_mod_filter(cpx* X, const cpx* D, const cpx* taps, int M, int K, int L, size_t frames, cudaStream_t s);
// receiver: R[k*M+m] = sum_i T[((i+h)%L)*M+m] * Y[((k+i-h) mod K)*M+m]
void launch_rx_filter(cpx* R, const cpx* Y, const cpx* taps, int M, int K, int L, size_t frames, cudaStream_t s);
void launch_eq_divide(cpx* out, const cpx* Y, const cpx* eq, size_t n, cudaStream_t s);
void launch_neighbor_sum(cpx* out, const cpx* td, int M, int K, size_t frames, cudaStream_t s);
void launch_ic_subtract(cpx* out, const cpx* fd, const cpx* F, const cpx* ic_taps, int M, int K, size_t frames, cudaStream_t s);
// one whole cancellation iteration (decide neighbours, re-modulate, subtract, back to time domain): y_out != y_in
bool sic_iter_supported(int M, int K, int n_points, size_t frames);
void launch_sic_iter(cpx* y_out, const cpx* y_in, const cpx* fb, const cpx* ic_taps, const unsigned char* active,
                     const cpx* points, int n_points, int rule, const DecideGrid& grid, int M, int K, size_t frames,
                     cudaStream_t s);
void launch_decide(cpx* out, const cpx* in, const unsigned char* active, const cpx* points, int n_points, int rule,
                   int M, int K, size_t frames, cudaStream_t s);
void launch_phase_rotate(cpx* R, const cpx* decided, const cpx* soft, const int* smap, int n_map, int M, int K,
                         size_t frames, cudaStream_t s);
void launch_map(cpx* out, const cpx* in, const int* inv_map, int M, int K, int A, bool per_timeslot, size_t n_in,
                size_t in_stride, size_t frames, cudaStream_t s);
void launch_demap(cpx* out, const cpx* in, const int* smap, int M, int K, int A, bool per_timeslot, size_t n_out,
                  size_t out_stride, size_t frames, cudaStream_t s);
void launch_add_cp(cpx* out, const cpx* in, int N, int cp, int cs, int ramp, const cpx* front, const cpx* back,
                   int shift, size_t out_stride, size_t frames, cudaStream_t s);
void launch_remove_cp(cpx* out, const cpx* in, int N, int cp, int cs, size_t frames, cudaStream_t s);
void launch_copy_rows(cpx* out, const cpx* row, int len, size_t out_stride, size_t frames, cudaStream_t s);
void launch_est_combine(cpx* H, const cpx* F, const cpx* inv0, const cpx* inv1, int K, size_t frames, cudaStream_t s);
void launch_est_filter(cpx* filt, const cpx* H, const float* g, int K, int A, int dc_free, size_t frames, cudaStream_t s);
void launch_est_interp(cpx* frame, const cpx* filt, int M, int K, int A, int dc_free, size_t frames, cudaStream_t s);
// estimate_frame as one kernel (power-of-two fft_len); tw = FftPlan(K).d_tw
bool est_fused_supported(int K, int A, int dc_free);
void launch_est_fused(cpx* frame, const cpx* rx, const cpx* tw, const cpx* inv0, const cpx* inv1, const float* g, int M,
                      int K, int A, int dc_free, size_t frames, cudaStream_t s);
void launch_est_snr(float* snr, float* cnrs, const cpx* F2, int K, int A, int dc_free, size_t frames, cudaStream_t s);
void launch_zf_prepare(cpx* out, const cpx* in, size_t n, cudaStream_t s);
void launch_energy(float* out, const cpx* in, size_t n, cudaStream_t s);

// ---------------------------------------------------------------- next_kernels.cu
// extract_burst_cc (lib/extract_burst_cc_impl.cc:117-242): one descriptor per produced burst
struct BurstDesc {
    long long start;       // index of the burst's first sample in the stream window (negative: zeros in front)
    double angle;              // CFO rotation per sample (radians), = arg(inc)        } filled on the device by
    double inc32_re, inc32_im; // cos/sin(32*angle): one warp step                     } burst_prepare_kernel
    float scale;               // power normalisation factor
    float pr_re, pr_im;        // the tag's phase_rotation
    int pad;
};
// desc: start, scale and phase_rotation set by the host; with cfo the rotation constants are derived on the device
// first (one more launch).  Returns the number of launches.
int launch_extract_burst(cpx* out, const cpx* in, BurstDesc* desc, int burst_len, bool cfo, int n_bursts, long long n_in,
                         cudaStream_t s);
// symbol mapping (python/pygfdm/symbolmapping.py:27-47): chunk = constellation point index, one byte per symbol
void launch_map_chunks(cpx* out, const unsigned char* chunks, const cpx* points, int n_points, size_t n, cudaStream_t s);
void launch_decide_chunks(unsigned char* chunks, const cpx* in, const cpx* points, int n_points, int rule,
                          const DecideGrid& grid, size_t n, cudaStream_t s);
void launch_bits2symbols(cpx* out, const unsigned char* bits, const cpx* points, int n_points, int bps, size_t n,
                         cudaStream_t s);
void launch_symbols2bits(unsigned char* bits, const cpx* in, const cpx* points, int n_points, int rule, int bps,
                         const DecideGrid& grid, size_t n, cudaStream_t s);
// sc16 host sample format: complex64 <-> interleaved int16 I/Q (n complex samples; aligned arrays, see next_kernels.cu)
void launch_cf32_to_sc16(short* out, const cpx* in, float scale, size_t n, cudaStream_t s);
void launch_sc16_to_cf32(cpx* out, const short* in, float scale, size_t n, cudaStream_t s);
// short_burst_shaper (lib/short_burst_shaper_impl.cc:161-182): [pre zeros | in * scale | post zeros] per burst of `len`
void launch_burst_shape(cpx* out, const cpx* in, int len, int pre, int post, cpx scale, size_t n_bursts, cudaStream_t s);
void launch_demap_chunks(unsigned char* out, const unsigned char* in, const int* smap, int M, int K, int A, bool per_timeslot,
                         size_t n_out, size_t frames, cudaStream_t s);

} // namespace gfdm
