// gfdm_b200.hpp -- host C++ layer of the B200 GFDM engine.
//
// Header-only classes in namespace gr::gfdm with the class names, constructor
// signatures and method names of gr-gfdm's GNU-Radio-free kernel layer, each a thin
// owner of one C-ABI handle (include/gfdm_b200.h).  A program written against the
// reference's headers compiles against these by switching the include path to
// <repo>/include (the forwarding headers include/gfdm/*.h carry the original file
// names) and linking libgfdm_b200.so instead of libgnuradio-gfdm.
//
//   reference class (header)                                   -> here
//   gfdm_kernel_utils        include/gfdm/gfdm_kernel_utils.h:36-52
//   modulator_kernel_cc      include/gfdm/modulator_kernel_cc.h:41-51
//   receiver_kernel_cc       include/gfdm/receiver_kernel_cc.h:53-89
//   advanced_receiver_kernel_cc  include/gfdm/advanced_receiver_kernel_cc.h:37-61
//   resource_mapper_kernel_cc    include/gfdm/resource_mapper_kernel_cc.h:38-58
//   add_cyclic_prefix_cc     include/gfdm/add_cyclic_prefix_cc.h:38-57
//   preamble_channel_estimator_cc include/gfdm/preamble_channel_estimator_cc.h:44-77
//   transmitter_kernel       include/gfdm/transmitter_kernel.h:43-69
//
// Error convention: GFDM_ERR_INVALID_ARGUMENT is rethrown as std::invalid_argument
// with the reference's message, everything else as std::runtime_error.
// Additions over the reference: every class has `*_batch(out, in, n_frames, mem)`
// (frame f at base + f*size, GFDM_MEM_HOST or GFDM_MEM_DEVICE), `set_stream`, `sync`.
// advanced_receiver_kernel_cc takes a plain `constellation` value instead of
// gr::digital::constellation_sptr (points() + decision rule).
#ifndef INCLUDED_GFDM_B200_HPP
#define INCLUDED_GFDM_B200_HPP

#include "gfdm_b200.h"

#include <complex>
#include <condition_variable>
#include <cstddef>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace gr {
namespace gfdm {

namespace detail {
inline void check(int status)
{
    if (status == GFDM_OK) return;
    const std::string msg = gfdm_last_error();
    if (status == GFDM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
typedef std::complex<float> cf;
inline const gfdm_complex* c(const cf* p) { return reinterpret_cast<const gfdm_complex*>(p); }
inline gfdm_complex* c(cf* p) { return reinterpret_cast<gfdm_complex*>(p); }

// owner of one C handle; movable, not copyable (the reference classes own FFTW plans the same way)
template <class H, void (*Destroy)(H*)>
class handle
{
public:
    handle() : d_h(nullptr) {}
    ~handle() { reset(); }
    handle(handle&& o) noexcept : d_h(o.d_h) { o.d_h = nullptr; }
    handle& operator=(handle&& o) noexcept
    {
        if (this != &o) {
            reset();
            d_h = o.d_h;
            o.d_h = nullptr;
        }
        return *this;
    }
    handle(const handle&) = delete;
    handle& operator=(const handle&) = delete;
    H* get() const { return d_h; }
    H** out() { return &d_h; }
    void reset()
    {
        if (d_h) Destroy(d_h);
        d_h = nullptr;
    }

private:
    H* d_h;
};
} // namespace detail

// ---------------------------------------------------------------------------
class gfdm_kernel_utils
{
public:
    typedef std::complex<float> gfdm_complex;
    // lib/gfdm_kernel_utils.cc:59-65
    float calculate_signal_energy(const gfdm_complex* p_in, const int ninput_size)
    {
        float e = 0.0f;
        detail::check(gfdm_calculate_signal_energy(&e, detail::c(p_in), ninput_size));
        return e;
    }
};

// mixin: stream control shared by all kernels (no reference counterpart)
template <class Derived>
class stream_control
{
public:
    void set_stream(void* cuda_stream) { detail::check(gfdm_set_stream(self()->raw(), cuda_stream)); }
    void sync() { detail::check(gfdm_sync(self()->raw())); }
    long long launch_count() const { return gfdm_launch_count(self()->raw()); }
    const char* last_kernel() const { return gfdm_last_kernel(self()->raw()); }

private:
    const Derived* self() const { return static_cast<const Derived*>(this); }
};

// ---------------------------------------------------------------------------
// lib/modulator_kernel_cc.cc:30-141
class modulator_kernel_cc : public gfdm_kernel_utils, public stream_control<modulator_kernel_cc>
{
public:
    modulator_kernel_cc(int n_timeslots, int n_subcarriers, int overlap, std::vector<gfdm_complex> frequency_taps)
        : d_n_taps(frequency_taps.size())
    {
        detail::check(gfdm_modulator_create(d_h.out(), n_timeslots, n_subcarriers, overlap,
                                            detail::c(frequency_taps.data()), (int)frequency_taps.size()));
    }
    void generic_work(gfdm_complex* p_out, const gfdm_complex* p_in)
    {
        detail::check(gfdm_modulator_work(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void generic_work_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_modulator_work_batch(d_h.get(), detail::c(p_out), detail::c(p_in), n_frames, mem));
    }
    int block_size() const { return gfdm_modulator_block_size(d_h.get()); }
    std::vector<gfdm_complex> filter_taps() const
    {
        std::vector<gfdm_complex> t(d_n_taps);
        detail::check(gfdm_modulator_filter_taps(d_h.get(), detail::c(t.data())));
        return t;
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_modulator, gfdm_modulator_destroy> d_h;
    size_t d_n_taps;
};

// ---------------------------------------------------------------------------
// lib/receiver_kernel_cc.cc:31-334
class receiver_kernel_cc : public gfdm_kernel_utils, public stream_control<receiver_kernel_cc>
{
public:
    receiver_kernel_cc(int n_timeslots, int n_subcarriers, int overlap, std::vector<gfdm_complex> frequency_taps)
        : d_n_taps(frequency_taps.size())
    {
        detail::check(gfdm_receiver_create(d_h.out(), n_timeslots, n_subcarriers, overlap,
                                           detail::c(frequency_taps.data()), (int)frequency_taps.size()));
    }
    void generic_work(gfdm_complex* out, const gfdm_complex* in)
    {
        detail::check(gfdm_receiver_work(d_h.get(), detail::c(out), detail::c(in)));
    }
    void generic_work_equalize(gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* f_eq_in)
    {
        detail::check(gfdm_receiver_work_equalize(d_h.get(), detail::c(out), detail::c(in), detail::c(f_eq_in)));
    }
    // f_eq_in may be null (no equalisation); it advances per frame like `in`
    void generic_work_batch(gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* f_eq_in, int n_frames,
                            int mem = GFDM_MEM_HOST)
    {
        detail::check(
            gfdm_receiver_work_batch(d_h.get(), detail::c(out), detail::c(in), detail::c(f_eq_in), n_frames, mem));
    }
    // frames read in place: frame f = block_size() samples at in + f*in_stride + in_offset (remove_prefix_cc fused in)
    void generic_work_strided_batch(gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* f_eq_in, size_t in_stride,
                                    size_t in_offset, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_receiver_work_strided_batch(d_h.get(), detail::c(out), detail::c(in), detail::c(f_eq_in), in_stride,
                                                       in_offset, n_frames, mem));
    }
    void fft_filter_downsample(gfdm_complex* p_out, const gfdm_complex* p_in)
    {
        detail::check(gfdm_receiver_fft_filter_downsample(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void fft_equalize_filter_downsample(gfdm_complex* p_out, const gfdm_complex* p_in, const gfdm_complex* f_eq_in)
    {
        detail::check(gfdm_receiver_fft_equalize_filter_downsample(d_h.get(), detail::c(p_out), detail::c(p_in),
                                                                   detail::c(f_eq_in)));
    }
    void fft_filter_downsample_batch(gfdm_complex* p_out, const gfdm_complex* p_in, const gfdm_complex* f_eq_in,
                                     int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_receiver_fft_filter_downsample_batch(d_h.get(), detail::c(p_out), detail::c(p_in),
                                                                detail::c(f_eq_in), n_frames, mem));
    }
    void transform_subcarriers_to_td(gfdm_complex* p_out, const gfdm_complex* p_in)
    {
        detail::check(gfdm_receiver_transform_subcarriers_to_td(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void transform_subcarriers_to_td_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int n_frames,
                                           int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_receiver_transform_subcarriers_to_td_batch(d_h.get(), detail::c(p_out), detail::c(p_in),
                                                                      n_frames, mem));
    }
    void cancel_sc_interference(gfdm_complex* p_out, const gfdm_complex* p_td_in, const gfdm_complex* p_fd_in)
    {
        detail::check(gfdm_receiver_cancel_sc_interference(d_h.get(), detail::c(p_out), detail::c(p_td_in),
                                                           detail::c(p_fd_in)));
    }
    void cancel_sc_interference_batch(gfdm_complex* p_out, const gfdm_complex* p_td_in, const gfdm_complex* p_fd_in,
                                      int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_receiver_cancel_sc_interference_batch(d_h.get(), detail::c(p_out), detail::c(p_td_in),
                                                                 detail::c(p_fd_in), n_frames, mem));
    }

    // ---- legacy 2-D interface (lib/receiver_kernel_cc.cc:130-163,194-209,227-272): the same
    //      stages on vector<vector<>>, [subcarrier][timeslot]; thin adapters over the flat calls.
    typedef std::vector<std::vector<gfdm_complex>> matrix;
    void filter_superposition(matrix& out, const gfdm_complex* in)
    {
        std::vector<gfdm_complex> flat((size_t)block_size());
        fft_filter_downsample(flat.data(), in);
        vectorize_2d(out, flat.data());
    }
    void demodulate_subcarrier(matrix& out, matrix& sc_fdomain)
    {
        std::vector<gfdm_complex> fd((size_t)block_size()), td((size_t)block_size());
        serialize_output(fd.data(), sc_fdomain);
        transform_subcarriers_to_td(td.data(), fd.data());
        vectorize_2d(out, td.data());
    }
    void serialize_output(gfdm_complex out[], matrix& sc_symbols)
    {
        const int M = timeslots(), K = subcarriers();
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) out[(size_t)k * M + m] = sc_symbols[k][m];
    }
    void vectorize_2d(matrix& out_vector, const gfdm_complex* p_in)
    {
        const int M = timeslots(), K = subcarriers();
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) out_vector[k][m] = p_in[(size_t)k * M + m];
    }
    void remove_sc_interference(matrix& sc_symbols, matrix& sc_fdomain)
    {
        std::vector<gfdm_complex> td((size_t)block_size()), fd((size_t)block_size()), res((size_t)block_size());
        serialize_output(td.data(), sc_symbols);
        serialize_output(fd.data(), sc_fdomain);
        cancel_sc_interference(res.data(), td.data(), fd.data());
        vectorize_2d(sc_symbols, res.data());
    }

    int block_size() const { return gfdm_receiver_block_size(d_h.get()); }
    int timeslots() const { return gfdm_receiver_timeslots(d_h.get()); }
    int subcarriers() const { return gfdm_receiver_subcarriers(d_h.get()); }
    int overlap() const { return gfdm_receiver_overlap(d_h.get()); }
    std::vector<gfdm_complex> filter_taps() const
    {
        std::vector<gfdm_complex> t(d_n_taps);
        detail::check(gfdm_receiver_filter_taps(d_h.get(), detail::c(t.data())));
        return t;
    }
    std::vector<gfdm_complex> ic_filter_taps() const
    {
        std::vector<gfdm_complex> t((size_t)timeslots());
        detail::check(gfdm_receiver_ic_filter_taps(d_h.get(), detail::c(t.data())));
        return t;
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_receiver, gfdm_receiver_destroy> d_h;
    size_t d_n_taps;
};

// ---------------------------------------------------------------------------
// GNU-Radio-free stand-in for gr::digital::constellation_sptr: the two things the reference
// uses are points() and decision_maker() (lib/advanced_receiver_kernel_cc.cc:114-120).
struct constellation {
    std::vector<std::complex<float>> points;
    int decision_rule; // gfdm_decision_rule
    static constellation qpsk() // gr::digital::constellation_qpsk: idx = 2*(im>0) + (re>0)
    {
        const float a = 0.70710678118654752440f;
        return constellation{ { { -a, -a }, { a, -a }, { -a, a }, { a, a } }, GFDM_DECISION_QPSK_SIGN };
    }
    static constellation nearest(std::vector<std::complex<float>> pts)
    {
        return constellation{ std::move(pts), GFDM_DECISION_NEAREST };
    }
};

// lib/advanced_receiver_kernel_cc.cc:32-123
class advanced_receiver_kernel_cc : public stream_control<advanced_receiver_kernel_cc>
{
public:
    typedef std::complex<float> gr_complex;
    advanced_receiver_kernel_cc(int timeslots, int subcarriers, int overlap, std::vector<gr_complex> frequency_taps,
                                std::vector<int> subcarrier_map, int ic_iter, const constellation& constellation_,
                                int do_phase_compensation)
    {
        gfdm_constellation c;
        c.points = detail::c(constellation_.points.data());
        c.n_points = (int)constellation_.points.size();
        c.decision_rule = constellation_.decision_rule;
        detail::check(gfdm_advanced_receiver_create(d_h.out(), timeslots, subcarriers, overlap,
                                                    detail::c(frequency_taps.data()), (int)frequency_taps.size(),
                                                    subcarrier_map.data(), (int)subcarrier_map.size(), ic_iter, &c,
                                                    do_phase_compensation));
    }
    void generic_work(gr_complex* p_out, const gr_complex* p_in)
    {
        detail::check(gfdm_advanced_receiver_work(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void generic_work_equalize(gr_complex* out, const gr_complex* in, const gr_complex* f_eq_in)
    {
        detail::check(
            gfdm_advanced_receiver_work_equalize(d_h.get(), detail::c(out), detail::c(in), detail::c(f_eq_in)));
    }
    void generic_work_batch(gr_complex* out, const gr_complex* in, const gr_complex* f_eq_in, int n_frames,
                            int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_advanced_receiver_work_batch(d_h.get(), detail::c(out), detail::c(in), detail::c(f_eq_in),
                                                        n_frames, mem));
    }
    void set_ic(int ic_iter) { detail::check(gfdm_advanced_receiver_set_ic(d_h.get(), ic_iter)); }
    int get_ic(void) { return gfdm_advanced_receiver_get_ic(d_h.get()); }
    int block_size() { return gfdm_advanced_receiver_block_size(d_h.get()); }
    void set_phase_compensation(int do_phase_compensation)
    {
        detail::check(gfdm_advanced_receiver_set_phase_compensation(d_h.get(), do_phase_compensation));
    }
    int get_phase_compensation() { return gfdm_advanced_receiver_get_phase_compensation(d_h.get()); }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_advanced_receiver, gfdm_advanced_receiver_destroy> d_h;
};

// ---------------------------------------------------------------------------
// lib/resource_mapper_kernel_cc.cc:30-162
class resource_mapper_kernel_cc : public stream_control<resource_mapper_kernel_cc>
{
public:
    typedef std::complex<float> gfdm_complex;
    resource_mapper_kernel_cc(int timeslots, int subcarriers, int active_subcarriers, std::vector<int> subcarrier_map,
                              bool per_timeslot = true, bool is_mapper = true)
    {
        detail::check(gfdm_resource_mapper_create(d_h.out(), timeslots, subcarriers, active_subcarriers,
                                                  subcarrier_map.data(), (int)subcarrier_map.size(),
                                                  per_timeslot ? 1 : 0, is_mapper ? 1 : 0));
    }
    size_t frame_size() { return gfdm_resource_mapper_frame_size(d_h.get()); }
    size_t block_size() { return gfdm_resource_mapper_block_size(d_h.get()); }
    size_t input_vector_size() { return gfdm_resource_mapper_input_vector_size(d_h.get()); }
    size_t output_vector_size() { return gfdm_resource_mapper_output_vector_size(d_h.get()); }
    void map_to_resources(gfdm_complex* p_out, const gfdm_complex* p_in, const size_t ninput_size)
    {
        detail::check(gfdm_resource_mapper_map_to_resources(d_h.get(), detail::c(p_out), detail::c(p_in), ninput_size));
    }
    void demap_from_resources(gfdm_complex* p_out, const gfdm_complex* p_in, const size_t noutput_size)
    {
        detail::check(
            gfdm_resource_mapper_demap_from_resources(d_h.get(), detail::c(p_out), detail::c(p_in), noutput_size));
    }
    void map_to_resources_batch(gfdm_complex* p_out, const gfdm_complex* p_in, size_t size_per_frame, int n_frames,
                                int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_resource_mapper_map_to_resources_batch(d_h.get(), detail::c(p_out), detail::c(p_in),
                                                                  size_per_frame, n_frames, mem));
    }
    void demap_from_resources_batch(gfdm_complex* p_out, const gfdm_complex* p_in, size_t size_per_frame, int n_frames,
                                    int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_resource_mapper_demap_from_resources_batch(d_h.get(), detail::c(p_out), detail::c(p_in),
                                                                      size_per_frame, n_frames, mem));
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_resource_mapper, gfdm_resource_mapper_destroy> d_h;
};

// ---------------------------------------------------------------------------
// lib/add_cyclic_prefix_cc.cc:30-104
class add_cyclic_prefix_cc : public stream_control<add_cyclic_prefix_cc>
{
public:
    typedef std::complex<float> gfdm_complex;
    add_cyclic_prefix_cc(int block_len, int cp_len, int cs_len, int ramp_len, std::vector<gfdm_complex> window_taps,
                         int cyclic_shift = 0)
    {
        detail::check(gfdm_cyclic_prefixer_create(d_h.out(), block_len, cp_len, cs_len, ramp_len,
                                                  detail::c(window_taps.data()), (int)window_taps.size(),
                                                  cyclic_shift));
    }
    void generic_work(gfdm_complex* p_out, const gfdm_complex* p_in)
    {
        detail::check(gfdm_cyclic_prefixer_work(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void add_cyclic_prefix(gfdm_complex* out, const gfdm_complex* in, const int cyclic_shift)
    {
        detail::check(gfdm_cyclic_prefixer_add_cyclic_prefix(d_h.get(), detail::c(out), detail::c(in), cyclic_shift));
    }
    void remove_cyclic_prefix(gfdm_complex* p_out, const gfdm_complex* p_in)
    {
        detail::check(gfdm_cyclic_prefixer_remove_cyclic_prefix(d_h.get(), detail::c(p_out), detail::c(p_in)));
    }
    void add_cyclic_prefix_batch(gfdm_complex* out, const gfdm_complex* in, int cyclic_shift, int n_frames,
                                 int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_cyclic_prefixer_add_cyclic_prefix_batch(d_h.get(), detail::c(out), detail::c(in),
                                                                   cyclic_shift, n_frames, mem));
    }
    void remove_cyclic_prefix_batch(gfdm_complex* out, const gfdm_complex* in, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(
            gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(d_h.get(), detail::c(out), detail::c(in), n_frames, mem));
    }
    int block_size() { return gfdm_cyclic_prefixer_block_size(d_h.get()); }
    int frame_size() { return gfdm_cyclic_prefixer_frame_size(d_h.get()); }
    int cyclic_shift() const { return gfdm_cyclic_prefixer_cyclic_shift(d_h.get()); }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_cyclic_prefixer, gfdm_cyclic_prefixer_destroy> d_h;
};

// ---------------------------------------------------------------------------
// lib/preamble_channel_estimator_cc.cc:34-294
class preamble_channel_estimator_cc : public gfdm_kernel_utils, public stream_control<preamble_channel_estimator_cc>
{
public:
    preamble_channel_estimator_cc(int timeslots, int fft_len, int active_subcarriers, bool is_dc_free,
                                  int which_estimator, std::vector<gfdm_complex> preamble)
    {
        detail::check(gfdm_channel_estimator_create(d_h.out(), timeslots, fft_len, active_subcarriers,
                                                    is_dc_free ? 1 : 0, which_estimator, detail::c(preamble.data()),
                                                    (int)preamble.size()));
    }
    int fft_len() { return gfdm_channel_estimator_fft_len(d_h.get()); }
    int timeslots() { return gfdm_channel_estimator_timeslots(d_h.get()); }
    int frame_len() { return gfdm_channel_estimator_frame_len(d_h.get()); }
    int active_subcarriers() { return gfdm_channel_estimator_active_subcarriers(d_h.get()); }
    bool is_dc_free() { return gfdm_channel_estimator_is_dc_free(d_h.get()) != 0; }
    std::vector<float> preamble_filter_taps()
    {
        std::vector<float> t(9);
        detail::check(gfdm_channel_estimator_preamble_filter_taps(d_h.get(), t.data()));
        return t;
    }
    void estimate_preamble_channel(gfdm_complex* fd_preamble_channel, const gfdm_complex* rx_preamble)
    {
        detail::check(gfdm_channel_estimator_estimate_preamble_channel(d_h.get(), detail::c(fd_preamble_channel),
                                                                       detail::c(rx_preamble)));
    }
    void filter_preamble_estimate(gfdm_complex* filtered, const gfdm_complex* estimate)
    {
        detail::check(
            gfdm_channel_estimator_filter_preamble_estimate(d_h.get(), detail::c(filtered), detail::c(estimate)));
    }
    void interpolate_frame(gfdm_complex* frame_estimate, const gfdm_complex* estimate)
    {
        detail::check(
            gfdm_channel_estimator_interpolate_frame(d_h.get(), detail::c(frame_estimate), detail::c(estimate)));
    }
    void estimate_frame(gfdm_complex* frame_estimate, const gfdm_complex* rx_preamble)
    {
        detail::check(
            gfdm_channel_estimator_estimate_frame(d_h.get(), detail::c(frame_estimate), detail::c(rx_preamble)));
    }
    void estimate_frame_batch(gfdm_complex* frame_estimate, const gfdm_complex* rx_preamble, int n_frames,
                              int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_channel_estimator_estimate_frame_batch(d_h.get(), detail::c(frame_estimate),
                                                                  detail::c(rx_preamble), n_frames, mem));
    }
    void prepare_for_zf(gfdm_complex* transformed_frame, const gfdm_complex* frame_estimate)
    {
        detail::check(
            gfdm_channel_estimator_prepare_for_zf(d_h.get(), detail::c(transformed_frame), detail::c(frame_estimate)));
    }
    float estimate_snr(std::vector<float>& cnrs, const gfdm_complex* rx_preamble)
    {
        float snr = 0.0f;
        cnrs.resize((size_t)active_subcarriers());
        detail::check(gfdm_channel_estimator_estimate_snr(d_h.get(), &snr, cnrs.data(), detail::c(rx_preamble)));
        return snr;
    }
    void estimate_snr_batch(float* snr_lin, float* cnrs, const gfdm_complex* rx_preamble, int n_frames,
                            int mem = GFDM_MEM_HOST)
    {
        detail::check(
            gfdm_channel_estimator_estimate_snr_batch(d_h.get(), snr_lin, cnrs, detail::c(rx_preamble), n_frames, mem));
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_channel_estimator, gfdm_channel_estimator_destroy> d_h;
};

// ---------------------------------------------------------------------------
// lib/short_burst_shaper_impl.cc:57-84 (ctor checks), :161-182 (sample path): [pre zeros | in * scale | post zeros]
class burst_shaper : public stream_control<burst_shaper>
{
public:
    typedef std::complex<float> gfdm_complex;
    burst_shaper(int pre_padding, int post_padding, gfdm_complex scale)
    {
        detail::check(gfdm_burst_shaper_create(d_h.out(), pre_padding, post_padding, scale.real(), scale.imag()));
    }
    int pre_padding() const { return gfdm_burst_shaper_pre_padding(d_h.get()); }
    int post_padding() const { return gfdm_burst_shaper_post_padding(d_h.get()); }
    // out[n_bursts][pre + burst_len + post]
    void work_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int burst_len, int n_bursts, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_burst_shaper_work_batch(d_h.get(), detail::c(p_out), detail::c(p_in), burst_len, n_bursts, mem));
    }
    gfdm_burst_shaper* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_burst_shaper, gfdm_burst_shaper_destroy> d_h;
};

// ---------------------------------------------------------------------------
// lib/transmitter_kernel.cc:34-107
class transmitter_kernel : public stream_control<transmitter_kernel>
{
public:
    typedef std::complex<float> gfdm_complex;
    transmitter_kernel(int timeslots, int subcarriers, int active_subcarriers, int cp_len, int cs_len, int ramp_len,
                       std::vector<int> subcarrier_map, bool per_timeslot, int overlap,
                       std::vector<gfdm_complex> frequency_taps, std::vector<gfdm_complex> window_taps,
                       std::vector<int> cyclic_shifts, std::vector<std::vector<gfdm_complex>> preambles)
        : d_cyclic_shifts(cyclic_shifts)
    {
        std::vector<const ::gfdm_complex*> ptrs;
        std::vector<int> sizes;
        for (const auto& p : preambles) {
            ptrs.push_back(detail::c(p.data()));
            sizes.push_back((int)p.size());
        }
        detail::check(gfdm_transmitter_create(
            d_h.out(), timeslots, subcarriers, active_subcarriers, cp_len, cs_len, ramp_len, subcarrier_map.data(),
            (int)subcarrier_map.size(), per_timeslot ? 1 : 0, overlap, detail::c(frequency_taps.data()),
            (int)frequency_taps.size(), detail::c(window_taps.data()), (int)window_taps.size(), cyclic_shifts.data(),
            (int)cyclic_shifts.size(), ptrs.data(), sizes.data(), (int)preambles.size()));
    }
    int input_vector_size() { return gfdm_transmitter_input_vector_size(d_h.get()); }
    int output_vector_size() { return gfdm_transmitter_output_vector_size(d_h.get()); }
    void generic_work(gfdm_complex* p_out, const gfdm_complex* p_in, const int ninput_size)
    {
        detail::check(gfdm_transmitter_work(d_h.get(), detail::c(p_out), detail::c(p_in), ninput_size));
    }
    void modulate(gfdm_complex* out, const gfdm_complex* in, const int ninput_size)
    {
        detail::check(gfdm_transmitter_modulate(d_h.get(), detail::c(out), detail::c(in), ninput_size));
    }
    void add_frame(gfdm_complex* out, const gfdm_complex* in, const int cyclic_shift)
    {
        detail::check(gfdm_transmitter_add_frame(d_h.get(), detail::c(out), detail::c(in), cyclic_shift));
    }
    // frames of cyclic_shifts()[0]: out[n_frames][output_vector_size()]
    void generic_work_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int ninput_size, int n_frames,
                            int mem = GFDM_MEM_HOST)
    {
        detail::check(
            gfdm_transmitter_work_batch(d_h.get(), detail::c(p_out), detail::c(p_in), ninput_size, n_frames, mem));
    }
    // every antenna (the loop of lib/transmitter_cc_impl.cc:165-177): out[n_shifts][n_frames][output_vector_size()]
    void generic_work_all_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int ninput_size, int n_frames,
                                int mem = GFDM_MEM_HOST)
    {
        detail::check(
            gfdm_transmitter_work_all_batch(d_h.get(), detail::c(p_out), detail::c(p_in), ninput_size, n_frames, mem));
    }
    // the chain with the short_burst_shaper as the kernel's epilogue: out[n_ant][n_frames][pre + output_vector_size() + post]
    void generic_work_shaped_batch(const burst_shaper& shaper, gfdm_complex* p_out, const gfdm_complex* p_in, int ninput_size,
                                   int n_frames, bool all_antennas = false, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_transmitter_work_shaped_batch(d_h.get(), shaper.raw(), detail::c(p_out), detail::c(p_in), ninput_size,
                                                         n_frames, all_antennas ? 1 : 0, mem));
    }
    const std::vector<int>& cyclic_shifts() const { return d_cyclic_shifts; }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_transmitter, gfdm_transmitter_destroy> d_h;
    std::vector<int> d_cyclic_shifts;
};

// ---------------------------------------------------------------------------
// The rows either side of the path.  The reference has them as GNU Radio blocks
// (remove_prefix_cc, extract_burst_cc) or stock gr-digital blocks / pygfdm helpers (symbol mapping); these
// classes are what the blocks' general_work() would own and call.

// lib/remove_prefix_cc_impl.cc:44-61,84-115
class remove_prefix : public stream_control<remove_prefix>
{
public:
    typedef std::complex<float> gfdm_complex;
    remove_prefix(int frame_len, int block_len, int offset) : d_block_len(block_len), d_frame_len(frame_len)
    {
        detail::check(gfdm_remove_prefix_create(d_h.out(), frame_len, block_len, offset));
    }
    int block_len() const { return d_block_len; }
    int frame_len() const { return d_frame_len; }
    void work_batch(gfdm_complex* p_out, const gfdm_complex* p_in, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_remove_prefix_work_batch(d_h.get(), detail::c(p_out), detail::c(p_in), n_frames, mem));
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_remove_prefix, gfdm_remove_prefix_destroy> d_h;
    int d_block_len, d_frame_len;
};

// lib/extract_burst_cc_impl.cc:43-242; one general_work() call per work()
class extract_burst : public stream_control<extract_burst>
{
public:
    typedef std::complex<float> gfdm_complex;
    struct result {
        int n_produced;        // bursts written
        long long n_consumed;  // items the block would consume_each()
    };
    extract_burst(int burst_len, int tag_backoff, bool activate_cfo_correction = false) : d_burst_len(burst_len)
    {
        detail::check(gfdm_extract_burst_create(d_h.out(), burst_len, tag_backoff, activate_cfo_correction ? 1 : 0));
    }
    int burst_len() const { return d_burst_len; }
    void activate_cfo_compensation(bool on) { detail::check(gfdm_extract_burst_activate_cfo_compensation(d_h.get(), on ? 1 : 0)); }
    // tags sorted by offset: burst_starts relative to p_in[0]; scale_factors / phase_rotations may be empty (1.0 / 1+0j)
    result work(gfdm_complex* p_out, int max_bursts, const gfdm_complex* p_in, long long n_in,
                const std::vector<long long>& burst_starts, const std::vector<float>& scale_factors = {},
                const std::vector<gfdm_complex>& phase_rotations = {}, int mem = GFDM_MEM_HOST)
    {
        if ((!scale_factors.empty() && scale_factors.size() != burst_starts.size()) ||
            (!phase_rotations.empty() && phase_rotations.size() != burst_starts.size()))
            throw std::invalid_argument("tag arrays MUST have the same length");
        result r{ 0, 0 };
        detail::check(gfdm_extract_burst_work(d_h.get(), detail::c(p_out), max_bursts, detail::c(p_in), n_in,
                                              burst_starts.data(), scale_factors.empty() ? nullptr : scale_factors.data(),
                                              phase_rotations.empty() ? nullptr : detail::c(phase_rotations.data()),
                                              (int)burst_starts.size(), &r.n_produced, &r.n_consumed, mem));
        return r;
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_extract_burst, gfdm_extract_burst_destroy> d_h;
    int d_burst_len;
};

// python/pygfdm/symbolmapping.py:27-47 (bits2symbols / symbols2bits), gr-digital chunks_to_symbols / constellation
// decoder (python/qa_advanced_receiver_sb_cc.py:97-99)
class symbol_mapper : public stream_control<symbol_mapper>
{
public:
    typedef std::complex<float> gfdm_complex;
    explicit symbol_mapper(const constellation& c)
    {
        const gfdm_constellation cc = { detail::c(c.points.data()), (int)c.points.size(), c.decision_rule };
        detail::check(gfdm_symbol_mapper_create(d_h.out(), &cc));
    }
    int n_points() const { return gfdm_symbol_mapper_n_points(d_h.get()); }
    int bits_per_symbol() const { return gfdm_symbol_mapper_bits_per_symbol(d_h.get()); }
    void map_chunks(gfdm_complex* p_out, const unsigned char* chunks, size_t n_symbols, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_symbol_mapper_map_chunks_batch(d_h.get(), detail::c(p_out), chunks, n_symbols, mem));
    }
    void decide(unsigned char* chunks_out, const gfdm_complex* p_in, size_t n_symbols, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_symbol_mapper_decide_batch(d_h.get(), chunks_out, detail::c(p_in), n_symbols, mem));
    }
    void bits2symbols(gfdm_complex* p_out, const unsigned char* bits, size_t n_symbols, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_symbol_mapper_bits2symbols_batch(d_h.get(), detail::c(p_out), bits, n_symbols, mem));
    }
    void symbols2bits(unsigned char* bits_out, const gfdm_complex* p_in, size_t n_symbols, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_symbol_mapper_symbols2bits_batch(d_h.get(), bits_out, detail::c(p_in), n_symbols, mem));
    }
    // the same mapping fused into the kernels of the path (one byte per symbol on the symbol side)
    void modulate_chunks(modulator_kernel_cc& mod, gfdm_complex* p_out, const unsigned char* chunks, int n_frames,
                         int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_modulator_work_chunks_batch(static_cast<gfdm_modulator*>(mod.raw()), d_h.get(), detail::c(p_out),
                                                       chunks, n_frames, mem));
    }
    void transmit_chunks(transmitter_kernel& tx, gfdm_complex* p_out, const unsigned char* chunks, int ninput_size,
                         int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_transmitter_work_chunks_batch(static_cast<gfdm_transmitter*>(tx.raw()), d_h.get(),
                                                         detail::c(p_out), chunks, ninput_size, n_frames, mem));
    }
    void demodulate_decide(receiver_kernel_cc& rx, unsigned char* chunks_out, const gfdm_complex* p_in,
                           const gfdm_complex* f_eq_in, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_receiver_work_decide_batch(static_cast<gfdm_receiver*>(rx.raw()), d_h.get(), chunks_out,
                                                      detail::c(p_in), detail::c(f_eq_in), n_frames, mem));
    }
    void demap_chunks(resource_mapper_kernel_cc& mapper, unsigned char* p_out, const unsigned char* p_in,
                      size_t size_per_frame, int n_frames, int mem = GFDM_MEM_HOST)
    {
        detail::check(gfdm_resource_mapper_demap_chunks_batch(static_cast<gfdm_resource_mapper*>(mapper.raw()), p_out, p_in,
                                                              size_per_frame, n_frames, mem));
    }
    void* raw() const { return d_h.get(); }

private:
    detail::handle<gfdm_symbol_mapper, gfdm_symbol_mapper_destroy> d_h;
};

// ---------------------------------------------------------------------------
// Single-process multi-GPU driver ("one host thread + one or two CUDA streams per GPU").
// Frames are independent (lib/simple_modulator_cc_impl.cc:72-76), so a batch is split contiguously over the devices and
// nothing is exchanged between them.  One persistent worker thread per device: device selection is per host thread
// (gfdm_set_device), every worker creates its OWN kernel object through `make` after selecting its device, and
// for_each_shard(n, fn) runs fn(kernel, first_frame, n_frames) for the worker's shard on that thread; it returns when
// every shard is done and rethrows the first exception a worker raised.
template <class Kernel>
class multi_gpu
{
public:
    template <class Factory>
    multi_gpu(const std::vector<int>& devices, Factory make) : d_workers(devices.size())
    {
        for (size_t i = 0; i < devices.size(); ++i) {
            worker& w = d_workers[i];
            w.device = devices[i];
            w.thread = std::thread([this, &w, make]() {
                try {
                    detail::check(gfdm_set_device(w.device));
                    w.kernel = make();
                } catch (...) {
                    w.error = std::current_exception();
                }
                signal_done(w);
                for (;;) {
                    std::unique_lock<std::mutex> lk(d_mutex);
                    d_cv.wait(lk, [&w]() { return w.has_job || w.stop; });
                    if (w.stop) break;
                    std::function<void(Kernel&)> job = w.job;
                    lk.unlock();
                    try {
                        if (w.kernel) job(*w.kernel);
                    } catch (...) {
                        w.error = std::current_exception();
                    }
                    signal_done(w);
                }
                w.kernel.reset(); // handles are destroyed on the thread that owns their device selection
            });
        }
        // construction errors (no such device, invalid kernel parameters) surface here; the workers are stopped and joined
        // first -- a constructor that throws does not run the destructor, and the condition variables must not be
        // destroyed while a worker still waits on them
        try {
            wait_all();
        } catch (...) {
            shutdown();
            throw;
        }
    }
    ~multi_gpu() { shutdown(); }
    multi_gpu(const multi_gpu&) = delete;
    multi_gpu& operator=(const multi_gpu&) = delete;
    size_t n_devices() const { return d_workers.size(); }
    // contiguous split, remainder spread over the first shards: [lo, hi) of shard `rank`
    static std::pair<size_t, size_t> shard_bounds(size_t n, size_t world, size_t rank)
    {
        const size_t base = n / world, rem = n % world;
        const size_t lo = rank * base + (rank < rem ? rank : rem);
        return { lo, lo + base + (rank < rem ? 1 : 0) };
    }
    template <class Fn>
    void for_each_shard(size_t n_frames, Fn fn)
    {
        {
            std::lock_guard<std::mutex> lk(d_mutex);
            for (size_t i = 0; i < d_workers.size(); ++i) {
                const std::pair<size_t, size_t> b = shard_bounds(n_frames, d_workers.size(), i);
                d_workers[i].job = [fn, b](Kernel& k) { fn(k, b.first, b.second - b.first); };
                d_workers[i].has_job = true;
                d_workers[i].done = false;
            }
        }
        d_cv.notify_all();
        wait_all();
    }

private:
    struct worker {
        int device = 0;
        std::thread thread;
        std::unique_ptr<Kernel> kernel;
        std::function<void(Kernel&)> job;
        bool has_job = false, stop = false, done = false;
        std::exception_ptr error;
    };
    void shutdown()
    {
        {
            std::lock_guard<std::mutex> lk(d_mutex);
            for (worker& w : d_workers) w.stop = true;
        }
        d_cv.notify_all();
        for (worker& w : d_workers)
            if (w.thread.joinable()) w.thread.join();
    }
    void signal_done(worker& w)
    {
        {
            std::lock_guard<std::mutex> lk(d_mutex);
            w.has_job = false;
            w.done = true;
        }
        d_cv_done.notify_all();
    }
    void wait_all()
    {
        std::unique_lock<std::mutex> lk(d_mutex);
        d_cv_done.wait(lk, [this]() {
            for (const worker& w : d_workers)
                if (!w.done) return false;
            return true;
        });
        std::exception_ptr first;
        for (worker& w : d_workers) {
            if (w.error && !first) first = w.error;
            w.error = nullptr; // the first error is reported, none is left behind for a later call
        }
        if (first) {
            lk.unlock();
            std::rethrow_exception(first);
        }
    }
    std::vector<worker> d_workers;
    std::mutex d_mutex;
    std::condition_variable d_cv, d_cv_done;
};

} // namespace gfdm
} // namespace gr


#endif /* INCLUDED_GFDM_B200_HPP */
