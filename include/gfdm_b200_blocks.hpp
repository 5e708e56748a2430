// gfdm_b200_blocks.hpp -- GNU Radio block shims over the B200 kernels.
//
// Header-only blocks with the reference blocks' names, constructor arguments and stream contracts whose work()
// forwards ALL frames of one scheduler call to the `*_batch` entry of the matching kernel class -- one host<->device
// round trip and one kernel launch per work() call instead of one kernel call per frame:
//
//   reference block (lib/*_impl.cc)                       per-frame loop replaced
//   simple_modulator_cc      simple_modulator_cc_impl.cc:60-82     generic_work per frame          -> generic_work_batch
//   simple_receiver_cc       simple_receiver_cc_impl.cc:60-77      generic_work per frame          -> generic_work_batch
//   advanced_receiver_sb_cc  advanced_receiver_sb_cc_impl.cc:87-124 generic_work[_equalize] per frame, eq advances per frame
//   transmitter_cc           transmitter_cc_impl.cc:128-196        modulate once + add_frame per antenna -> generic_work_all_batch
//   channel_estimator_cc     channel_estimator_cc_impl.cc:96-131   estimate_frame + estimate_snr (tags snr_lin / cnr)
//   resource_mapper_cc       resource_mapper_cc_impl.cc:84-104     map_to_resources per frame
//   resource_demapper_cc     resource_demapper_cc_impl.cc:84-106   demap_from_resources per frame
//   cyclic_prefixer_cc       cyclic_prefixer_cc_impl.cc:87-106     generic_work per frame
//   short_burst_shaper       short_burst_shaper_impl.cc:161-182    sample path (padding + scale); the timed-command / message
//                                                                  part of that block is radio control and stays with GNU Radio
//
// The blocks live in gr::gfdm::b200 so that they can be loaded next to the original module.  They need GNU Radio's
// <gnuradio/sync_block.h>, <gnuradio/block.h>, <gnuradio/tagged_stream_block.h>, <gnuradio/io_signature.h> and <pmt/pmt.h>
// (3.9 API).  GNU Radio is not installable in this repository's build image: tests/test_block_shims.py compiles this
// header against a minimal stand-in of those headers (tests/stub_gnuradio/) and drives every work() on the GPU.
#ifndef INCLUDED_GFDM_B200_BLOCKS_HPP
#define INCLUDED_GFDM_B200_BLOCKS_HPP

#include <gnuradio/block.h>
#include <gnuradio/io_signature.h>
#include <gnuradio/sync_block.h>
#include <gnuradio/tagged_stream_block.h>
#include <pmt/pmt.h>

#include "gfdm_b200.hpp"

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace gr {
namespace gfdm {
namespace b200 {

typedef std::complex<float> gr_complex_t; // == gr_complex

// lib/simple_modulator_cc_impl.cc
class simple_modulator_cc : public gr::sync_block
{
public:
    typedef std::shared_ptr<simple_modulator_cc> sptr;
    static sptr make(int n_timeslots, int n_subcarriers, int overlap, std::vector<gr_complex_t> frequency_taps)
    {
        return gnuradio::make_block_sptr<simple_modulator_cc>(n_timeslots, n_subcarriers, overlap, frequency_taps);
    }
    simple_modulator_cc(int n_timeslots, int n_subcarriers, int overlap, std::vector<gr_complex_t> frequency_taps)
        : gr::sync_block("simple_modulator_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                         gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_kernel(new modulator_kernel_cc(n_timeslots, n_subcarriers, overlap, frequency_taps))
    {
        set_output_multiple(d_kernel->block_size());
    }
    int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const int n_blocks = noutput_items / d_kernel->block_size();
        d_kernel->generic_work_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], n_blocks);
        return noutput_items;
    }
    modulator_kernel_cc& kernel() { return *d_kernel; }

private:
    std::unique_ptr<modulator_kernel_cc> d_kernel;
};

// lib/simple_receiver_cc_impl.cc
class simple_receiver_cc : public gr::sync_block
{
public:
    typedef std::shared_ptr<simple_receiver_cc> sptr;
    static sptr make(int n_timeslots, int n_subcarriers, int overlap, std::vector<gr_complex_t> frequency_taps)
    {
        return gnuradio::make_block_sptr<simple_receiver_cc>(n_timeslots, n_subcarriers, overlap, frequency_taps);
    }
    simple_receiver_cc(int n_timeslots, int n_subcarriers, int overlap, std::vector<gr_complex_t> frequency_taps)
        : gr::sync_block("simple_receiver_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                         gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_kernel(new receiver_kernel_cc(n_timeslots, n_subcarriers, overlap, frequency_taps))
    {
        set_output_multiple(d_kernel->block_size());
    }
    int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const int n_blocks = noutput_items / d_kernel->block_size();
        d_kernel->generic_work_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], nullptr, n_blocks);
        return noutput_items;
    }
    receiver_kernel_cc& kernel() { return *d_kernel; }

private:
    std::unique_ptr<receiver_kernel_cc> d_kernel;
};

// lib/advanced_receiver_sb_cc_impl.cc: one or two inputs (samples [, per-frame channel]); tags of the deciding input
// are forwarded unchanged (:110-122)
class advanced_receiver_sb_cc : public gr::sync_block
{
public:
    typedef std::shared_ptr<advanced_receiver_sb_cc> sptr;
    // `constellation`: anything with points() (gr::digital::constellation_sptr in the reference); the decision rule is
    // the QPSK sign rule when the points are gr::digital::constellation_qpsk's, nearest point otherwise
    template <class ConstellationSptr>
    static sptr make(int n_timeslots, int n_subcarriers, int overlap, int ic_iter, std::vector<gr_complex_t> frequency_taps,
                     ConstellationSptr constellation, std::vector<int> subcarrier_map, int do_phase_compensation)
    {
        return gnuradio::make_block_sptr<advanced_receiver_sb_cc>(n_timeslots, n_subcarriers, overlap, ic_iter, frequency_taps,
                                                                  to_constellation(constellation->points()), subcarrier_map,
                                                                  do_phase_compensation);
    }
    advanced_receiver_sb_cc(int n_timeslots, int n_subcarriers, int overlap, int ic_iter,
                            std::vector<gr_complex_t> frequency_taps, const gr::gfdm::constellation& constellation,
                            std::vector<int> subcarrier_map, int do_phase_compensation)
        : gr::sync_block("advanced_receiver_sb_cc", gr::io_signature::make(1, 2, sizeof(gr_complex_t)),
                         gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_adv_kernel(new advanced_receiver_kernel_cc(n_timeslots, n_subcarriers, overlap, frequency_taps, subcarrier_map,
                                                       ic_iter, constellation, do_phase_compensation))
    {
        set_output_multiple(d_adv_kernel->block_size());
        set_tag_propagation_policy(TPP_DONT);
    }
    static gr::gfdm::constellation to_constellation(const std::vector<gr_complex_t>& points)
    {
        const gr::gfdm::constellation q = gr::gfdm::constellation::qpsk();
        gr::gfdm::constellation c;
        c.points = points;
        c.decision_rule = (points == q.points) ? q.decision_rule : (int)GFDM_DECISION_NEAREST;
        return c;
    }
    int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const int bs = d_adv_kernel->block_size();
        const int n_blocks = noutput_items / bs;
        const gr_complex_t* in_eq = input_items.size() > 1 ? (const gr_complex_t*)input_items[1] : nullptr;
        d_adv_kernel->generic_work_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], in_eq, n_blocks);
        std::vector<tag_t> tags;
        get_tags_in_window(tags, in_eq ? 1 : 0, 0, n_blocks * bs);
        for (auto t : tags) add_item_tag(0, t);
        return n_blocks * bs;
    }
    advanced_receiver_kernel_cc& kernel() { return *d_adv_kernel; }

private:
    std::unique_ptr<advanced_receiver_kernel_cc> d_adv_kernel;
};

// lib/transmitter_cc_impl.cc: one output per cyclic shift; every frame is modulated once and framed per antenna --
// here all of it is one kernel per work() call
class transmitter_cc : public gr::block
{
public:
    typedef std::shared_ptr<transmitter_cc> sptr;
    static sptr make(int timeslots, int subcarriers, int active_subcarriers, int cp_len, int cs_len, int ramp_len,
                     std::vector<int> subcarrier_map, bool per_timeslot, int overlap, std::vector<gr_complex_t> frequency_taps,
                     std::vector<gr_complex_t> window_taps, std::vector<int> cyclic_shifts,
                     std::vector<std::vector<gr_complex_t>> preambles, const std::string& tsb_tag_key = "")
    {
        return gnuradio::make_block_sptr<transmitter_cc>(timeslots, subcarriers, active_subcarriers, cp_len, cs_len, ramp_len,
                                                         subcarrier_map, per_timeslot, overlap, frequency_taps, window_taps,
                                                         cyclic_shifts, preambles, tsb_tag_key);
    }
    transmitter_cc(int timeslots, int subcarriers, int active_subcarriers, int cp_len, int cs_len, int ramp_len,
                   std::vector<int> subcarrier_map, bool per_timeslot, int overlap, std::vector<gr_complex_t> frequency_taps,
                   std::vector<gr_complex_t> window_taps, std::vector<int> cyclic_shifts,
                   std::vector<std::vector<gr_complex_t>> preambles, const std::string& tsb_tag_key)
        : gr::block("transmitter_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                    gr::io_signature::make((int)cyclic_shifts.size(), (int)cyclic_shifts.size(), sizeof(gr_complex_t))),
          d_length_tag_key_str(tsb_tag_key),
          d_length_tag_key(pmt::string_to_symbol(tsb_tag_key)),
          d_kernel(new transmitter_kernel(timeslots, subcarriers, active_subcarriers, cp_len, cs_len, ramp_len, subcarrier_map,
                                          per_timeslot, overlap, frequency_taps, window_taps, cyclic_shifts, preambles))
    {
        set_relative_rate(1.0 * d_kernel->output_vector_size() / d_kernel->input_vector_size());
        set_fixed_rate(true);
        set_output_multiple(d_kernel->output_vector_size());
    }
    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        for (size_t i = 0; i < ninput_items_required.size(); ++i)
            ninput_items_required[i] = fixed_rate_noutput_to_ninput(noutput_items);
    }
    int fixed_rate_ninput_to_noutput(int ninput) override
    {
        return (ninput / d_kernel->input_vector_size()) * d_kernel->output_vector_size();
    }
    int fixed_rate_noutput_to_ninput(int noutput) override
    {
        return (noutput / d_kernel->output_vector_size()) * d_kernel->input_vector_size();
    }
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) override
    {
        const int os = d_kernel->output_vector_size(), is = d_kernel->input_vector_size();
        const int n_frames = std::min(noutput_items / os, ninput_items[0] / is);
        if (!d_length_tag_key_str.empty()) {
            std::vector<tag_t> tags;
            get_tags_in_range(tags, 0, nitems_read(0), nitems_read(0) + (uint64_t)n_frames * is, d_length_tag_key);
            for (auto tag : tags)
                if (pmt::eqv(tag.key, d_length_tag_key)) remove_item_tag(0, tag);
        }
        const size_t n_ant = output_items.size();
        if (n_ant == 1) {
            d_kernel->generic_work_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], is, n_frames);
        } else {
            // the batch entry lays the antennas out as [antenna][frame][sample]; GNU Radio hands one buffer per port
            d_all.resize(n_ant * (size_t)n_frames * os);
            d_kernel->generic_work_all_batch(d_all.data(), (const gr_complex_t*)input_items[0], is, n_frames);
            for (size_t a = 0; a < n_ant; ++a)
                std::memcpy(output_items[a], d_all.data() + a * (size_t)n_frames * os, sizeof(gr_complex_t) * (size_t)n_frames * os);
        }
        consume_each(n_frames * is);
        if (!d_length_tag_key_str.empty())
            for (int i = 0; i < n_frames; ++i)
                for (unsigned port = 0; port < output_items.size(); ++port)
                    add_item_tag(port, nitems_written(port) + (uint64_t)i * os, d_length_tag_key, pmt::from_long(os));
        return n_frames * os;
    }
    transmitter_kernel& kernel() { return *d_kernel; }

private:
    std::string d_length_tag_key_str;
    pmt::pmt_t d_length_tag_key;
    std::unique_ptr<transmitter_kernel> d_kernel;
    std::vector<gr_complex_t> d_all;
};

// lib/channel_estimator_cc_impl.cc: 2*fft_len preamble samples in, timeslots*fft_len channel bins out, SNR tags per frame
class channel_estimator_cc : public gr::block
{
public:
    typedef std::shared_ptr<channel_estimator_cc> sptr;
    static sptr make(int timeslots, int fft_len, int active_subcarriers, bool is_dc_free, int which_estimator,
                     std::vector<gr_complex_t> preamble)
    {
        return gnuradio::make_block_sptr<channel_estimator_cc>(timeslots, fft_len, active_subcarriers, is_dc_free,
                                                               which_estimator, preamble);
    }
    channel_estimator_cc(int timeslots, int fft_len, int active_subcarriers, bool is_dc_free, int which_estimator,
                         std::vector<gr_complex_t> preamble)
        : gr::block("channel_estimator_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                    gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_estimator_kernel(new preamble_channel_estimator_cc(timeslots, fft_len, active_subcarriers, is_dc_free,
                                                               which_estimator, preamble))
    {
        set_relative_rate(timeslots / 2.0);
        set_fixed_rate(true);
        set_output_multiple(fft_len * timeslots);
    }
    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        for (size_t i = 0; i < ninput_items_required.size(); ++i)
            ninput_items_required[i] = fixed_rate_noutput_to_ninput(noutput_items);
    }
    int fixed_rate_ninput_to_noutput(int ninput) override { return ninput * d_estimator_kernel->timeslots() / 2; }
    int fixed_rate_noutput_to_ninput(int noutput) override { return 2 * noutput / d_estimator_kernel->timeslots(); }
    int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) override
    {
        const int invec_len = 2 * d_estimator_kernel->fft_len();
        const int frame_len = d_estimator_kernel->frame_len();
        const int n_frames = noutput_items / frame_len;
        const int A = d_estimator_kernel->active_subcarriers();
        const gr_complex_t* in = (const gr_complex_t*)input_items[0];
        d_estimator_kernel->estimate_frame_batch((gr_complex_t*)output_items[0], in, n_frames);
        d_snr.resize((size_t)n_frames);
        d_cnr.resize((size_t)n_frames * A);
        d_estimator_kernel->estimate_snr_batch(d_snr.data(), d_cnr.data(), in, n_frames);
        for (int i = 0; i < n_frames; ++i) {
            const std::vector<float> cnrs(d_cnr.begin() + (size_t)i * A, d_cnr.begin() + (size_t)(i + 1) * A);
            add_item_tag(0, nitems_written(0) + (uint64_t)i * frame_len, pmt::intern("snr_lin"), pmt::from_float(d_snr[i]));
            add_item_tag(0, nitems_written(0) + (uint64_t)i * frame_len, pmt::intern("cnr"), pmt::init_f32vector(cnrs.size(), cnrs));
        }
        consume_each(n_frames * invec_len);
        return n_frames * frame_len;
    }
    preamble_channel_estimator_cc& kernel() { return *d_estimator_kernel; }

private:
    std::unique_ptr<preamble_channel_estimator_cc> d_estimator_kernel;
    std::vector<float> d_snr, d_cnr;
};

// lib/resource_mapper_cc_impl.cc / lib/resource_demapper_cc_impl.cc
template <bool IS_MAPPER>
class resource_mapping_block : public gr::block
{
public:
    resource_mapping_block(int timeslots, int subcarriers, int active_subcarriers, std::vector<int> subcarrier_map, bool per_timeslot)
        : gr::block(IS_MAPPER ? "resource_mapper_cc" : "resource_demapper_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                    gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_kernel(new resource_mapper_kernel_cc(timeslots, subcarriers, active_subcarriers, subcarrier_map, per_timeslot, IS_MAPPER))
    {
        set_relative_rate(1.0 * d_kernel->output_vector_size() / d_kernel->input_vector_size());
        set_fixed_rate(true);
        set_output_multiple((int)d_kernel->output_vector_size());
    }
    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        ninput_items_required[0] = fixed_rate_noutput_to_ninput(noutput_items);
    }
    int fixed_rate_ninput_to_noutput(int ninput) override
    {
        return (ninput / (int)d_kernel->input_vector_size()) * (int)d_kernel->output_vector_size();
    }
    int fixed_rate_noutput_to_ninput(int noutput) override
    {
        return (noutput / (int)d_kernel->output_vector_size()) * (int)d_kernel->input_vector_size();
    }
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) override
    {
        const int is = (int)d_kernel->input_vector_size(), os = (int)d_kernel->output_vector_size();
        const int n_frames = std::min(noutput_items / os, ninput_items[0] / is);
        if (IS_MAPPER)
            d_kernel->map_to_resources_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], (size_t)is, n_frames);
        else
            d_kernel->demap_from_resources_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0], (size_t)os, n_frames);
        consume_each(n_frames * is);
        return n_frames * os;
    }
    resource_mapper_kernel_cc& kernel() { return *d_kernel; }

private:
    std::unique_ptr<resource_mapper_kernel_cc> d_kernel;
};
class resource_mapper_cc : public resource_mapping_block<true>
{
public:
    typedef std::shared_ptr<resource_mapper_cc> sptr;
    using resource_mapping_block<true>::resource_mapping_block;
    static sptr make(int timeslots, int subcarriers, int active_subcarriers, std::vector<int> subcarrier_map, bool per_timeslot = true)
    {
        return gnuradio::make_block_sptr<resource_mapper_cc>(timeslots, subcarriers, active_subcarriers, subcarrier_map, per_timeslot);
    }
};
class resource_demapper_cc : public resource_mapping_block<false>
{
public:
    typedef std::shared_ptr<resource_demapper_cc> sptr;
    using resource_mapping_block<false>::resource_mapping_block;
    static sptr make(int timeslots, int subcarriers, int active_subcarriers, std::vector<int> subcarrier_map, bool per_timeslot = true)
    {
        return gnuradio::make_block_sptr<resource_demapper_cc>(timeslots, subcarriers, active_subcarriers, subcarrier_map, per_timeslot);
    }
};

// lib/cyclic_prefixer_cc_impl.cc
class cyclic_prefixer_cc : public gr::block
{
public:
    typedef std::shared_ptr<cyclic_prefixer_cc> sptr;
    static sptr make(int block_len, int cp_len, int cs_len, int ramp_len, std::vector<gr_complex_t> window_taps, int cyclic_shift = 0)
    {
        return gnuradio::make_block_sptr<cyclic_prefixer_cc>(block_len, cp_len, cs_len, ramp_len, window_taps, cyclic_shift);
    }
    cyclic_prefixer_cc(int block_len, int cp_len, int cs_len, int ramp_len, std::vector<gr_complex_t> window_taps, int cyclic_shift)
        : gr::block("cyclic_prefixer_cc", gr::io_signature::make(1, 1, sizeof(gr_complex_t)),
                    gr::io_signature::make(1, 1, sizeof(gr_complex_t))),
          d_kernel(new add_cyclic_prefix_cc(block_len, cp_len, cs_len, ramp_len, window_taps, cyclic_shift))
    {
        set_relative_rate(1.0 * d_kernel->frame_size() / d_kernel->block_size());
        set_fixed_rate(true);
        set_output_multiple(d_kernel->frame_size());
    }
    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        for (size_t i = 0; i < ninput_items_required.size(); ++i)
            ninput_items_required[i] = fixed_rate_noutput_to_ninput(noutput_items);
    }
    int fixed_rate_ninput_to_noutput(int ninput) override { return (ninput / d_kernel->block_size()) * d_kernel->frame_size(); }
    int fixed_rate_noutput_to_ninput(int noutput) override { return (noutput / d_kernel->frame_size()) * d_kernel->block_size(); }
    int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const int n_frames = noutput_items / d_kernel->frame_size();
        d_kernel->add_cyclic_prefix_batch((gr_complex_t*)output_items[0], (const gr_complex_t*)input_items[0],
                                          d_kernel->cyclic_shift(), n_frames);
        consume_each(n_frames * d_kernel->block_size());
        return n_frames * d_kernel->frame_size();
    }
    add_cyclic_prefix_cc& kernel() { return *d_kernel; }

private:
    std::unique_ptr<add_cyclic_prefix_cc> d_kernel;
};

// lib/short_burst_shaper_impl.cc: the sample path of the tagged-stream block (one burst per port and work() call)
class short_burst_shaper : public gr::tagged_stream_block
{
public:
    typedef std::shared_ptr<short_burst_shaper> sptr;
    static sptr make(int pre_padding, int post_padding, gr_complex_t scale, const unsigned nports = 1,
                     const std::string& length_tag_name = "packet_len")
    {
        return gnuradio::make_block_sptr<short_burst_shaper>(pre_padding, post_padding, scale, nports, length_tag_name);
    }
    short_burst_shaper(int pre_padding, int post_padding, gr_complex_t scale, const unsigned nports, const std::string& length_tag_name)
        : gr::tagged_stream_block("short_burst_shaper", gr::io_signature::make((int)nports, (int)nports, sizeof(gr_complex_t)),
                                  gr::io_signature::make((int)nports, (int)nports, sizeof(gr_complex_t)), length_tag_name),
          d_pre_padding(pre_padding),
          d_post_padding(post_padding),
          d_kernel(new burst_shaper(pre_padding, post_padding, scale))
    {
    }
    int calculate_output_stream_length(const gr_vector_int& ninput_items) override
    {
        return ninput_items[0] + d_pre_padding + d_post_padding;
    }
    int work(int, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        for (unsigned port = 0; port < input_items.size(); ++port)
            d_kernel->work_batch((gr_complex_t*)output_items[port], (const gr_complex_t*)input_items[port], ninput_items[port], 1);
        return ninput_items[0] + d_pre_padding + d_post_padding;
    }

private:
    int d_pre_padding, d_post_padding;
    std::unique_ptr<burst_shaper> d_kernel;
};

} // namespace b200
} // namespace gfdm
} // namespace gr

#endif /* INCLUDED_GFDM_B200_BLOCKS_HPP */
