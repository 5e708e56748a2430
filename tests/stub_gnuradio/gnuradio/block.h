// Minimal stand-in for <gnuradio/block.h> (test infrastructure, see pmt/pmt.h): the members of gr::block the shims call,
// with just enough state for a test to act as the scheduler (item counters, a tag store per port).
#ifndef STUB_GR_BLOCK_H
#define STUB_GR_BLOCK_H
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>

#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gr {
struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value, srcid;
};
class block
{
public:
    enum tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2 };
    block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out), d_read(8, 0), d_written(8, 0), d_in_tags(8), d_out_tags(8) {}
    virtual ~block() {}
    virtual int general_work(int, gr_vector_int&, gr_vector_const_void_star&, gr_vector_void_star&) { return 0; }
    virtual void forecast(int noutput_items, gr_vector_int& req) { for (auto& r : req) r = noutput_items; }
    virtual int fixed_rate_ninput_to_noutput(int n) { return n; }
    virtual int fixed_rate_noutput_to_ninput(int n) { return n; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    void set_relative_rate(double r) { d_relative_rate = r; }
    double relative_rate() const { return d_relative_rate; }
    void set_fixed_rate(bool f) { d_fixed_rate = f; }
    void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tpp = p; }
    void consume_each(int n) { d_consumed = n; }
    int consumed() const { return d_consumed; }
    uint64_t nitems_read(unsigned port) const { return d_read[port]; }
    uint64_t nitems_written(unsigned port) const { return d_written[port]; }
    void add_item_tag(unsigned port, uint64_t offset, const pmt::pmt_t& key, const pmt::pmt_t& value)
    {
        tag_t t; t.offset = offset; t.key = key; t.value = value; d_out_tags[port].push_back(t);
    }
    void add_item_tag(unsigned port, const tag_t& t) { d_out_tags[port].push_back(t); }
    void remove_item_tag(unsigned port, const tag_t& t)
    {
        auto& v = d_in_tags[port];
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].offset == t.offset && pmt::eqv(v[i].key, t.key)) { v.erase(v.begin() + i); return; }
    }
    void get_tags_in_range(std::vector<tag_t>& out, unsigned port, uint64_t lo, uint64_t hi, const pmt::pmt_t& key)
    {
        out.clear();
        for (const tag_t& t : d_in_tags[port])
            if (t.offset >= lo && t.offset < hi && pmt::eqv(t.key, key)) out.push_back(t);
    }
    void get_tags_in_window(std::vector<tag_t>& out, unsigned port, uint64_t lo, uint64_t hi)
    {
        out.clear();
        for (const tag_t& t : d_in_tags[port])
            if (t.offset >= d_read[port] + lo && t.offset < d_read[port] + hi) out.push_back(t);
    }
    const std::string& name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }
    // test-side access (the "scheduler")
    std::vector<tag_t>& test_input_tags(unsigned port) { return d_in_tags[port]; }
    std::vector<tag_t>& test_output_tags(unsigned port) { return d_out_tags[port]; }
    void test_advance(unsigned in_port, uint64_t n_read, unsigned out_port, uint64_t n_written) { d_read[in_port] += n_read; d_written[out_port] += n_written; }

private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    int d_output_multiple = 1, d_consumed = 0;
    double d_relative_rate = 1.0;
    bool d_fixed_rate = false;
    tag_propagation_policy_t d_tpp = TPP_ALL_TO_ALL;
    std::vector<uint64_t> d_read, d_written;
    std::vector<std::vector<tag_t>> d_in_tags, d_out_tags;
};
} // namespace gr

namespace gnuradio {
template <class T, class... Args>
std::shared_ptr<T> make_block_sptr(Args&&... args) { return std::shared_ptr<T>(new T(std::forward<Args>(args)...)); }
} // namespace gnuradio
#endif
