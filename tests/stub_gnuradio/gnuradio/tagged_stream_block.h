// Minimal stand-in for <gnuradio/tagged_stream_block.h> (test infrastructure, see pmt/pmt.h).
#ifndef STUB_GR_TAGGED_STREAM_BLOCK_H
#define STUB_GR_TAGGED_STREAM_BLOCK_H
#include <gnuradio/block.h>
namespace gr {
class tagged_stream_block : public block
{
public:
    tagged_stream_block(const std::string& name, io_signature::sptr in, io_signature::sptr out, const std::string& length_tag_key)
        : block(name, in, out), d_length_tag_key_str(length_tag_key) {}
    virtual int calculate_output_stream_length(const gr_vector_int& ninput_items) = 0;
    virtual int work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) = 0;
    const std::string& length_tag_key() const { return d_length_tag_key_str; }

private:
    std::string d_length_tag_key_str;
};
} // namespace gr
#endif
