// Minimal stand-in for <gnuradio/sync_block.h> (test infrastructure, see pmt/pmt.h).
#ifndef STUB_GR_SYNC_BLOCK_H
#define STUB_GR_SYNC_BLOCK_H
#include <gnuradio/block.h>
namespace gr {
class sync_block : public block
{
public:
    sync_block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
    virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
    int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& in, gr_vector_void_star& out) override
    {
        const int r = work(noutput_items, in, out);
        if (r > 0) consume_each(r);
        return r;
    }
};
} // namespace gr
#endif
