// Minimal stand-in for <gnuradio/io_signature.h> (test infrastructure, see pmt/pmt.h).
#ifndef STUB_GR_IO_SIGNATURE_H
#define STUB_GR_IO_SIGNATURE_H
#include <memory>
namespace gr {
class io_signature
{
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    {
        return sptr(new io_signature{ min_streams, max_streams, sizeof_stream_item });
    }
    int min_streams, max_streams, sizeof_stream_item;
};
} // namespace gr
#endif
