/*
 * Shim <gnuradio/digital/constellation.h>: the two members of
 * gr::digital::constellation that lib/advanced_receiver_kernel_cc.cc:114-120
 * uses -- points() and decision_maker().  GNU Radio is not installed here; its
 * gr-digital sources are not part of /root/reference.  The QPSK rule below
 * restates gr::digital::constellation_qpsk of GNU Radio 3.9
 * (points {-1-j, 1-j, -1+j, 1+j} * 0.707107, index = 2*(im>0) + (re>0));
 * any other constellation decides by nearest point (first minimum), which is
 * what gr::digital::constellation::get_closest_point does.
 * TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORACLE_SHIM_GR_DIGITAL_CONSTELLATION_H
#define ORACLE_SHIM_GR_DIGITAL_CONSTELLATION_H
#include <gnuradio/gr_complex.h>
#include <memory>
#include <vector>
namespace gr {
namespace digital {
class constellation
{
public:
    constellation(std::vector<gr_complex> pts, int rule) : d_points(pts), d_rule(rule) {}
    virtual ~constellation() {}
    std::vector<gr_complex> points() { return d_points; }
    virtual unsigned int decision_maker(const gr_complex* sample)
    {
        if (d_rule == 1) return 2 * (sample->imag() > 0) + (sample->real() > 0);
        unsigned int best = 0;
        float dmin = 0.0f;
        for (unsigned int i = 0; i < d_points.size(); ++i) {
            const float dr = sample->real() - d_points[i].real();
            const float di = sample->imag() - d_points[i].imag();
            const float d = dr * dr + di * di;
            if (i == 0 || d < dmin) {
                dmin = d;
                best = i;
            }
        }
        return best;
    }

protected:
    std::vector<gr_complex> d_points;
    int d_rule;
};
typedef std::shared_ptr<constellation> constellation_sptr;
} // namespace digital
} // namespace gr
#endif
