/* Forwarding header: keeps gr-gfdm's include name <gfdm/preamble_channel_estimator_cc.h> working against the
 * B200 engine.  The class gr::gfdm::preamble_channel_estimator_cc lives in gfdm_b200.hpp. */
#ifndef INCLUDED_GFDM_B200_FWD_PREAMBLE_CHANNEL_ESTIMATOR_CC_H
#define INCLUDED_GFDM_B200_FWD_PREAMBLE_CHANNEL_ESTIMATOR_CC_H
#include "../gfdm_b200.hpp"
#endif
