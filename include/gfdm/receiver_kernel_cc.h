/* Forwarding header: keeps gr-gfdm's include name <gfdm/receiver_kernel_cc.h> working against the
 * B200 engine.  The class gr::gfdm::receiver_kernel_cc lives in gfdm_b200.hpp. */
#ifndef INCLUDED_GFDM_B200_FWD_RECEIVER_KERNEL_CC_H
#define INCLUDED_GFDM_B200_FWD_RECEIVER_KERNEL_CC_H
#include "../gfdm_b200.hpp"
#endif
