/* Forwarding header: keeps gr-gfdm's include name <gfdm/gfdm_kernel_utils.h> working against the
 * B200 engine.  The class gr::gfdm::gfdm_kernel_utils lives in gfdm_b200.hpp. */
#ifndef INCLUDED_GFDM_B200_FWD_GFDM_KERNEL_UTILS_H
#define INCLUDED_GFDM_B200_FWD_GFDM_KERNEL_UTILS_H
#include "../gfdm_b200.hpp"
#endif
