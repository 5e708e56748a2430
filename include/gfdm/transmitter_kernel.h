/* Forwarding header: keeps gr-gfdm's include name <gfdm/transmitter_kernel.h> working against the
 * B200 engine.  The class gr::gfdm::transmitter_kernel lives in gfdm_b200.hpp. */
#ifndef INCLUDED_GFDM_B200_FWD_TRANSMITTER_KERNEL_H
#define INCLUDED_GFDM_B200_FWD_TRANSMITTER_KERNEL_H
#include "../gfdm_b200.hpp"
#endif
