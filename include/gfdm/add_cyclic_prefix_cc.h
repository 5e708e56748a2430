/* Forwarding header: keeps gr-gfdm's include name <gfdm/add_cyclic_prefix_cc.h> working against the
 * B200 engine.  The class gr::gfdm::add_cyclic_prefix_cc lives in gfdm_b200.hpp. */
#ifndef INCLUDED_GFDM_B200_FWD_ADD_CYCLIC_PREFIX_CC_H
#define INCLUDED_GFDM_B200_FWD_ADD_CYCLIC_PREFIX_CC_H
#include "../gfdm_b200.hpp"
#endif
