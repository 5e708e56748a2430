/* Shim: included but unused by lib/gfdm_kernel_utils.cc:21, lib/transmitter_kernel.cc:25,
 * lib/advanced_receiver_kernel_cc.cc:25.  The real header transitively provides the
 * C/C++ standard headers below (advanced_receiver_kernel_cc.cc relies on it for memset).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GR_IO_SIGNATURE_H
#define ORACLE_SHIM_GR_IO_SIGNATURE_H
#include <gnuradio/gr_complex.h>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#endif
