#ifndef ORACLE_SHIM_GR_COMPLEX_H
#define ORACLE_SHIM_GR_COMPLEX_H
#include <complex>
typedef std::complex<float> gr_complex;
#endif
