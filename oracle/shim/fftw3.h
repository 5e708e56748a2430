/*
 * Shim <fftw3.h> for building the UNMODIFIED gr-gfdm kernel sources as a test
 * oracle (oracle/_ref).  FFTW3 is not installed in this image and there is no
 * network, so the handful of FFTW entry points the reference calls
 * (lib/gfdm_kernel_utils.cc:32-57: plan_dft_1d / execute / destroy_plan /
 * wisdom import+export) are supplied by oracle/shim/shim_fft.cc: a scalar
 * single-precision mixed-radix DFT with double-precision-generated twiddles.
 * Same semantics as FFTW: out-of-place, unnormalised, sign = -1 forward.
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_import_wisdom_from_file(FILE* f);
void fftwf_export_wisdom_to_file(FILE* f);
#ifdef __cplusplus
}
#endif
#endif
