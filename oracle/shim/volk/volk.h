/*
 * Shim <volk/volk.h>: the ten VOLK kernels gr-gfdm's kernel sources call
 * (SURVEY.md section 2.3), as the obvious scalar loops with VOLK's generic
 * arithmetic.  VOLK is not installed here.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORACLE_SHIM_VOLK_H
#define ORACLE_SHIM_VOLK_H
#include <complex>
typedef std::complex<float> lv_32fc_t;

static inline lv_32fc_t shim_cmul(const lv_32fc_t a, const lv_32fc_t b)
{
    return lv_32fc_t(a.real() * b.real() - a.imag() * b.imag(),
                     a.real() * b.imag() + a.imag() * b.real());
}

static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b,
                                              unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = shim_cmul(a[i], b[i]);
}
static inline void volk_32f_x2_add_32f(float* c, const float* a, const float* b, unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = a[i] + b[i];
}
static inline void volk_32f_x2_subtract_32f(float* c, const float* a, const float* b, unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = a[i] - b[i];
}
static inline void volk_32fc_s32fc_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a,
                                                 const lv_32fc_t scalar, unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = shim_cmul(a[i], scalar);
}
static inline void volk_32fc_x2_conjugate_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* a,
                                                        const lv_32fc_t* b, unsigned int n)
{
    lv_32fc_t acc(0.0f, 0.0f);
    for (unsigned int i = 0; i < n; ++i) acc += shim_cmul(a[i], std::conj(b[i]));
    *result = acc;
}
static inline void volk_32fc_x2_divide_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b,
                                            unsigned int n)
{
    // volk generic: c = a * conj(b) / |b|^2
    for (unsigned int i = 0; i < n; ++i) {
        const lv_32fc_t num = shim_cmul(a[i], std::conj(b[i]));
        const float den = b[i].real() * b[i].real() + b[i].imag() * b[i].imag();
        c[i] = lv_32fc_t(num.real() / den, num.imag() / den);
    }
}
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* a,
                                               const float* taps, unsigned int n)
{
    float re = 0.0f, im = 0.0f;
    for (unsigned int i = 0; i < n; ++i) {
        re += a[i].real() * taps[i];
        im += a[i].imag() * taps[i];
    }
    *result = lv_32fc_t(re, im);
}
static inline void volk_32f_s32f_multiply_32f(float* c, const float* a, const float scalar,
                                              unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = a[i] * scalar;
}
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t* out, const lv_32fc_t* in,
                                                   const lv_32fc_t phase_inc, lv_32fc_t* phase,
                                                   unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) {
        out[i] = shim_cmul(in[i], *phase);
        *phase = shim_cmul(*phase, phase_inc);
    }
}
static inline void volk_32fc_conjugate_32fc(lv_32fc_t* c, const lv_32fc_t* a, unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i) c[i] = std::conj(a[i]);
}
#endif
