// Scalar float mixed-radix DFT behind the fftw3.h shim (see fftw3.h here).
// Decimation-in-time recursion over the prime-ish factorisation of n with
// dedicated radix-2/3/4/5 butterflies and an O(p^2) butterfly for any other
// prime factor.  Twiddles are generated in double precision.
// TEST INFRASTRUCTURE ONLY.
#include "fftw3.h"
#include <cmath>
#include <complex>
#include <vector>

typedef std::complex<float> cpx;

struct fftwf_plan_s {
    int n;
    int sign;
    const cpx* in;
    cpx* out;
    std::vector<int> factors; // pairs (radix, remaining length)
    std::vector<cpx> tw;      // tw[j] = exp(sign * 2 pi i j / n)
    std::vector<cpx> tmp;
};

static void factorize(int n, std::vector<int>& f)
{
    int p = 4;
    const double floor_sqrt = std::floor(std::sqrt((double)n));
    do {
        while (n % p) {
            switch (p) {
            case 4: p = 2; break;
            case 2: p = 3; break;
            default: p += 2; break;
            }
            if (p > floor_sqrt) p = n;
        }
        n /= p;
        f.push_back(p);
        f.push_back(n);
    } while (n > 1);
}

static inline cpx cmul(const cpx a, const cpx b)
{
    return cpx(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}

static void bfly2(cpx* F, const size_t fstride, const fftwf_plan_s* st, int m)
{
    const cpx* tw = st->tw.data();
    for (int u = 0; u < m; ++u) {
        const cpx t = cmul(F[u + m], tw[u * fstride]);
        F[u + m] = F[u] - t;
        F[u] += t;
    }
}

static void bfly4(cpx* F, const size_t fstride, const fftwf_plan_s* st, int m)
{
    const cpx* tw = st->tw.data();
    const bool inv = st->sign > 0;
    for (int u = 0; u < m; ++u) {
        const cpx s0 = cmul(F[u + m], tw[u * fstride]);
        const cpx s1 = cmul(F[u + 2 * m], tw[u * fstride * 2]);
        const cpx s2 = cmul(F[u + 3 * m], tw[u * fstride * 3]);
        const cpx s5 = F[u] - s1;
        const cpx a0 = F[u] + s1;
        const cpx s3 = s0 + s2;
        const cpx s4 = s0 - s2;
        F[u + 2 * m] = a0 - s3;
        F[u] = a0 + s3;
        // multiply s4 by -i (forward) or +i (inverse)
        const cpx r = inv ? cpx(-s4.imag(), s4.real()) : cpx(s4.imag(), -s4.real());
        F[u + m] = s5 + r;
        F[u + 3 * m] = s5 - r;
    }
}

static void bfly3(cpx* F, const size_t fstride, const fftwf_plan_s* st, int m)
{
    const cpx* tw = st->tw.data();
    const float epi3_i = st->tw[fstride * m].imag();
    for (int u = 0; u < m; ++u) {
        const cpx s1 = cmul(F[u + m], tw[u * fstride]);
        const cpx s2 = cmul(F[u + 2 * m], tw[u * fstride * 2]);
        const cpx s3 = s1 + s2;
        const cpx s0 = (s1 - s2) * epi3_i;
        const cpx c = F[u] - 0.5f * s3;
        F[u] += s3;
        F[u + 2 * m] = cpx(c.real() + s0.imag(), c.imag() - s0.real());
        F[u + m] = cpx(c.real() - s0.imag(), c.imag() + s0.real());
    }
}

static void bfly5(cpx* F, const size_t fstride, const fftwf_plan_s* st, int m)
{
    const cpx* tw = st->tw.data();
    const cpx ya = tw[fstride * m];
    const cpx yb = tw[fstride * 2 * m];
    for (int u = 0; u < m; ++u) {
        const cpx s0 = F[u];
        const cpx s1 = cmul(F[u + m], tw[u * fstride]);
        const cpx s2 = cmul(F[u + 2 * m], tw[2 * u * fstride]);
        const cpx s3 = cmul(F[u + 3 * m], tw[3 * u * fstride]);
        const cpx s4 = cmul(F[u + 4 * m], tw[4 * u * fstride]);
        const cpx s7 = s1 + s4, s10 = s1 - s4, s8 = s2 + s3, s9 = s2 - s3;
        F[u] = s0 + s7 + s8;
        const cpx s5(s0.real() + s7.real() * ya.real() + s8.real() * yb.real(),
                     s0.imag() + s7.imag() * ya.real() + s8.imag() * yb.real());
        const cpx s6(s10.imag() * ya.imag() + s9.imag() * yb.imag(),
                     -s10.real() * ya.imag() - s9.real() * yb.imag());
        F[u + m] = s5 - s6;
        F[u + 4 * m] = s5 + s6;
        const cpx s11(s0.real() + s7.real() * yb.real() + s8.real() * ya.real(),
                      s0.imag() + s7.imag() * yb.real() + s8.imag() * ya.real());
        const cpx s12(-s10.imag() * yb.imag() + s9.imag() * ya.imag(),
                      s10.real() * yb.imag() - s9.real() * ya.imag());
        F[u + 2 * m] = s11 + s12;
        F[u + 3 * m] = s11 - s12;
    }
}

static void bfly_generic(cpx* F, const size_t fstride, fftwf_plan_s* st, int m, int p)
{
    const cpx* tw = st->tw.data();
    const int n = st->n;
    cpx* scratch = st->tmp.data();
    for (int u = 0; u < m; ++u) {
        int k = u;
        for (int q1 = 0; q1 < p; ++q1) {
            scratch[q1] = F[k];
            k += m;
        }
        k = u;
        for (int q1 = 0; q1 < p; ++q1) {
            size_t twidx = 0;
            cpx acc = scratch[0];
            for (int q = 1; q < p; ++q) {
                twidx += fstride * k;
                if (twidx >= (size_t)n) twidx -= n;
                acc += cmul(scratch[q], tw[twidx]);
            }
            F[k] = acc;
            k += m;
        }
    }
}

static void work(cpx* Fout, const cpx* f, const size_t fstride, const int* factors, fftwf_plan_s* st)
{
    const int p = factors[0];
    const int m = factors[1];
    if (m == 1) {
        for (int k = 0; k < p; ++k) Fout[k] = f[k * fstride];
    } else {
        for (int k = 0; k < p; ++k) work(Fout + k * m, f + k * fstride, fstride * p, factors + 2, st);
    }
    switch (p) {
    case 2: bfly2(Fout, fstride, st, m); break;
    case 3: bfly3(Fout, fstride, st, m); break;
    case 4: bfly4(Fout, fstride, st, m); break;
    case 5: bfly5(Fout, fstride, st, m); break;
    default: bfly_generic(Fout, fstride, st, m, p); break;
    }
}

extern "C" {

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned)
{
    fftwf_plan_s* p = new fftwf_plan_s;
    p->n = n;
    p->sign = sign;
    p->in = reinterpret_cast<const cpx*>(in);
    p->out = reinterpret_cast<cpx*>(out);
    factorize(n, p->factors);
    p->tw.resize(n);
    const double s = sign < 0 ? -1.0 : 1.0;
    for (int j = 0; j < n; ++j) {
        const double ph = s * 2.0 * M_PI * (double)j / (double)n;
        p->tw[j] = cpx((float)std::cos(ph), (float)std::sin(ph));
    }
    int pmax = 1;
    for (size_t i = 0; i < p->factors.size(); i += 2) pmax = std::max(pmax, p->factors[i]);
    p->tmp.resize(pmax);
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    if (p->n == 1) {
        p->out[0] = p->in[0];
        return;
    }
    work(p->out, p->in, 1, p->factors.data(), p);
}

void fftwf_destroy_plan(fftwf_plan p) { delete p; }
int fftwf_import_wisdom_from_file(FILE*) { return 1; }
void fftwf_export_wisdom_to_file(FILE*) {}
}
