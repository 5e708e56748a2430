// Host-side unit test of csrc/regfft.cuh (the same templates run in registers on the GPU).
//   nvcc -std=c++17 --expt-relaxed-constexpr -I gr-gfdm_b200/csrc tools/regfft_host_test.cu -o /tmp/regfft_test && /tmp/regfft_test
#include "regfft.cuh"
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
using namespace gfdm;

template <int N, int DIR>
double check()
{
    cpx v[N];
    std::complex<double> x[N];
    for (int i = 0; i < N; ++i) {
        v[i] = cmake((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
        x[i] = std::complex<double>(v[i].x, v[i].y);
    }
    rf::FFTN<N, DIR>::run(v);
    double err = 0, nrm = 0;
    for (int k = 0; k < N; ++k) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n) acc += x[n] * std::polar(1.0, DIR * 2.0 * M_PI * (double)((long)n * k % N) / N);
        err += std::norm(acc - std::complex<double>(v[k].x, v[k].y));
        nrm += std::norm(acc);
    }
    return std::sqrt(err / nrm);
}
#define T(N)                                                                                  \
    {                                                                                         \
        double a = check<N, -1>(), b = check<N, 1>();                                         \
        printf("N=%3d fwd %.2e inv %.2e %s\n", N, a, b, (a < 4e-7 && b < 4e-7) ? "ok" : "FAIL"); \
        if (!(a < 4e-7 && b < 4e-7)) bad++;                                                   \
    }
int main()
{
    int bad = 0;
    T(1) T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(9) T(10) T(11) T(12) T(13) T(15) T(16) T(18) T(20) T(21) T(25) T(27) T(32) T(45) T(64)
    return bad;
}
