// regfft.cuh -- compile-time-unrolled DFTs that live entirely in registers.
//
// FFTN<N, DIR>::run(v) transforms the local array v[0..N) in place (natural order
// in and out, unnormalised, DIR = -1: exp(-j..) forward, +1: backward).  Every
// index and every twiddle factor is a compile-time constant:
//   * N = 2^a            radix-2 decimation in time, trivial twiddles folded away
//   * N = P*Q, gcd = 1   Good-Thomas prime-factor mapping (no twiddles; 15 = 3 x 5)
//   * N = P*Q otherwise  Cooley-Tukey with twiddles (9 = 3 x 3, 25 = 5 x 5)
//   * N prime            symmetric-pair DFT (3, 5, 7, 11, 13, ...)
// Twiddles are evaluated in double precision by constexpr series and rounded to
// float once, so accuracy matches a table generated on the host in double.
// The functions are __host__ __device__ so the identical code is unit-tested on
// the CPU (tools/regfft_host_test.cu).
#pragma once
#include "common.cuh"

#include <type_traits>
#include <utility>

namespace gfdm {
namespace rf {

#define RF_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------- constexpr trig
constexpr double kPi = 3.141592653589793238462643383279502884;

constexpr double sin_small(double x) // |x| <= pi/4
{
    const double x2 = x * x;
    double term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / ((2.0 * i) * (2.0 * i + 1.0));
        sum += term;
    }
    return sum;
}
constexpr double cos_small(double x) // |x| <= pi/4
{
    const double x2 = x * x;
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / ((2.0 * i - 1.0) * (2.0 * i));
        sum += term;
    }
    return sum;
}
// cos / sin of 2*pi*k/n for integers, octant-reduced
constexpr double cos2pi(long k, long n)
{
    k %= n;
    if (k < 0) k += n;
    if (2 * k > n) k = n - k; // even symmetry -> angle in [0, pi]
    const double x = 2.0 * kPi * (double)k / (double)n;
    if (8 * k <= n) return cos_small(x);
    if (8 * k <= 3 * n) return -sin_small(x - kPi / 2);
    return -cos_small(kPi - x);
}
constexpr double sin2pi(long k, long n)
{
    k %= n;
    if (k < 0) k += n;
    double sgn = 1.0;
    if (2 * k > n) {
        k = n - k;
        sgn = -1.0;
    }
    const double x = 2.0 * kPi * (double)k / (double)n;
    if (8 * k <= n) return sgn * sin_small(x);
    if (8 * k <= 3 * n) return sgn * cos_small(x - kPi / 2);
    return sgn * sin_small(kPi - x);
}

constexpr int smallest_factor(int n)
{
    for (int p = 2; p * p <= n; ++p)
        if (n % p == 0) return p;
    return n;
}
constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
constexpr bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
constexpr int mod_inverse(int a, int m) // a^-1 mod m (m small)
{
    a %= m;
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 1;
}

// ---------------------------------------------------------------- compile-time loop
template <int I, int E>
struct StaticFor {
    template <class F>
    static RF_HD void run(F&& f)
    {
        f(std::integral_constant<int, I>{});
        StaticFor<I + 1, E>::run(f);
    }
};
template <int E>
struct StaticFor<E, E> {
    template <class F>
    static RF_HD void run(F&&) {}
};

// ---------------------------------------------------------------- twiddles
// a * exp(DIR * 2 pi i K / N)
template <int N, int K, int DIR>
RF_HD cpx mul_tw(cpx a)
{
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (2 * k == N) {
        return cmake(-a.x, -a.y);
    } else if constexpr (4 * k == N) { // exp(DIR * i pi/2) = DIR * i
        return DIR < 0 ? cmake(a.y, -a.x) : cmake(-a.y, a.x);
    } else if constexpr (4 * k == 3 * N) { // exp(DIR * i 3pi/2) = -DIR * i
        return DIR < 0 ? cmake(-a.y, a.x) : cmake(a.y, -a.x);
    } else if constexpr ((8 * k) % N == 0) { // odd multiples of pi/4: (+-1 +- i)/sqrt(2)
        constexpr float c = 0.70710678118654752440f;
        constexpr float wr = (8 * k == N || 8 * k == 7 * N) ? c : -c;
        constexpr float wi0 = (8 * k == N || 8 * k == 3 * N) ? c : -c; // sin(2 pi k / N)
        constexpr float wi = DIR < 0 ? -wi0 : wi0;
        return cmake(wr * a.x - wi * a.y, wr * a.y + wi * a.x);
    } else {
        constexpr float wr = (float)cos2pi(k, N);
        constexpr float wi = (float)(DIR * sin2pi(k, N));
        return cmake(wr * a.x - wi * a.y, wr * a.y + wi * a.x);
    }
}

template <int N, int DIR>
struct FFTN;

// ---------------------------------------------------------------- prime sizes
template <int P, int DIR>
struct PrimeDFT {
    static RF_HD void run(cpx (&v)[P])
    {
        if constexpr (P == 1) {
        } else if constexpr (P == 2) {
            const cpx a = v[0], b = v[1];
            v[0] = cadd(a, b);
            v[1] = csub(a, b);
        } else {
            constexpr int H = (P - 1) / 2;
            cpx s[H], d[H];
            StaticFor<0, H>::run([&](auto J) {
                constexpr int j = J.value;
                s[j] = cadd(v[j + 1], v[P - 1 - j]);
                d[j] = csub(v[j + 1], v[P - 1 - j]);
            });
            const cpx x0 = v[0];
            cpx sum = x0;
            StaticFor<0, H>::run([&](auto J) { sum = cadd(sum, s[J.value]); });
            v[0] = sum;
            StaticFor<1, H + 1>::run([&](auto KK) {
                constexpr int k = KK.value;
                cpx a = x0, b = cmake(0.f, 0.f);
                StaticFor<0, H>::run([&](auto J) {
                    constexpr int j = J.value + 1;
                    constexpr float c = (float)cos2pi((long)j * k, P);
                    constexpr float sn = (float)sin2pi((long)j * k, P);
                    a.x = fmaf(c, s[j - 1].x, a.x);
                    a.y = fmaf(c, s[j - 1].y, a.y);
                    b.x = fmaf(sn, d[j - 1].x, b.x);
                    b.y = fmaf(sn, d[j - 1].y, b.y);
                });
                // forward: X[k] = a - i b, X[P-k] = a + i b ; backward: swapped
                const cpx ib = cmake(-b.y, b.x);
                if (DIR < 0) {
                    v[k] = csub(a, ib);
                    v[P - k] = cadd(a, ib);
                } else {
                    v[k] = cadd(a, ib);
                    v[P - k] = csub(a, ib);
                }
            });
        }
    }
};

// ---------------------------------------------------------------- composite sizes
template <int N, int DIR>
struct FFTN {
    static constexpr int P = smallest_factor(N);
    static constexpr int Q = N / P;

    static RF_HD void run(cpx (&v)[N])
    {
        if constexpr (Q == 1) {
            PrimeDFT<N, DIR>::run(v);
        } else if constexpr (cgcd(P, Q) == 1) {
            // Good-Thomas: n = (Q n1 + P n2) mod N, k = (k1 Q Qi + k2 P Pi) mod N
            constexpr int Qi = mod_inverse(Q, P), Pi = mod_inverse(P, Q);
            cpx S[P][Q];
            StaticFor<0, P>::run([&](auto N1) {
                constexpr int n1 = N1.value;
                cpx sub[Q];
                StaticFor<0, Q>::run([&](auto N2) { sub[N2.value] = v[(Q * n1 + P * N2.value) % N]; });
                FFTN<Q, DIR>::run(sub);
                StaticFor<0, Q>::run([&](auto K2) { S[n1][K2.value] = sub[K2.value]; });
            });
            StaticFor<0, Q>::run([&](auto K2) {
                constexpr int k2 = K2.value;
                cpx col[P];
                StaticFor<0, P>::run([&](auto N1) { col[N1.value] = S[N1.value][k2]; });
                PrimeDFT<P, DIR>::run(col);
                StaticFor<0, P>::run([&](auto K1) {
                    v[(K1.value * Q * Qi + k2 * P * Pi) % N] = col[K1.value];
                });
            });
        } else {
            // Cooley-Tukey DIT: n = P n2 + n1, k = k2 + Q k1
            cpx S[P][Q];
            StaticFor<0, P>::run([&](auto N1) {
                constexpr int n1 = N1.value;
                cpx sub[Q];
                StaticFor<0, Q>::run([&](auto N2) { sub[N2.value] = v[P * N2.value + n1]; });
                FFTN<Q, DIR>::run(sub);
                StaticFor<0, Q>::run([&](auto K2) {
                    S[n1][K2.value] = mul_tw<N, n1 * K2.value, DIR>(sub[K2.value]);
                });
            });
            StaticFor<0, Q>::run([&](auto K2) {
                constexpr int k2 = K2.value;
                cpx col[P];
                StaticFor<0, P>::run([&](auto N1) { col[N1.value] = S[N1.value][k2]; });
                PrimeDFT<P, DIR>::run(col);
                StaticFor<0, P>::run([&](auto K1) { v[k2 + Q * K1.value] = col[K1.value]; });
            });
        }
    }
};

template <int DIR>
struct FFTN<1, DIR> {
    static RF_HD void run(cpx (&)[1]) {}
};

#undef RF_HD

} // namespace rf
} // namespace gfdm
