// common.cuh -- shared device/host helpers of the B200 GFDM engine.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace gfdm {

typedef float2 cpx; // layout == std::complex<float> == gfdm_complex

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define GFDM_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw ::gfdm::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                    __FILE__ + ":" + std::to_string(__LINE__) + ")");           \
    } while (0)

__host__ __device__ __forceinline__ cpx cmake(float a, float b) { return make_float2(a, b); }
__host__ __device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cpx cmul(cpx a, cpx b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cpx cmulc(cpx a, cpx b)
{
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cpx cscale(cpx a, float s) { return make_float2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
// acc + a*b
__host__ __device__ __forceinline__ cpx cfma(cpx a, cpx b, cpx acc)
{
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y)));
}
// a / b the way volk_32fc_x2_divide_32fc does it: a * conj(b) / |b|^2
__device__ __forceinline__ cpx cdiv(cpx a, cpx b)
{
    const float den = b.x * b.x + b.y * b.y;
    const cpx num = cmulc(a, b);
    return make_float2(__fdiv_rn(num.x, den), __fdiv_rn(num.y, den));
}
// unfused complex multiply in the reference's operand order (bit-exact window path)
__device__ __forceinline__ cpx cmul_rn(cpx a, cpx b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// Hard decision of gr::digital::constellation::decision_maker (call site
// lib/advanced_receiver_kernel_cc.cc:114-120): rule 1 = constellation_qpsk's sign rule, otherwise
// nearest point, first minimum wins (unfused arithmetic so that ties break like the CPU oracle).
__device__ __forceinline__ int decide_symbol(cpx s, const cpx* __restrict__ points, int n_points, int rule)
{
    if (rule == 1) return 2 * (s.y > 0.f) + (s.x > 0.f);
    int best = 0;
    float dmin = 0.f;
    for (int i = 0; i < n_points; ++i) {
        const float dr = __fsub_rn(s.x, points[i].x), di = __fsub_rn(s.y, points[i].y);
        const float d = __fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di));
        if (i == 0 || d < dmin) {
            dmin = d;
            best = i;
        }
    }
    return best;
}

// out-of-line copy for the rare fallback of decide_symbol_grid: keeps the unrolled decision sites of the fused
// kernels small (30 inlined searches would more than double their code size)
static __device__ __noinline__ int decide_symbol_slow(float sx, float sy, const cpx* points, int n_points)
{
    return decide_symbol(cmake(sx, sy), points, n_points, 0);
}

// Uniform rectangular grid view of a constellation (square/rectangular QAM, PSK-2/4 on the axes' lattice),
// detected on the host (make_decide_grid): point lut[i*n_im + q] = (re0 + i*d_re) + j(im0 + q*d_im).  n_re == 0: no
// such structure, every decision is the exhaustive search.
struct DecideGrid {
    float re0 = 0.f, inv_dre = 0.f, im0 = 0.f, inv_dim = 0.f;
    int n_re = 0, n_im = 0;
    int identity = 0; // lut[c] == c for every cell: the point index IS the cell index (no table load)
    unsigned char lut[64] = { 0 };
};
// Nearest-point decision in O(1) on a grid constellation: quantise each axis, and fall back to the exhaustive
// search (decide_symbol, the oracle's arithmetic) whenever the symbol lies within 1e-4 level spacings of a
// decision boundary or more than 8 spacings away from the grid origin on either axis.  Inside that window the two
// agree exactly -- the squared distances of the two nearest points differ by >= 2e-4 d^2 while their fp32 rounding
// error is <= 1.5 ulp(128 d^2) ~ 1e-5 d^2 -- so the result is bit-identical to decide_symbol everywhere, ties
// ("first minimum wins") and far-out symbols whose distances round to the same float included.
__device__ __forceinline__ int decide_symbol_grid(cpx s, const cpx* __restrict__ points, int n_points, int rule,
                                                  const DecideGrid& g, const unsigned char* __restrict__ lut)
{
    if (rule == 1) return 2 * (s.y > 0.f) + (s.x > 0.f);
    if (g.n_re > 0) {
        // round to nearest by the 1.5*2^23 trick: FADD/IADD only (FRND and F2I run on the quarter-rate conversion pipe);
        // exact for |t| < 2^22, and the window test below rejects everything else (NaN and infinities included)
        const float magic = 12582912.f;
        const float tr = (s.x - g.re0) * g.inv_dre, ti = (s.y - g.im0) * g.inv_dim;
        const float ur = tr + magic, ui = ti + magic;
        const float rr = ur - magic, ri = ui - magic;
        const float lim = 0.5f - 1e-4f;
        if (fabsf(tr - rr) < lim && fabsf(ti - ri) < lim && fabsf(tr) < 8.f && fabsf(ti) < 8.f) { // false for NaN
            const int i = min(max(__float_as_int(ur) - 0x4B400000, 0), g.n_re - 1);
            const int q = min(max(__float_as_int(ui) - 0x4B400000, 0), g.n_im - 1);
            const int cell = i * g.n_im + q;
            return g.identity ? cell : (int)lut[cell];
        }
    }
    return decide_symbol_slow(s.x, s.y, points, n_points);
}

// Nearest grid point for the interference-cancellation loops: clamp, round to nearest (even), no boundary window and no
// search fallback -- ~12 instructions per symbol.  Differs from decide_symbol only for a soft symbol that lies EXACTLY on a
// decision boundary (first-minimum-wins vs round-half-even) or is NaN (-> cell 0); the loop's output is the soft symbol
// vector, which fp32 rounding moves across such a boundary anyway ("Bit-exactness").
__device__ __forceinline__ int decide_symbol_grid_fast(cpx s, const DecideGrid& g, const unsigned char* __restrict__ lut)
{
    const float magic = 12582912.f;
    const float tr = fminf(fmaxf((s.x - g.re0) * g.inv_dre, 0.f), (float)(g.n_re - 1));
    const float ti = fminf(fmaxf((s.y - g.im0) * g.inv_dim, 0.f), (float)(g.n_im - 1));
    const int i = __float_as_int(tr + magic) - 0x4B400000, q = __float_as_int(ti + magic) - 0x4B400000;
    const int cell = i * g.n_im + q;
    return g.identity ? cell : (int)lut[cell];
}

// M decisions of one thread at once (the epilogue of the fused receiver): the quantiser runs branch-free over all M
// symbols and collects a bit mask of the symbols that need the exhaustive search, so the common case is straight-line
// code and the rare one a single divergent region.  Same results as M calls of decide_symbol_grid.
template <int M, bool IDENT>
__device__ __forceinline__ unsigned decide_block_grid(const cpx (&v)[M], int (&dec)[M], const DecideGrid& g,
                                                      const unsigned char* __restrict__ lut)
{
    const float magic = 12582912.f, lim = 0.5f - 1e-4f;
    const float re0 = g.re0, im0 = g.im0, ir = g.inv_dre, ii = g.inv_dim;
    const int nre1 = g.n_re - 1, nim1 = g.n_im - 1, nim = g.n_im;
    unsigned bad = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float tr = (v[m].x - re0) * ir, ti = (v[m].y - im0) * ii;
        const float ur = tr + magic, ui = ti + magic;
        const float er = fabsf(tr - (ur - magic)), ei = fabsf(ti - (ui - magic));
        const bool ok = fmaxf(er, ei) < lim && fmaxf(fabsf(tr), fabsf(ti)) < 8.f; // false for NaN (fmaxf drops one NaN:
        const bool nan = !(tr == tr) || !(ti == ti);                              //  test it explicitly)
        bad |= (ok && !nan) ? 0u : (1u << m);
        const int i = min(max(__float_as_int(ur) - 0x4B400000, 0), nre1);
        const int q = min(max(__float_as_int(ui) - 0x4B400000, 0), nim1);
        const int cell = i * nim + q;
        dec[m] = IDENT ? cell : (int)lut[cell];
    }
    return bad;
}
template <int M>
__device__ __forceinline__ void decide_block(const cpx (&v)[M], unsigned char* __restrict__ dst, const cpx* __restrict__ points,
                                             int n_points, int rule, const DecideGrid& g,
                                             const unsigned char* __restrict__ lut)
{
    static_assert(M <= 32, "one mask bit per symbol");
    int dec[M];
    if (rule == 1) {
#pragma unroll
        for (int m = 0; m < M; ++m) dec[m] = 2 * (v[m].y > 0.f) + (v[m].x > 0.f);
    } else {
        unsigned bad = (M == 32) ? 0xffffffffu : ((1u << M) - 1u); // no grid: every symbol takes the search
        if (g.n_re > 0) bad = g.identity ? decide_block_grid<M, true>(v, dec, g, lut) : decide_block_grid<M, false>(v, dec, g, lut);
        if (bad) {
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (bad & (1u << m)) dec[m] = decide_symbol_slow(v[m].x, v[m].y, points, n_points);
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) dst[m] = (unsigned char)dec[m];
}

// growable device buffer owned by a handle
struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void ensure(size_t n)
    {
        if (n <= bytes) return;
        release();
        GFDM_CUDA_CHECK(cudaMalloc(&p, n));
        bytes = n;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

template <class T>
inline T* upload(const std::vector<T>& v)
{
    T* d = nullptr;
    GFDM_CUDA_CHECK(cudaMalloc(&d, sizeof(T) * (v.empty() ? 1 : v.size())));
    if (!v.empty()) GFDM_CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return d;
}

// host: find the grid structure of a constellation (exact float levels, uniform spacing, every combination once)
inline DecideGrid make_decide_grid(const std::vector<cpx>& pts)
{
    DecideGrid g;
    const size_t n = pts.size();
    if (n < 1 || n > 64) return g;
    std::vector<float> re, im;
    for (const cpx& p : pts) {
        bool fr = false, fi = false;
        for (float v : re) fr = fr || v == p.x;
        for (float v : im) fi = fi || v == p.y;
        if (!fr) re.push_back(p.x);
        if (!fi) im.push_back(p.y);
    }
    if (re.size() * im.size() != n) return g;
    auto sort_check = [](std::vector<float>& v, float& v0, float& inv_d) {
        for (size_t a = 0; a < v.size(); ++a)
            for (size_t b = a + 1; b < v.size(); ++b)
                if (v[b] < v[a]) { const float t = v[a]; v[a] = v[b]; v[b] = t; }
        v0 = v[0];
        inv_d = 0.f;
        if (v.size() == 1) return true;
        const double d = ((double)v.back() - (double)v[0]) / (double)(v.size() - 1);
        if (!(d > 0.0)) return false;
        for (size_t a = 0; a < v.size(); ++a)
            if (std::fabs(((double)v[a] - ((double)v[0] + d * (double)a)) / d) > 1e-5) return false; // uniform spacing
        inv_d = (float)(1.0 / d);
        return true;
    };
    if (!sort_check(re, g.re0, g.inv_dre) || !sort_check(im, g.im0, g.inv_dim)) return g;
    // an axis with a single level is measured in the other axis' spacing (the distance window needs a scale)
    if (g.inv_dre == 0.f) g.inv_dre = g.inv_dim;
    if (g.inv_dim == 0.f) g.inv_dim = g.inv_dre;
    if (g.inv_dre == 0.f) return g; // one point: nothing to decide
    std::vector<int> seen(n, 0);
    for (size_t idx = 0; idx < n; ++idx) {
        size_t i = 0, q = 0;
        while (re[i] != pts[idx].x) ++i;
        while (im[q] != pts[idx].y) ++q;
        if (seen[i * im.size() + q]++) return g; // duplicate point
        g.lut[i * im.size() + q] = (unsigned char)idx;
    }
    g.n_re = (int)re.size();
    g.n_im = (int)im.size();
    g.identity = 1;
    for (size_t c = 0; c < n; ++c)
        if (g.lut[c] != c) g.identity = 0;
    return g;
}

inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

} // namespace gfdm
