// common.cuh -- shared device/host helpers of the B200 GFDM engine.
#pragma once
#include <cuda_runtime.h>

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

inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

} // namespace gfdm
