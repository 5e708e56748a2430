// fft_engine.cu -- any-length batched DFT for shapes without a fused kernel.
//
// Replaces gfdm_kernel_utils::initialize_fft + fftwf_execute
// (lib/gfdm_kernel_utils.cc:32-57): unnormalised, out-of-place c2c fp32.
// One Stockham autosort pass per radix; one thread per output element, so any
// prime factor works (O(p) MACs per output).  Twiddles come from a table
// generated in double precision (no __sincosf) to hold rel-L2 <= 1e-5.
#include "engine.h"

#include <cmath>

namespace gfdm {

void FftPlan::init(int n_)
{
    n = n_;
    radices.clear();
    int r = n;
    while (r % 4 == 0) { radices.push_back(4); r /= 4; }
    while (r % 2 == 0) { radices.push_back(2); r /= 2; }
    for (int p = 3; r > 1; p += 2) {
        while (r % p == 0) { radices.push_back(p); r /= p; }
        if ((long)p * p > r && r > 1) { radices.push_back(r); r = 1; }
    }
    if (radices.empty()) radices.push_back(1);
    std::vector<cpx> tw(n);
    for (int j = 0; j < n; ++j) {
        const double ph = -2.0 * M_PI * (double)j / (double)n;
        tw[j] = make_float2((float)std::cos(ph), (float)std::sin(ph));
    }
    d_tw = upload(tw);
}

void FftPlan::destroy()
{
    if (d_tw) cudaFree(d_tw);
    d_tw = nullptr;
}

// One Stockham pass of radix p; Ns = product of the radices already applied.
template <int P, bool INV>
__global__ void __launch_bounds__(256) stockham_pass(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                     const cpx* __restrict__ tw, int n, int p_rt, int Ns,
                                                     size_t total, float scale)
{
    const int p = P > 0 ? P : p_rt;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const size_t b = gid / n;
    const int o = (int)(gid - b * n);
    const int span = Ns * p;
    const int q = o / span;
    const int rem = o - q * span;
    const int t = rem / Ns;
    const int k = rem - t * Ns;
    const int j = q * Ns + k;
    const int stride = n / p;
    const int e = (k + t * Ns) * (n / span); // < n
    const cpx* x = in + b * n + j;
    cpx acc = x[0];
    int idx = 0;
#pragma unroll
    for (int r = 1; r < p; ++r) {
        idx += e;
        if (idx >= n) idx -= n;
        cpx w = tw[idx];
        if (INV) w.y = -w.y;
        acc = cfma(x[(size_t)r * stride], w, acc);
    }
    out[gid] = cscale(acc, scale);
}

template <bool INV>
static void launch_pass(int p, cpx* out, const cpx* in, const cpx* tw, int n, int Ns, size_t total, float scale,
                        cudaStream_t s)
{
    const unsigned th = 256, bl = blocks_for(total, th);
    switch (p) {
    case 2: stockham_pass<2, INV><<<bl, th, 0, s>>>(out, in, tw, n, p, Ns, total, scale); break;
    case 3: stockham_pass<3, INV><<<bl, th, 0, s>>>(out, in, tw, n, p, Ns, total, scale); break;
    case 4: stockham_pass<4, INV><<<bl, th, 0, s>>>(out, in, tw, n, p, Ns, total, scale); break;
    case 5: stockham_pass<5, INV><<<bl, th, 0, s>>>(out, in, tw, n, p, Ns, total, scale); break;
    default: stockham_pass<0, INV><<<bl, th, 0, s>>>(out, in, tw, n, p, Ns, total, scale); break;
    }
}

int fft_exec(const FftPlan& pl, cpx* out, const cpx* in, cpx* scratch, size_t batch, bool inverse, float scale,
             cudaStream_t s)
{
    if (batch == 0) return 0;
    const int P = (int)pl.radices.size();
    const size_t total = batch * (size_t)pl.n;
    int Ns = 1;
    const cpx* src = in;
    for (int i = 0; i < P; ++i) {
        // alternate so that the last pass lands in `out`
        cpx* dst = ((P - 1 - i) % 2 == 0) ? out : scratch;
        const float sc = (i == P - 1) ? scale : 1.0f;
        if (inverse)
            launch_pass<true>(pl.radices[i], dst, src, pl.d_tw, pl.n, Ns, total, sc, s);
        else
            launch_pass<false>(pl.radices[i], dst, src, pl.d_tw, pl.n, Ns, total, sc, s);
        Ns *= pl.radices[i];
        src = dst;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return P;
}

} // namespace gfdm
