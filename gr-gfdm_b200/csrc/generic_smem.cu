// generic_smem.cu -- modulator and receiver for ANY (M, K, L) whose frame fits shared memory twice (N <= 12288 complex),
// i.e. every shape outside the table of fused kernels (K not a power of two, M = 127, M = 25, ...).
//
// One CTA owns a frame: it is read from HBM once, every stage of the reference's algorithm runs between two shared-memory
// buffers, and the result is written once -- 16 N bytes of HBM traffic per frame (24 N with a per-frame channel) instead
// of one HBM round trip per stage and per radix of the staged path (api.cu: fft_engine + stage kernels, ~10 passes).
// The transforms are Stockham autosort passes between the two buffers: one thread per radix-P butterfly held in
// registers for P <= 16 (composite radices from regfft.cuh, so N = 2400 is three passes: 15 x 10 x 16), one thread per
// output element (O(p) multiply-adds) for prime factors above 16; twiddles from the tables of fft_engine.cu's plans.
//
//   modulator (lib/modulator_kernel_cc.cc:98-141):  D_k = FFT_M(d_k) -> X = scatter-add of taps * D -> x = IFFT_N(X) / N
//   receiver  (lib/receiver_kernel_cc.cc:165-225,301-334):  Y = FFT_N(x) [ / H ] -> R_k = sum_i taps * Y -> y_k = IFFT_M(R_k) / M
#include "engine.h"
#include "regfft.cuh"

#include <algorithm>
#include <vector>

namespace gfdm {

// threads per CTA: a template parameter GT (256 x 3 CTAs per SM for short frames, 384 x 2 and 768 x 1 for longer ones)
static constexpr int MAX_RAD = 16;   // N <= 12288 < 2^14: at most 13 prime factors
static constexpr int MAX_BFLY = 16;  // largest radix with a register butterfly

struct PassArgs {
    int p, Ns, nb, step;    // radix, product of the radices before it, butterflies per transform n/p, n/(Ns p)
    float inv_nb, inv_ns;
};

struct GenArgs {
    int M, K, L, N;
    int n_rad_m, n_rad_n;
    int tw_smem; // > 0: elements behind the two frame buffers for the compact twiddle table of a one-output-per-thread pass
    PassArgs pm[MAX_RAD], pn[MAX_RAD];
    const cpx* tw_m; // W_M^j
    const cpx* tw_n; // W_N^j
    const cpx* taps; // L*M, normalised
};

// floor(a / d) for 0 <= a < 2^20, 1 <= d < 2^20 without an integer division: (a + 0.5) / d is never closer than
// 0.5 / d >= 4.7e-7 to an integer, far above the rounding of one fp32 multiply at these magnitudes
__device__ __forceinline__ int fdiv(int a, float inv_d) { return __float2int_rd(((float)a + 0.5f) * inv_d); }

__device__ __forceinline__ void cp_async8(cpx* smem_dst, const cpx* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// one Stockham pass of radix p over `batch` contiguous transforms of length n with one thread per OUTPUT element (O(p)
// multiply-adds per output: any radix works; used for prime factors above 16 only -- fft_engine.cu: stockham_pass)
template <bool INV, int GT>
__device__ __forceinline__ void smem_pass_any(cpx* __restrict__ out, const cpx* __restrict__ in, const cpx* __restrict__ tw,
                                              cpx* __restrict__ tw_sm, int n, const PassArgs& pa, int total, int tid)
{
    const int p = pa.p, Ns = pa.Ns, span = Ns * p, stride = pa.nb, step = pa.step;
    const float inv_n = 1.0f / (float)n, inv_span = 1.0f / (float)span;
    const bool one = total == n;
    // p twiddles per output: W_span^{r rem} from a COMPACT copy of W_span in shared memory (consecutive lanes read with
    // stride r; an L1 gather out of the plan's table would cost a sector per lane and lookup)
    if (tw_sm) {
        for (int j = tid; j < span; j += GT) tw_sm[j] = __ldg(reinterpret_cast<const float2*>(tw) + j * step);
        __syncthreads();
    }
    for (int gid = tid; gid < total; gid += GT) {
        const int b = one ? 0 : fdiv(gid, inv_n), o = gid - b * n;
        const int q = fdiv(o, inv_span), rem = o - q * span;
        const int k = Ns == 1 ? 0 : rem - fdiv(rem, pa.inv_ns) * Ns;
        const cpx* x = in + b * n + q * Ns + k;
        cpx acc = x[0];
        int idx = 0;
#pragma unroll 4
        for (int r = 1; r < p; ++r) {
            idx += rem;
            if (idx >= span) idx -= span;
            cpx w = tw_sm ? tw_sm[idx] : __ldg(reinterpret_cast<const float2*>(tw) + idx * step);
            if (INV) w.y = -w.y;
            acc = cfma(x[r * stride], w, acc);
        }
        out[gid] = acc;
    }
}

// one Stockham pass with one thread per radix-P BUTTERFLY: P strided reads (consecutive lanes on consecutive elements),
// the twiddles W_span^{r k} = W_n^{r k n/span} from one lookup in the plan's table, a register-resident DFT_P (regfft.cuh), P writes
//   out[q span + t Ns + k] = sum_r in[q Ns + k + r n/P] W_span^{r k} W_P^{r t},  span = Ns P, butterfly j = q Ns + k
template <bool INV, int P, int GT>
__device__ __forceinline__ void smem_bfly(cpx* __restrict__ out, const cpx* __restrict__ in, const cpx* __restrict__ tw, int n,
                                          const PassArgs& pa, int batch, int tid)
{
    const int nb = pa.nb, Ns = pa.Ns, total = batch * nb;
    for (int g = tid; g < total; g += GT) {
        const int b = batch == 1 ? 0 : fdiv(g, pa.inv_nb), j = g - b * nb;
        const int q = Ns == 1 ? j : fdiv(j, pa.inv_ns), k = j - q * Ns;
        const cpx* x = in + b * n + j;
        cpx v[P];
#pragma unroll
        for (int r = 0; r < P; ++r) v[r] = x[r * nb];
        if (Ns > 1) {
            // ONE table lookup (a gather: every lane its own sector), the other powers by products of depth <= log2(P)
            // (w^r = w^{r/2} w^{r - r/2}: at most 4 roundings deep, far inside the 1e-5 budget); only w^1 .. w^{P/2} stay live
            cpx w[P / 2 + 1];
            w[1] = __ldg(reinterpret_cast<const float2*>(tw) + k * pa.step); // r k step < P Ns step = n
            if (INV) w[1].y = -w[1].y;
#pragma unroll
            for (int r = 2; r < P; ++r) {
                const cpx wr = cmul(w[r / 2], w[r - r / 2]);
                if (r <= P / 2) w[r] = wr;
                v[r] = cmul(v[r], wr);
            }
            v[1] = cmul(v[1], w[1]);
        }
        rf::FFTN<P, INV ? +1 : -1>::run(v);
        cpx* o = out + b * n + q * (Ns * P) + k;
#pragma unroll
        for (int t = 0; t < P; ++t) o[t * Ns] = v[t];
    }
}
template <bool INV, int GT>
__device__ __forceinline__ void smem_pass(cpx* out, const cpx* in, const cpx* tw, cpx* tw_sm, int n, const PassArgs& pa, int batch, int tid)
{
    switch (pa.p) {
#define GFDM_BFLY(P_) case P_: smem_bfly<INV, P_, GT>(out, in, tw, n, pa, batch, tid); break;
    GFDM_BFLY(2) GFDM_BFLY(3) GFDM_BFLY(4) GFDM_BFLY(5) GFDM_BFLY(6) GFDM_BFLY(7) GFDM_BFLY(8) GFDM_BFLY(9)
    GFDM_BFLY(10) GFDM_BFLY(11) GFDM_BFLY(12) GFDM_BFLY(13) GFDM_BFLY(14) GFDM_BFLY(15) GFDM_BFLY(16)
#undef GFDM_BFLY
    default: smem_pass_any<INV, GT>(out, in, tw, tw_sm, n, pa, batch * n, tid); break;
    }
}
// whole transform between the two buffers; returns the buffer that holds the result
template <bool INV, int GT>
__device__ __forceinline__ cpx* smem_fft(cpx* src, cpx* dst, const cpx* tw, cpx* tw_sm, int n, const PassArgs* pa, int n_rad, int batch,
                                         int tid)
{
    for (int i = 0; i < n_rad; ++i) {
        smem_pass<INV, GT>(dst, src, tw, tw_sm, n, pa[i], batch, tid);
        __syncthreads();
        cpx* t = src;
        src = dst;
        dst = t;
    }
    return src;
}

// frame -> shared memory without a register round trip (all copies of a thread in flight at once)
template <int GT>
__device__ __forceinline__ void load_frame(cpx* dst, const cpx* src, int N, int tid)
{
    for (int i = tid; i < N; i += GT) cp_async8(dst + i, src + i);
    cp_async_wait_all();
    __syncthreads();
}

template <int GT>
__global__ void __launch_bounds__(GT, 768 / GT) generic_smem_mod_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, int n_frames,
                                                                        const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(16) unsigned char gsm[];
    cpx* A = reinterpret_cast<cpx*>(gsm);
    cpx* B = A + a.N;
    const int tid = threadIdx.x, M = a.M, K = a.K, L = a.L, N = a.N, h = L / 2;
    cpx* TW = a.tw_smem ? B + N : nullptr; // twiddles of a one-output-per-thread pass (plans with a large prime factor)
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M;
    const int tp_first = (L - 1 + h) % L; // i = L-1 first: subcarrier (b - i + h) mod K, tap block (i + h) mod L
    const float inv_n = 1.0f / (float)N, inv_mf = 1.0f / (float)M;
    const int LM = L * M;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        load_frame<GT>(A, in + (size_t)f * N, N, tid);
        cpx* D = smem_fft<false, GT>(A, B, a.tw_m, TW, M, a.pm, a.n_rad_m, K, tid); // D_k = FFT_M(d_k), [k][m]
        cpx* X = D == A ? B : A;
        for (int r = tid; r < N; r += GT) {
            const int b = fdiv(r, inv_mf), m = r - b * M;
            cpx acc = cmake(0.f, 0.f);
            if (m < part_len) {
                // the same accumulation order as the reference's k-loop produces for this bin (:113-135): i = L-1 first,
                // subcarrier (b - i + h) mod K, tap block (i + h) mod L; both walked as element offsets
                int k = b + h - (L - 1); // > -K: L <= K (generic_smem_supported)
                if (k < 0) k += K;
                int di = k * M + m, ti = tp_first * M + m;
                for (int i = L - 1; i >= 0; --i) {
                    acc = cadd(acc, cmul(D[di], __ldg(reinterpret_cast<const float2*>(a.taps) + ti)));
                    di += M;
                    if (di >= N) di -= N;
                    ti -= M;
                    if (ti < 0) ti += LM;
                }
            }
            X[r] = acc;
        }
        __syncthreads();
        const cpx* y = smem_fft<true, GT>(X, D, a.tw_n, TW, N, a.pn, a.n_rad_n, 1, tid);
        cpx* o = out + (size_t)f * N;
        for (int i = tid; i < N; i += GT) o[i] = cscale(y[i], inv_n);
        __syncthreads(); // the buffers are reused by the next frame
    }
}

// mode 0: soft symbols y (generic_work[_equalize]); mode 1: R (fft_[equalize_]filter_downsample)
template <int GT>
__global__ void __launch_bounds__(GT, 768 / GT) generic_smem_rx_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                                       const cpx* __restrict__ eq, int mode, int n_frames,
                                                                       const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(16) unsigned char gsm[];
    cpx* A = reinterpret_cast<cpx*>(gsm);
    cpx* B = A + a.N;
    const int tid = threadIdx.x, M = a.M, K = a.K, L = a.L, N = a.N, h = L / 2;
    cpx* TW = a.tw_smem ? B + N : nullptr; // twiddles of a one-output-per-thread pass (plans with a large prime factor)
    const float inv_m = 1.0f / (float)M;
    const int LM = L * M;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        load_frame<GT>(A, in + (size_t)f * N, N, tid);
        cpx* Y = smem_fft<false, GT>(A, B, a.tw_n, TW, N, a.pn, a.n_rad_n, 1, tid);
        if (eq != nullptr) { // volk_32fc_x2_divide_32fc (:315)
            const cpx* hq = eq + (size_t)f * N;
            for (int i = tid; i < N; i += GT) Y[i] = cdiv(Y[i], hq[i]);
            __syncthreads();
        }
        cpx* R = Y == A ? B : A;
        for (int r = tid; r < N; r += GT) { // filter_subcarriers_and_downsample_fd (:165-192)
            const int k = fdiv(r, inv_m), m = r - k * M;
            cpx acc = cmake(0.f, 0.f);
            int kk = k - h; // i = 0: subcarrier (k - h) mod K, tap block h mod L = h; h < K: L <= K
            if (kk < 0) kk += K;
            int yi = kk * M + m, ti = h * M + m;
            for (int i = 0; i < L; ++i) {
                acc = cadd(acc, cmul(__ldg(reinterpret_cast<const float2*>(a.taps) + ti), Y[yi]));
                yi += M;
                if (yi >= N) yi -= N;
                ti += M;
                if (ti >= LM) ti -= LM;
            }
            R[r] = acc;
        }
        __syncthreads();
        cpx* o = out + (size_t)f * N;
        if (mode == 1) {
            for (int i = tid; i < N; i += GT) o[i] = R[i];
        } else {
            const cpx* y = smem_fft<true, GT>(R, Y, a.tw_m, TW, M, a.pm, a.n_rad_m, K, tid); // transform_subcarriers_to_td (:211-225)
            for (int i = tid; i < N; i += GT) o[i] = cscale(y[i], inv_m);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Radix plan of the in-kernel transforms: the prime factors up to 13 packed into as few radices <= 16 as first-fit
// decreasing finds (2400 = 15 x 10 x 16, 96 = 12 x 8), odd radices first (the first pass writes with a stride of P
// elements: odd strides are free of bank conflicts); every prime factor above 16 is a pass of its own.
static std::vector<int> smem_radices(int n)
{
    std::vector<int> small, big, bins;
    int r = n;
    for (int p = 2; r > 1; ++p) {
        while (r % p == 0) { (p <= MAX_BFLY ? small : big).push_back(p); r /= p; }
        if ((long)p * p > r && r > 1) { (r <= MAX_BFLY ? small : big).push_back(r); r = 1; }
    }
    std::sort(small.begin(), small.end(), [](int x, int y) { return x > y; });
    for (int p : small) {
        bool placed = false;
        for (int& b : bins)
            if (b * p <= MAX_BFLY) { b *= p; placed = true; break; }
        if (!placed) bins.push_back(p);
    }
    std::stable_sort(bins.begin(), bins.end(), [](int x, int y) { return (x & 1) > (y & 1); });
    big.insert(big.end(), bins.begin(), bins.end());
    if (big.empty()) big.push_back(1);
    return big;
}

static bool bfly_radix(int r) { return r >= 2 && r <= MAX_BFLY; }

bool generic_smem_supported(int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n)
{
    const size_t N = (size_t)M * K;
    return N >= 1 && N <= 12288 && L >= 1 && L <= K && smem_radices(M).size() <= (size_t)MAX_RAD &&
           smem_radices((int)N).size() <= (size_t)MAX_RAD && fft_m.d_tw != nullptr && fft_n.d_tw != nullptr;
}

static int fill_passes(PassArgs* pa, const std::vector<int>& rad, int n, int& need)
{
    int Ns = 1;
    for (size_t i = 0; i < rad.size(); ++i) {
        const int p = rad[i];
        pa[i].p = p; pa[i].Ns = Ns; pa[i].nb = n / p; pa[i].step = n / (Ns * p);
        pa[i].inv_nb = 1.0f / (float)(n / p); pa[i].inv_ns = 1.0f / (float)Ns;
        Ns *= p;
        if (!bfly_radix(p) && p > 1 && Ns > need) need = Ns; // span of a one-output-per-thread pass
    }
    return (int)rad.size();
}

static GenArgs make_args(int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n, const cpx* d_taps)
{
    GenArgs a{};
    a.M = M; a.K = K; a.L = L; a.N = M * K;
    int need = 0;
    a.n_rad_m = fill_passes(a.pm, smem_radices(M), M, need);
    a.n_rad_n = fill_passes(a.pn, smem_radices(M * K), M * K, need);
    a.tw_m = fft_m.d_tw; a.tw_n = fft_n.d_tw; a.taps = d_taps;
    a.tw_smem = (need > 0 && sizeof(cpx) * (2 * (size_t)a.N + need) <= (size_t)200 * 1024) ? need : 0;
    return a;
}
static size_t smem_bytes(const GenArgs& a) { return sizeof(cpx) * (2 * (size_t)a.N + a.tw_smem); }

template <class Kern>
static int grid_for_frames(Kern kern, int threads, size_t smem, size_t frames)
{
    int dev = 0, sms = 0, per_sm = 0;
    GFDM_CUDA_CHECK(cudaGetDevice(&dev));
    GFDM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GFDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GFDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) throw CudaError("generic shared-memory kernel does not fit on this device");
    const size_t cap = (size_t)sms * per_sm;
    return (int)(frames < cap ? frames : cap);
}

// 768 threads per SM (80 registers each): three CTAs of 256 while three frame pairs fit, two of 384, else one of 768
static int cta_threads(size_t smem) { return smem <= (size_t)72 * 1024 ? 256 : (smem <= (size_t)110 * 1024 ? 384 : 768); }

int launch_generic_smem_mod(cpx* out, const cpx* in, int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n,
                            const cpx* d_taps, size_t frames, cudaStream_t s)
{
    const GenArgs a = make_args(M, K, L, fft_m, fft_n, d_taps);
    const size_t smem = smem_bytes(a), max_chunk = (size_t)1 << 20;
    const int gt = cta_threads(smem);
    int launches = 0;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const size_t nf = frames - f0 < max_chunk ? frames - f0 : max_chunk;
        cpx* o = out + f0 * a.N;
        const cpx* i = in + f0 * a.N;
        if (gt == 256) generic_smem_mod_kernel<256><<<grid_for_frames(generic_smem_mod_kernel<256>, 256, smem, nf), 256, smem, s>>>(o, i, (int)nf, a);
        else if (gt == 384) generic_smem_mod_kernel<384><<<grid_for_frames(generic_smem_mod_kernel<384>, 384, smem, nf), 384, smem, s>>>(o, i, (int)nf, a);
        else generic_smem_mod_kernel<768><<<grid_for_frames(generic_smem_mod_kernel<768>, 768, smem, nf), 768, smem, s>>>(o, i, (int)nf, a);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

int launch_generic_smem_rx(cpx* out, const cpx* in, const cpx* eq, int mode, int M, int K, int L, const FftPlan& fft_m,
                           const FftPlan& fft_n, const cpx* d_taps, size_t frames, cudaStream_t s)
{
    const GenArgs a = make_args(M, K, L, fft_m, fft_n, d_taps);
    const size_t smem = smem_bytes(a), max_chunk = (size_t)1 << 20;
    const int gt = cta_threads(smem);
    int launches = 0;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const size_t nf = frames - f0 < max_chunk ? frames - f0 : max_chunk;
        cpx* o = out + f0 * a.N;
        const cpx* i = in + f0 * a.N;
        const cpx* e = eq ? eq + f0 * a.N : nullptr;
        if (gt == 256) generic_smem_rx_kernel<256><<<grid_for_frames(generic_smem_rx_kernel<256>, 256, smem, nf), 256, smem, s>>>(o, i, e, mode, (int)nf, a);
        else if (gt == 384) generic_smem_rx_kernel<384><<<grid_for_frames(generic_smem_rx_kernel<384>, 384, smem, nf), 384, smem, s>>>(o, i, e, mode, (int)nf, a);
        else generic_smem_rx_kernel<768><<<grid_for_frames(generic_smem_rx_kernel<768>, 768, smem, nf), 768, smem, s>>>(o, i, e, mode, (int)nf, a);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

} // namespace gfdm
