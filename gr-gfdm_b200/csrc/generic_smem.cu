// generic_smem.cu -- modulator and receiver for ANY (M, K, L) whose frame fits shared memory twice (N <= 12288 complex),
// i.e. every shape outside the table of fused kernels (K not a power of two, M = 127, M = 25, ...).
//
// One CTA owns a frame: it is read from HBM once, every stage of the reference's algorithm runs between two shared-memory
// buffers, and the result is written once -- 16 N bytes of HBM traffic per frame (24 N with a per-frame channel) instead
// of one HBM round trip per stage and per radix of the staged path (api.cu: fft_engine + stage kernels, ~10 passes).
// The transforms are Stockham autosort passes with one thread per output element (O(p) multiply-adds per output, so
// any prime factor works), the same arithmetic as fft_engine.cu's global-memory passes, twiddles from the same tables.
//
//   modulator (lib/modulator_kernel_cc.cc:98-141):  D_k = FFT_M(d_k) -> X = scatter-add of taps * D -> x = IFFT_N(X) / N
//   receiver  (lib/receiver_kernel_cc.cc:165-225,301-334):  Y = FFT_N(x) [ / H ] -> R_k = sum_i taps * Y -> y_k = IFFT_M(R_k) / M
#include "engine.h"

namespace gfdm {

static constexpr int GT = 256;       // threads per CTA
static constexpr int MAX_RAD = 16;   // N <= 12288 < 2^14: at most 13 prime factors

struct GenArgs {
    int M, K, L, N;
    int n_rad_m, n_rad_n;
    int rad_m[MAX_RAD], rad_n[MAX_RAD];
    const cpx* tw_m; // W_M^j
    const cpx* tw_n; // W_N^j
    const cpx* taps; // L*M, normalised
};

// floor(a / d) for 0 <= a < 2^20, 1 <= d < 2^20 without an integer division (~25 instructions): (a + 0.5) / d is never
// closer than 0.5 / d >= 4.7e-7 (relative: >= 2^-20 / quotient ... far above fp32 rounding of one multiply) to an integer
__device__ __forceinline__ int fdiv(int a, float inv_d) { return __float2int_rd(((float)a + 0.5f) * inv_d); }

// one Stockham pass of radix p over `batch` contiguous transforms of length n (fft_engine.cu: stockham_pass); P > 0: the
// radix as a compile-time constant (unrolled multiply-add chain), P = 0: any radix
template <bool INV, int P>
__device__ __forceinline__ void smem_pass(cpx* __restrict__ out, const cpx* __restrict__ in, const cpx* __restrict__ tw, int n,
                                          int p_rt, int Ns, int total, int tid)
{
    const int p = P > 0 ? P : p_rt;
    const int span = Ns * p, stride = n / p, step = n / span;
    const float inv_n = 1.0f / (float)n, inv_span = 1.0f / (float)span, inv_ns = 1.0f / (float)Ns;
    const bool one = total == n;
    for (int gid = tid; gid < total; gid += GT) {
        const int b = one ? 0 : fdiv(gid, inv_n), o = gid - b * n;
        const int q = fdiv(o, inv_span), rem = o - q * span;
        const int t = Ns == 1 ? rem : fdiv(rem, inv_ns), k = rem - t * Ns;
        const int e = (k + t * Ns) * step; // < n
        const cpx* x = in + b * n + q * Ns + k;
        cpx acc = x[0];
        int idx = 0;
#pragma unroll
        for (int r = 1; r < p; ++r) {
            idx += e;
            if (idx >= n) idx -= n;
            cpx w = __ldg(reinterpret_cast<const float2*>(tw) + idx);
            if (INV) w.y = -w.y;
            acc = cfma(x[r * stride], w, acc);
        }
        out[gid] = acc;
    }
}
template <bool INV>
__device__ __forceinline__ void smem_pass_any(cpx* out, const cpx* in, const cpx* tw, int n, int p, int Ns, int total, int tid)
{
    switch (p) {
    case 2: smem_pass<INV, 2>(out, in, tw, n, p, Ns, total, tid); break;
    case 3: smem_pass<INV, 3>(out, in, tw, n, p, Ns, total, tid); break;
    case 4: smem_pass<INV, 4>(out, in, tw, n, p, Ns, total, tid); break;
    case 5: smem_pass<INV, 5>(out, in, tw, n, p, Ns, total, tid); break;
    default: smem_pass<INV, 0>(out, in, tw, n, p, Ns, total, tid); break;
    }
}
// whole transform between the two buffers; returns the buffer that holds the result
template <bool INV>
__device__ __forceinline__ cpx* smem_fft(cpx* src, cpx* dst, const cpx* tw, int n, const int* rad, int n_rad, int batch, int tid)
{
    int Ns = 1;
    for (int i = 0; i < n_rad; ++i) {
        smem_pass_any<INV>(dst, src, tw, n, rad[i], Ns, batch * n, tid);
        __syncthreads();
        Ns *= rad[i];
        cpx* t = src;
        src = dst;
        dst = t;
    }
    return src;
}

__global__ void __launch_bounds__(GT) generic_smem_mod_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, int n_frames,
                                                              const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(16) unsigned char gsm[];
    cpx* A = reinterpret_cast<cpx*>(gsm);
    cpx* B = A + a.N;
    const int tid = threadIdx.x, M = a.M, K = a.K, L = a.L, N = a.N, h = L / 2;
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M;
    const float inv_n = 1.0f / (float)N;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const cpx* x = in + (size_t)f * N;
        for (int i = tid; i < N; i += GT) A[i] = x[i];
        __syncthreads();
        cpx* D = smem_fft<false>(A, B, a.tw_m, M, a.rad_m, a.n_rad_m, K, tid); // D_k = FFT_M(d_k), [k][m]
        cpx* X = D == A ? B : A;
        const float inv_mf = 1.0f / (float)M;
        for (int r = tid; r < N; r += GT) {
            const int b = fdiv(r, inv_mf), m = r - b * M;
            cpx acc = cmake(0.f, 0.f);
            if (m < part_len) {
                // the same accumulation order as the reference's k-loop produces for this bin (:113-135)
                int k = b + h, tp = (L - 1 + h) % L; // i = L-1 first: k = (b - i + h) mod K, tap block (i + h) mod L
                k -= L - 1;
                k %= K;
                if (k < 0) k += K;
                for (int i = L - 1; i >= 0; --i) {
                    acc = cadd(acc, cmul(D[k * M + m], __ldg(reinterpret_cast<const float2*>(a.taps) + tp * M + m)));
                    k = k + 1 == K ? 0 : k + 1;
                    tp = tp == 0 ? L - 1 : tp - 1;
                }
            }
            X[r] = acc;
        }
        __syncthreads();
        const cpx* y = smem_fft<true>(X, D, a.tw_n, N, a.rad_n, a.n_rad_n, 1, tid);
        cpx* o = out + (size_t)f * N;
        for (int i = tid; i < N; i += GT) o[i] = cscale(y[i], inv_n);
        __syncthreads(); // the buffers are reused by the next frame
    }
}

// mode 0: soft symbols y (generic_work[_equalize]); mode 1: R (fft_[equalize_]filter_downsample)
__global__ void __launch_bounds__(GT) generic_smem_rx_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                             const cpx* __restrict__ eq, int mode, int n_frames,
                                                             const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(16) unsigned char gsm[];
    cpx* A = reinterpret_cast<cpx*>(gsm);
    cpx* B = A + a.N;
    const int tid = threadIdx.x, M = a.M, K = a.K, L = a.L, N = a.N, h = L / 2;
    const float inv_m = 1.0f / (float)M;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const cpx* x = in + (size_t)f * N;
        for (int i = tid; i < N; i += GT) A[i] = x[i];
        __syncthreads();
        cpx* Y = smem_fft<false>(A, B, a.tw_n, N, a.rad_n, a.n_rad_n, 1, tid);
        if (eq != nullptr) { // volk_32fc_x2_divide_32fc (:315)
            const cpx* hq = eq + (size_t)f * N;
            for (int i = tid; i < N; i += GT) Y[i] = cdiv(Y[i], hq[i]);
            __syncthreads();
        }
        cpx* R = Y == A ? B : A;
        const float inv_mf = 1.0f / (float)M;
        for (int r = tid; r < N; r += GT) { // filter_subcarriers_and_downsample_fd (:165-192)
            const int k = fdiv(r, inv_mf), m = r - k * M;
            cpx acc = cmake(0.f, 0.f);
            int kk = (k + K - h % K) % K, tp = h % L; // i = 0: subcarrier (k - h) mod K, tap block h mod L
            for (int i = 0; i < L; ++i) {
                acc = cadd(acc, cmul(__ldg(reinterpret_cast<const float2*>(a.taps) + tp * M + m), Y[kk * M + m]));
                kk = kk + 1 == K ? 0 : kk + 1;
                tp = tp + 1 == L ? 0 : tp + 1;
            }
            R[r] = acc;
        }
        __syncthreads();
        cpx* o = out + (size_t)f * N;
        if (mode == 1) {
            for (int i = tid; i < N; i += GT) o[i] = R[i];
        } else {
            const cpx* y = smem_fft<true>(R, Y, a.tw_m, M, a.rad_m, a.n_rad_m, K, tid); // transform_subcarriers_to_td (:211-225)
            for (int i = tid; i < N; i += GT) o[i] = cscale(y[i], inv_m);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
bool generic_smem_supported(int M, int K, const FftPlan& fft_m, const FftPlan& fft_n)
{
    const size_t N = (size_t)M * K;
    return N >= 1 && N <= 12288 && fft_m.radices.size() <= (size_t)MAX_RAD && fft_n.radices.size() <= (size_t)MAX_RAD &&
           fft_m.d_tw != nullptr && fft_n.d_tw != nullptr;
}

static GenArgs make_args(int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n, const cpx* d_taps)
{
    GenArgs a{};
    a.M = M; a.K = K; a.L = L; a.N = M * K;
    a.n_rad_m = (int)fft_m.radices.size();
    a.n_rad_n = (int)fft_n.radices.size();
    for (int i = 0; i < a.n_rad_m; ++i) a.rad_m[i] = fft_m.radices[i];
    for (int i = 0; i < a.n_rad_n; ++i) a.rad_n[i] = fft_n.radices[i];
    a.tw_m = fft_m.d_tw; a.tw_n = fft_n.d_tw; a.taps = d_taps;
    return a;
}

template <class Kern>
static int grid_for_frames(Kern kern, size_t smem, size_t frames)
{
    int dev = 0, sms = 0, per_sm = 0;
    GFDM_CUDA_CHECK(cudaGetDevice(&dev));
    GFDM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GFDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GFDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GT, smem));
    if (per_sm < 1) throw CudaError("generic shared-memory kernel does not fit on this device");
    const size_t cap = (size_t)sms * per_sm;
    return (int)(frames < cap ? frames : cap);
}

int launch_generic_smem_mod(cpx* out, const cpx* in, int M, int K, int L, const FftPlan& fft_m, const FftPlan& fft_n,
                            const cpx* d_taps, size_t frames, cudaStream_t s)
{
    const GenArgs a = make_args(M, K, L, fft_m, fft_n, d_taps);
    const size_t smem = sizeof(cpx) * 2 * (size_t)a.N, max_chunk = (size_t)1 << 20;
    int launches = 0;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const size_t nf = frames - f0 < max_chunk ? frames - f0 : max_chunk;
        const int grid = grid_for_frames(generic_smem_mod_kernel, smem, nf);
        generic_smem_mod_kernel<<<grid, GT, smem, s>>>(out + f0 * a.N, in + f0 * a.N, (int)nf, a);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

int launch_generic_smem_rx(cpx* out, const cpx* in, const cpx* eq, int mode, int M, int K, int L, const FftPlan& fft_m,
                           const FftPlan& fft_n, const cpx* d_taps, size_t frames, cudaStream_t s)
{
    const GenArgs a = make_args(M, K, L, fft_m, fft_n, d_taps);
    const size_t smem = sizeof(cpx) * 2 * (size_t)a.N, max_chunk = (size_t)1 << 20;
    int launches = 0;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const size_t nf = frames - f0 < max_chunk ? frames - f0 : max_chunk;
        const int grid = grid_for_frames(generic_smem_rx_kernel, smem, nf);
        generic_smem_rx_kernel<<<grid, GT, smem, s>>>(out + f0 * a.N, in + f0 * a.N, eq ? eq + f0 * a.N : nullptr, mode, (int)nf, a);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

} // namespace gfdm
