// stage_kernels.cu -- coalesced gather/scatter and elementwise stages.
//
// These kernels are (a) the final implementation of the integer / copy paths
// (resource mapper, cyclic prefix, preamble copy, channel-estimator stages) and
// (b) the building blocks of the any-shape modulator/receiver pipeline used
// when no fused kernel exists for (M, K, L).  One thread per OUTPUT element, so
// every global store is coalesced; scatters of the reference are restated as
// gathers (no atomics).
#include "engine.h"
#include "regfft.cuh"

namespace gfdm {

static constexpr unsigned TH = 256;

// The copy / gather kernels below move E elements per thread: a CTA owns TH*E consecutive output elements, thread t
// takes t, t+TH, ... (every access of a warp is one contiguous 256-byte run), all E loads are issued before the
// first store (enough bytes in flight to cover HBM latency), and the (frame, offset) split costs one 64-bit
// division per thread instead of one per element.
static constexpr int EPT = 4;
struct RowPos {
    size_t f; // frame
    int i;    // element inside the frame's row of length W
    __device__ __forceinline__ RowPos(size_t gid, int W) : f(gid / (size_t)W), i((int)(gid - (gid / (size_t)W) * (size_t)W)) {}
    __device__ __forceinline__ void advance(int step, int W)
    {
        i += step;
        while (i >= W) {
            i -= W;
            ++f;
        }
    }
};
static inline unsigned blocks_ept(size_t total) { return blocks_for(total, TH * EPT); }

// ---------------------------------------------------------------------------
// modulator_kernel_cc::generic_work, lib/modulator_kernel_cc.cc:107-134, as a gather:
// X[b*M+m] = sum_i T[((i+h)%L)*M+m] * D[((b-i+h) mod K)*M+m]   for m < part_len, else 0
__global__ void __launch_bounds__(TH) mod_filter_kernel(cpx* __restrict__ X, const cpx* __restrict__ D,
                                                        const cpx* __restrict__ taps, int M, int K, int L,
                                                        int part_len, size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int N = M * K;
    const size_t f = gid / N;
    const int r = (int)(gid - f * N);
    const int b = r / M, m = r - b * M;
    cpx acc = cmake(0.f, 0.f);
    if (m < part_len) {
        const int h = L / 2;
        const cpx* Df = D + f * N;
        // same accumulation order as the reference's k-loop would produce for this bin
        for (int i = L - 1; i >= 0; --i) {
            int k = (b - i + h) % K;
            if (k < 0) k += K;
            acc = cadd(acc, cmul(Df[k * M + m], taps[((i + h) % L) * M + m]));
        }
    }
    X[gid] = acc;
}

void launch_mod_filter(cpx* X, const cpx* D, const cpx* taps, int M, int K, int L, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M;
    mod_filter_kernel<<<blocks_for(total, TH), TH, 0, s>>>(X, D, taps, M, K, L, part_len, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// receiver_kernel_cc::filter_subcarriers_and_downsample_fd, lib/receiver_kernel_cc.cc:165-192
__global__ void __launch_bounds__(TH) rx_filter_kernel(cpx* __restrict__ R, const cpx* __restrict__ Y,
                                                       const cpx* __restrict__ taps, int M, int K, int L,
                                                       size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int N = M * K;
    const size_t f = gid / N;
    const int r = (int)(gid - f * N);
    const int k = r / M, m = r - k * M;
    const int h = L / 2;
    const cpx* Yf = Y + f * N;
    cpx acc = cmake(0.f, 0.f);
    for (int i = 0; i < L; ++i) {
        const int src = ((k + i + K - h) % K) * M;
        acc = cadd(acc, cmul(taps[((i + h) % L) * M + m], Yf[src + m]));
    }
    R[gid] = acc;
}

void launch_rx_filter(cpx* R, const cpx* Y, const cpx* taps, int M, int K, int L, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    rx_filter_kernel<<<blocks_for(total, TH), TH, 0, s>>>(R, Y, taps, M, K, L, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// volk_32fc_x2_divide_32fc, lib/receiver_kernel_cc.cc:315
__global__ void __launch_bounds__(TH) eq_divide_kernel(cpx* __restrict__ out, const cpx* __restrict__ Y,
                                                       const cpx* __restrict__ eq, size_t n)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) out[gid] = cdiv(Y[gid], eq[gid]);
}
void launch_eq_divide(cpx* out, const cpx* Y, const cpx* eq, size_t n, cudaStream_t s)
{
    if (!n) return;
    eq_divide_kernel<<<blocks_for(n, TH), TH, 0, s>>>(out, Y, eq, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// cancel_sc_interference, lib/receiver_kernel_cc.cc:279-286: td[(k-1)%K] + td[(k+1)%K]
__global__ void __launch_bounds__(TH) neighbor_sum_kernel(cpx* __restrict__ out, const cpx* __restrict__ td, int M,
                                                          int K, size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int N = M * K;
    const size_t f = gid / N;
    const int r = (int)(gid - f * N);
    const int k = r / M, m = r - k * M;
    const int prev = (k - 1 + K) % K, next = (k + 1) % K;
    const cpx* t = td + f * N;
    out[gid] = cadd(t[prev * M + m], t[next * M + m]);
}
void launch_neighbor_sum(cpx* out, const cpx* td, int M, int K, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    neighbor_sum_kernel<<<blocks_for(total, TH), TH, 0, s>>>(out, td, M, K, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// lib/receiver_kernel_cc.cc:288-297: out = fd - ic_taps * F
__global__ void __launch_bounds__(TH) ic_subtract_kernel(cpx* __restrict__ out, const cpx* __restrict__ fd,
                                                         const cpx* __restrict__ F, const cpx* __restrict__ ic, int M,
                                                         size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int m = (int)(gid % M);
    out[gid] = csub(fd[gid], cmul(ic[m], F[gid]));
}
void launch_ic_subtract(cpx* out, const cpx* fd, const cpx* F, const cpx* ic_taps, int M, int K, size_t frames,
                        cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    ic_subtract_kernel<<<blocks_for(total, TH), TH, 0, s>>>(out, fd, F, ic_taps, M, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// One interference-cancellation iteration as ONE kernel, for the shapes whose whole loop is not resident in the
// fused receiver (lib/advanced_receiver_kernel_cc.cc:56-76 = map_symbols_to_constellation_points :109-123,
// cancel_sc_interference lib/receiver_kernel_cc.cc:274-299, transform_subcarriers_to_td :211-225):
//   y'_k = IFFT_M( R_k - ic (.) FFT_M( dec(y_{k-1}) + dec(y_{k+1}) ) ) / M,   dec = hard decision (0 on unused subcarriers)
// A CTA owns SB consecutive subcarriers of one frame: their soft symbols plus one halo record either side and their
// kept frequency blocks R are staged in shared memory with full-line accesses, thread <-> subcarrier does the two
// M-point transforms in registers, results leave through shared memory again.  Traffic: read y and R once, write y'
// once (24 bytes per symbol and iteration) instead of five kernels and four intermediate arrays.
static constexpr int SB = 128;
template <int M>
__global__ void __launch_bounds__(SB) sic_iter_kernel(cpx* __restrict__ y_out, const cpx* __restrict__ y_in,
                                                      const cpx* __restrict__ fb, const cpx* __restrict__ ic_taps,
                                                      const unsigned char* __restrict__ active,
                                                      const cpx* __restrict__ points, int n_points, int rule, int K,
                                                      const __grid_constant__ DecideGrid grid)
{
    extern __shared__ __align__(16) unsigned char sic_smem[];
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = grid.lut[threadIdx.x];
    cpx* ys = reinterpret_cast<cpx*>(sic_smem); // [SB + 2][M]: record 0 = subcarrier k0-1, record SB+1 = k0+SB
    cpx* rs = ys + (SB + 2) * M;                // [SB][M]: R in, y' out
    __shared__ cpx pts[64];
    __shared__ cpx ic_s[M];
    const int tid = threadIdx.x;
    const int blocks_per_frame = (K + SB - 1) / SB;
    const size_t f = blockIdx.x / blocks_per_frame;
    const int k0 = (int)(blockIdx.x - f * blocks_per_frame) * SB;
    const int nb = min(SB, K - k0);
    const cpx* yf = y_in + f * (size_t)K * M;
    for (int i = tid; i < 64; i += SB) pts[i] = i < n_points ? points[i] : cmake(0.f, 0.f);
    for (int i = tid; i < M; i += SB) ic_s[i] = ic_taps[i];
    for (int i = tid; i < nb * M; i += SB) {
        ys[M + i] = yf[(size_t)k0 * M + i];
        rs[i] = fb[(f * K + k0) * (size_t)M + i];
    }
    const int kp = k0 == 0 ? K - 1 : k0 - 1, kn = k0 + nb >= K ? 0 : k0 + nb;
    for (int i = tid; i < M; i += SB) {
        ys[i] = yf[(size_t)kp * M + i];
        ys[(nb + 1) * M + i] = yf[(size_t)kn * M + i];
    }
    __syncthreads();
    if (tid < nb) {
        const int k = k0 + tid;
        const bool ap = active[k == 0 ? K - 1 : k - 1] != 0, an = active[k == K - 1 ? 0 : k + 1] != 0;
        const cpx* prev = ys + tid * M;       // record of subcarrier k-1
        const cpx* next = ys + (tid + 2) * M; // record of subcarrier k+1
        cpx d[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const cpx a = ap ? pts[decide_symbol_grid(prev[m], pts, n_points, rule, grid, lut)] : cmake(0.f, 0.f);
            const cpx b = an ? pts[decide_symbol_grid(next[m], pts, n_points, rule, grid, lut)] : cmake(0.f, 0.f);
            d[m] = cadd(a, b);
        }
        rf::FFTN<M, -1>::run(d);
        cpx* mine = rs + tid * M;
        const float inv_m = 1.0f / (float)M;
#pragma unroll
        for (int m = 0; m < M; ++m) d[m] = csub(mine[m], cmul(ic_s[m], d[m]));
        rf::FFTN<M, +1>::run(d);
#pragma unroll
        for (int m = 0; m < M; ++m) mine[m] = cscale(d[m], inv_m);
    }
    __syncthreads();
    cpx* of = y_out + (f * K + k0) * (size_t)M;
    for (int i = tid; i < nb * M; i += SB) of[i] = rs[i];
}
template <int M>
static void launch_sic_iter_m(cpx* y_out, const cpx* y_in, const cpx* fb, const cpx* ic_taps, const unsigned char* active,
                              const cpx* points, int n_points, int rule, const DecideGrid& grid, int K, size_t frames,
                              cudaStream_t s)
{
    const size_t smem = sizeof(cpx) * (size_t)(2 * SB + 2) * M;
    if (smem > 48 * 1024)
        GFDM_CUDA_CHECK(cudaFuncSetAttribute(sic_iter_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t blocks = frames * (size_t)((K + SB - 1) / SB);
    sic_iter_kernel<M><<<(unsigned)blocks, SB, smem, s>>>(y_out, y_in, fb, ic_taps, active, points, n_points, rule, K, grid);
    GFDM_CUDA_CHECK(cudaGetLastError());
}
bool sic_iter_supported(int M, int K, int n_points, size_t frames)
{
    switch (M) {
    case 2: case 3: case 4: case 5: case 7: case 8: case 9: case 15: case 16: case 21: case 25: break;
    default: return false;
    }
    return K >= 2 && n_points >= 1 && n_points <= 64 && frames * (size_t)((K + SB - 1) / SB) < ((size_t)1 << 31);
}
void launch_sic_iter(cpx* y_out, const cpx* y_in, const cpx* fb, const cpx* ic_taps, const unsigned char* active,
                     const cpx* points, int n_points, int rule, const DecideGrid& grid, int M, int K, size_t frames,
                     cudaStream_t s)
{
    if (!frames) return;
#define GFDM_SIC_CASE(MM) \
    case MM: launch_sic_iter_m<MM>(y_out, y_in, fb, ic_taps, active, points, n_points, rule, grid, K, frames, s); break;
    switch (M) {
        GFDM_SIC_CASE(2) GFDM_SIC_CASE(3) GFDM_SIC_CASE(4) GFDM_SIC_CASE(5) GFDM_SIC_CASE(7) GFDM_SIC_CASE(8)
        GFDM_SIC_CASE(9) GFDM_SIC_CASE(15) GFDM_SIC_CASE(16) GFDM_SIC_CASE(21) GFDM_SIC_CASE(25)
    default: throw std::invalid_argument("sic_iter: unsupported timeslot count");
    }
#undef GFDM_SIC_CASE
}

// map_symbols_to_constellation_points, lib/advanced_receiver_kernel_cc.cc:109-123 (decide_symbol: common.cuh)
__global__ void __launch_bounds__(TH) decide_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                    const unsigned char* __restrict__ active,
                                                    const cpx* __restrict__ points, int n_points, int rule, int M,
                                                    int K, size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int k = (int)((gid / M) % K);
    cpx v = cmake(0.f, 0.f);
    if (active[k]) v = points[decide_symbol(in[gid], points, n_points, rule)];
    out[gid] = v;
}
void launch_decide(cpx* out, const cpx* in, const unsigned char* active, const cpx* points, int n_points, int rule,
                   int M, int K, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    decide_kernel<<<blocks_for(total, TH), TH, 0, s>>>(out, in, active, points, n_points, rule, M, K, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// calculate_phase_offset + rotator, lib/advanced_receiver_kernel_cc.cc:61-91.
// One CTA per frame: block-reduce mean(arg(decided) - arg(soft)) over the map
// (duplicates in the map count twice, as in the reference), then rotate the
// kept frequency block in place.
__global__ void __launch_bounds__(256) phase_rotate_kernel(cpx* __restrict__ R, const cpx* __restrict__ decided,
                                                           const cpx* __restrict__ soft, const int* __restrict__ smap,
                                                           int n_map, int M, int K)
{
    __shared__ float red[256];
    __shared__ cpx rot;
    const size_t f = blockIdx.x;
    const int N = M * K;
    const cpx* d = decided + f * N;
    const cpx* y = soft + f * N;
    float acc = 0.f;
    const int total = n_map * M;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int a = i / M, m = i - a * M;
        const int pos = smap[a] * M + m;
        acc += atan2f(d[pos].y, d[pos].x) - atan2f(y[pos].y, y[pos].x);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float ph = red[0] / (float)((size_t)n_map * (size_t)M);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        rot = cmake(cs, sn);
    }
    __syncthreads();
    const cpx w = rot;
    cpx* r = R + f * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) r[i] = cmul(r[i], w);
}
void launch_phase_rotate(cpx* R, const cpx* decided, const cpx* soft, const int* smap, int n_map, int M, int K,
                         size_t frames, cudaStream_t s)
{
    if (!frames) return;
    phase_rotate_kernel<<<(unsigned)frames, 256, 0, s>>>(R, decided, soft, smap, n_map, M, K);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// resource_mapper_kernel_cc::map_to_resources, lib/resource_mapper_kernel_cc.cc:74-134,
// as a gather over the K x M grid: inv_map[k] = position of k in the sorted map or -1.
__global__ void __launch_bounds__(TH) map_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                 const int* __restrict__ inv_map, int M, int K, int A,
                                                 int per_timeslot, size_t n_in, size_t in_stride, size_t total)
{
    const int N = M * K;
    const size_t gid0 = (size_t)blockIdx.x * (TH * EPT) + threadIdx.x;
    RowPos p(gid0, N);
    cpx v[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        v[j] = cmake(0.f, 0.f);
        if (gid0 + (size_t)j * TH < total) {
            const int k = p.i / M, t = p.i - k * M;
            const int a = inv_map[k];
            if (a >= 0) {
                const size_t src = per_timeslot ? (size_t)t * A + a : (size_t)a * M + t;
                if (src < n_in) v[j] = in[p.f * in_stride + src];
            }
        }
        p.advance(TH, N);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j)
        if (gid0 + (size_t)j * TH < total) out[gid0 + (size_t)j * TH] = v[j];
}
void launch_map(cpx* out, const cpx* in, const int* inv_map, int M, int K, int A, bool per_timeslot, size_t n_in,
                size_t in_stride, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    map_kernel<<<blocks_ept(total), TH, 0, s>>>(out, in, inv_map, M, K, A, per_timeslot ? 1 : 0, n_in, in_stride,
                                                    total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// demap_from_resources, lib/resource_mapper_kernel_cc.cc:91-106,136-162.  Writes exactly
// n_out symbols per frame (the reference's one-past write in the per-subcarrier
// branch is an overflow of the caller's buffer and is deliberately not reproduced).
__global__ void __launch_bounds__(TH) demap_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                   const int* __restrict__ smap, int M, int K, int A, int per_timeslot,
                                                   size_t n_out, size_t out_stride, size_t total)
{
    const size_t gid0 = (size_t)blockIdx.x * (TH * EPT) + threadIdx.x;
    const int W = (int)n_out;
    RowPos p(gid0, W);
    cpx v[EPT];
    size_t dst[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        dst[j] = p.f * out_stride + p.i;
        if (gid0 + (size_t)j * TH < total) {
            int a, t;
            if (per_timeslot) {
                t = p.i / A;
                a = p.i - t * A;
            } else {
                a = p.i / M;
                t = p.i - a * M;
            }
            v[j] = in[p.f * (size_t)M * K + (size_t)M * smap[a] + t];
        }
        p.advance(TH, W);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j)
        if (gid0 + (size_t)j * TH < total) out[dst[j]] = v[j];
}
// per-timeslot order is a transpose between the grid ([k][t]) and the compact vector ([t][a]).  For wide grids a CTA
// moves a tile of TS slots through shared memory so that both sides see contiguous runs (grid side: whole records,
// compact side: TS consecutive slots of one timeslot); measured on B200: 56.6 -> 65 % of the copy roofline at K = 1024,
// but slower than the 4-elements-per-thread gather for small grids (K = 64: 85 -> 50 %), hence the K >= 512 switch.
// The same tiling gained nothing for map_to_resources (73 -> 74 %) and is not used there.
static constexpr int TS = 64;
__global__ void __launch_bounds__(TH) demap_tiled_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                         const int* __restrict__ smap, int M, int K, int A, size_t n_out,
                                                         size_t out_stride)
{
    extern __shared__ __align__(16) unsigned char map_smem[];
    cpx* tile = reinterpret_cast<cpx*>(map_smem); // [TS][M]
    __shared__ int sub[TS];
    const int tiles = (A + TS - 1) / TS;
    const size_t f = blockIdx.x / tiles;
    const int a0 = (int)(blockIdx.x - f * tiles) * TS;
    const int na = min(TS, A - a0);
    if ((int)threadIdx.x < TS) sub[threadIdx.x] = (int)threadIdx.x < na ? smap[a0 + threadIdx.x] : 0;
    __syncthreads();
    const cpx* src = in + f * (size_t)M * K;
    // records of the tile's subcarriers: consecutive threads on consecutive elements of a record
    int j = threadIdx.x / M, t = threadIdx.x - j * M;
    const int qj = (int)TH / M, rt = (int)TH - qj * M;
    for (int i = threadIdx.x; i < na * M; i += TH) {
        tile[i] = src[(size_t)sub[j] * M + t];
        t += rt;
        j += qj;
        if (t >= M) {
            t -= M;
            ++j;
        }
    }
    __syncthreads();
    cpx* dst = out + f * out_stride;
    for (int i = threadIdx.x; i < M * TS; i += TH) {
        const int tt = i / TS, jj = i - tt * TS;
        const size_t e = (size_t)tt * A + a0 + jj;
        if (jj < na && e < n_out) dst[e] = tile[jj * M + tt];
    }
}
void launch_demap(cpx* out, const cpx* in, const int* smap, int M, int K, int A, bool per_timeslot, size_t n_out,
                  size_t out_stride, size_t frames, cudaStream_t s)
{
    const size_t total = frames * n_out;
    if (!total) return;
    const size_t tiled_blocks = frames * (size_t)((A + TS - 1) / TS);
    // static `sub[TS]` + the dynamic tile must stay inside the 48 KB that need no opt-in
    if (per_timeslot && K >= 512 && sizeof(cpx) * TS * M + sizeof(int) * TS <= 48 * 1024 && A > 0 && tiled_blocks < ((size_t)1 << 31)) {
        demap_tiled_kernel<<<(unsigned)tiled_blocks, TH, sizeof(cpx) * TS * M, s>>>(out, in, smap, M, K, A, n_out, out_stride);
        GFDM_CUDA_CHECK(cudaGetLastError());
        return;
    }
    demap_kernel<<<blocks_ept(total), TH, 0, s>>>(out, in, smap, M, K, A, per_timeslot ? 1 : 0, n_out, out_stride,
                                                      total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// add_cyclic_prefix_cc::add_cyclic_prefix, lib/add_cyclic_prefix_cc.cc:67-98:
// o[i] = x[(i + N - cp - s) mod N], first/last ramp samples times the ramp
// (complex x complex, unfused, reference operand order -> bit-exact).
__global__ void __launch_bounds__(TH) add_cp_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, int N, int cp,
                                                    int cs, int ramp, const cpx* __restrict__ front,
                                                    const cpx* __restrict__ back, int shift, size_t out_stride,
                                                    size_t total)
{
    const int W = N + cp + cs;
    const size_t gid0 = (size_t)blockIdx.x * (TH * EPT) + threadIdx.x;
    RowPos p(gid0, W);
    cpx v[EPT];
    size_t dst[EPT];
    int pos[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        dst[j] = p.f * out_stride + p.i;
        pos[j] = p.i;
        if (gid0 + (size_t)j * TH < total) {
            int src = p.i + N - cp - shift;
            while (src >= N) src -= N;
            v[j] = in[p.f * N + src];
        }
        p.advance(TH, W);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j)
        if (gid0 + (size_t)j * TH < total) {
            cpx x = v[j];
            const int i = pos[j];
            if (i < ramp) x = cmul_rn(x, front[i]);
            if (i >= W - ramp) x = cmul_rn(x, back[i - (W - ramp)]);
            out[dst[j]] = x;
        }
}
void launch_add_cp(cpx* out, const cpx* in, int N, int cp, int cs, int ramp, const cpx* front, const cpx* back,
                   int shift, size_t out_stride, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)(N + cp + cs);
    if (!total) return;
    add_cp_kernel<<<blocks_ept(total), TH, 0, s>>>(out, in, N, cp, cs, ramp, front, back, shift, out_stride, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// remove_cyclic_prefix, lib/add_cyclic_prefix_cc.cc:100-104
__global__ void __launch_bounds__(TH) remove_cp_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, int N,
                                                       int cp, int W, size_t total)
{
    const size_t gid0 = (size_t)blockIdx.x * (TH * EPT) + threadIdx.x;
    RowPos p(gid0, N);
    cpx v[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        if (gid0 + (size_t)j * TH < total) v[j] = in[p.f * W + cp + p.i];
        p.advance(TH, N);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j)
        if (gid0 + (size_t)j * TH < total) out[gid0 + (size_t)j * TH] = v[j];
}
void launch_remove_cp(cpx* out, const cpx* in, int N, int cp, int cs, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)N;
    if (!total) return;
    remove_cp_kernel<<<blocks_ept(total), TH, 0, s>>>(out, in, N, cp, N + cp + cs, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// transmitter_kernel::insert_preamble, lib/transmitter_kernel.cc:86-90
__global__ void __launch_bounds__(TH) copy_rows_kernel(cpx* __restrict__ out, const cpx* __restrict__ row, int len,
                                                       size_t out_stride, size_t total)
{
    const size_t gid0 = (size_t)blockIdx.x * (TH * EPT) + threadIdx.x;
    RowPos p(gid0, len);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        if (gid0 + (size_t)j * TH < total) out[p.f * out_stride + p.i] = row[p.i];
        p.advance(TH, len);
    }
}
void launch_copy_rows(cpx* out, const cpx* row, int len, size_t out_stride, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)len;
    if (!total) return;
    copy_rows_kernel<<<blocks_ept(total), TH, 0, s>>>(out, row, len, out_stride, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// preamble_channel_estimator_cc, lib/preamble_channel_estimator_cc.cc
// :121-143  H = F0 * inv0 + F1 * inv1   (F = [frames][2][K])
__global__ void __launch_bounds__(TH) est_combine_kernel(cpx* __restrict__ H, const cpx* __restrict__ F,
                                                         const cpx* __restrict__ inv0, const cpx* __restrict__ inv1,
                                                         int K, size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const size_t f = gid / K;
    const int q = (int)(gid - f * K);
    const cpx a = cmul_rn(F[f * 2 * K + q], inv0[q]);
    const cpx b = cmul_rn(F[f * 2 * K + K + q], inv1[q]);
    H[gid] = cmake(__fadd_rn(b.x, a.x), __fadd_rn(b.y, a.y));
}
void launch_est_combine(cpx* H, const cpx* F, const cpx* inv0, const cpx* inv1, int K, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)K;
    if (!total) return;
    est_combine_kernel<<<blocks_for(total, TH), TH, 0, s>>>(H, F, inv0, inv1, K, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// :145-185  reorder + edge replicate + 9-tap Gaussian correlation
__device__ __forceinline__ cpx est_padded(const cpx* __restrict__ H, int j, int K, int A, int off)
{
    const int G2 = 4, Ah = A / 2;
    if (j < G2) return H[K - Ah];
    j -= G2;
    if (j < Ah) return H[j + K - Ah];
    if (off && j == Ah) {
        const cpx a = H[K - 1], b = H[1];
        return cmake((a.x + b.x) / 2.0f, (a.y + b.y) / 2.0f);
    }
    j -= Ah + off;
    if (j < Ah) return H[off + j];
    return H[off + Ah - 1];
}
__global__ void __launch_bounds__(TH) est_filter_kernel(cpx* __restrict__ filt, const cpx* __restrict__ H,
                                                        const float* __restrict__ g, int K, int A, int off,
                                                        size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int n_taps = A + off;
    const size_t f = gid / n_taps;
    const int i = (int)(gid - f * n_taps);
    const cpx* Hf = H + f * K;
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const cpx v = est_padded(Hf, i + t, K, A, off);
        re = __fadd_rn(re, __fmul_rn(v.x, g[t]));
        im = __fadd_rn(im, __fmul_rn(v.y, g[t]));
    }
    filt[gid] = cmake(re, im);
}
void launch_est_filter(cpx* filt, const cpx* H, const float* g, int K, int A, int dc_free, size_t frames,
                       cudaStream_t s)
{
    const int off = dc_free ? 1 : 0;
    const size_t total = frames * (size_t)(A + off);
    if (!total) return;
    est_filter_kernel<<<blocks_for(total, TH), TH, 0, s>>>(filt, H, g, K, A, off, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// :238-274  piecewise-linear K -> N bins; one thread per output bin replays the
// reference's running float sum (factor += inc, j times) so results match bit
// for bit; later loops of the reference overwrite earlier ones, hence the
// reverse priority below.  Bins no loop writes are left untouched.
__global__ void __launch_bounds__(TH) est_interp_kernel(cpx* __restrict__ frame, const cpx* __restrict__ filt, int M,
                                                        int K, int A, int off, size_t total)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int N = M * K;
    const size_t f = gid / N;
    const int b = (int)(gid - f * N);
    const int n_est = A + off;
    const int center = N / 2;
    const int dead_half = M * (K - A) / 2;
    const cpx* e = filt + f * n_est;
    const int half = n_est / 2;
    int seg = -1, j = 0;
    cpx fill;
    bool write = false, is_fill = false;
    if (b < (n_est - 1 - half) * M) { // last loop: i in [half, n_est-1)
        seg = half + b / M;
        j = b % M;
        write = true;
    } else if (b >= center + dead_half && b < center + dead_half + half * M) { // i in [0, half)
        seg = (b - center - dead_half) / M;
        j = (b - center - dead_half) % M;
        write = true;
    } else if (b >= M * A / 2 && b < center) {
        fill = e[n_est - 1];
        write = is_fill = true;
    } else if (b >= center && b < center + dead_half) {
        fill = e[0];
        write = is_fill = true;
    }
    if (!write) return;
    if (is_fill) {
        frame[gid] = fill;
        return;
    }
    const float step = 1.0f / (float)M;
    const cpx d = csub(e[seg + 1], e[seg]);
    // (estimate[i+1] - estimate[i]) * gfdm_complex(step, 0): complex x complex, unfused
    const cpx inc = cmake(__fsub_rn(__fmul_rn(d.x, step), __fmul_rn(d.y, 0.0f)),
                          __fadd_rn(__fmul_rn(d.x, 0.0f), __fmul_rn(d.y, step)));
    cpx factor = e[seg];
    for (int t = 0; t < j; ++t) factor = cmake(__fadd_rn(factor.x, inc.x), __fadd_rn(factor.y, inc.y));
    frame[gid] = factor;
}
void launch_est_interp(cpx* frame, const cpx* filt, int M, int K, int A, int dc_free, size_t frames, cudaStream_t s)
{
    const size_t total = frames * (size_t)M * K;
    if (!total) return;
    est_interp_kernel<<<blocks_for(total, TH), TH, 0, s>>>(frame, filt, M, K, A, dc_free ? 1 : 0, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

__device__ __forceinline__ void stg_stream_cpx(cpx* p, cpx v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------
// estimate_frame (:285-294) as ONE kernel for power-of-two fft_len: a CTA keeps a frame's preamble halves in shared
// memory, runs both K-point transforms there (Stockham autosort, radix 4 (+ one radix-2 pass), double-precision twiddle table),
// forms H, the 9-tap filtered estimate, and streams the N interpolated bins out -- 8*(2K + N) bytes of HBM traffic
// per frame, against four kernels with three intermediate arrays.  Arithmetic after the transforms is the unfused
// fp32 of the stage kernels above (same bits); grid-stride over frames, CTAs sized by occupancy.
__global__ void __launch_bounds__(TH) est_fused_kernel(cpx* __restrict__ frame, const cpx* __restrict__ rx,
                                                       const cpx* __restrict__ tw, const cpx* __restrict__ inv0,
                                                       const cpx* __restrict__ inv1, const float* __restrict__ g, int M,
                                                       int K, int A, int off, int fpc, size_t frames)
{
    // fpc = frames per CTA pass: short transforms are batched so that every pass keeps all TH threads busy
    extern __shared__ __align__(16) unsigned char est_smem[];
    const int HH = 2 * fpc;                     // preamble halves in flight
    cpx* xa = reinterpret_cast<cpx*>(est_smem); // [HH][K] ping
    cpx* xb = xa + HH * K;                      // [HH][K] pong; later H [fpc][K] and the filtered estimates
    cpx* tws = xb + HH * K;                     // [K] twiddles W_K^e
    const int tid = threadIdx.x;
    const int N = M * K, n_est = A + off, half = n_est / 2;
    const int center = N / 2, dead_half = M * (K - A) / 2;
    float gt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) gt[t] = g[t];
    const int so0 = tid / M, j0 = tid - so0 * M, qTH = (int)TH / M, rTH = (int)TH - qTH * M; // walk of the interpolation loops
    for (int i = tid; i < K; i += TH) tws[i] = tw[i];
    const int Q = K / 4, lq = 31 - __clz(Q), lk2 = lq + 1; // K is a power of two
    for (size_t f0 = (size_t)blockIdx.x * fpc; f0 < frames; f0 += (size_t)gridDim.x * fpc) {
        const int nf = (int)(frames - f0 < (size_t)fpc ? frames - f0 : (size_t)fpc);
        const cpx* in = rx + f0 * 2 * (size_t)K;
        for (int i = tid; i < nf * 2 * K; i += TH) xa[i] = in[i];
        __syncthreads();
        // all transforms at once (h = which preamble half of which frame).  Radix-4 Stockham passes while they keep every
        // thread busy, radix-2 passes for the rest (one for K = 2 * 4^n); twiddles W_K^e from the shared-memory copy of
        // the double-precision table.  Measured on B200: K = 1024 radix 4: 0.194 -> 0.159 ms; K = 256 needs fpc = 2 for it.
        cpx* src = xa;
        cpx* dst = xb;
        int Ns = 1;
        for (; HH * Q >= (int)TH && Ns * 4 <= K; Ns <<= 2) {
            const int sh = K / (4 * Ns);
            for (int w = tid; w < HH * Q; w += TH) {
                const int h = w >> lq, j = w & (Q - 1);
                const int k = j & (Ns - 1);
                const cpx* x = src + h * K + j;
                const cpx v0 = x[0];
                const cpx v1 = cmul(x[Q], tws[k * sh]);
                const cpx v2 = cmul(x[2 * Q], tws[2 * k * sh]);
                const cpx v3 = cmul(x[3 * Q], tws[3 * k * sh]);
                const cpx a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3);
                const cpx t = csub(v1, v3);
                const cpx d = cmake(t.y, -t.x); // -j (v1 - v3)
                cpx* y = dst + h * K + ((j - k) << 2) + k;
                y[0] = cadd(a, c);
                y[Ns] = cadd(b, d);
                y[2 * Ns] = csub(a, c);
                y[3 * Ns] = csub(b, d);
            }
            __syncthreads();
            cpx* t2 = src;
            src = dst;
            dst = t2;
        }
        for (; Ns < K; Ns <<= 1) {
            for (int w = tid; w < HH * (K / 2); w += TH) {
                const int h = w >> lk2, j = w & (K / 2 - 1);
                const int k = j & (Ns - 1);
                const cpx a = src[h * K + j];
                const cpx b = cmul(src[h * K + j + K / 2], tws[k * (K / (2 * Ns))]);
                const int j0b = ((j - k) << 1) + k;
                dst[h * K + j0b] = cadd(a, b);
                dst[h * K + j0b + Ns] = csub(a, b);
            }
            __syncthreads();
            cpx* t2 = src;
            src = dst;
            dst = t2;
        }
        // H = F0 * inv0 + F1 * inv1 (:121-143) -> dst[fi][0..K)
        for (int q = tid; q < fpc * K; q += TH) {
            const int fi = q >> (lk2 + 1), qq = q & (K - 1);
            const cpx a = cmul_rn(src[(2 * fi) * K + qq], inv0[qq]);
            const cpx b = cmul_rn(src[(2 * fi + 1) * K + qq], inv1[qq]);
            dst[fi * K + qq] = cmake(__fadd_rn(b.x, a.x), __fadd_rn(b.y, a.y));
        }
        __syncthreads();
        // reorder + edge replicate + 9-tap correlation (:145-185) -> src[fi][0..n_est)
        for (int i = tid; i < fpc * n_est; i += TH) {
            const int fi = i / n_est, ii = i - fi * n_est;
            float re = 0.f, im = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const cpx v = est_padded(dst + fi * K, ii + t, K, A, off);
                re = __fadd_rn(re, __fmul_rn(v.x, gt[t]));
                im = __fadd_rn(im, __fmul_rn(v.y, gt[t]));
            }
            src[fi * K + ii] = cmake(re, im);
        }
        __syncthreads();
        // piecewise-linear interpolation to N bins (:238-274): the cases of est_interp_kernel as four division-free
        // loops (thread t walks bins t, t+TH, ...: segment and offset advance by TH/M and TH%M); the reference's
        // running sum e[i] + inc + ... + inc (j times) is evaluated as fma(j, inc, e[i]) -- at most j ulps apart
        const float step = 1.0f / (float)M;
        for (int fi = 0; fi < nf; ++fi) {
            const cpx* e = src + fi * K;
            cpx* o = frame + (f0 + fi) * (size_t)N;
            auto ramp = [&](int lo, int n_bins, int seg0) {
                int seg = seg0 + so0, j = j0;
                for (int r = tid; r < n_bins; r += TH) {
                    const cpx e0 = e[seg], d = csub(e[seg + 1], e0);
                    const float fj = (float)j;
                    stg_stream_cpx(o + lo + r, cmake(fmaf(fj, d.x * step, e0.x), fmaf(fj, d.y * step, e0.y)));
                    j += rTH;
                    seg += qTH;
                    if (j >= M) {
                        j -= M;
                        ++seg;
                    }
                }
            };
            ramp(0, (n_est - 1 - half) * M, half);          // last loop of the reference: i in [half, n_est-1)
            ramp(center + dead_half, half * M, 0);          // i in [0, half)
            const cpx hi = e[n_est - 1], lo = e[0];
            for (int b2 = M * A / 2 + tid; b2 < center; b2 += TH) stg_stream_cpx(o + b2, hi);
            for (int b2 = center + tid; b2 < center + dead_half; b2 += TH) stg_stream_cpx(o + b2, lo);
        }
        __syncthreads(); // src/dst are reused by the next pass
    }
}
bool est_fused_supported(int K, int A, int dc_free)
{
    // even A: the four bin ranges of the interpolation are then disjoint, so their loops need no ordering
    return K >= 8 && K <= 4096 && (K & (K - 1)) == 0 && A >= 2 && A % 2 == 0 && A + (dc_free ? 1 : 0) <= K;
}
void launch_est_fused(cpx* frame, const cpx* rx, const cpx* tw, const cpx* inv0, const cpx* inv1, const float* g, int M,
                      int K, int A, int dc_free, size_t frames, cudaStream_t s)
{
    if (!frames) return;
    int fpc = (2 * (int)TH) / K; // short transforms: batch frames until a radix-4 pass has TH butterflies
    fpc = fpc < 1 ? 1 : (fpc > 8 ? 8 : fpc);
    const size_t smem = sizeof(cpx) * (size_t)(4 * fpc + 1) * (size_t)K;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        GFDM_CUDA_CHECK(cudaGetDevice(&dev));
        GFDM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (smem > 48 * 1024)
        GFDM_CUDA_CHECK(cudaFuncSetAttribute(est_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    GFDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, est_fused_kernel, (int)TH, smem));
    const size_t cap = (size_t)sms * (per_sm > 0 ? per_sm : 1);
    const size_t passes = (frames + fpc - 1) / fpc;
    est_fused_kernel<<<(unsigned)(passes < cap ? passes : cap), TH, smem, s>>>(frame, rx, tw, inv0, inv1, g, M, K, A,
                                                                                dc_free ? 1 : 0, fpc, frames);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// :187-235  one CTA per frame; F2 = FFT_2K(rx)
__global__ void __launch_bounds__(256) est_snr_kernel(float* __restrict__ snr, float* __restrict__ cnrs,
                                                      const cpx* __restrict__ F2, int K, int A, int off)
{
    __shared__ float rs[256], rn[256];
    __shared__ float s_scale;
    const size_t f = blockIdx.x;
    const cpx* F = F2 + f * 2 * K;
    const int half = A / 2;
    const int low_offset = (K - A) / 2 + K / 2;
    float se = 0.f, ne = 0.f;
    for (int i = threadIdx.x; i < 2 * half; i += blockDim.x) {
        const int pos = (i < half) ? 2 * (i + off) : 2 * (i - half + low_offset);
        const cpx a = F[pos], b = F[pos + 1];
        se += a.x * a.x + a.y * a.y;
        ne += b.x * b.x + b.y * b.y;
    }
    rs[threadIdx.x] = se;
    rn[threadIdx.x] = ne;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if ((int)threadIdx.x < sft) {
            rs[threadIdx.x] += rs[threadIdx.x + sft];
            rn[threadIdx.x] += rn[threadIdx.x + sft];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float v = (rs[0] - rn[0]) / rn[0];
        snr[f] = v;
        s_scale = v / (rs[0] / (float)A);
    }
    __syncthreads();
    if (cnrs) {
        const float sc = s_scale;
        for (int i = threadIdx.x; i < 2 * half; i += blockDim.x) {
            const int pos = (i < half) ? 2 * (i + off) : 2 * (i - half + low_offset);
            const cpx a = F[pos];
            cnrs[f * A + i] = (a.x * a.x + a.y * a.y) * sc;
        }
    }
}
void launch_est_snr(float* snr, float* cnrs, const cpx* F2, int K, int A, int dc_free, size_t frames, cudaStream_t s)
{
    if (!frames) return;
    est_snr_kernel<<<(unsigned)frames, 256, 0, s>>>(snr, cnrs, F2, K, A, dc_free ? 1 : 0);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// prepare_for_zf, :276-282: conj(1 / h)
__global__ void __launch_bounds__(TH) zf_prepare_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, size_t n)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) out[gid] = cconj(cdiv(cmake(1.f, 0.f), in[gid]));
}
void launch_zf_prepare(cpx* out, const cpx* in, size_t n, cudaStream_t s)
{
    if (!n) return;
    zf_prepare_kernel<<<blocks_for(n, TH), TH, 0, s>>>(out, in, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// gfdm_kernel_utils::calculate_signal_energy, lib/gfdm_kernel_utils.cc:59-65 (single CTA)
__global__ void __launch_bounds__(1024) energy_kernel(float* __restrict__ out, const cpx* __restrict__ in, size_t n)
{
    __shared__ float red[1024];
    float acc = 0.f;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += in[i].x * in[i].x + in[i].y * in[i].y;
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int sft = 512; sft > 0; sft >>= 1) {
        if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}
void launch_energy(float* out, const cpx* in, size_t n, cudaStream_t s)
{
    energy_kernel<<<1, 1024, 0, s>>>(out, in, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

} // namespace gfdm
