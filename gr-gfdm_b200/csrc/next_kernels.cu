// next_kernels.cu -- the rows either side of the hot path as coalesced
// gather kernels: burst extraction (lib/extract_burst_cc_impl.cc:117-242) fed by a descriptor array instead of
// stream tags, and the byte-wide symbol mapping (python/pygfdm/symbolmapping.py:27-47, gr-digital's
// chunks_to_symbols / constellation decoder).  remove_prefix_cc (lib/remove_prefix_cc_impl.cc:84-115) is the
// strided copy of remove_cp_kernel (stage_kernels.cu) with cp = offset.
// All of them are HBM-bound: one thread per (group of) output element(s), 128-bit accesses where the layout
// allows, grid-stride loops sized to the SM count.
#include "engine.h"

namespace gfdm {

static constexpr unsigned TH = 256;

__device__ __forceinline__ cpx ldg_stream_cpx(const cpx* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

static unsigned grid_for(size_t items, unsigned threads)
{
    // enough CTAs to fill the device (148 SMs x 8 resident CTAs of 256 threads), grid-stride beyond that
    const size_t want = (items + threads - 1) / threads;
    const size_t cap = (size_t)148 * 8;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// extract_burst_cc: burst b, sample i:  out[b][i] = scale_b * in[start_b + i] * inc_b^i   (zeros where start_b + i < 0)
// The rotation inc^i is evaluated directly, not by VOLK's recursive fp32 rotator (no drift, any sample can be computed
// independently): a warp owns 128 consecutive samples of a burst, lane l takes samples l, l+32, l+64, l+96 (every
// access is a full 256-byte line), computes one sincos in double at its first sample and steps by inc^32 in double.
__global__ void __launch_bounds__(TH) extract_burst_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                           const BurstDesc* __restrict__ desc, int burst_len, int cfo,
                                                           int n_bursts, long long n_in)
{
    const int tiles = (burst_len + 127) / 128; // warp tiles per burst
    const size_t total = (size_t)n_bursts * tiles;
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
        const int b = (int)(w / tiles);
        const int i0 = (int)(w - (size_t)b * tiles) * 128 + lane;
        const BurstDesc d = desc[b];
        double pr = 1.0, pi = 0.0;
        if (cfo) sincos(d.angle * (double)i0, &pi, &pr);
        cpx* o = out + (size_t)b * burst_len;
        cpx x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + 32 * j;
            const long long src = d.start + i;
            x[j] = (i < burst_len && src >= 0 && src < n_in) ? ldg_stream_cpx(in + src) : cmake(0.f, 0.f); // never outside the window
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + 32 * j;
            cpx v = cmake(__fmul_rn(x[j].x, d.scale), __fmul_rn(x[j].y, d.scale)); // volk_32f_s32f_multiply_32f
            if (cfo) {
                const double re = (double)v.x * pr - (double)v.y * pi;
                const double im = (double)v.x * pi + (double)v.y * pr;
                v = cmake((float)re, (float)im);
                const double nr = pr * d.inc32_re - pi * d.inc32_im;
                pi = pr * d.inc32_im + pi * d.inc32_re;
                pr = nr;
            }
            if (i < burst_len) o[i] = v;
        }
    }
}
// get_phase_rotation (lib/extract_burst_cc_impl.cc:88-96): inc = conj(pr)/|pr| rounded to complex<float>; the
// extraction kernel rotates by the angle of that number.  One thread per burst (keeps the trigonometry off the host,
// which would otherwise pace the calls).
__global__ void __launch_bounds__(TH) burst_prepare_kernel(BurstDesc* __restrict__ desc, int n_bursts)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bursts) return;
    const float pr = desc[b].pr_re, pi = desc[b].pr_im;
    const double scale = 1.0 / (double)hypotf(pr, pi);
    const float ir = (float)(scale * (double)pr), ii = (float)(-1.0f * scale * (double)pi);
    const double angle = atan2((double)ii, (double)ir);
    desc[b].angle = angle;
    sincos(32.0 * angle, &desc[b].inc32_im, &desc[b].inc32_re);
}
int launch_extract_burst(cpx* out, const cpx* in, BurstDesc* desc, int burst_len, bool cfo, int n_bursts, long long n_in,
                         cudaStream_t s)
{
    if (n_bursts <= 0) return 0;
    if (cfo) burst_prepare_kernel<<<blocks_for((size_t)n_bursts, TH), TH, 0, s>>>(desc, n_bursts);
    const size_t total = (size_t)n_bursts * ((burst_len + 127) / 128) * 32; // threads
    extract_burst_kernel<<<grid_for(total, TH), TH, 0, s>>>(out, in, desc, burst_len, cfo ? 1 : 0, n_bursts, n_in);
    GFDM_CUDA_CHECK(cudaGetLastError());
    return cfo ? 2 : 1;
}

// ---------------------------------------------------------------------------------------------
// symbol mapping.  The constellation (<= 256 points) sits in shared memory.
// Both directions walk the symbols in tiles of TILE = 16*TH: the byte side of a tile moves as one 16-byte access per
// thread (when the tile is 16-byte aligned in memory), the complex side as 16 accesses per thread with consecutive
// lanes on consecutive symbols (256-byte lines, 16 independent requests in flight), shared memory in between.
static constexpr int TILE = 16 * TH;

__global__ void __launch_bounds__(TH) map_chunks_kernel(cpx* __restrict__ out, const unsigned char* __restrict__ chunks,
                                                        const cpx* __restrict__ points, int n_points, size_t n)
{
    __shared__ cpx pts[256];
    __shared__ __align__(16) unsigned char tile[TILE];
    for (int i = threadIdx.x; i < 256; i += TH) pts[i] = i < n_points ? points[i] : cmake(0.f, 0.f);
    const bool vec = (reinterpret_cast<uintptr_t>(chunks) & 15) == 0; // tiles start at multiples of 4096
    const size_t n_tiles = (n + TILE - 1) / TILE;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = t * TILE;
        const int cnt = (int)(n - base < (size_t)TILE ? n - base : (size_t)TILE);
        __syncthreads(); // the previous tile has been consumed (and pts is in place)
        if (vec && cnt == TILE) {
            reinterpret_cast<uint4*>(tile)[threadIdx.x] = reinterpret_cast<const uint4*>(chunks + base)[threadIdx.x];
        } else {
            for (int i = threadIdx.x; i < cnt; i += TH) tile[i] = chunks[base + i];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = threadIdx.x + j * TH;
            if (i < cnt) out[base + i] = pts[tile[i]];
        }
    }
}
void launch_map_chunks(cpx* out, const unsigned char* chunks, const cpx* points, int n_points, size_t n, cudaStream_t s)
{
    if (!n) return;
    map_chunks_kernel<<<grid_for((n + 15) / 16, TH), TH, 0, s>>>(out, chunks, points, n_points, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(TH) decide_chunks_kernel(unsigned char* __restrict__ chunks, const cpx* __restrict__ in,
                                                           const cpx* __restrict__ points, int n_points, int rule,
                                                           const __grid_constant__ DecideGrid grid, size_t n)
{
    __shared__ cpx pts[256];
    __shared__ unsigned char lut[64];
    __shared__ __align__(16) unsigned char tile[TILE];
    for (int i = threadIdx.x; i < 256; i += TH) pts[i] = i < n_points ? points[i] : cmake(0.f, 0.f);
    if (threadIdx.x < 64) lut[threadIdx.x] = grid.lut[threadIdx.x];
    const bool vec = (reinterpret_cast<uintptr_t>(chunks) & 15) == 0;
    const size_t n_tiles = (n + TILE - 1) / TILE;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = t * TILE;
        const int cnt = (int)(n - base < (size_t)TILE ? n - base : (size_t)TILE);
        cpx x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = threadIdx.x + j * TH;
            x[j] = i < cnt ? ldg_stream_cpx(in + base + i) : cmake(0.f, 0.f);
        }
        __syncthreads(); // the previous tile has been written out (and pts / lut are in place)
#pragma unroll
        for (int j = 0; j < 16; ++j)
            tile[threadIdx.x + j * TH] = (unsigned char)decide_symbol_grid(x[j], pts, n_points, rule, grid, lut);
        __syncthreads();
        if (vec && cnt == TILE) {
            reinterpret_cast<uint4*>(chunks + base)[threadIdx.x] = reinterpret_cast<const uint4*>(tile)[threadIdx.x];
        } else {
            for (int i = threadIdx.x; i < cnt; i += TH) chunks[base + i] = tile[i];
        }
    }
}
void launch_decide_chunks(unsigned char* chunks, const cpx* in, const cpx* points, int n_points, int rule,
                          const DecideGrid& grid, size_t n, cudaStream_t s)
{
    if (!n) return;
    decide_chunks_kernel<<<grid_for((n + 15) / 16, TH), TH, 0, s>>>(chunks, in, points, n_points, rule, grid, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// pack_bits (MSB first) + constellation lookup: one thread per symbol
__global__ void __launch_bounds__(TH) bits2symbols_kernel(cpx* __restrict__ out, const unsigned char* __restrict__ bits,
                                                          const cpx* __restrict__ points, int n_points, int bps, size_t n)
{
    __shared__ cpx pts[256];
    for (int i = threadIdx.x; i < 256; i += TH) pts[i] = i < n_points ? points[i] : cmake(0.f, 0.f);
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int v = 0;
        for (int b = 0; b < bps; ++b) v = (v << 1) | (bits[i * bps + b] & 1);
        out[i] = pts[v];
    }
}
void launch_bits2symbols(cpx* out, const unsigned char* bits, const cpx* points, int n_points, int bps, size_t n,
                         cudaStream_t s)
{
    if (!n) return;
    bits2symbols_kernel<<<grid_for(n, TH), TH, 0, s>>>(out, bits, points, n_points, bps, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// decision + unpackbits (MSB first): one thread per symbol
__global__ void __launch_bounds__(TH) symbols2bits_kernel(unsigned char* __restrict__ bits, const cpx* __restrict__ in,
                                                          const cpx* __restrict__ points, int n_points, int rule, int bps,
                                                          const __grid_constant__ DecideGrid grid, size_t n)
{
    __shared__ cpx pts[256];
    __shared__ unsigned char lut[64];
    for (int i = threadIdx.x; i < 256; i += TH) pts[i] = i < n_points ? points[i] : cmake(0.f, 0.f);
    if (threadIdx.x < 64) lut[threadIdx.x] = grid.lut[threadIdx.x];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int v = decide_symbol_grid(in[i], pts, n_points, rule, grid, lut);
        for (int b = 0; b < bps; ++b) bits[i * bps + b] = (unsigned char)((v >> (bps - 1 - b)) & 1);
    }
}
void launch_symbols2bits(unsigned char* bits, const cpx* in, const cpx* points, int n_points, int rule, int bps,
                         const DecideGrid& grid, size_t n, cudaStream_t s)
{
    if (!n) return;
    symbols2bits_kernel<<<grid_for(n, TH), TH, 0, s>>>(bits, in, points, n_points, rule, bps, grid, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// demap_from_resources (lib/resource_mapper_kernel_cc.cc:136-162) on a grid of chunks
__global__ void __launch_bounds__(TH) demap_chunks_kernel(unsigned char* __restrict__ out, const unsigned char* __restrict__ in,
                                                          const int* __restrict__ smap, int M, int K, int A, int per_timeslot,
                                                          size_t n_out, size_t total)
{
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t f = gid / n_out;
        const size_t i = gid - f * n_out;
        int a, t;
        if (per_timeslot) {
            t = (int)(i / A);
            a = (int)(i - (size_t)t * A);
        } else {
            a = (int)(i / M);
            t = (int)(i - (size_t)a * M);
        }
        out[gid] = in[f * (size_t)M * K + (size_t)M * smap[a] + t];
    }
}
void launch_demap_chunks(unsigned char* out, const unsigned char* in, const int* smap, int M, int K, int A, bool per_timeslot,
                         size_t n_out, size_t frames, cudaStream_t s)
{
    const size_t total = frames * n_out;
    if (!total) return;
    demap_chunks_kernel<<<grid_for(total, TH), TH, 0, s>>>(out, in, smap, M, K, A, per_timeslot ? 1 : 0, n_out, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// sc16 host sample format (include/gfdm_b200.h, "sc16 sample format"): complex64 <-> interleaved int16 I/Q.
// Two samples (16 bytes of complex64, 8 bytes of sc16) per thread and step; HBM-bound streaming passes that run behind
// / in front of the path kernels of a HOST batch, where the PCIe leg is 100x slower than they are.
__device__ __forceinline__ short quant_sc16(float v, float scale)
{
    const int q = __float2int_rn(v * scale); // round to nearest even; NaN -> 0, +-inf saturate in the conversion
    return (short)max(-32768, min(32767, q));
}
__global__ void __launch_bounds__(TH) cf32_to_sc16_kernel(short2* __restrict__ out, const cpx* __restrict__ in, float scale, size_t n)
{
    const size_t pairs = n / 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(in + 2 * p));
        short4 q;
        q.x = quant_sc16(v.x, scale); q.y = quant_sc16(v.y, scale); q.z = quant_sc16(v.z, scale); q.w = quant_sc16(v.w, scale);
        reinterpret_cast<short4*>(out)[p] = q;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const cpx v = in[n - 1];
        out[n - 1] = make_short2(quant_sc16(v.x, scale), quant_sc16(v.y, scale));
    }
}
__global__ void __launch_bounds__(TH) sc16_to_cf32_kernel(cpx* __restrict__ out, const short2* __restrict__ in, float inv_scale, size_t n)
{
    const size_t pairs = n / 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
        const short4 q = __ldg(reinterpret_cast<const short4*>(in) + p);
        reinterpret_cast<float4*>(out)[p] =
            make_float4((float)q.x * inv_scale, (float)q.y * inv_scale, (float)q.z * inv_scale, (float)q.w * inv_scale);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const short2 q = in[n - 1];
        out[n - 1] = cmake((float)q.x * inv_scale, (float)q.y * inv_scale);
    }
}
// both need 16-byte aligned complex64 and 8-byte aligned sc16 arrays (the library's own staging buffers, or frames of
// an even number of samples behind an aligned base)
void launch_cf32_to_sc16(short* out, const cpx* in, float scale, size_t n, cudaStream_t s)
{
    if (!n) return;
    cf32_to_sc16_kernel<<<grid_for((n + 1) / 2, TH), TH, 0, s>>>(reinterpret_cast<short2*>(out), in, scale, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}
void launch_sc16_to_cf32(cpx* out, const short* in, float scale, size_t n, cudaStream_t s)
{
    if (!n) return;
    sc16_to_cf32_kernel<<<grid_for((n + 1) / 2, TH), TH, 0, s>>>(out, reinterpret_cast<const short2*>(in), 1.0f / scale, n);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// short_burst_shaper (lib/short_burst_shaper_impl.cc:161-182): out = [pre zeros | in * scale | post zeros] per burst;
// the multiply is volk_32fc_s32fc_multiply_32fc's plain complex product (unfused, reference operand order).
__global__ void __launch_bounds__(TH) burst_shape_kernel(cpx* __restrict__ out, const cpx* __restrict__ in, int len, int pre,
                                                         int post, cpx scale, size_t total)
{
    const size_t row = (size_t)pre + len + post;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (size_t)gridDim.x * blockDim.x) {
        const size_t b = gid / row;
        const int i = (int)(gid - b * row) - pre;
        cpx v = cmake(0.f, 0.f);
        if (i >= 0 && i < len) v = cmul_rn(ldg_stream_cpx(in + b * (size_t)len + i), scale);
        out[gid] = v;
    }
}
void launch_burst_shape(cpx* out, const cpx* in, int len, int pre, int post, cpx scale, size_t n_bursts, cudaStream_t s)
{
    const size_t total = n_bursts * ((size_t)pre + len + post);
    if (!total) return;
    burst_shape_kernel<<<grid_for(total, TH), TH, 0, s>>>(out, in, len, pre, post, scale, total);
    GFDM_CUDA_CHECK(cudaGetLastError());
}

} // namespace gfdm
