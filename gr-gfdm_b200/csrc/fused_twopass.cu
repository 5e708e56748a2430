// fused_twopass.cu -- modulator and receiver for frames LARGER than shared memory (K = 2*K1;
// BASELINE config 5: K = 2048, M = 15, N = 30720 complex = 245,760 B > 227 KB per SM).
//
// One CTA still owns a whole frame, but visits it in two passes.  The K-point transform over the
// subcarrier index is split by decimation in frequency (tools/fused_math_check.py, "two-pass"):
// pass p in {0,1} computes the K1-point transform whose results are the outputs of parity p, so a pass
// needs a K1-row buffer only (the C3 kernel's shared-memory budget).  HBM traffic stays at the algorithmic
// 16*N bytes per frame.  Two generations of kernels live here:
//   * fused_mod2p_kernel / fused_rx2p_kernel (shipped): the frame is read ONCE; what pass 1 needs (the M-point results of
//     the second parity) and what pass 0 produced for a location pass 1 completes (even samples / even records) is
//     thread-private and waits in TENSOR MEMORY (4 x 2M columns per thread), so pass 1 has no loads, no transforms of
//     the inputs and writes whole 16-byte pairs / one contiguous bulk store per item; the table columns, which used to fill
//     the tensor memory, are read from L2 one step ahead of their use.
//   * fused_mod2_kernel / fused_rx2_kernel (round 1; GFDM_MOD2_SCRATCH=1 / GFDM_RX2_REREAD=1 select them for A/B runs):
//     both passes read the frame (pass 1 from L2) and re-derive the M-point transforms, the modulator's even samples
//     wait in an L2-resident scratch, the receiver writes M-element runs with M-element gaps.
//
//   modulator (lib/modulator_kernel_cc.cc:98-141), W = e^{+j2pi/K}
//     E^p_b'        = (d_b' + (-1)^p d_{b'+K1}) W^{p b'}              radix-2 butterfly on the raw symbols
//     Z_m[2n'+p]    = IFFT_K1 over b' of FFT_M(E^p_b')[m]             stage A, row FFTs (stage B)
//     x[2n'+p+K n2] = IFFT_M over m of C_tx[m][2n'+p] Z_m[2n'+p]      stage C
//   receiver (lib/receiver_kernel_cc.cc:165-225,301-307), W = e^{-j2pi/K}
//     U_n1          = FFT_M over n2 of x[n1 + K n2]                   both halves n1 = n', n'+K1 per pass
//     B^p_m[n']     = Tlo^p[m][n'] U_n'[m] + Thi^p[m][n'] U_{n'+K1}[m]
//                     Tlo^p = C_rx[m][n'] W^{p n'},  Thi^p = (-1)^p C_rx[m][n'+K1] W^{p n'}
//     R_{2k'+p}[m]  = FFT_K1 over n' of B^p_m[n'];   y_{2k'+p} = IFFT_M(R_{2k'+p}) / M
//
// Shared memory per CTA (same carve-up as fused_modem.cu): row buffer `buf` (M rows of K1, padded),
// the row-FFT twiddles and the staging region P.  Input moves by cp.async.bulk (TMA 1D) and is prefetched
// one pass ahead wherever a region is free; the pieces that cannot be resident early are fetched one
// compute step ahead.
// Equalisation (receiver_kernel_cc.cc:309-320) at this shape, overlap 2: fused_rx2p_kernel<S, true>.  With Y/H between FFT and
// filter the taps cannot be folded into the table, and R_k = G1 Yeq[k-1] + G0 Yeq[k] mixes a block of each parity, i.e. of
// each pass: pass 0 divides its (even) blocks by the channel and parks them in tensor memory, pass 1 divides the odd ones,
// gets block 2k'-1 from the previous lane and combines both parities in registers -- 24 N bytes of HBM traffic per frame.
// The interference-cancelling entry points still use the staged path of api.cu at this shape.
#include "fused.h"
#include "fused_dev.cuh"

#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace gfdm {

// ----------------------------------------------------------------------------------------
// Loop-invariant table columns of a thread in tensor memory (see fused_dev.cuh): 4 sets of 2M 32-bit columns per
// thread, T = 512 threads -> 4 thread groups per lane quarter -> 4 * 4 * 2M = 480 of the 512 columns.  The two-pass
// kernels have no MMA and run one CTA per SM, so the whole tensor memory is theirs; the table reads leave L2 and the
// long-scoreboard stalls they caused in the step loops go with them.
template <class S>
struct Tmem4 {
    static constexpr int SET = 2 * S::M;
    static constexpr int PER_THREAD = 4 * SET;
    static constexpr int COLS = 512;
    static_assert(((S::T / 32 + 3) / 4) * PER_THREAD <= COLS, "table columns do not fit in tensor memory");
    // col(set, m) of this thread <- src(set, m); returns the allocation base and this thread's first column
    template <class Src>
    static __device__ __forceinline__ void setup(uint32_t* slot, int tid, uint32_t& base, uint32_t& mine, Src src)
    {
        if (tid < 32) tmem_alloc(slot, COLS);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        base = *slot;
        mine = base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)(tid >> 7) * PER_THREAD;
#pragma unroll
        for (int set = 0; set < 4; ++set) {
            float tf[SET];
#pragma unroll
            for (int m = 0; m < S::M; ++m) {
                const cpx c = src(set, m);
                tf[2 * m] = c.x;
                tf[2 * m + 1] = c.y;
            }
            tmem_st<SET>(mine + set * SET, tf);
        }
        tmem_wait_st();
    }
    static __device__ __forceinline__ void load(cpx (&tc)[S::M], uint32_t mine, int set)
    {
        float tf[SET];
        tmem_ld<SET>(tf, mine + set * SET);
        tmem_wait_ld();
#pragma unroll
        for (int m = 0; m < S::M; ++m) tc[m] = cmake(tf[2 * m], tf[2 * m + 1]);
    }
    static __device__ __forceinline__ void release(uint32_t base, int tid)
    {
        tmem_fence_before_sync();
        __syncthreads();
        if (tid < 32) tmem_dealloc(base, COLS);
    }
};

// ----------------------------------------------------------------------------------------
// modulator.  in/out: [n_frames][N], N = M*2*K1; tableP: [2][M][K1] = C_tx[m][2n'+p]; tw: row-FFT
// twiddles of the K1-point transform; w2: W^{b'} = e^{+j2pi b'/K}, b' < K1.
// Pieces of the staged frame (T*M elements each): LO0 = k < T, LO1 = T <= k < K1, HI0 = K1 <= k < K1+T,
// HI1 = k >= K1+T.  P holds LO0 and the first HA elements of HI1, buf holds LO1|HI0 (contiguous in
// HBM: one copy); the rest of HI1 lands on LO0's place as soon as step 0 has consumed it.
template <class S>
__global__ void __launch_bounds__(S::T, 1) fused_mod2_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                            const cpx* __restrict__ tableP,
                                                            const cpx* __restrict__ tw, const cpx* __restrict__ w2,
                                                            cpx* __restrict__ scratch, int n_frames)
{
    constexpr int M = S::M, K1 = S::K, K = 2 * K1, N = M * K, T = S::T, RS = S::RS;
    static_assert(S::F == 1 && S::IPT == 2 && S::TWO_PASS, "two-pass kernels: one half-frame per CTA pass, two items per thread");
    constexpr int PIECE = T * M;                                      // elements of one piece
    constexpr int HA = ((S::P_ELEMS - PIECE) / (2 * M)) * (2 * M);    // head of HI1 that fits beside LO0 (whole record pairs)
    constexpr int HB = PIECE - HA;                                    // late part of HI1
    static_assert(HA > 0 && HB > 0 && HB <= PIECE && 2 * PIECE <= S::BUF_ELEMS, "staging does not fit");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* pre = tw_s + S::TW_ELEMS + S::TBL_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS);
    uint64_t* bar_p = bars;     // LO0 + head of HI1 (region P)
    uint64_t* bar_r = bars + 1; // LO1 | HI0 (region buf)
    uint64_t* bar_l = bars + 2; // late part of HI1
    const int tid = threadIdx.x;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_r, 1);
        mbar_init(bar_l, 1);
    }
    const cpx wj0 = w2[tid], wj1 = w2[tid + T];
    // table columns C_tx[m][2n'+p] of this thread's two items, both passes: set = 2*p + j
    uint32_t tmem_base = 0, tmem_mine = 0;
    Tmem4<S>::setup(reinterpret_cast<uint32_t*>(bars + 3), tid, tmem_base, tmem_mine, [&](int set, int m) {
        return ldg_nc(tableP + ((size_t)(set >> 1) * M + m) * K1 + tid + (set & 1) * T);
    });
    // the frame is read twice (pass 0 keeps it in L2, pass 1 releases it); the even samples of pass 0
    // wait in this CTA's L2-resident scratch so that pass 1 can write whole 16-byte sample pairs
    // (the policies are re-created where they are used: one instruction each, no live registers)
    __syncthreads();

    // `first`: pass 0 of the frame reads it from HBM (keep it in L2); pass 1 reads it again (then drop it)
    auto load_p = [&](int g, bool first) {
        const cpx* f = in + (size_t)g * N;
        const uint64_t pol = first ? l2_policy_evict_last() : l2_policy_evict_first();
        mbar_expect_tx(bar_p, (uint32_t)(PIECE + HA) * sizeof(cpx));
        bulk_load_hint(pre, f, PIECE * sizeof(cpx), bar_p, pol);
        bulk_load_hint(pre + PIECE, f + 3 * PIECE, HA * sizeof(cpx), bar_p, pol);
        if (first) bulk_prefetch_l2(f + 3 * PIECE + HA, HB * sizeof(cpx)); // the late piece has one step to arrive
    };
    auto load_r = [&](int g, bool first) {
        mbar_expect_tx(bar_r, (uint32_t)(2 * PIECE) * sizeof(cpx));
        bulk_load_hint(buf, in + (size_t)g * N + PIECE, 2 * PIECE * sizeof(cpx), bar_r,
                       first ? l2_policy_evict_last() : l2_policy_evict_first());
    };
    auto load_l = [&](int g, bool first) {
        mbar_expect_tx(bar_l, (uint32_t)HB * sizeof(cpx));
        bulk_load_hint(pre, in + (size_t)g * N + 3 * PIECE + HA, HB * sizeof(cpx), bar_l,
                       first ? l2_policy_evict_last() : l2_policy_evict_first());
    };

    int g = blockIdx.x;
    if (tid == 0 && g < n_frames) {
        load_p(g, true);
        load_r(g, true);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_frames; g += gridDim.x) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            const int gn = p == 0 ? g : g + gridDim.x; // frame of the next pass
            const bool has_next = gn < n_frames;
            const float sgn = p ? -1.f : 1.f;
            mbar_wait(bar_p, phase);
            mbar_wait(bar_r, phase);
            STAGE_MARK(0) // wait for the bulk loads
            cpx v[2][M];
            // ---- step 0: b' = tid, lo record in P (LO0), hi record in buf (HI0)
            {
                const cpx* lo = pre + tid * M;
                const cpx* hi = buf + PIECE + tid * M;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const cpx a = lo[m], b = hi[m];
                    v[0][m] = cmake(fmaf(sgn, b.x, a.x), fmaf(sgn, b.y, a.y));
                }
            }
            __syncthreads(); // LO0 consumed: its place takes the late part of HI1
            if (tid == 0) {
                fence_proxy_async();
                load_l(g, p == 0);
            }
            if (p) {
#pragma unroll
                for (int m = 0; m < M; ++m) v[0][m] = cmul(v[0][m], wj0);
            }
            rf::FFTN<M, -1>::run(v[0]);
            STAGE_MARK(1) // step 0
            mbar_wait(bar_l, phase);
            // ---- step 1: b' = T + tid, lo record in buf (LO1), hi record in P (head after LO0, late part at 0)
            {
                const cpx* lo = buf + tid * M;
                const int e = tid * M;
                const cpx* hi = e < HA ? pre + PIECE + e : pre + (e - HA);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const cpx a = lo[m], b = hi[m];
                    v[1][m] = cmake(fmaf(sgn, b.x, a.x), fmaf(sgn, b.y, a.y));
                }
            }
            __syncthreads(); // staging fully consumed: P takes the next pass, buf takes the rows
            if (tid == 0 && has_next) {
                fence_proxy_async();
                load_p(gn, p == 1);
            }
            if (p) {
#pragma unroll
                for (int m = 0; m < M; ++m) v[1][m] = cmul(v[1][m], wj1);
            }
            rf::FFTN<M, -1>::run(v[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                cpx* dst = buf + S::swz(tid + j * T);
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m * RS] = v[j][m];
            }
            __syncthreads();
            STAGE_MARK(2) // step 1 + row writes
            // ---- stage B: K1-point inverse FFT of every row
            row_fft<S, +1>(buf, tw_s, tid);
            STAGE_MARK(3) // row FFT
            __syncthreads();
            cpx tc[M];
            Tmem4<S>::load(tc, tmem_mine, 2 * p);
            // ---- stage C: column n' of all rows -> registers
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const cpx* src = buf + tid + j * T;
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = src[m * RS];
            }
            __syncthreads(); // rows are dead: buf takes LO1|HI0 of the next pass
            if (tid == 0 && has_next) {
                fence_proxy_async();
                load_r(gn, p == 1);
            }
            STAGE_MARK(4) // stage C reads
            cpx* sc = scratch + (size_t)blockIdx.x * (M * K1) + tid;
#pragma unroll
            for (int m = 0; m < M; ++m) v[0][m] = cmul(v[0][m], tc[m]);
            Tmem4<S>::load(tc, tmem_mine, 2 * p + 1);
            rf::FFTN<M, +1>::run(v[0]);
            if (p == 0) {
                // even samples -> scratch [n2][n'] (whole lines, kept in L2 until pass 1 collects them)
                const uint64_t pol_keep = l2_policy_evict_last();
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) stg_hint(sc + n2 * K1, v[0][n2], pol_keep);
#pragma unroll
                for (int m = 0; m < M; ++m) v[1][m] = cmul(v[1][m], tc[m]);
                rf::FFTN<M, +1>::run(v[1]);
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) stg_hint(sc + T + n2 * K1, v[1][n2], pol_keep);
            } else {
                // odd samples join the even ones: one 16-byte store per sample pair.  The even samples of item 0
                // are fetched while item 1's M-point IFFT runs, those of item 1 while item 0 is stored.
                const uint64_t pol_drop = l2_policy_evict_first();
#pragma unroll
                for (int m = 0; m < M; ++m) v[1][m] = cmul(v[1][m], tc[m]);
                cpx ev[M];
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) ev[n2] = ldg_hint(sc + n2 * K1, pol_drop);
                rf::FFTN<M, +1>::run(v[1]);
                cpx* dst = out + (size_t)g * N + 2 * tid;
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) {
                    stg_stream4(dst + (size_t)n2 * K, ev[n2], v[0][n2]);
                    ev[n2] = ldg_hint(sc + T + n2 * K1, pol_drop);
                }
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) stg_stream4(dst + 2 * T + (size_t)n2 * K, ev[n2], v[1][n2]);
            }
            STAGE_MARK(5) // stage C compute + stores
            phase ^= 1;
        }
    }
    Tmem4<S>::release(tmem_base, tid);
}

// ----------------------------------------------------------------------------------------
// Tensor memory as a per-thread PARKING area (the two-pass modulator below): 4 sets of 2M columns per thread, nothing
// stored at allocation.  Same addressing as Tmem4.
template <class S>
struct TmemPark {
    static constexpr int SET = 2 * S::M;
    static constexpr int PER_THREAD = 4 * SET;
    static constexpr int COLS = 512;
    static_assert(((S::T / 32 + 3) / 4) * PER_THREAD <= COLS, "parked sets do not fit in tensor memory");
    static __device__ __forceinline__ void alloc(uint32_t* slot, int tid, uint32_t& base, uint32_t& mine)
    {
        if (tid < 32) tmem_alloc(slot, COLS);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        base = *slot;
        mine = base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)(tid >> 7) * PER_THREAD;
    }
    static __device__ __forceinline__ void st(uint32_t mine, int set, const cpx (&v)[S::M])
    {
        float tf[SET];
#pragma unroll
        for (int m = 0; m < S::M; ++m) {
            tf[2 * m] = v[m].x;
            tf[2 * m + 1] = v[m].y;
        }
        tmem_st<SET>(mine + set * SET, tf);
        tmem_wait_st();
    }
    static __device__ __forceinline__ void ld(cpx (&v)[S::M], uint32_t mine, int set)
    {
        float tf[SET];
        tmem_ld<SET>(tf, mine + set * SET);
        tmem_wait_ld();
#pragma unroll
        for (int m = 0; m < S::M; ++m) v[m] = cmake(tf[2 * m], tf[2 * m + 1]);
    }
    static __device__ __forceinline__ void release(uint32_t base, int tid)
    {
        tmem_fence_before_sync();
        __syncthreads();
        if (tid < 32) tmem_dealloc(base, COLS);
    }
};

// ----------------------------------------------------------------------------------------
// modulator, second version: the frame is read ONCE and nothing goes through an L2 scratch.  Everything a thread has to
// carry from pass 0 to pass 1 is its own: the M-point transforms of the odd-parity inputs E^1_b' of its two subcarrier
// pairs (stage A computes both parities from one read of the symbols) and the even samples of its two output columns
// (pass 1 writes 16-byte sample pairs).  Both wait in tensor memory (4 x 2M = 120 of the 128 columns a thread can have
// at T = 512), so pass 1 has no loads, no M-point input transforms and no L2 round trip; the table columns, which used to
// occupy the tensor memory, are read from L2 (coalesced, issued before the column reads that hide their latency).
// Staging as in fused_mod2_kernel, but each region is filled once per frame: P (LO0 + head of HI1, then the late part
// of HI1) is free again after stage A and takes the NEXT frame at once; buf (LO1|HI0) is refilled when pass 1 has read
// its columns.
template <class S>
__global__ void __launch_bounds__(S::T, 1) fused_mod2p_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                             const cpx* __restrict__ tableP, const cpx* __restrict__ tw,
                                                             const cpx* __restrict__ w2, int n_frames)
{
    constexpr int M = S::M, K1 = S::K, K = 2 * K1, N = M * K, T = S::T, RS = S::RS;
    static_assert(S::F == 1 && S::IPT == 2 && S::TWO_PASS, "two-pass kernels: one half-frame per CTA pass, two items per thread");
    constexpr int PIECE = T * M;
    constexpr int HA = ((S::P_ELEMS - PIECE) / (2 * M)) * (2 * M);
    constexpr int HB = PIECE - HA;
    static_assert(HA > 0 && HB > 0 && HB <= PIECE && 2 * PIECE <= S::BUF_ELEMS, "staging does not fit");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* pre = tw_s + S::TW_ELEMS + S::TBL_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS);
    uint64_t* bar_p = bars;     // LO0 + head of HI1 (region P)
    uint64_t* bar_r = bars + 1; // HI0 (upper half of the staging in buf): step 0
    uint64_t* bar_l = bars + 2; // late part of HI1
    uint64_t* bar_q = bars + 3; // LO1 (lower half of the staging in buf): step 1
    const int tid = threadIdx.x;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_r, 1);
        mbar_init(bar_l, 1);
        mbar_init(bar_q, 1);
    }
    const cpx wj0 = w2[tid], wj1 = w2[tid + T];
    uint32_t tmem_base = 0, tmem_mine = 0;
    TmemPark<S>::alloc(reinterpret_cast<uint32_t*>(bars + 4), tid, tmem_base, tmem_mine);
    __syncthreads();

    auto load_p = [&](int g) {
        const cpx* f = in + (size_t)g * N;
        const uint64_t pol = l2_policy_evict_first();
        mbar_expect_tx(bar_p, (uint32_t)(PIECE + HA) * sizeof(cpx));
        bulk_load_hint(pre, f, PIECE * sizeof(cpx), bar_p, pol);
        bulk_load_hint(pre + PIECE, f + 3 * PIECE, HA * sizeof(cpx), bar_p, pol);
        bulk_prefetch_l2(f + 3 * PIECE + HA, HB * sizeof(cpx)); // the late piece and LO1|HI0 are fetched later: pull them
        bulk_prefetch_l2(f + PIECE, 2 * PIECE * sizeof(cpx));   // into L2 now
    };
    auto load_r = [&](int g) {
        // two copies: step 0 needs HI0 only, LO1 may land while it runs
        mbar_expect_tx(bar_r, (uint32_t)PIECE * sizeof(cpx));
        bulk_load_hint(buf + PIECE, in + (size_t)g * N + 2 * PIECE, PIECE * sizeof(cpx), bar_r, l2_policy_evict_first());
        mbar_expect_tx(bar_q, (uint32_t)PIECE * sizeof(cpx));
        bulk_load_hint(buf, in + (size_t)g * N + PIECE, PIECE * sizeof(cpx), bar_q, l2_policy_evict_first());
    };
    auto load_l = [&](int g) {
        mbar_expect_tx(bar_l, (uint32_t)HB * sizeof(cpx));
        bulk_load_hint(pre, in + (size_t)g * N + 3 * PIECE + HA, HB * sizeof(cpx), bar_l, l2_policy_evict_first());
    };

    int g = blockIdx.x;
    if (tid == 0 && g < n_frames) {
        load_p(g);
        load_r(g);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_frames; g += gridDim.x) {
        const int gn = g + gridDim.x;
        const bool has_next = gn < n_frames;
        mbar_wait(bar_p, phase);
        mbar_wait(bar_r, phase);
        STAGE_MARK(0) // wait for the bulk loads
        {
            cpx v[M], u[M];
            // ---- step 0: b' = tid, lo record in P (LO0), hi record in buf (HI0); both parities of the radix-2 butterfly
            {
                const cpx* lo = pre + tid * M;
                const cpx* hi = buf + PIECE + tid * M;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const cpx a = lo[m], b = hi[m];
                    v[m] = cadd(a, b);
                    u[m] = cmul(csub(a, b), wj0);
                }
            }
            __syncthreads(); // LO0 consumed: its place takes the late part of HI1
            if (tid == 0) {
                fence_proxy_async();
                load_l(g);
            }
            rf::FFTN<M, -1>::run(u);
            TmemPark<S>::st(tmem_mine, 0, u); // E^1 of item 0 waits for pass 1
            rf::FFTN<M, -1>::run(v);
            TmemPark<S>::st(tmem_mine, 2, v); // E^0 of item 0 waits for the row buffer (still staging) in an output slot
            STAGE_MARK(1) // step 0
            mbar_wait(bar_l, phase);
            mbar_wait(bar_q, phase);
            // ---- step 1: b' = T + tid, lo record in buf (LO1), hi record in P (head after LO0, late part at 0)
            {
                const cpx* lo = buf + tid * M;
                const int e = tid * M;
                const cpx* hi = e < HA ? pre + PIECE + e : pre + (e - HA);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const cpx a = lo[m], b = hi[m];
                    v[m] = cadd(a, b);
                    u[m] = cmul(csub(a, b), wj1);
                }
            }
            __syncthreads(); // staging fully consumed: P takes the next frame, buf takes the rows
            if (tid == 0 && has_next) {
                fence_proxy_async();
                load_p(gn);
            }
            rf::FFTN<M, -1>::run(u);
            TmemPark<S>::st(tmem_mine, 1, u);
            rf::FFTN<M, -1>::run(v);
            {
                cpx* dst = buf + S::swz(tid + T);
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m * RS] = v[m];
            }
            TmemPark<S>::ld(u, tmem_mine, 2);
            {
                cpx* dst = buf + S::swz(tid);
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m * RS] = u[m];
            }
        }
        __syncthreads();
        STAGE_MARK(2) // step 1 + row writes
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            if (p == 1) {
                // the odd-parity rows come out of tensor memory
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    cpx u[M];
                    TmemPark<S>::ld(u, tmem_mine, j);
                    cpx* dst = buf + S::swz(tid + j * T);
#pragma unroll
                    for (int m = 0; m < M; ++m) dst[m * RS] = u[m];
                }
                __syncthreads();
                STAGE_MARK(6) // pass 1: rows out of tensor memory
            }
            // ---- stage B: K1-point inverse FFT of every row
            row_fft<S, +1>(buf, tw_s, tid);
            STAGE_MARK(3) // row FFT
            __syncthreads();
            STAGE_MARK(7) // barrier after the row FFT
            // ---- stage C: table column of item 0 (L2; in flight during the column reads), column n' of all rows
            cpx tc[M], c0[M], c1[M];
            const cpx* tp = tableP + (size_t)p * M * K1 + tid;
            {
                const cpx* src = buf + tid;
#pragma unroll
                for (int m = 0; m < M; ++m) c0[m] = src[m * RS];
#pragma unroll
                for (int m = 0; m < M; ++m) c1[m] = src[T + m * RS];
            }
#pragma unroll
            for (int m = 0; m < M; ++m) tc[m] = ldg_nc(tp + (size_t)m * K1); // in flight across the barrier
            STAGE_MARK(8) // table loads issued, column reads
            __syncthreads(); // rows are dead
            if (p == 1 && tid == 0 && has_next) {
                fence_proxy_async();
                load_r(gn); // buf takes LO1|HI0 of the next frame
            }
            STAGE_MARK(4) // stage C reads
#pragma unroll
            for (int m = 0; m < M; ++m) c0[m] = cmul(c0[m], tc[m]);
#pragma unroll
            for (int m = 0; m < M; ++m) tc[m] = ldg_nc(tp + T + (size_t)m * K1);
            rf::FFTN<M, +1>::run(c0);
            if (p == 0) {
                TmemPark<S>::st(tmem_mine, 2, c0); // even samples wait for their odd neighbours
#pragma unroll
                for (int m = 0; m < M; ++m) c1[m] = cmul(c1[m], tc[m]);
                rf::FFTN<M, +1>::run(c1);
                TmemPark<S>::st(tmem_mine, 3, c1);
            } else {
                cpx ev[M];
                TmemPark<S>::ld(ev, tmem_mine, 2);
                cpx* dst = out + (size_t)g * N + 2 * tid;
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) stg_stream4(dst + (size_t)n2 * K, ev[n2], c0[n2]);
#pragma unroll
                for (int m = 0; m < M; ++m) c1[m] = cmul(c1[m], tc[m]);
                rf::FFTN<M, +1>::run(c1);
                TmemPark<S>::ld(ev, tmem_mine, 3);
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) stg_stream4(dst + 2 * T + (size_t)n2 * K, ev[n2], c1[n2]);
            }
            if (p == 0) { STAGE_MARK(5) } else { STAGE_MARK(9) } // stage C compute; parking (pass 0) / stores (pass 1)
        }
        phase ^= 1;
    }
    TmemPark<S>::release(tmem_base, tid);
}

// ----------------------------------------------------------------------------------------
// receiver.  in: [n_frames][N] time samples; out: [n_frames][N]; mode 0: soft symbols, mode 1: R.
// tables: [2 (half)][M][K1] = C_rx[m][n' + half*K1]; w2: W^{n'}.  A pass has four steps (item j, half h): the thread's
// sample column n1 = tid + j*T + h*K1, i.e. M quarter rows of T samples, each moved by one bulk copy.
// Homes of the quarter rows (n2 = sample row) -- chosen so that every copy is issued at least one compute
// step before its data is needed, except step 1 of the next pass (issued when the output staging is done):
//   step 0 (j0,lo): P[n2*T]                                           issued after step 3 of the previous pass
//   step 1 (j0,hi): upper half of row-buffer row n2                   issued at the end of the previous pass
//   step 2 (j1,lo): n2 < NE: P[M*T + n2*T]   (issued after step 2 of the previous pass)
//                   n2 >= NE: P[(n2-NE)*T]   (issued after step 0 of this pass)
//   step 3 (j1,hi): n2 < NE: P[(M-NE)*T + n2*T]   (after step 0);  n2 >= NE: upper half of row n2-NE (after step 1)
// Item 0's rows occupy the lower halves of the padded rows, item 1's the upper halves.
template <class S>
__global__ void __launch_bounds__(S::T, 1) fused_rx2_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                           const cpx* __restrict__ tables,
                                                           const cpx* __restrict__ tw, const cpx* __restrict__ w2,
                                                           int mode, int n_frames)
{
    constexpr int M = S::M, K1 = S::K, K = 2 * K1, N = M * K, T = S::T, RS = S::RS;
    static_assert(S::F == 1 && S::IPT == 2 && S::TWO_PASS, "two-pass kernels: one half-frame per CTA pass, two items per thread");
    constexpr int UP = RS / 2;                          // first slot of the upper half of a padded row
    constexpr int NE = (S::P_ELEMS - M * T) / T;        // quarter rows of step 2 that fit beside step 0 in P
    static_assert(S::swz(T - 1) < UP && UP + T <= RS, "row halves");
    constexpr int N1A = (T * M - UP - T) / RS + 1;      // rows whose upper half lies inside item 0's output staging
    static_assert(NE >= 1 && NE < M && M <= T / 32 && M * T * 2 <= S::BUF_ELEMS, "staging does not fit");
    static_assert(N1A >= 1 && N1A < M && (N1A - 1) * RS + UP + T <= T * M && (M - NE - 1) * RS + UP + T <= S::BUF_ELEMS, "row halves");
    constexpr uint32_t QBYTES = T * sizeof(cpx);
    constexpr uint32_t STEP_BYTES = M * QBYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* pre = tw_s + S::TW_ELEMS + S::TBL_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS); // one per step
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_m = 1.0f / (float)M;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if (tid == 0)
        for (int s = 0; s < 4; ++s) mbar_init(bars + s, 1);
    // With a = C_rx[m][n'] U_n'[m] and b = C_rx[m][n'+K1] U_{n'+K1}[m] the two passes need the same table columns:
    // B^0 = a + b, B^1 = (a - b) W^{n'}.  They live in tensor memory: set = 2*j + h (item j, half h).
    uint32_t tmem_base = 0, tmem_mine = 0;
    Tmem4<S>::setup(reinterpret_cast<uint32_t*>(bars + 4), tid, tmem_base, tmem_mine, [&](int set, int m) {
        return ldg_nc(tables + ((size_t)(set & 1) * M + m) * K1 + tid + (set >> 1) * T);
    });
    const cpx wj[2] = { w2[tid], w2[tid + T] };
    __syncthreads();

    // quarter row n2 of step s of frame gg -> its home; issued by lane 0 of warp n2
    auto qsrc = [&](int gg, int s, int n2) {
        return in + (size_t)gg * N + (size_t)n2 * K + (s & 1) * K1 + (s >> 1) * T;
    };
    auto home = [&](int s, int n2) -> cpx* {
        switch (s) {
        case 0: return pre + n2 * T;
        case 1: return buf + n2 * RS + UP;
        case 2: return n2 < NE ? pre + (M + n2) * T : pre + (n2 - NE) * T;
        default: return n2 < NE ? pre + (M - NE + n2) * T : buf + (n2 - NE) * RS + UP;
        }
    };
    // issue the copies [n_lo, n_hi) of step s; `arm`: first issue point of this step's phase
    // `first`: pass 0 of frame gg (from HBM, keep in L2 for pass 1); otherwise the second and last read
    auto issue = [&](int gg, int s, int n_lo, int n_hi, bool arm, bool first) {
        if (arm && tid == 0) mbar_expect_tx(bars + s, STEP_BYTES);
        if (lane == 0 && warp >= n_lo && warp < n_hi) {
            fence_proxy_async();
            bulk_load_hint(home(s, warp), qsrc(gg, s, warp), QBYTES, bars + s,
                           first ? l2_policy_evict_last() : l2_policy_evict_first());
        }
    };
    // copy-out index math: element i = tid + q*T of a staged item is element e of record r
    const int r0 = tid / M, e0 = tid - r0 * M;

    int g = blockIdx.x;
    if (g < n_frames) {
        issue(g, 0, 0, M, true, true);
        issue(g, 1, 0, M, true, true);
        issue(g, 2, 0, NE, true, true);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_frames; g += gridDim.x) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            const int gn = p == 0 ? g : g + gridDim.x;
            const bool has_next = gn < n_frames;
            cpx v[2][M];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int j = s >> 1, h = s & 1;
                cpx tc[M], x[M];
                mbar_wait(bars + s, phase);
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) x[n2] = home(s, n2)[tid];
                __syncthreads(); // step s consumed
                if (s == 0) {
                    issue(g, 2, NE, M, false, p == 0);
                    issue(g, 3, 0, NE, true, p == 0);
                } else if (s == 1) {
                    issue(g, 3, NE, M, false, p == 0);
                } else if (s == 2) {
                    if (has_next) issue(gn, 2, 0, NE, true, p == 1);
                } else {
                    if (has_next) issue(gn, 0, 0, M, true, p == 1);
                }
                rf::FFTN<M, -1>::run(x);
                Tmem4<S>::load(tc, tmem_mine, s);
                if (h == 0) {
#pragma unroll
                    for (int m = 0; m < M; ++m) v[j][m] = cmul(x[m], tc[m]);
                } else {
                    if (p == 0) {
#pragma unroll
                        for (int m = 0; m < M; ++m) v[j][m] = cfma(x[m], tc[m], v[j][m]);
                    } else {
#pragma unroll
                        for (int m = 0; m < M; ++m) v[j][m] = cmul(csub(v[j][m], cmul(x[m], tc[m])), wj[j]);
                    }
                    cpx* dst = buf + S::swz(tid + j * T);
#pragma unroll
                    for (int m = 0; m < M; ++m) dst[m * RS] = v[j][m];
                }
                STAGE_MARK(20 + s) // step s
            }
            __syncthreads();
            STAGE_MARK(16) // barrier after the row writes
            // ---- stage B: K1-point forward FFT of every row
            row_fft<S, -1>(buf, tw_s, tid);
            STAGE_MARK(17) // row FFT
            __syncthreads();
            // ---- stage C': column k' of all rows = the M bins of subcarrier 2k'+p
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const cpx* src = buf + tid + j * T;
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = src[m * RS];
            }
            __syncthreads();
            STAGE_MARK(18) // stage C' reads
            // output: records of item j staged at buf[j*T*M ...] in [k'][m] order, then written with
            // consecutive lanes on consecutive elements (records of the other parity lie in between)
            cpx* of = out + (size_t)g * N;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (mode == 0) {
                    rf::FFTN<M, +1>::run(v[j]);
#pragma unroll
                    for (int m = 0; m < M; ++m) v[j][m] = cscale(v[j][m], inv_m);
                }
                cpx* st = buf + j * T * M;
#pragma unroll
                for (int m = 0; m < M; ++m) st[tid * M + m] = v[j][m];
                __syncthreads();
                // item 0's staging has been copied out: the upper halves of the rows that lie inside it are free
                if (j == 1 && has_next) issue(gn, 1, 0, N1A, true, p == 1);
                cpx* ob = of + (2 * j * T + p) * M; // record 2*(j*T + r) + p starts at ob + 2*M*r
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    // i = tid + q*T = M*(r0 + (T/M)*q) + e0 + (T%M)*q
                    constexpr int TQ = T / M, TR = T % M;
                    int e = e0 + TR * q, r = r0 + TQ * q;
#pragma unroll
                    for (int c = 0; c < (M - 1 + TR * (M - 1)) / M; ++c)
                        if (e >= M) { e -= M; ++r; }
                    stg_stream(ob + 2 * M * r + e, st[tid + q * T]);
                }
            }
            __syncthreads();
            if (has_next) issue(gn, 1, N1A, M, false, p == 1);
            STAGE_MARK(19) // M-IFFT + output staging + stores
            phase ^= 1;
        }
    }
    Tmem4<S>::release(tmem_base, tid);
}

// ----------------------------------------------------------------------------------------
// receiver, second version: the samples are read ONCE.  Pass 0 forms both parities of the radix-2 step from one set of
// M-point transforms and table products (B^0 = a + b into the row buffer, B^1 = (a - b) W^{n'} parked in tensor memory),
// so pass 1 has no loads, no transforms and no table products; the even-subcarrier records of pass 0 wait in tensor
// memory as well, and pass 1 writes each thread's record PAIR (2k', 2k'+1: 2M contiguous elements) through a staging
// area with ONE bulk store per item -- whole lines instead of M-element runs with M-element gaps.  The table columns come
// from L2 (issued at the top of each step, in flight while the samples are read and transformed).
// Homes and issue points of the quarter rows as in fused_rx2_kernel, with "next pass" = pass 0 of the next frame; the
// quarter rows of step 1 (upper halves of the row buffer) follow the last bulk store of the frame.
template <class S, bool EQ>
__global__ void __launch_bounds__(S::T, 1) fused_rx2p_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                            const cpx* __restrict__ tables, const cpx* __restrict__ tw,
                                                            const cpx* __restrict__ w2, const cpx* __restrict__ eq,
                                                            const cpx* __restrict__ taps, int mode, int n_frames)
{
    constexpr int M = S::M, K1 = S::K, K = 2 * K1, N = M * K, T = S::T, RS = S::RS;
    static_assert(S::F == 1 && S::IPT == 2 && S::TWO_PASS, "two-pass kernels: one half-frame per CTA pass, two items per thread");
    constexpr int UP = RS / 2;
    constexpr int NE = (S::P_ELEMS - M * T) / T;
    static_assert(S::swz(T - 1) < UP && UP + T <= RS, "row halves");
    static_assert(NE >= 1 && NE < M && M <= T / 32 && M * T * 2 <= S::BUF_ELEMS, "staging does not fit");
    static_assert((M - NE - 1) * RS + UP + T <= S::BUF_ELEMS, "row halves");
    constexpr uint32_t QBYTES = T * sizeof(cpx);
    constexpr uint32_t STEP_BYTES = M * QBYTES;
    constexpr int ITEM = 2 * T * M; // elements of one item's record pairs: a contiguous piece of the output frame
    static_assert(!EQ || S::BUF_ELEMS - ITEM >= 2 * (T / 32) * M, "no room for the hand-over array behind the staging area");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* pre = tw_s + S::TW_ELEMS + S::TBL_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS); // one per step
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_m = 1.0f / (float)M;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if (tid == 0)
        for (int s = 0; s < 4; ++s) mbar_init(bars + s, 1);
    uint32_t tmem_base = 0, tmem_mine = 0;
    TmemPark<S>::alloc(reinterpret_cast<uint32_t*>(bars + 4), tid, tmem_base, tmem_mine);
    const cpx wj[2] = { w2[tid], w2[tid + T] };
    // equalising variant (overlap 2): receive taps in shared memory, one more barrier for the channel copies
    cpx* taps_s = pre + S::P_ELEMS;
    uint64_t* bar_h = bars + 5;
    uint32_t phase_h = 0;
    if constexpr (EQ) {
        if (tid < 2 * M) taps_s[tid] = taps[tid];
        if (tid == 0) mbar_init(bar_h, 1);
    }
    __syncthreads();

    auto qsrc = [&](int gg, int s, int n2) {
        return in + (size_t)gg * N + (size_t)n2 * K + (s & 1) * K1 + (s >> 1) * T;
    };
    auto home = [&](int s, int n2) -> cpx* {
        switch (s) {
        case 0: return pre + n2 * T;
        case 1: return buf + n2 * RS + UP;
        case 2: return n2 < NE ? pre + (M + n2) * T : pre + (n2 - NE) * T;
        default: return n2 < NE ? pre + (M - NE + n2) * T : buf + (n2 - NE) * RS + UP;
        }
    };
    auto issue = [&](int gg, int s, int n_lo, int n_hi, bool arm) {
        if (arm && tid == 0) mbar_expect_tx(bars + s, STEP_BYTES);
        if (lane == 0 && warp >= n_lo && warp < n_hi) {
            fence_proxy_async();
            bulk_load_hint(home(s, warp), qsrc(gg, s, warp), QBYTES, bars + s, l2_policy_evict_first());
        }
    };

    int g = blockIdx.x;
    if (g < n_frames) {
        issue(g, 0, 0, M, true);
        issue(g, 1, 0, M, true);
        issue(g, 2, 0, NE, true);
    }
    uint32_t phase = 0;
    bool first = true; // first frame of this CTA: step 1 was issued up front, no store is in flight
    STAGE_INIT();
    for (; g < n_frames; g += gridDim.x) {
        const int gn = g + gridDim.x;
        const bool has_next = gn < n_frames;
        if constexpr (EQ) { // the frame's channel is read twice (one parity per pass): pull it into L2 now
            if (tid < 2) bulk_prefetch_l2(eq + (size_t)g * N + (size_t)tid * ITEM, ITEM * sizeof(cpx));
        }
        {
            // Both h = 0 steps first: they read the prefetch region P only, so they run while the previous frame's last
            // bulk store still drains the staging area; a_j (table x transform) waits in the record slots of tensor memory.
            constexpr int ORDER[4] = { 0, 2, 1, 3 };
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = ORDER[i], j = s >> 1, h = s & 1;
                cpx tc[M], x[M];
                const cpx* tp = tables + (size_t)h * M * K1 + tid + j * T;
#pragma unroll
                for (int m = 0; m < M; ++m) tc[m] = ldg_nc(tp + (size_t)m * K1);
                mbar_wait(bars + s, phase);
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) x[n2] = home(s, n2)[tid];
                __syncthreads(); // step s consumed
                if (s == 0) {
                    issue(g, 2, NE, M, false);
                    issue(g, 3, 0, NE, true);
                } else if (s == 2) {
                    if (has_next) issue(gn, 2, 0, NE, true);
                    if (!first) {
                        // the quarter rows of step 1 live in the row buffer, which the previous frame's last store reads
                        if (tid == 0) bulk_wait_read();
                        __syncthreads();
                        issue(g, 1, 0, M, true);
                    }
                } else if (s == 1) {
                    issue(g, 3, NE, M, false);
                } else {
                    if (has_next) {
                        issue(gn, 0, 0, M, true);
                        if (lane == 0 && warp < M) bulk_prefetch_l2(qsrc(gn, 1, warp), QBYTES); // fetched after the stores
                    }
                }
                rf::FFTN<M, -1>::run(x);
                if (h == 0) {
#pragma unroll
                    for (int m = 0; m < M; ++m) x[m] = cmul(x[m], tc[m]);
                    TmemPark<S>::st(tmem_mine, 2 + j, x);
                } else {
                    cpx a[M];
                    TmemPark<S>::ld(a, tmem_mine, 2 + j);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const cpx b = cmul(x[m], tc[m]);
                        x[m] = cmul(csub(a[m], b), wj[j]); // B^1: waits for pass 1
                        a[m] = cadd(a[m], b);              // B^0: this pass
                    }
                    TmemPark<S>::st(tmem_mine, j, x);
                    cpx* dst = buf + S::swz(tid + j * T);
#pragma unroll
                    for (int m = 0; m < M; ++m) dst[m * RS] = a[m];
                }
                STAGE_MARK(20 + s) // step s
            }
        }
        __syncthreads();
        STAGE_MARK(16) // barrier after the row writes
        cpx* of = out + (size_t)g * N;
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            if (p == 1) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    cpx u[M];
                    TmemPark<S>::ld(u, tmem_mine, j);
                    cpx* dst = buf + S::swz(tid + j * T);
#pragma unroll
                    for (int m = 0; m < M; ++m) dst[m * RS] = u[m];
                }
                __syncthreads();
                STAGE_MARK(24) // pass 1: rows out of tensor memory
            }
            // ---- stage B: K1-point forward FFT of every row
            row_fft<S, -1>(buf, tw_s, tid);
            STAGE_MARK(17) // row FFT
            __syncthreads();
            // ---- stage C': column k' of all rows = the M bins of subcarrier 2k'+p
            cpx c[2][M];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const cpx* src = buf + tid + j * T;
#pragma unroll
                for (int m = 0; m < M; ++m) c[j][m] = src[m * RS];
            }
            __syncthreads(); // rows are dead
            STAGE_MARK(18) // stage C' reads
            if constexpr (EQ) {
                // Equalisation with overlap 2 (receiver_kernel_cc.cc:165-192,309-320): the rows now hold the true bins
                // Y[(2k'+p) M + m].  The channel records of an item's subcarrier pairs arrive by one bulk copy into the (free)
                // row buffer, in the order their owners read them; division in registers.
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (tid == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(bar_h, (uint32_t)ITEM * sizeof(cpx));
                        bulk_load(buf, eq + (size_t)g * N + (size_t)j * ITEM, (uint32_t)ITEM * sizeof(cpx), bar_h);
                    }
                    mbar_wait(bar_h, phase_h);
                    phase_h ^= 1;
                    const cpx* hh = buf + (size_t)tid * 2 * M + p * M;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const cpx h1 = hh[m];
                        const float rden = __fdividef(1.0f, h1.x * h1.x + h1.y * h1.y);
                        const cpx num = cmulc(c[j][m], h1); // y * conj(h) / |h|^2, volk_32fc_x2_divide_32fc
                        c[j][m] = cmake(num.x * rden, num.y * rden);
                    }
                    __syncthreads(); // channel consumed
                    if (p == 0) TmemPark<S>::st(tmem_mine, 2 + j, c[j]); // equalised even block waits for its odd neighbours
                }
                if (p == 1) {
                    // R_k[m] = taps[M+m] Yeq_{k-1}[m] + taps[m] Yeq_k[m].  Block 2k'+1 needs 2k' (parked, own), block 2k' needs
                    // 2k'-1 = the odd block of the previous thread: a shuffle; lane 0 takes it from the hand-over array
                    // behind the staging area (k' = 0 wraps to the last odd block of the frame).
                    cpx* bnd = buf + ITEM; // [item][warp][M]
                    if (lane == 31) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            cpx* b = bnd + (size_t)(j * (T / 32) + warp) * M;
#pragma unroll
                            for (int m = 0; m < M; ++m) b[m] = c[j][m];
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        cpx ev[M];
                        TmemPark<S>::ld(ev, tmem_mine, 2 + j);
                        const int wprev = (j * (T / 32) + warp + 2 * (T / 32) - 1) % (2 * (T / 32)); // warp 0 of item 0: last of item 1
                        const cpx* b = bnd + (size_t)wprev * M;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            cpx pv = cmake(__shfl_up_sync(0xffffffffu, c[j][m].x, 1), __shfl_up_sync(0xffffffffu, c[j][m].y, 1));
                            if (lane == 0) pv = b[m];
                            const cpx re = cadd(cmul(taps_s[M + m], pv), cmul(taps_s[m], ev[m]));
                            c[j][m] = cadd(cmul(taps_s[M + m], ev[m]), cmul(taps_s[m], c[j][m]));
                            ev[m] = re;
                        }
                        if (mode == 0) {
                            rf::FFTN<M, +1>::run(ev);
                            rf::FFTN<M, +1>::run(c[j]);
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                ev[m] = cscale(ev[m], inv_m);
                                c[j][m] = cscale(c[j][m], inv_m);
                            }
                        }
                        if (j == 1) { // item 0's store must have read the staging area
                            if (tid == 0) bulk_wait_read();
                            __syncthreads();
                        }
                        float4* st = reinterpret_cast<float4*>(buf + (size_t)tid * 2 * M);
#pragma unroll
                        for (int q = 0; q < M; ++q) {
                            const cpx e0 = 2 * q < M ? ev[2 * q] : c[j][2 * q - M];
                            const cpx e1 = 2 * q + 1 < M ? ev[2 * q + 1] : c[j][2 * q + 1 - M];
                            st[q] = make_float4(e0.x, e0.y, e1.x, e1.y);
                        }
                        fence_proxy_async();
                        __syncthreads();
                        if (tid == 0) bulk_store(of + (size_t)j * ITEM, buf, (uint32_t)ITEM * sizeof(cpx));
                    }
                }
            } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (mode == 0) {
                    rf::FFTN<M, +1>::run(c[j]);
#pragma unroll
                    for (int m = 0; m < M; ++m) c[j][m] = cscale(c[j][m], inv_m);
                }
                if (p == 0) {
                    TmemPark<S>::st(tmem_mine, 2 + j, c[j]); // record 2k' waits for record 2k'+1
                } else {
                    cpx ev[M];
                    TmemPark<S>::ld(ev, tmem_mine, 2 + j);
                    if (j == 1) { // item 0's store must have read the staging area
                        if (tid == 0) bulk_wait_read();
                        __syncthreads();
                    }
                    // records 2k' and 2k'+1 of k' = tid + j*T: 2M contiguous elements = M 16-byte stores (lane stride
                    // 2M elements = 240 B at M = 15: conflict-free per quarter warp)
                    float4* st = reinterpret_cast<float4*>(buf + (size_t)tid * 2 * M);
#pragma unroll
                    for (int q = 0; q < M; ++q) {
                        const cpx e0 = 2 * q < M ? ev[2 * q] : c[j][2 * q - M];
                        const cpx e1 = 2 * q + 1 < M ? ev[2 * q + 1] : c[j][2 * q + 1 - M];
                        st[q] = make_float4(e0.x, e0.y, e1.x, e1.y);
                    }
                    fence_proxy_async();
                    __syncthreads();
                    if (tid == 0) bulk_store(of + (size_t)j * ITEM, buf, (uint32_t)ITEM * sizeof(cpx));
                }
            }
            }
            if (p == 0) { STAGE_MARK(19) } else { STAGE_MARK(25) } // M-IFFT + parking (pass 0) / staging + bulk stores (pass 1)
        }
        // (the staging area is still being read by the last store: the next frame issues the quarter rows of its step 1,
        // whose homes are the upper halves of the rows, after its two P-only steps)
        first = false;
        phase ^= 1;
    }
    if (tid == 0) bulk_wait_all();
    TmemPark<S>::release(tmem_base, tid);
}

// ----------------------------------------------------------------------------------------
// host side
struct TwoPass {
    int M = 0, K = 0, L = 0;
    bool tx = false;
    bool parked = true; // second parity + first-pass results parked in tensor memory (GFDM_MOD2_SCRATCH / GFDM_RX2_REREAD: first versions)
    cpx* d_table = nullptr;
    cpx* d_tw = nullptr;
    cpx* d_w2 = nullptr;
    cpx* d_scratch = nullptr; // modulator: grid_cap x M*K1 even samples of pass 0 (L2 resident)
    cpx* d_table_eq = nullptr; // receiver, overlap 2: plain N-point twiddle in the same [h][M][K1] layout (equalising variant)
    cpx* d_taps = nullptr;     // receiver, overlap 2: the L*M receive taps
    int grid_cap = 0;
    size_t smem = 0;
    std::string name;
};

typedef Shape<15, 32, 32, 512, 2, 1, false> S15x1024;

bool twopass_supported(int M, int K) { return M == S15x1024::M && K == 2 * S15x1024::K; }

static void twopass_free(TwoPass* t)
{
    if (!t) return;
    if (t->d_table) cudaFree(t->d_table);
    if (t->d_tw) cudaFree(t->d_tw);
    if (t->d_w2) cudaFree(t->d_w2);
    if (t->d_scratch) cudaFree(t->d_scratch);
    if (t->d_table_eq) cudaFree(t->d_table_eq);
    if (t->d_taps) cudaFree(t->d_taps);
    delete t;
}
void twopass_destroy(TwoPass* t) { twopass_free(t); }

TwoPass* twopass_create_tx(int M, int K, int L, const std::vector<std::complex<float>>& taps)
{
    if (!twopass_supported(M, K) || L < 1) return nullptr;
    typedef S15x1024 S;
    const int K1 = K / 2;
    TwoPass* t = new TwoPass;
    t->M = M; t->K = K; t->L = L; t->tx = true;
    t->smem = S::SMEM_BYTES;
    t->name = "fused_mod2_kernel<M=15,K=2x32x32,T=512>";
    try {
        t->parked = std::getenv("GFDM_MOD2_SCRATCH") == nullptr;
        t->grid_cap = fused_grid_cap((const void*)&fused_mod2_kernel<S>, S::T, S::SMEM_BYTES);
        const int cap_p = fused_grid_cap((const void*)&fused_mod2p_kernel<S>, S::T, S::SMEM_BYTES);
        if (t->parked) t->grid_cap = cap_p;
        if (t->parked) t->name = "fused_mod2p_kernel<M=15,K=2x32x32,T=512>";
        const std::vector<cpx> C = make_fold_table(M, K, L, taps, +1, true); // [m][n1]
        std::vector<cpx> P((size_t)2 * M * K1);
        for (int p = 0; p < 2; ++p)
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < K1; ++n) P[((size_t)p * M + m) * K1 + n] = C[(size_t)m * K + 2 * n + p];
        t->d_table = upload(P);
        t->d_tw = upload(make_row_twiddles(S::R1, S::R2));
        std::vector<cpx> w((size_t)K1);
        for (int b = 0; b < K1; ++b) {
            const double ph = 2.0 * M_PI * (double)b / (double)K;
            w[b] = make_float2((float)std::cos(ph), (float)std::sin(ph));
        }
        t->d_w2 = upload(w);
        GFDM_CUDA_CHECK(cudaMalloc(&t->d_scratch, sizeof(cpx) * (size_t)t->grid_cap * M * K1));
    } catch (...) {
        twopass_free(t);
        throw;
    }
    return t;
}

TwoPass* twopass_create_rx(int M, int K, int L, const std::vector<std::complex<float>>& taps)
{
    if (!twopass_supported(M, K) || L < 2) return nullptr;
    typedef S15x1024 S;
    const int K1 = K / 2;
    TwoPass* t = new TwoPass;
    t->M = M; t->K = K; t->L = L; t->tx = false;
    t->smem = S::SMEM_BYTES;
    t->name = "fused_rx2_kernel<M=15,K=2x32x32,T=512>";
    try {
        t->parked = std::getenv("GFDM_RX2_REREAD") == nullptr; // (set: the first version, which reads the frame in both passes)
        t->grid_cap = fused_grid_cap((const void*)&fused_rx2_kernel<S>, S::T, S::SMEM_BYTES);
        const int cap_p = fused_grid_cap((const void*)&fused_rx2p_kernel<S, false>, S::T, S::SMEM_BYTES);
        fused_grid_cap((const void*)&fused_rx2p_kernel<S, true>, S::T, S::SMEM_BYTES); // opts the equalising variant into its shared memory
        if (t->parked) t->grid_cap = cap_p;
        if (t->parked) t->name = "fused_rx2p_kernel<M=15,K=2x32x32,T=512>";
        if (t->parked && L == 2) { // the equalising variant combines exactly two blocks in registers
            const std::vector<std::complex<double>> E = make_fold_table_d(M, K, L, taps, -1, false);
            std::vector<cpx> Pe((size_t)2 * M * K1), tp((size_t)L * M);
            for (int h = 0; h < 2; ++h)
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < K1; ++n) {
                        const std::complex<double> c = E[(size_t)m * K + n + h * K1];
                        Pe[((size_t)h * M + m) * K1 + n] = make_float2((float)c.real(), (float)c.imag());
                    }
            for (int i = 0; i < L * M; ++i) tp[i] = make_float2(taps[i].real(), taps[i].imag());
            t->d_table_eq = upload(Pe);
            t->d_taps = upload(tp);
        }
        // C_rx[m][n'] and C_rx[m][n'+K1] as [h][m][n']; the pass-1 twiddle W^{n'} = e^{-j2pi n'/K} is applied in the kernel
        const std::vector<std::complex<double>> C = make_fold_table_d(M, K, L, taps, -1, true);
        std::vector<cpx> P((size_t)2 * M * K1);
        for (int h = 0; h < 2; ++h)
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < K1; ++n) {
                    const std::complex<double> c = C[(size_t)m * K + n + h * K1];
                    P[((size_t)h * M + m) * K1 + n] = make_float2((float)c.real(), (float)c.imag());
                }
        std::vector<cpx> w((size_t)K1);
        for (int n = 0; n < K1; ++n) {
            const double ph = -2.0 * M_PI * (double)n / (double)K;
            w[n] = make_float2((float)std::cos(ph), (float)std::sin(ph));
        }
        t->d_w2 = upload(w);
        t->d_table = upload(P);
        t->d_tw = upload(make_row_twiddles(S::R1, S::R2));
    } catch (...) {
        twopass_free(t);
        throw;
    }
    return t;
}

int twopass_modulate(TwoPass* t, cpx* out, const cpx* in, size_t frames, cudaStream_t s)
{
    typedef S15x1024 S;
    int launches = 0;
    const size_t N = (size_t)t->M * t->K, max_chunk = (size_t)1 << 20;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int grid = nf < t->grid_cap ? nf : t->grid_cap;
        if (t->parked)
            fused_mod2p_kernel<S><<<grid, S::T, S::SMEM_BYTES, s>>>(out + f0 * N, in + f0 * N, t->d_table, t->d_tw, t->d_w2, nf);
        else
            fused_mod2_kernel<S><<<grid, S::T, S::SMEM_BYTES, s>>>(out + f0 * N, in + f0 * N, t->d_table, t->d_tw, t->d_w2, t->d_scratch, nf);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

bool twopass_supports_eq(const TwoPass* t) { return t && !t->tx && t->d_table_eq != nullptr; }

int twopass_demodulate(TwoPass* t, cpx* out, const cpx* in, const cpx* eq, int mode, size_t frames, cudaStream_t s)
{
    typedef S15x1024 S;
    if (eq && !twopass_supports_eq(t)) throw std::invalid_argument("the two-pass receiver kernel equalises with overlap 2 only");
    int launches = 0;
    const size_t N = (size_t)t->M * t->K, max_chunk = (size_t)1 << 20;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int grid = nf < t->grid_cap ? nf : t->grid_cap;
        if (eq)
            fused_rx2p_kernel<S, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out + f0 * N, in + f0 * N, t->d_table_eq, t->d_tw, t->d_w2,
                                                                        eq + f0 * N, t->d_taps, mode, nf);
        else if (t->parked)
            fused_rx2p_kernel<S, false><<<grid, S::T, S::SMEM_BYTES, s>>>(out + f0 * N, in + f0 * N, t->d_table, t->d_tw, t->d_w2, nullptr,
                                                                         nullptr, mode, nf);
        else
            fused_rx2_kernel<S><<<grid, S::T, S::SMEM_BYTES, s>>>(out + f0 * N, in + f0 * N, t->d_table, t->d_tw, t->d_w2, mode, nf);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

const char* twopass_name(const TwoPass* t) { return t->name.c_str(); }

#ifdef GFDM_PROFILE_STAGES
extern "C" __attribute__((visibility("default"))) int gfdm_debug_stage_cycles2(unsigned long long* out32, int reset)
{
    if (out32 && cudaMemcpyFromSymbol(out32, g_stage_cycles, sizeof(g_stage_cycles)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[32] = { 0 };
        if (cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

} // namespace gfdm
