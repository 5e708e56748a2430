// fused_modem.cu -- single-kernel GFDM modulator and receiver (sm_100a).
//
// Factorisation (checked in NumPy by tools/fused_math_check.py).
// Bin index b*M+m, sample index n1 + K*n2, symbols d[k*M+t].  The K-point transform over the
// subcarrier index and the M-point transforms over the sub-symbol index commute, so
//
//   modulator  (replaces lib/modulator_kernel_cc.cc:98-141)
//     E_t[n1]      = IFFT_K over k of d[k][t]                      row phase, IN PLACE on the [k][t] frame
//     x[n1+K*n2]   = IFFT_M(m->n2){ C_tx[m][n1] * FFT_M(t->m){ E_t[n1] } }   column phase, registers
//   receiver   (replaces lib/receiver_kernel_cc.cc:165-225,301-334), no equalisation
//     Q_t[n1]      = IFFT_M(m->t){ C_rx[m][n1] * FFT_M(n2->m){ x[n1+K*n2] } } / M   column phase
//     y[k][t]      = FFT_K over n1 of Q_t[n1]                      row phase, in place, result = output layout
//   receiver with equalisation: the column phase stops after the table multiply (plain twiddle), the
//     row phase yields the true bins Y[b*M+m] in place, then Y/H, the L-part tap combination and the
//     M-point IFFT run per subcarrier (the division sits between FFT and filter, receiver_kernel_cc.cc:315-319).
//   C_tx / C_rx fold the L-fold spectral repetition, the filter taps, the scatter-add into the N-bin
//   grid (a circular shift over b is a phase ramp over n1), the N-point twiddle and 1/N into one table.
//
// Shared memory: ONE frame-group array S in the [k][t] order of the symbol side of each kernel
// (modulator input, receiver output), the row-FFT twiddles, and a prefetch region P filled by TMA
// (cp.async.bulk + mbarrier) with the next group's input while the current group is processed.
// A K = R*R row is transformed by R lanes of one warp in two radix-R register passes; the exchange
// between the passes stays inside the row's own strided slots (slot R*k1 + (n0 ^ k1): conflict-free
// for odd M in both directions), so the row phase needs no CTA barrier and no second buffer.
// [k][t]-ordered arrays cross HBM through bulk copies (128 B lines); time samples are accessed with
// lanes on consecutive n1, two columns per thread (128-bit loads/stores).  Grid = SMs x occupancy,
// persistent loop over frame groups.
#include "fused.h"
#include "regfft.cuh"

#include <cmath>
#include <cstring>

namespace gfdm {

// ----------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1D bulk async copy (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// streaming global accesses: every input byte is read once, every output byte written once
__device__ __forceinline__ cpx ldg_stream(const cpx* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream4(const cpx* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(cpx* p, cpx v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream4(cpx* p, cpx a, cpx b)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x),
                 "f"(b.y)
                 : "memory");
}

// CPT consecutive complex values at p (16-byte aligned when CPT == 2)
template <int CPT>
__device__ __forceinline__ void ld_cols(cpx (&dst)[CPT], const cpx* p)
{
    if constexpr (CPT == 2) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        dst[0] = cmake(q.x, q.y);
        dst[1] = cmake(q.z, q.w);
    } else {
        dst[0] = *p;
    }
}
template <int CPT>
__device__ __forceinline__ void ldg_cols(cpx (&dst)[CPT], const cpx* p)
{
    if constexpr (CPT == 2) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        dst[0] = cmake(q.x, q.y);
        dst[1] = cmake(q.z, q.w);
    } else {
        dst[0] = __ldg(reinterpret_cast<const float2*>(p));
    }
}
template <int CPT>
__device__ __forceinline__ void ldg_stream_cols(cpx (&dst)[CPT], const cpx* p)
{
    if constexpr (CPT == 2) {
        const float4 q = ldg_stream4(p);
        dst[0] = cmake(q.x, q.y);
        dst[1] = cmake(q.z, q.w);
    } else {
        dst[0] = ldg_stream(p);
    }
}

// ----------------------------------------------------------------------------------------
// optional per-stage cycle counters (build with -DGFDM_PROFILE_STAGES; tools/stage_profile.py)
#ifdef GFDM_PROFILE_STAGES
__device__ unsigned long long g_stage_cycles[32];
#define STAGE_INIT() long long t_last = clock64()
#define STAGE_MARK(i)                                                        \
    if (threadIdx.x == 0) {                                                  \
        const long long t_now = clock64();                                   \
        atomicAdd(&g_stage_cycles[i], (unsigned long long)(t_now - t_last)); \
        t_last = t_now;                                                      \
    }
#else
#define STAGE_INIT()
#define STAGE_MARK(i)
#endif

// ----------------------------------------------------------------------------------------
// compile-time shape of one fused kernel pair
//   M  sub-symbols, K = R*R subcarriers (SINGLE: K = R, one register pass per row)
//   T  threads, F frames per CTA pass, CPT sample columns per thread, MINB CTAs per SM
template <int M_, int R_, bool SINGLE_, int T_, int F_, int CPT_, int MINB_>
struct Shape {
    static constexpr int M = M_, R = R_, T = T_, F = F_, CPT = CPT_, MINB = MINB_;
    static constexpr bool SINGLE = SINGLE_;
    static constexpr int K = SINGLE ? R : R * R;
    static constexpr int N = M * K;
    static constexpr int GROUP = F * N;                          // elements per CTA pass
    static constexpr int ROW_SLOTS = SINGLE ? F * M : F * M * R; // lane slots of the row phase
    static constexpr int ROW_ROUNDS = (ROW_SLOTS + T - 1) / T;
    static constexpr int COLS = K / CPT;      // column items per frame
    static constexpr int COL_ITEMS = F * COLS;
    static constexpr int COL_ROUNDS = (COL_ITEMS + T - 1) / T; // processed one after the other, highest n1 first
    static_assert(M % 2 == 1, "in-place [k][t] addressing is bank-conflict free for odd M only");
    static_assert(32 % R == 0 || SINGLE, "a row's R lanes must sit inside one warp");
    static_assert(K % CPT == 0 && (CPT == 1 || CPT == 2), "one or two columns per thread");
    static_assert(T % 32 == 0, "whole warps");
    // Row-phase skew (R == 32, one row per warp): the second half of the row warps starts one step later, so
    // that shared-memory steps of one half overlap the fp32 steps of the other; the warp without a row
    // (if any) issues the bulk loads so that no row warp waits on the copy engine.
    static constexpr int ROW_WARPS = SINGLE ? 0 : (ROW_SLOTS + 31) / 32;
    static constexpr bool SKEW = !SINGLE && R == 32 && ROW_ROUNDS == 1 && ROW_WARPS >= 4;
    static constexpr int SKEW_A = (ROW_WARPS + 1) / 2;                    // warps [0, SKEW_A) go first
    static constexpr bool DMA_WARP = SKEW && ROW_WARPS * 32 < T;          // an idle warp exists in the row phase
    static constexpr int DMA_TID = ROW_WARPS * 32;
    static constexpr int TW_ELEMS = SINGLE ? 0 : K;
    static constexpr int TAPS_ELEMS = 64; // receive taps of the equalising path (L*M <= 64)
    // per-CTA shared memory budget in complex elements (228 KB per SM, 1 KB per CTA reserved)
    static constexpr int BUDGET_ELEMS = ((233472 / MINB) - 1024) / 8 - 8 - TAPS_ELEMS;
    static constexpr int P_MAX = BUDGET_ELEMS - GROUP - TW_ELEMS;
    static_assert(P_MAX >= 0, "frame group does not fit in shared memory");
    // modulator prefetch: the whole next group when it fits, else (F == 1) its first KH subcarriers;
    // KH is a multiple of R so that "head or tail" is decided by the register index of pass 1
    static constexpr bool MOD_FULL = P_MAX >= GROUP;
    static constexpr int KH = MOD_FULL ? F * K : (P_MAX / (R * M)) * R;
    static constexpr int PF = KH * M;
    static_assert(MOD_FULL || (F == 1 && !SINGLE && ROW_ROUNDS == 1 && KH >= R), "split prefetch needs one frame, one round");
    // split prefetch: the first column round (highest n1) must cover the tail subcarriers [KH, K)
    static_assert(MOD_FULL || (COL_ITEMS % T == 0 && KH >= K - T * CPT), "tail columns must fall into the first column round");
    // receiver prefetch: the first PR sample rows (n2) of every frame of the next group
    static constexpr int PR = M < P_MAX / (F * K) ? M : P_MAX / (F * K);
    static_assert(PR >= 1, "no room for the receiver prefetch");
    static constexpr int XR = M - PR; // sample rows prefetched into registers instead
    static constexpr int P_ELEMS = PF > F * PR * K ? PF : F * PR * K;
    static constexpr size_t SMEM_BYTES = sizeof(cpx) * (size_t)(GROUP + TW_ELEMS + P_ELEMS + TAPS_ELEMS) + 64;
};

// ----------------------------------------------------------------------------------------
// Row phase: K-point transforms (direction DIR) of the rows t of every frame held in s as
// s[f*N + q*M + t], natural order in and out, in place.  MOD_SRC: pass 1 reads the staged modulator
// input (head part from `head`, tail part already in s) instead of s.
// The caller provides the CTA barriers before and after; `after_reads` runs once all pass-1 reads of
// the CTA are done (split prefetch only: the staged input is then dead).
template <class S, int DIR, bool MOD_SRC, class AfterReads>
__device__ __forceinline__ void row_phase(cpx* __restrict__ s, const cpx* __restrict__ head, const cpx* __restrict__ tw_s,
                                          int tid, uint64_t* bar_skew, uint64_t* bar_consumed, uint32_t parity, int zero,
                                          AfterReads&& after_reads)
{
    constexpr int M = S::M, R = S::R, K = S::K, N = S::N, T = S::T;
    constexpr bool SPLIT = MOD_SRC && !S::MOD_FULL;
    constexpr int HEAD_BLOCKS = S::KH / R; // split prefetch: register index i < HEAD_BLOCKS <=> element in `head`
    constexpr int TWC = 8;   // twiddles in flight per chunk (register pressure of the radix-32 passes)
#pragma unroll 1
    for (int round = 0; round < S::ROW_ROUNDS; ++round) {
        const int slot0 = round * T + (tid & ~31);
        const bool warp_on = slot0 < S::ROW_SLOTS; // whole warps drop out together
        // surplus lanes of a partly filled warp redo the last slot (same reads, same values written):
        // no divergence around the warp barriers
        int slot = round * T + tid;
        slot = slot < S::ROW_SLOTS ? slot : S::ROW_SLOTS - 1;
        if constexpr (S::SINGLE) {
            if (warp_on) {
                const int f = slot / M, t = slot - f * M;
                cpx* row = s + f * N + t;
                const cpx* src = MOD_SRC ? head + f * N + t : row;
                cpx a[R];
#pragma unroll
                for (int i = 0; i < R; ++i) a[i] = src[i * M];
                rf::FFTN<R, DIR>::run(a);
#pragma unroll
                for (int i = 0; i < R; ++i) row[i * M] = a[i];
            }
        } else {
            const int rowi = slot / R, j = slot - rowi * R; // j: n0 in pass 1, k1 in pass 2
            const int f = rowi / M, t = rowi - f * M;
            cpx* row = s + f * N + t;
            cpx a[R];
            const int warp = tid >> 5;
            if constexpr (S::SKEW) {
                if (warp >= S::SKEW_A && warp_on) mbar_wait(bar_skew, parity); // second half: one step behind
            }
            if (warp_on) {
                // pass 1 reads: a[i] = x[R*i + j]
                const int e0 = (f * K + j) * M + t;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const cpx* src = !MOD_SRC ? s : ((!SPLIT || i < HEAD_BLOCKS) ? head : s);
                    a[i] = src[e0 + i * R * M];
                }
            }
            if constexpr (S::SKEW) {
                // first half: signal once the loads have RETURNED (the address depends on the last value read;
                // `zero` is a run-time 0 the compiler cannot fold)
                if (warp < S::SKEW_A && (tid & 31) == 0)
                    mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(bar_skew) +
                                                            (__float_as_int(a[R - 1].y) & zero)));
            }
            if constexpr (SPLIT && S::DMA_WARP) {
                // a row's slots are private to its warp, so the only CTA-wide event is "the staged input in P has
                // been read": every row warp signals, the idle warp waits and refills P
                __syncwarp();
                if (warp_on && (tid & 31) == 0) mbar_arrive(bar_consumed);
                if (tid == S::DMA_TID) {
                    mbar_wait(bar_consumed, parity);
                    after_reads();
                }
            } else if constexpr (SPLIT) {
                __syncthreads(); // every pass-1 read of the CTA has retired: the staged input is dead
                if (tid == 0) after_reads();
            } else if constexpr (!MOD_SRC) {
                __syncwarp(); // the row's lanes have read before any of them overwrites the row
            }
            if (warp_on) {
                rf::FFTN<R, DIR>::run(a);
                // twiddle W_K^{j*k1} + exchange write to slot R*k1 + (j ^ k1), in chunks: the empty asm
                // keeps the chunks in program order so that only a few twiddles are live next to the data
                row[(size_t)j * M] = a[0];
#pragma unroll
                for (int c = 0; c < R; c += TWC) {
                    // the lane index is laundered per chunk: otherwise the compiler hoists all 2*R swizzled
                    // offsets out of the frame loop and spills them
                    int jc = j;
                    asm volatile("" : "+r"(jc));
                    cpx w[TWC];
#pragma unroll
                    for (int i = 0; i < TWC; ++i)
                        if (c + i > 0 && c + i < R) w[i] = tw_s[(c + i) * R + j];
#pragma unroll
                    for (int i = 0; i < TWC; ++i)
                        if (c + i > 0 && c + i < R) {
                            if (DIR > 0) w[i].y = -w[i].y;
                            row[(R * (c + i) + (jc ^ (c + i))) * M] = cmul(a[c + i], w[i]);
                        }
                    asm volatile("" ::: "memory");
                }
            }
            __syncwarp();
            if (warp_on) {
                // pass 2 reads: b[n0] = A[n0][k1 = j]
                // (in chunks of 8 again: the swizzled offsets are computed, not immediate)
                cpx* blk = row + (size_t)R * j * M;
#pragma unroll
                for (int c = 0; c < R; c += 8) {
                    int jc = j;
                    asm volatile("" : "+r"(jc));
#pragma unroll
                    for (int i = c; i < c + 8 && i < R; ++i) a[i] = blk[(i ^ jc) * M];
                    asm volatile("" ::: "memory");
                }
                rf::FFTN<R, DIR>::run(a);
            }
            __syncwarp();
            if (warp_on) {
                // natural-order write: X[j + R*k0]
#pragma unroll
                for (int i = 0; i < R; ++i) row[(size_t)(j + R * i) * M] = a[i];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// Fused modulator.  in/out: [n_frames][N]; table: C_tx [M][K]; tw: W_K^{n0*k1} as [k1][n0].
template <class S>
__global__ void __launch_bounds__(S::T, S::MINB) fused_mod_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                                  const cpx* __restrict__ table,
                                                                  const cpx* __restrict__ tw, int n_frames)
{
    constexpr int M = S::M, K = S::K, N = S::N, T = S::T, F = S::F, CPT = S::CPT, PF = S::PF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = s + S::GROUP;
    cpx* pre = tw_s + S::TW_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS);
    uint64_t* bar_p = bars;     // head of the staged input (region P)
    uint64_t* bar_t = bars + 1; // tail of the staged input (lands in S)
    uint64_t* bar_c = bars + 2; // "P has been read" (row warps -> the warp that refills P)
    uint64_t* bar_k = bars + 3; // row-phase skew
    const int tid = threadIdx.x;
    const int n_groups = (n_frames + F - 1) / F;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_t, 1);
        mbar_init(bar_c, S::ROW_WARPS > 0 ? S::ROW_WARPS : 1);
        mbar_init(bar_k, S::SKEW_A > 0 ? S::SKEW_A : 1);
    }
    __syncthreads();
    const int zero = n_frames >> 31; // 0 at run time, unknown at compile time

    // each called by ONE thread
    auto load_head = [&](int gg) {
        const int el = min(F, n_frames - gg * F) * N;
        const uint32_t bytes = (uint32_t)min(el, PF) * sizeof(cpx);
        mbar_expect_tx(bar_p, bytes);
        bulk_load(pre, in + (size_t)gg * F * N, bytes, bar_p);
    };
    auto load_tail = [&](int gg) { // split prefetch only: elements [PF, N) go to their home slots in S
        constexpr uint32_t bytes = (uint32_t)(S::GROUP - PF) * sizeof(cpx);
        mbar_expect_tx(bar_t, bytes);
        bulk_load(s + PF, in + (size_t)gg * F * N + PF, bytes, bar_t);
    };

    int g = blockIdx.x;
    if (tid == 0 && g < n_groups) {
        load_head(g);
        if constexpr (!S::MOD_FULL) load_tail(g);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_groups; g += gridDim.x) {
        const int fh = min(F, n_frames - g * F);
        const int gn = g + gridDim.x;
        mbar_wait(bar_p, phase);
        if constexpr (!S::MOD_FULL) mbar_wait(bar_t, phase);
        phase ^= 1;
        STAGE_MARK(0) // wait for the bulk loads
        // ---- row phase: E_t = IFFT_K over k, in place
        row_phase<S, +1, true>(s, pre, tw_s, tid, bar_k, bar_c, phase ^ 1, zero, [&] {
            if (gn < n_groups) { // split prefetch (one thread): P is free as soon as pass 1 has read it
                fence_proxy_async();
                load_head(gn);
            }
        });
        // table column of the first column round: requested before the barrier, used after the M-point FFT
        cpx tc[M][CPT];
        auto load_table = [&](int r) {
            const int item = (S::COL_ROUNDS - 1 - r) * T + tid;
            const int n1 = (item % S::COLS) * CPT;
            if (item < S::COL_ITEMS) {
#pragma unroll
                for (int m = 0; m < M; ++m) ldg_cols<CPT>(tc[m], table + m * K + n1);
            }
        };
        load_table(0);
        __syncthreads();
        STAGE_MARK(1) // row phase
        if constexpr (S::MOD_FULL) {
            if (tid == 0 && gn < n_groups) {
                fence_proxy_async();
                load_head(gn);
            }
        }
        // ---- column phase: CPT adjacent sample columns n1 per thread and round, highest columns first
#pragma unroll
        for (int r = 0; r < S::COL_ROUNDS; ++r) {
            const int item = (S::COL_ROUNDS - 1 - r) * T + tid;
            const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
            const bool col_on = item < S::COL_ITEMS;
            cpx v[CPT][M];
            if (col_on) {
                const cpx* src = s + (size_t)f * N + (size_t)n1 * M;
                if constexpr (CPT == 2) {
                    // 2*M contiguous values: (col0 t=0..M-1)(col1 t=0..M-1), read as M float4
                    cpx lin[2 * M];
#pragma unroll
                    for (int q = 0; q < M; ++q) {
                        const float4 u = reinterpret_cast<const float4*>(src)[q];
                        lin[2 * q] = cmake(u.x, u.y);
                        lin[2 * q + 1] = cmake(u.z, u.w);
                    }
#pragma unroll
                    for (int t = 0; t < M; ++t) {
                        v[0][t] = lin[t];
                        v[1][t] = lin[M + t];
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < M; ++t) v[0][t] = src[t];
                }
            }
            if constexpr (!S::MOD_FULL) {
                if (r == 0) {
                    __syncthreads(); // the tail slots of S are dead: fetch the tail of the next group
                    if (tid == 0 && gn < n_groups) {
                        fence_proxy_async();
                        load_tail(gn);
                    }
                }
            } else {
                if (r == S::COL_ROUNDS - 1) __syncthreads(); // S is dead: the next pass may overwrite it
            }
            if (col_on) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) rf::FFTN<M, -1>::run(v[c]);
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int c = 0; c < CPT; ++c) v[c][m] = cmul(v[c][m], tc[m][c]);
            }
            if (r + 1 < S::COL_ROUNDS) load_table(r + 1); // next round's table column: in flight during the IFFT + stores
            if (col_on) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) rf::FFTN<M, +1>::run(v[c]);
                if (f < fh) {
                    cpx* dst = out + ((size_t)g * F + f) * N + n1;
#pragma unroll
                    for (int n2 = 0; n2 < M; ++n2) {
                        if constexpr (CPT == 2)
                            stg_stream4(dst + (size_t)n2 * K, v[0][n2], v[1][n2]);
                        else
                            stg_stream(dst + (size_t)n2 * K, v[0][n2]);
                    }
                }
            }
        }
        STAGE_MARK(3) // column compute + stores
    }
}

// ----------------------------------------------------------------------------------------
// Fused receiver.  in: [n_frames][N] time samples; out: [n_frames][N];
// mode 0: soft symbols y (generic_work[_equalize]); mode 1: R (fft_[equalize_]filter_downsample).
// EQ: eq holds the per-bin channel of every frame, `table` is the plain twiddle; otherwise `table`
// carries the receive taps as well.
template <class S, bool EQ>
__global__ void __launch_bounds__(S::T, S::MINB) fused_rx_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                                 const cpx* __restrict__ eq,
                                                                 const cpx* __restrict__ table,
                                                                 const cpx* __restrict__ tw,
                                                                 const cpx* __restrict__ taps, int L, int mode,
                                                                 int n_frames)
{
    constexpr int M = S::M, K = S::K, N = S::N, T = S::T, F = S::F, CPT = S::CPT, PR = S::PR;
    constexpr int XR = S::XR > 0 ? S::XR : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* s = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = s + S::GROUP;
    cpx* pre = tw_s + S::TW_ELEMS;
    cpx* taps_s = pre + S::P_ELEMS;
    uint64_t* bar_p = reinterpret_cast<uint64_t*>(taps_s + S::TAPS_ELEMS);
    uint64_t* bar_s = bar_p + 1; // "the previous bulk store has finished reading S" (signalled by thread 0)
    uint64_t* bar_k = bar_p + 2; // row-phase skew
    const int tid = threadIdx.x;
    const int n_groups = (n_frames + F - 1) / F;
    const float inv_m = 1.0f / (float)M;
    constexpr int ROUNDS = S::COL_ROUNDS;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    if constexpr (EQ)
        for (int i = tid; i < L * M && i < S::TAPS_ELEMS; i += T) taps_s[i] = taps[i];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_k, S::SKEW_A > 0 ? S::SKEW_A : 1);
    }
    __syncthreads();
    const int zero = n_frames >> 31; // 0 at run time, unknown at compile time

    // sample rows n2 < PR of every frame of group gg -> P (one bulk copy per group, or one per frame)
    auto load_head = [&](int gg) { // called by one thread
        const int fhh = min(F, n_frames - gg * F);
        if constexpr (PR == M) {
            const uint32_t bytes = (uint32_t)fhh * N * sizeof(cpx);
            mbar_expect_tx(bar_p, bytes);
            bulk_load(pre, in + (size_t)gg * F * N, bytes, bar_p);
        } else {
            constexpr uint32_t bytes = (uint32_t)PR * K * sizeof(cpx);
            mbar_expect_tx(bar_p, bytes * fhh);
            for (int ff = 0; ff < fhh; ++ff)
                bulk_load(pre + (size_t)ff * PR * K, in + ((size_t)gg * F + ff) * N, bytes, bar_p);
        }
    };
    // sample rows n2 >= PR -> registers (one set per column round)
    cpx xr[ROUNDS][XR][CPT];
    auto load_rest = [&](int gg) {
        if constexpr (S::XR > 0) {
            const int fhh = min(F, n_frames - gg * F);
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int item = r * T + tid;
                const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
                const cpx* src = in + ((size_t)gg * F + f) * N + n1;
#pragma unroll
                for (int q = 0; q < S::XR; ++q) {
                    if (item < S::COL_ITEMS && f < fhh) {
                        ldg_stream_cols<CPT>(xr[r][q], src + (size_t)(PR + q) * K);
                    } else {
#pragma unroll
                        for (int c = 0; c < CPT; ++c) xr[r][q][c] = cmake(0.f, 0.f);
                    }
                }
            }
        }
    };

    int g = blockIdx.x;
    if (g < n_groups) {
        if (tid == 0) load_head(g);
        load_rest(g);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_groups; g += gridDim.x) {
        const int fh = min(F, n_frames - g * F);
        const int gn = g + gridDim.x;
        mbar_wait(bar_p, phase);
        STAGE_MARK(16) // wait for the bulk load
        // ---- column phase: x[n1 + K*n2] -> registers (lanes = consecutive columns) for every round first, so
        // that P is free again (and refilled) as early as possible
        cpx v[ROUNDS][CPT][M];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int item = r * T + tid;
            const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
            if (item < S::COL_ITEMS) {
                const cpx* src = pre + (size_t)f * PR * K + n1;
#pragma unroll
                for (int n2 = 0; n2 < M; ++n2) {
                    cpx two[CPT];
                    if (n2 < PR) {
                        ld_cols<CPT>(two, src + n2 * K);
                    } else {
#pragma unroll
                        for (int c = 0; c < CPT; ++c) two[c] = xr[r][n2 - PR < XR ? n2 - PR : 0][c];
                    }
#pragma unroll
                    for (int c = 0; c < CPT; ++c) v[r][c][n2] = two[c];
                }
            }
        }
        __syncthreads(); // P consumed: it is refilled below, after the first round's table loads are on their way
        STAGE_MARK(17) // column reads
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int item = r * T + tid;
            const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
            const bool col_on = item < S::COL_ITEMS;
            if (col_on) {
                // the table column is requested first: its latency hides behind the M-point FFT
                cpx tc[M][CPT];
#pragma unroll
                for (int m = 0; m < M; ++m) ldg_cols<CPT>(tc[m], table + m * K + n1);
#pragma unroll
                for (int c = 0; c < CPT; ++c) rf::FFTN<M, -1>::run(v[r][c]);
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int c = 0; c < CPT; ++c) v[r][c][m] = cmul(v[r][c][m], tc[m][c]);
                if (!EQ && mode == 0) {
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        rf::FFTN<M, +1>::run(v[r][c]);
#pragma unroll
                        for (int t = 0; t < M; ++t) v[r][c][t] = cscale(v[r][c][t], inv_m);
                    }
                }
            }
            if (r == 0) {
                // S is written from here on: thread 0 signals through an mbarrier once the previous bulk store
                // has finished reading S (long done by now)
                STAGE_MARK(21) // column round 0 FFTs
                if (tid == 0) {
                    if (gn < n_groups) {
                        fence_proxy_async();
                        load_head(gn);
                    }
                    bulk_wait_read();
                    mbar_arrive(bar_s);
                }
                mbar_wait(bar_s, phase);
                STAGE_MARK(22) // wait until the previous bulk store has read S
            }
            if (col_on) {
                // column write: S[f*N + n1*M + t], CPT*M contiguous values per thread
                cpx* dst = s + (size_t)f * N + (size_t)n1 * M;
                if constexpr (CPT == 2) {
                    cpx lin[2 * M];
#pragma unroll
                    for (int t = 0; t < M; ++t) {
                        lin[t] = v[r][0][t];
                        lin[M + t] = v[r][1][t];
                    }
#pragma unroll
                    for (int q = 0; q < M; ++q)
                        reinterpret_cast<float4*>(dst)[q] =
                            make_float4(lin[2 * q].x, lin[2 * q].y, lin[2 * q + 1].x, lin[2 * q + 1].y);
                } else {
#pragma unroll
                    for (int t = 0; t < M; ++t) dst[t] = v[r][0][t];
                }
            }
        }
        phase ^= 1;
        if (gn < n_groups) load_rest(gn); // tail sample rows of the next group: in flight during the row phase
        __syncthreads();
        STAGE_MARK(18) // column phase
        // ---- row phase: FFT_K over n1, in place; the result is the [k][t] (or [k][m]) output order
        row_phase<S, -1, false>(s, nullptr, tw_s, tid, bar_k, nullptr, phase ^ 1, zero, [] {});
        STAGE_MARK(19) // row phase
        if constexpr (EQ) {
            __syncthreads();
            // Y[b*M+m] / H[b*M+m], consecutive lanes on consecutive bin pairs; every channel load is
            // issued before the first division so that their latencies overlap
            {
                constexpr int PAIRS = S::GROUP / 2;
                constexpr int PER = (PAIRS + T - 1) / T;
                const float4* eqg = reinterpret_cast<const float4*>(eq + (size_t)g * F * N);
                float4* s4 = reinterpret_cast<float4*>(s);
                float4 hq[PER];
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int i = tid + q * T;
                    hq[q] = (i < PAIRS && 2 * i < fh * N) ? __ldcs(eqg + i) : make_float4(1.f, 0.f, 1.f, 0.f);
                }
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int i = tid + q * T;
                    if (i < PAIRS) {
                        const float4 y = s4[i], hh = hq[q];
                        const float r0 = __fdividef(1.0f, hh.x * hh.x + hh.y * hh.y);
                        const float r1 = __fdividef(1.0f, hh.z * hh.z + hh.w * hh.w);
                        const cpx a = cmulc(cmake(y.x, y.y), cmake(hh.x, hh.y)); // y * conj(h) / |h|^2
                        const cpx b = cmulc(cmake(y.z, y.w), cmake(hh.z, hh.w));
                        s4[i] = make_float4(a.x * r0, a.y * r0, b.x * r1, b.y * r1);
                    }
                }
            }
            __syncthreads();
            // per subcarrier k (same thread <-> column index mapping as above): combine the L neighbouring parts
            const int h = L / 2;
            cpx res[ROUNDS][CPT][M];
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int item = r * T + tid;
                const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
                if (item < S::COL_ITEMS) {
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
#pragma unroll
                        for (int m = 0; m < M; ++m) res[r][c][m] = cmake(0.f, 0.f);
                        for (int i = 0; i < L; ++i) {
                            int kk = n1 + c + i - h;
                            kk = kk < 0 ? kk + K : (kk >= K ? kk - K : kk);
                            const cpx* src = s + (size_t)f * N + (size_t)kk * M;
                            const cpx* tp = taps_s + ((i + h) % L) * M;
#pragma unroll
                            for (int m = 0; m < M; ++m) res[r][c][m] = cadd(res[r][c][m], cmul(tp[m], src[m]));
                        }
                        if (mode == 0) {
                            rf::FFTN<M, +1>::run(res[r][c]);
#pragma unroll
                            for (int t = 0; t < M; ++t) res[r][c][t] = cscale(res[r][c][t], inv_m);
                        }
                    }
                }
            }
            __syncthreads(); // every neighbour read is done: overwrite S with the result
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int item = r * T + tid;
                const int f = item / S::COLS, n1 = (item - f * S::COLS) * CPT;
                if (item < S::COL_ITEMS) {
                    cpx* dst = s + (size_t)f * N + (size_t)n1 * M;
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
#pragma unroll
                        for (int t = 0; t < M; ++t) dst[c * M + t] = res[r][c][t];
                }
            }
        }
#ifdef GFDM_RX_STG_STORE
        __syncthreads();
        {   // experiment: drain S with ordinary 128-bit stores instead of a bulk store
            constexpr int Q4 = S::GROUP / 2;
            const float4* s4 = reinterpret_cast<const float4*>(s);
            cpx* og = out + (size_t)g * F * N;
#pragma unroll 5
            for (int i = tid; i < Q4; i += T)
                if (2 * i < fh * N) {
                    const float4 u = s4[i];
                    stg_stream4(og + 2 * i, cmake(u.x, u.y), cmake(u.z, u.w));
                }
        }
        __syncthreads();
#else
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) bulk_store(out + (size_t)g * F * N, s, (uint32_t)fh * N * sizeof(cpx));
#endif
        STAGE_MARK(20) // equalise/combine + store issue
    }
    if (tid == 0) bulk_wait_all();
}

// ----------------------------------------------------------------------------------------
// host side
typedef void (*mod_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, int, int, cudaStream_t);
typedef void (*rx_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, const cpx*, const cpx*, int, int, int, int,
                            cudaStream_t);

template <class S>
static void launch_mod(cpx* out, const cpx* in, const cpx* table, const cpx* tw, int n_frames, int grid,
                       cudaStream_t s)
{
    fused_mod_kernel<S><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, table, tw, n_frames);
}
template <class S>
static void launch_rx(cpx* out, const cpx* in, const cpx* eq, const cpx* table, const cpx* tw, const cpx* taps, int L,
                      int mode, int n_frames, int grid, cudaStream_t s)
{
    if (eq)
        fused_rx_kernel<S, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, mode, n_frames);
    else
        fused_rx_kernel<S, false><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, mode, n_frames);
}

struct ShapeEntry {
    int M, K, R, T, F;
    bool single;
    size_t smem;
    const char* mod_name;
    const char* rx_name;
    mod_launch_t mod;
    rx_launch_t rx;
    const void* mod_fn;
    const void* rx_fn;
    const void* rxeq_fn;
};

template <class S>
static ShapeEntry make_entry(const char* mn, const char* rn)
{
    ShapeEntry e;
    e.M = S::M; e.K = S::K; e.R = S::R; e.T = S::T; e.F = S::F;
    e.single = S::SINGLE;
    e.smem = S::SMEM_BYTES;
    e.mod_name = mn;
    e.rx_name = rn;
    e.mod = &launch_mod<S>;
    e.rx = &launch_rx<S>;
    e.mod_fn = (const void*)&fused_mod_kernel<S>;
    e.rx_fn = (const void*)&fused_rx_kernel<S, false>;
    e.rxeq_fn = (const void*)&fused_rx_kernel<S, true>;
    return e;
}

#define GFDM_SHAPE(M, R, SINGLE, T, F, CPT, MINB)                                                          \
    make_entry<Shape<M, R, SINGLE, T, F, CPT, MINB>>("fused_mod_kernel<M=" #M ",R=" #R ",T=" #T ",F=" #F ">", \
                                                      "fused_rx_kernel<M=" #M ",R=" #R ",T=" #T ",F=" #F ">")

static const std::vector<ShapeEntry>& shape_table()
{
    static const std::vector<ShapeEntry> t = {
        GFDM_SHAPE(5, 16, true, 256, 32, 2, 3),  // K=16   (BASELINE config 1)
        GFDM_SHAPE(9, 8, false, 256, 8, 2, 3),   // K=64   (config 2)
        GFDM_SHAPE(15, 16, false, 256, 1, 1, 3), // K=256  (config 4)
        GFDM_SHAPE(15, 32, false, 512, 1, 1, 1), // K=1024 (config 3, headline)
    };
    return t;
}

struct FusedImpl {
    const ShapeEntry* e = nullptr;
    int M = 0, K = 0, L = 0;
    cpx* d_table = nullptr;    // tx: C_tx ; rx: C_rx (taps folded)
    cpx* d_table_eq = nullptr; // rx only: plain twiddle
    cpx* d_tw = nullptr;
    cpx* d_taps = nullptr;
    int mod_grid_cap = 0, rx_grid_cap = 0;
};

static const ShapeEntry* find_shape(int M, int K)
{
    for (const ShapeEntry& e : shape_table())
        if (e.M == M && e.K == K) return &e;
    return nullptr;
}

static int grid_cap(const void* fn, int threads, size_t smem)
{
    int dev = 0, sms = 0, per_sm = 0;
    GFDM_CUDA_CHECK(cudaGetDevice(&dev));
    GFDM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GFDM_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GFDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) throw CudaError("fused kernel does not fit on this device");
    return sms * per_sm;
}

// W_K^{n0*k1} as [k1][n0]
static std::vector<cpx> make_tw(int R, bool single)
{
    if (single) return std::vector<cpx>(1, make_float2(1.f, 0.f));
    const int K = R * R;
    std::vector<cpx> tw((size_t)K);
    for (int k1 = 0; k1 < R; ++k1)
        for (int n0 = 0; n0 < R; ++n0) {
            const double ph = -2.0 * M_PI * (double)((long)n0 * k1 % K) / (double)K;
            tw[(size_t)k1 * R + n0] = make_float2((float)std::cos(ph), (float)std::sin(ph));
        }
    return tw;
}

// sign = +1: modulator table (incl. 1/N and part_len); sign = -1: receiver table; with_taps = false: twiddle only
static std::vector<cpx> make_table(int M, int K, int L, const std::vector<std::complex<float>>& taps, int sign,
                                   bool with_taps)
{
    const int N = M * K, h = L / 2;
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M;
    std::vector<cpx> t((size_t)N);
    for (int m = 0; m < M; ++m)
        for (int n1 = 0; n1 < K; ++n1) {
            std::complex<double> G(1.0, 0.0);
            if (with_taps) {
                G = 0.0;
                for (int i = 0; i < L; ++i) {
                    const std::complex<double> tp(taps[((i + h) % L) * M + m].real(), taps[((i + h) % L) * M + m].imag());
                    long e = ((long)(i - h) * n1) % K;
                    if (e < 0) e += K;
                    G += tp * std::polar(1.0, sign * 2.0 * M_PI * (double)e / (double)K);
                }
                if (sign > 0 && m >= part_len) G = 0.0;
            }
            const std::complex<double> w = std::polar(1.0, sign * 2.0 * M_PI * (double)((long)m * n1 % N) / (double)N);
            std::complex<double> c = G * w;
            if (sign > 0) c /= (double)N;
            t[(size_t)m * K + n1] = make_float2((float)c.real(), (float)c.imag());
        }
    return t;
}

static std::vector<cpx> to_cpx(const std::vector<std::complex<float>>& v)
{
    std::vector<cpx> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = make_float2(v[i].real(), v[i].imag());
    return o;
}

void FusedModem::init_tx(int M, int K, int L, const std::vector<std::complex<float>>& taps)
{
    destroy();
    const ShapeEntry* e = find_shape(M, K);
    if (!e || L < 1) return;
    FusedImpl* p = new FusedImpl;
    p->e = e; p->M = M; p->K = K; p->L = L;
    try {
        p->mod_grid_cap = grid_cap(e->mod_fn, e->T, e->smem);
        p->d_table = upload(make_table(M, K, L, taps, +1, true));
        p->d_tw = upload(make_tw(e->R, e->single));
    } catch (...) {
        impl_ = p;
        destroy();
        throw;
    }
    impl_ = p;
}

void FusedModem::init_rx(int M, int K, int L, const std::vector<std::complex<float>>& taps,
                         const std::vector<std::complex<float>>&)
{
    destroy();
    const ShapeEntry* e = find_shape(M, K);
    if (!e || L < 2 || L * M > 64) return; // the kernel keeps at most 64 receive taps in shared memory
    FusedImpl* p = new FusedImpl;
    p->e = e; p->M = M; p->K = K; p->L = L;
    try {
        p->rx_grid_cap = std::min(grid_cap(e->rx_fn, e->T, e->smem), grid_cap(e->rxeq_fn, e->T, e->smem));
        p->d_table = upload(make_table(M, K, L, taps, -1, true));
        p->d_table_eq = upload(make_table(M, K, L, taps, -1, false));
        p->d_tw = upload(make_tw(e->R, e->single));
        p->d_taps = upload(to_cpx(taps));
    } catch (...) {
        impl_ = p;
        destroy();
        throw;
    }
    impl_ = p;
}

int FusedModem::modulate(cpx* out, const cpx* in, size_t frames, cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    int launches = 0;
    const size_t max_chunk = (size_t)1 << 20; // keep frame counts in int range
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->mod_grid_cap ? groups : impl_->mod_grid_cap;
        e->mod(out + f0 * (size_t)e->M * e->K, in + f0 * (size_t)e->M * e->K, impl_->d_table, impl_->d_tw, nf, grid, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

int FusedModem::demodulate(cpx* out_td, cpx* out_fd, const cpx* in, const cpx* eq, size_t frames, cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    int launches = 0;
    const size_t N = (size_t)e->M * e->K;
    const size_t max_chunk = (size_t)1 << 20;
    for (int pass = 0; pass < 2; ++pass) {
        cpx* out = pass == 0 ? out_td : out_fd;
        if (!out) continue;
        for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
            const int nf = (int)std::min(max_chunk, frames - f0);
            const int groups = (nf + e->F - 1) / e->F;
            const int grid = groups < impl_->rx_grid_cap ? groups : impl_->rx_grid_cap;
            e->rx(out + f0 * N, in + f0 * N, eq ? eq + f0 * N : nullptr, eq ? impl_->d_table_eq : impl_->d_table,
                  impl_->d_tw, impl_->d_taps, impl_->L, pass, nf, grid, s);
            ++launches;
        }
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

const char* FusedModem::mod_name() const { return impl_ ? impl_->e->mod_name : "none"; }
const char* FusedModem::rx_name() const { return impl_ ? impl_->e->rx_name : "none"; }

#ifdef GFDM_PROFILE_STAGES
extern "C" __attribute__((visibility("default"))) int gfdm_debug_stage_cycles(unsigned long long* out32, int reset)
{
    if (out32 && cudaMemcpyFromSymbol(out32, g_stage_cycles, sizeof(g_stage_cycles)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[32] = { 0 };
        if (cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

void FusedModem::destroy()
{
    if (!impl_) return;
    if (impl_->d_table) cudaFree(impl_->d_table);
    if (impl_->d_table_eq) cudaFree(impl_->d_table_eq);
    if (impl_->d_tw) cudaFree(impl_->d_tw);
    if (impl_->d_taps) cudaFree(impl_->d_taps);
    delete impl_;
    impl_ = nullptr;
}

} // namespace gfdm
