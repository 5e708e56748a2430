// fused_dev.cuh -- device-side building blocks shared by the fused kernels (fused_modem.cu: frame
// resident in shared memory; fused_twopass.cu: frame larger than shared memory): mbarrier / bulk-copy
// (TMA) PTX wrappers, streaming global accesses, the compile-time Shape and the shared-memory row FFT.
#pragma once
#include "common.cuh"
#include "regfft.cuh"

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
// L2 eviction policies (createpolicy): evict_last keeps data that is read again soon (second pass of the
// two-pass kernels, their per-CTA scratch), evict_first marks data that is dead after this access
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// bring a global range into L2 ahead of the bulk load that will need it at short notice
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void stg_hint(cpx* p, cpx v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ cpx ldg_hint(const cpx* p, uint64_t pol)
{
    float2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stg_stream4(cpx* p, cpx a, cpx b)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
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

__device__ __forceinline__ cpx ldg_nc(const cpx* p)
{
    return __ldg(reinterpret_cast<const float2*>(p));
}
// streaming global accesses: every input byte is read once, every output byte written once
__device__ __forceinline__ cpx ldg_stream(const cpx* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(cpx* p, cpx v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ----------------------------------------------------------------------------------------
// Tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bit) as a per-thread constant store.
// The fused kernels have no MMA, so TMEM is idle; the 32x32b access shape gives every thread of a warp its own
// lane (lane = 32*(warp%4) + laneid) and N consecutive columns -- a software-managed extension of the register
// file with ~12 cycles of latency on a datapath that is neither the LSU nor L2.  Used for the loop-invariant
// folded filter/twiddle table column(s) of each thread when the table does not fit in shared memory.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) // one converged warp; ncols = 2^n >= 32
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) // the warp that allocated
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x2(float* r, uint32_t a)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st_x2(uint32_t a, const float* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "f"(r[0]), "f"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(float* r, uint32_t a)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st_x4(uint32_t a, const float* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(float* r, uint32_t a)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t a, const float* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(float* r, uint32_t a)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t a, const float* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}
// N (even) consecutive 32-bit columns <-> registers, as power-of-two pieces (whole warp, converged)
template <int N>
__device__ __forceinline__ void tmem_ld(float* r, uint32_t a)
{
    static_assert(N % 2 == 0 && N >= 0, "column count must be even");
    if constexpr (N >= 16) { tmem_ld_x16(r, a); tmem_ld<N - 16>(r + 16, a + 16); }
    else if constexpr (N >= 8) { tmem_ld_x8(r, a); tmem_ld<N - 8>(r + 8, a + 8); }
    else if constexpr (N >= 4) { tmem_ld_x4(r, a); tmem_ld<N - 4>(r + 4, a + 4); }
    else if constexpr (N >= 2) { tmem_ld_x2(r, a); tmem_ld<N - 2>(r + 2, a + 2); }
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t a, const float* r)
{
    static_assert(N % 2 == 0 && N >= 0, "column count must be even");
    if constexpr (N >= 16) { tmem_st_x16(a, r); tmem_st<N - 16>(a + 16, r + 16); }
    else if constexpr (N >= 8) { tmem_st_x8(a, r); tmem_st<N - 8>(a + 8, r + 8); }
    else if constexpr (N >= 4) { tmem_st_x4(a, r); tmem_st<N - 4>(a + 4, r + 4); }
    else if constexpr (N >= 2) { tmem_st_x2(a, r); tmem_st<N - 2>(a + 2, r + 2); }
}

// ----------------------------------------------------------------------------------------
// optional per-stage cycle counters (build with -DGFDM_PROFILE_STAGES; tools/stage_profile.py)
#ifdef GFDM_PROFILE_STAGES
static __device__ unsigned long long g_stage_cycles[32];
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
// compile-time shape of one fused kernel
// TW_TMEM_ = false keeps the pass-1 row-FFT twiddles in shared memory even when the table lives in tensor memory, and
// KEEP_COLS_ reserves that many extra tensor-memory columns per thread: the interference-cancelling receiver of the
// shapes with two subcarriers per thread parks its kept frequency blocks there (fused_kernels.cuh).
template <int M_, int R1_, int R2_, int T_, int IPT_, int MINB_, bool TMEM_ = true, bool TW_TMEM_ = true, int KEEP_COLS_ = 0>
struct Shape {
    static constexpr int M = M_, R1 = R1_, R2 = R2_, T = T_, IPT = IPT_, MINB = MINB_;
    static constexpr int K = R1 * R2;
    static constexpr int N = M * K;
    static constexpr int F = IPT * T / K; // frames per CTA pass
    static_assert(IPT * T % K == 0 && F >= 1, "threads x items must cover whole frames");
    static_assert(32 % R1 == 0, "R1 must divide the warp");
    static_assert(R2 == 1 || ((R2 & (R2 - 1)) == 0 && R2 <= 32), "R2 must be a power of two <= 32");
    static constexpr bool TWO_PASS = R2 > 1;
    // Rows of the K-point stage.  Two-pass rows are padded by one element per R2 block: element
    // n = R2*q + r lives at (R2+1)*q + r, which makes pass 1 (lanes = r), pass 2 (lanes = q) and the
    // producer (lanes = consecutive n) bank-conflict free with compile-time (immediate) offsets -- an
    // XOR swizzle would save the padding but costs 32 live address registers in the radix-32 passes.
    // Single-pass rows get an odd stride.
    static constexpr int RS = TWO_PASS ? R1 * (R2 + 1) : (K | 1);
    static constexpr int ROWS = F * M;
    static constexpr int ROW_ELEMS = ROWS * RS;
    static constexpr int STAGE_ELEMS = F * N;
    static constexpr int BUF_ELEMS = ROW_ELEMS > STAGE_ELEMS ? ROW_ELEMS : STAGE_ELEMS;
    // per-CTA shared memory budget in complex elements (228 KB per SM, 1 KB per CTA reserved)
    // small constants of the receiver: [0,64) receive taps of the equalising path (L*M <= 64),
    // [64,96) interference-cancellation taps, [96,160) constellation points, [160,192) reduction scratch
    static constexpr int TAPS_ELEMS = 192;
    static constexpr int IC_OFF = 64, PTS_OFF = 96, RED_OFF = 160, MAX_POINTS = 64;
    static constexpr int BUDGET_ELEMS = ((233472 / MINB) - 1024) / 8 - 8 - TAPS_ELEMS;
    // the folded filter/twiddle table stays resident in shared memory when the whole next group fits beside it
    // (and beside the row-FFT twiddles) ...
    static constexpr bool TBL_SMEM = BUDGET_ELEMS - BUF_ELEMS - (TWO_PASS ? K : 0) >= F * N + N;
    // ... otherwise every thread parks its own loop-invariant constants in tensor memory: its IPT table columns
    // (2M 32-bit columns each) and, for two-pass rows, its R1 pass-1 twiddles W_K^{n0*k1}, n0 = tid % R2.
    // Warp w owns lanes 32*(w%4).., the warps sharing a lane quarter take consecutive column blocks.
    static constexpr bool TBL_TMEM = !TBL_SMEM && TMEM_; // the two-pass kernels (fused_twopass.cu) opt out
    static constexpr bool TW_TMEM = TBL_TMEM && TWO_PASS && T % R2 == 0 && TW_TMEM_;
    static constexpr int TW_ELEMS = (TWO_PASS && !TW_TMEM) ? K : 0;
    static constexpr int P_MAX = BUDGET_ELEMS - BUF_ELEMS - TW_ELEMS;
    static_assert(P_MAX >= 0, "frame group does not fit in shared memory");
    static constexpr int TBL_ELEMS = TBL_SMEM ? N : 0;
    static constexpr int P_AVAIL = P_MAX - TBL_ELEMS;
    // prefetch region P: modulator -> the first PF staged elements of the next group,
    //                    receiver  -> the first PR sample rows (n2) of every frame of the next group
    static constexpr int PF = (F * N) < (P_AVAIL / (2 * M)) * 2 * M ? (F * N) : (P_AVAIL / (2 * M)) * 2 * M;
    static constexpr int PR = M < P_AVAIL / (F * K) ? M : P_AVAIL / (F * K);
    static constexpr int P_ELEMS = PF > F * PR * K ? PF : F * PR * K;
    static constexpr int TMEM_TBL_COLS = IPT * 2 * M;             // per thread
    static constexpr int TMEM_TW_COLS = TW_TMEM ? 2 * R1 : 0;     // per thread
    static constexpr int TMEM_KEEP_COLS = KEEP_COLS_;             // per thread
    static constexpr int TMEM_PER_THREAD = TMEM_TBL_COLS + TMEM_TW_COLS + TMEM_KEEP_COLS;
    static constexpr int TMEM_USED = ((T / 32 + 3) / 4) * TMEM_PER_THREAD;
    static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static_assert(!TBL_TMEM || (TMEM_USED <= 512 && TMEM_COLS * MINB <= 512), "constants do not fit in tensor memory");
    static constexpr size_t SMEM_BYTES = sizeof(cpx) * (size_t)(BUF_ELEMS + TW_ELEMS + TBL_ELEMS + P_ELEMS + TAPS_ELEMS) + 64;
    __host__ __device__ static constexpr int swz(int n) { return TWO_PASS ? n + n / R2 : n; } // row-FFT input slot
};

// Tensor-memory setup of a fused kernel: allocate, then every thread stores its own table columns (and pass-1
// twiddles).  Returns the allocation base (for the dealloc) and this thread's first column.
template <class S>
__device__ __forceinline__ void tmem_setup(uint32_t* slot, const cpx* __restrict__ table, const cpx* __restrict__ tw, int tid,
                                           uint32_t& base, uint32_t& mine)
{
    constexpr int M = S::M, K = S::K, T = S::T;
    if (tid < 32) tmem_alloc(slot, S::TMEM_COLS);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    base = *slot;
    mine = base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)(tid >> 7) * S::TMEM_PER_THREAD;
#pragma unroll
    for (int j = 0; j < S::IPT; ++j) {
        const int n1 = (tid + j * T) % K;
        float tf[2 * M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const cpx c = ldg_nc(table + m * K + n1);
            tf[2 * m] = c.x;
            tf[2 * m + 1] = c.y;
        }
        tmem_st<2 * M>(mine + j * 2 * M, tf);
    }
    if constexpr (S::TW_TMEM) {
        static_assert(S::R1 % 8 == 0, "pass-1 twiddles are fetched in groups of eight");
        const int n0 = tid % S::R2;
        float wf[2 * S::R1];
#pragma unroll
        for (int k1 = 0; k1 < S::R1; ++k1) {
            const cpx c = ldg_nc(tw + k1 * S::R2 + n0);
            wf[2 * k1] = c.x;
            wf[2 * k1 + 1] = c.y;
        }
        tmem_st<2 * S::R1>(mine + S::TMEM_TBL_COLS, wf);
    }
    tmem_wait_st();
}

// Row FFTs over the K-long rows held in shared memory (in place; input at swz(), output natural).
template <class S, int DIR>
__device__ __forceinline__ void row_fft(cpx* __restrict__ buf, const cpx* __restrict__ tw_s, int tid, uint32_t tmem_tw = 0)
{
    constexpr int R1 = S::R1, R2 = S::R2, RS = S::RS, T = S::T;
    if constexpr (!S::TWO_PASS) {
        for (int it = tid; it < S::ROWS; it += T) {
            cpx* row = buf + it * RS;
            cpx a[R1];
#pragma unroll
            for (int i = 0; i < R1; ++i) a[i] = row[i];
            rf::FFTN<R1, DIR>::run(a);
#pragma unroll
            for (int i = 0; i < R1; ++i) row[i] = a[i];
        }
    } else {
        // pass 1: item (row, n0): radix-R1 over x[R2*n1 + n0], then W_K^{n0*k1}; in place
        constexpr int ITEMS1 = S::ROWS * R2;
        // Twiddles in tensor memory: tcgen05.ld is a warp-wide (.sync.aligned) instruction, so a warp whose last round is
        // only partly populated (ITEMS1 not a multiple of 32) must still execute it with every lane -- the surplus lanes
        // run the round without touching shared memory.  Whole warps still drop out.
        constexpr bool PARTIAL_WARP = S::TW_TMEM && ITEMS1 % 32 != 0;
        constexpr int ROUNDS1 = (ITEMS1 + T - 1) / T;
#pragma unroll 1
        for (int rd = 0; rd < ROUNDS1; ++rd) {
            int it = tid + rd * T;
            bool live = true;
            if constexpr (PARTIAL_WARP) {
                if ((it & ~31) >= ITEMS1) break; // no lane of this warp has an item
                live = it < ITEMS1;
                it = live ? it : ITEMS1 - 1;
            } else {
                if (it >= ITEMS1) break;
            }
            const int row = it / R2, n0 = it - row * R2;
            cpx* p = buf + row * RS + n0;
            cpx a[R1];
#pragma unroll
            for (int i = 0; i < R1; ++i) a[i] = live ? p[(R2 + 1) * i] : cmake(0.f, 0.f); // (surplus lanes touch no data)
            rf::FFTN<R1, DIR>::run(a);
            // twiddle + store in chunks of 8; the empty asm keeps the chunks in program order so that
            // at most 8 twiddles are live next to the 2*R1 data registers (no spills in this hot loop)
            if (live) p[0] = a[0];
#pragma unroll
            for (int c = 0; c < R1; c += 8) {
                cpx w[8];
                if constexpr (S::TW_TMEM) {
                    float wf[16]; // this thread's twiddles k1 = c .. c+7 out of tensor memory
                    tmem_ld_x16(wf, tmem_tw + 2 * c);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = cmake(wf[2 * i], wf[2 * i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c + i > 0 && c + i < R1) w[i] = tw_s[(c + i) * R2 + n0];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c + i > 0 && c + i < R1) {
                        if (DIR > 0) w[i].y = -w[i].y;
                        if (live) p[(R2 + 1) * (c + i)] = cmul(a[c + i], w[i]);
                    }
                asm volatile("" ::: "memory");
            }
        }
        if constexpr (R1 == R2) __syncwarp(); else __syncthreads();
        // pass 2: item (row, k1): radix-R2 over A[n0][k1]; result X[k1 + R1*k0] stored at natural index.
        // A row is owned by one warp (R1 divides 32), so a warp-level barrier orders its reads and writes.
        constexpr int ITEMS2 = S::ROWS * R1;
        constexpr int ITERS2 = (ITEMS2 + T - 1) / T;
#pragma unroll 1
        for (int ii = 0; ii < ITERS2; ++ii) {
            int it = tid + ii * T;
            if constexpr (ITEMS2 % 32 == 0) {
                if (it >= ITEMS2) break; // whole warps drop out together
            } else {
                // whole warps without an item drop out; the surplus lanes of the one partly populated warp redo the
                // last item (same reads, same values written, same warp as its owner): no divergence around the warp
                // barrier and b[] stays in registers
                if ((it & ~31) >= ITEMS2) break;
                it = it < ITEMS2 ? it : ITEMS2 - 1;
            }
            const int row = it / R1, k1 = it - row * R1;
            cpx* r = buf + row * RS;
            cpx b[R2];
#pragma unroll
            for (int i = 0; i < R2; ++i) b[i] = r[(R2 + 1) * k1 + i];
            rf::FFTN<R2, DIR>::run(b);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < R2; ++i) r[k1 + R1 * i] = b[i];
        }
    }
}

} // namespace gfdm
