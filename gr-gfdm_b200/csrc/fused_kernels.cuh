// fused_kernels.cuh -- single-kernel GFDM modulator and receiver, one frame group
// resident in shared memory per CTA (sm_100a).
//
// Factorisation (checked in NumPy by tools/fused_math_check.py).
// With bin index b*M+m and sample index n1 + K*n2:
//
//   modulator  (replaces lib/modulator_kernel_cc.cc:98-141)
//     D_b[m]   = FFT_M(d_b)                       stage A   thread <-> subcarrier, registers
//     Z_m[n1]  = IFFT_K over b of D_b[m]          stage B   row FFTs in shared memory
//     x[n1+K*n2] = IFFT_M over m of C_tx[m][n1]*Z_m[n1]     stage C   thread <-> n1, registers
//   where C_tx[m][n1] = (sum_i T[((i+h)%L)M+m] e^{+j2pi(i-h)n1/K}) e^{+j2pi m n1/N} / N folds the
//   L-fold spectral repetition, the pulse-shaping taps, the scatter-add into the N-bin grid,
//   the N-point twiddle and the 1/N scale into ONE table multiply (a circular shift over b is a
//   phase ramp over n1).
//
//   receiver   (replaces lib/receiver_kernel_cc.cc:165-225,301-334)
//     U_n1[m]  = FFT_M over n2 of x[n1+K*n2]      stage A'  thread <-> n1, registers
//     V_m[k]   = FFT_K over n1 of C_rx[m][n1]*U_n1[m]       stage B
//     y_k      = IFFT_M(R_k)/M, R_k[m] = V_m[k]             stage C'  thread <-> subcarrier
//   without equalisation C_rx carries the receive taps as well; with equalisation C_rx is the
//   plain twiddle, Y = V / H_eq is formed bin by bin and the L neighbouring parts are combined
//   explicitly (the division sits between FFT and filter, receiver_kernel_cc.cc:315-319).
//
// Data movement: [k][m]-ordered arrays (modulator input, receiver output) cross HBM through
// cp.async.bulk (TMA 1D) into / out of shared memory, so every byte moves in 128 B lines;
// [n2][n1]-ordered arrays (time samples) are accessed directly, lanes = consecutive n1.
// The CTA is persistent (grid = SMs x occupancy) and the modulator issues the bulk load of the
// next frame group as soon as the last shared-memory read of the current one has retired.#pragma once
#include "fused_dev.cuh"
#include "fused_shapes.h"

#include <cmath>
#include <cstring>
#include <string>
#include <type_traits>

namespace gfdm {

// transmitter chain, stage C: the few output positions that carry a window ramp and/or are written twice (cyclic
// prefix / suffix).  Out of line so that the unrolled common case stays one compare and one store.
static __device__ __noinline__ void tx_store_edge(cpx* o, int i, cpx v, int N, int W, int ramp, const cpx* front,
                                                  const cpx* back, bool shaped, cpx scale)
{
    for (; i < W; i += N) {
        cpx val = v;
        if (i < ramp) val = cmul_rn(val, __ldg(front + i));
        if (i >= W - ramp) val = cmul_rn(val, __ldg(back + (i - (W - ramp))));
        if (shaped) val = cmul_rn(val, scale); // short_burst_shaper: after the window, as the block chain orders them
        stg_stream(o + i, val);
    }
}

// ----------------------------------------------------------------------------------------
// Fused modulator.  in/out: [n_frames][N]; table: C_tx [M][K]; tw: W_K^{n0*k1} as [k1][n0].
// Shared memory: R = row buffer (also holds the tail of the staged input), P = prefetch region
// holding the head of the NEXT group's staged input, loaded by TMA while this group is processed.
//
// TXF = true is the whole transmitter_kernel::generic_work chain (lib/transmitter_kernel.cc:78-107) in this one
// kernel: the resource mapper (lib/resource_mapper_kernel_cc.cc:108-134) becomes the gather of stage A out of the
// staged COMPACT symbol vectors (in: [n_frames][n_in]), and preamble insertion + cyclic prefix/suffix + window
// (lib/add_cyclic_prefix_cc.cc:67-98) become the store pattern of stage C (out: [n_ant][n_frames][P+cp+N+cs]),
// so a frame costs 8*(n_in + n_ant*(P+cp+N+cs)) bytes of HBM traffic instead of four kernels' worth.
//
// CHK = true: the symbol side arrives as CHUNKS, one byte per symbol = index of a constellation point
// (python/pygfdm/symbolmapping.py:34-38 / gr-digital chunks_to_symbols); `in` then points to bytes, the whole
// group is staged in P (F*EL bytes) and stage A looks the points up in shared memory, so the symbols cross HBM
// as EL instead of 8*EL bytes per frame.
template <class S, bool TXF, bool CHK = false>
__global__ void __launch_bounds__(S::T, S::MINB) fused_mod_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                                  const cpx* __restrict__ table,
                                                                  const cpx* __restrict__ tw, int n_frames,
                                                                  const __grid_constant__ TxArgs tx)
{
    constexpr int M = S::M, K = S::K, N = S::N, T = S::T, IPT = S::IPT, F = S::F, RS = S::RS, PF = S::PF;
    const int EL = TXF ? tx.n_in : N; // staged elements per frame
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* tbl_s = tw_s + S::TW_ELEMS;
    cpx* pre = tbl_s + S::TBL_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + S::P_ELEMS + S::TAPS_ELEMS);
    uint64_t* bar_p = bars;     // head of the staged input (region P)
    uint64_t* bar_r = bars + 1; // tail of the staged input (region R)
    const int tid = threadIdx.x;
    const int n_groups = (n_frames + F - 1) / F;
    // chunk input: staged bytes in P, constellation in the (otherwise unused) small-constant region
    const unsigned char* in_b = reinterpret_cast<const unsigned char*>(in);
    const unsigned char* pre_b = reinterpret_cast<const unsigned char*>(pre);
    cpx* pts_s = pre + S::P_ELEMS + S::PTS_OFF;
    if constexpr (CHK) {
        static_assert((size_t)F * N <= (size_t)S::P_ELEMS * sizeof(cpx), "a group of chunks must fit in region P");
        for (int i = tid; i < S::MAX_POINTS; i += T) pts_s[i] = i < tx.n_points ? tx.points[i] : cmake(0.f, 0.f);
    }
    const int n_pts = CHK ? tx.n_points : 0;
    auto lookup = [&](unsigned char c) { return (int)c < n_pts ? pts_s[c] : cmake(0.f, 0.f); }; // chunk >= n_points: 0+0j

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    for (int i = tid; i < S::TBL_ELEMS; i += T) tbl_s[i] = table[i];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_r, 1);
    }
    // table columns of this thread -> tensor memory (once per CTA)
    uint32_t tmem_base = 0, tmem_mine = 0;
    if constexpr (S::TBL_TMEM) {
        uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
        tmem_setup<S>(slot, table, tw, tid, tmem_base, tmem_mine);
    }
    __syncthreads();

    // issue the bulk loads of group gg: head -> P, tail -> R
    auto load_head = [&](int gg) {
        const int el = min(F, n_frames - gg * F) * EL;
        if constexpr (CHK) {
            mbar_expect_tx(bar_p, (uint32_t)el);
            if (el) bulk_load(pre, in_b + (size_t)gg * F * EL, (uint32_t)el, bar_p);
            return;
        }
        const uint32_t bytes = (uint32_t)min(el, PF) * sizeof(cpx);
        mbar_expect_tx(bar_p, bytes);
        if (bytes) bulk_load(pre, in + (size_t)gg * F * EL, bytes, bar_p);
        // the tail of the group can only be staged once the row buffer is free (after stage C's reads), which leaves it
        // little time to arrive (stage profile: 922 of 15.3k cycles per frame waiting at the loop top): bring it into L2
        // now so that the later bulk load is an L2 hit
        if (el > PF) bulk_prefetch_l2(in + (size_t)gg * F * EL + PF, (uint32_t)(el - PF) * sizeof(cpx));
    };
    auto load_tail = [&](int gg) {
        if constexpr (CHK) return;
        const int el = min(F, n_frames - gg * F) * EL;
        const uint32_t bytes = (uint32_t)max(el - PF, 0) * sizeof(cpx);
        mbar_expect_tx(bar_r, bytes);
        if (bytes) bulk_load(buf, in + (size_t)gg * F * EL + PF, bytes, bar_r);
    };
    // transmitter: position of this thread's subcarrier(s) in the sorted subcarrier map (-1: unused)
    // The gather of stage A is loop invariant: staged index of timeslot 0, stride over timeslots, and how many
    // timeslots lie inside the n_in symbols the caller supplied (the rest, and unused subcarriers, are zero).
    int g_e0[IPT], g_mv[IPT];
    const int g_st = tx.per_timeslot ? tx.A : 1;
    if constexpr (TXF) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T, f = it / K;
            const int a = __ldg(tx.inv_map + (it - f * K));
            const int s0 = tx.per_timeslot ? a : a * M; // src of timeslot 0
            int mv = 0;
            if (a >= 0 && s0 < tx.n_in) mv = min(M, (tx.n_in - s0 + g_st - 1) / g_st);
            g_e0[j] = f * tx.n_in + s0;
            g_mv[j] = mv;
        }
    }

    int g = blockIdx.x;
    if (tid == 0 && g < n_groups) {
        load_head(g);
        load_tail(g);
    }
    uint32_t phase = 0;
    STAGE_INIT();
    for (; g < n_groups; g += gridDim.x) {
        const int fh = min(F, n_frames - g * F);
        const int gn = g + gridDim.x;
        mbar_wait(bar_p, phase);
        // the tail of the staged group (region R) is only read by the items whose records lie behind PF: the plain
        // modulator waits for it right before the first of those (the transmitter's gather may touch it anywhere)
        constexpr int J_TAIL = PF / (T * M); // first item with a record at or behind PF
        constexpr bool LATE_TAIL = !TXF && !CHK && PF < F * N && J_TAIL >= 1 && J_TAIL < IPT;
        if (!CHK && !LATE_TAIL && PF < F * N) mbar_wait(bar_r, phase);
        STAGE_MARK(0) // wait for the bulk loads

        cpx v[IPT][M];
        // ---- stage A: subcarrier symbols from the staged [k][m] block -> registers
        if constexpr (!TXF && CHK) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const unsigned char* src = pre_b + (tid + j * T) * M;
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = lookup(src[m]);
            }
        } else if constexpr (!TXF) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                if (LATE_TAIL && j == J_TAIL) mbar_wait(bar_r, phase);
                const int e = (tid + j * T) * M; // (f*K + k)*M
                const cpx* src = (e < PF) ? pre + e : buf + (e - PF);
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = src[m];
            }
        } else {
            // map_to_resources as a gather: symbol (slot a, timeslot m) sits at m*A + a (per timeslot) or
            // a*M + m (per subcarrier) of the frame's compact vector; beyond n_in and on unused subcarriers: zero
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                int e = g_e0[j];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    // branch-free: out-of-range timeslots read staged element 0 and are zeroed afterwards
                    const bool ok = m < g_mv[j];
                    const int ec = ok ? e : 0;
                    cpx val;
                    if constexpr (CHK) val = lookup(pre_b[ec]);
                    else if constexpr (PF >= F * N) val = pre[ec]; // the whole staged group lives in P
                    else val = *((ec < PF ? pre : buf - PF) + ec);
                    v[j][m] = ok ? val : cmake(0.f, 0.f);
                    e += g_st;
                }
            }
        }
        phase ^= 1;
        __syncthreads(); // staging fully consumed: P may be refilled, R may take the rows
        // Issuing the bulk copy costs its warp a few hundred cycles and every other warp waits for it at the next barrier.
        // When the last warp has no row-FFT items (ROWS*R2 <= T-32, e.g. 480 of 512 threads at K = 1024) it issues the
        // copy during that phase instead, off everybody's critical path; the data still has three phases to arrive.
        constexpr bool IDLE_WARP = S::TWO_PASS && S::ROWS * S::R2 + 32 <= T;
        if (!IDLE_WARP && tid == 0 && gn < n_groups) {
            fence_proxy_async();
            load_head(gn);
        }
        STAGE_MARK(1) // stage A reads
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, k = it - f * K;
            rf::FFTN<M, -1>::run(v[j]);
            cpx* dst = buf + (size_t)f * M * RS + S::swz(k);
#pragma unroll
            for (int m = 0; m < M; ++m) dst[m * RS] = v[j][m];
        }
        __syncthreads();
        STAGE_MARK(2) // stage A FFT + row writes
        // ---- stage B: K-point inverse FFT of every row (over the subcarrier index)
        if (IDLE_WARP && tid == T - 32 && gn < n_groups) {
            fence_proxy_async();
            load_head(gn);
        }
        row_fft<S, +1>(buf, tw_s, tid, tmem_mine + S::TMEM_TBL_COLS);
        STAGE_MARK(3) // row FFT (warp 0's own time)
        __syncthreads();
        STAGE_MARK(4) // barrier after row FFT
        // table column of the first item (unless resident): issued after the barrier so that the
        // compiler cannot hoist it into the register-hungry radix-32 passes; the column reads hide it
        // ---- stage C: column n1 of all rows -> registers
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, n1 = it - f * K;
            const cpx* src = buf + (size_t)f * M * RS + n1;
#pragma unroll
            for (int m = 0; m < M; ++m) v[j][m] = src[m * RS];
        }
        STAGE_MARK(7) // stage C reads (own loads issued)
        __syncthreads(); // R is dead: fetch the tail of the next group while stage C computes and stores
        STAGE_MARK(8) // barrier behind the column reads
        // (this second bulk copy of the group costs thread 0 ~1400 cycles while the head copy is still in flight -- stage
        // profile r02h -- and every warp waits for it at the next barrier; reading the tail records straight from global
        // memory instead was measured slower, experiments/README.md)
        if (tid == 0 && gn < n_groups) {
            fence_proxy_async();
            load_tail(gn);
        }
        STAGE_MARK(5) // issue of the tail load
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, n1 = it - f * K;
            if constexpr (S::TBL_SMEM) {
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cmul(v[j][m], tbl_s[m * K + n1]);
            } else {
                float tf[2 * M]; // this item's table column out of tensor memory
                tmem_ld<2 * M>(tf, tmem_mine + j * 2 * M);
                tmem_wait_ld();
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cmul(v[j][m], cmake(tf[2 * m], tf[2 * m + 1]));
            }
            rf::FFTN<M, +1>::run(v[j]);
            if constexpr (!TXF) {
                if (f < fh) {
                    cpx* dst = out + ((size_t)g * F + f) * N + n1;
#pragma unroll
                    for (int n2 = 0; n2 < M; ++n2) stg_stream(dst + (size_t)n2 * K, v[j][n2]);
                }
            } else if (f < fh) {
                // add_cyclic_prefix: o[i] = x[(i + N - cp - s) mod N], i < W = N + cp + cs  <=>  sample n goes to
                // i = (n + cp + s) mod N and again to i + N while that is < W; ramps on the first / last samples.
                // Positions lo <= i < hi are neither ramped nor duplicated: one compare, one store.
                const int W = N + tx.cp + tx.cs, os = tx.pre_pad + tx.P + W + tx.post_pad, dup = tx.cp + tx.cs;
                const int lo = max(tx.ramp, dup), hi = max(lo, min(N, W - tx.ramp));
                const bool shaped = tx.shaped != 0;
                const cpx scale = cmake(tx.sc_re, tx.sc_im);
                // two copies of the store loop: the burst shaper's scaling must not cost the plain chain a select per store
                auto store_all = [&](auto SH) {
                    constexpr bool sh = decltype(SH)::value;
                    for (int a = 0; a < tx.n_ant; ++a) {
                        cpx* o = out + (size_t)a * tx.ant_stride + ((size_t)g * F + f) * os + tx.pre_pad + tx.P;
                        int i = (n1 + tx.cp + tx.shift[a]) % N;
#pragma unroll
                        for (int n2 = 0; n2 < M; ++n2) {
                            if ((unsigned)(i - lo) < (unsigned)(hi - lo)) stg_stream(o + i, sh ? cmul_rn(v[j][n2], scale) : v[j][n2]);
                            else tx_store_edge(o, i, v[j][n2], N, W, tx.ramp, tx.front, tx.back, sh, scale);
                            i += K;
                            i = (int)min((unsigned)i, (unsigned)(i - N)); // wrap at N without a branch (i < 2N)
                        }
                    }
                };
                if (shaped) store_all(std::true_type{});
                else store_all(std::false_type{});
            }
        }
        if constexpr (TXF) {
            // insert_preamble (lib/transmitter_kernel.cc:86-90): the shift's preamble in front of every frame
            const int W = N + tx.cp + tx.cs, os = tx.pre_pad + tx.P + W + tx.post_pad;
            const bool shaped = tx.shaped != 0;
            const cpx scale = cmake(tx.sc_re, tx.sc_im);
            for (int a = 0; a < tx.n_ant; ++a) {
                const cpx* p = tx.preambles + (size_t)tx.pre_idx[a] * tx.P;
                // 16-byte copies when every row start is 16-byte aligned (even P, even row length, aligned bases)
                const bool vec = !shaped && ((tx.P | os) & 1) == 0 && (tx.ant_stride & 1) == 0 &&
                                 ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(p)) & 15) == 0;
                for (int f = 0; f < fh; ++f) {
                    cpx* o = out + (size_t)a * tx.ant_stride + ((size_t)g * F + f) * os + tx.pre_pad;
                    if (shaped) {
                        // short_burst_shaper: zero padding either side of the burst, preamble scaled like the frame
                        for (int i = tid; i < tx.pre_pad; i += T) stg_stream(o - tx.pre_pad + i, cmake(0.f, 0.f));
                        for (int i = tid; i < tx.post_pad; i += T) stg_stream(o + tx.P + W + i, cmake(0.f, 0.f));
                        for (int i = tid; i < tx.P; i += T) stg_stream(o + i, cmul_rn(ldg_nc(p + i), scale));
                    } else if (vec) {
                        for (int i = tid; i < tx.P / 2; i += T) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(p) + i);
                            stg_stream4(o + 2 * i, cmake(q.x, q.y), cmake(q.z, q.w));
                        }
                    } else {
                        for (int i = tid; i < tx.P; i += T) stg_stream(o + i, ldg_nc(p + i));
                    }
                }
            }
        }
        STAGE_MARK(6) // stage C compute + stores
    }
    if constexpr (S::TBL_TMEM) {
        tmem_fence_before_sync();
        __syncthreads();
        if (tid < 32) tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------
// Fused receiver.  in: [n_frames][N] time samples; eq: per-bin channel or nullptr;
// out: [n_frames][N]; mode 0: soft symbols y (generic_work[_equalize]); mode 1: R (fft_[equalize_]filter_downsample).
// Shared memory: R = row buffer / output staging (bulk-stored), P = the first PR sample rows of every
// frame of the NEXT group (TMA prefetch); the remaining rows are prefetched into registers.

// DEC = true: the output is the hard decision of every soft symbol, one byte per symbol (chunks) -- see the epilogue.
// EQ = true: the equalising variants (eq != nullptr).  A template parameter, not a run-time branch: the plain receiver is the
// headline kernel and must not carry the equaliser's registers, barrier and branches (measured: 0.223 -> 0.237 ms at C3 with
// the equaliser compiled into the same kernel).
template <class S, bool SIC, bool DEC = false, bool EQ = false>
__global__ void __launch_bounds__(S::T, S::MINB) fused_rx_kernel(cpx* __restrict__ out, const cpx* __restrict__ in,
                                                                 const cpx* __restrict__ eq,
                                                                 const cpx* __restrict__ table,
                                                                 const cpx* __restrict__ tw,
                                                                 const cpx* __restrict__ taps, int L, int mode,
                                                                 int n_frames, SicArgs sic)
{
    constexpr int M = S::M, K = S::K, N = S::N, T = S::T, IPT = S::IPT, F = S::F, RS = S::RS, PR = S::PR;
    constexpr int XR = M - PR > 0 ? M - PR : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(smem_raw);
    cpx* tw_s = buf + S::BUF_ELEMS;
    cpx* tbl_s = tw_s + S::TW_ELEMS;
    cpx* pre = tbl_s + S::TBL_ELEMS;
    cpx* taps_s = pre + S::P_ELEMS;
    uint64_t* bar_p = reinterpret_cast<uint64_t*>(taps_s + S::TAPS_ELEMS);
    uint64_t* bar_h = bar_p + 2; // channel of the current group (equalising variant); bar_p + 1 holds the tensor-memory slot
    const int tid = threadIdx.x;
    const int n_groups = (n_frames + F - 1) / F;
    const float inv_m = 1.0f / (float)M;
    // frame f starts at in + f*fstride: N for a packed batch; larger when the frames still carry preamble / cyclic prefix /
    // suffix around the block (remove_prefix_cc, lib/remove_prefix_cc_impl.cc:84-115, fused into these loads)
    const size_t fstride = sic.in_stride ? sic.in_stride : (size_t)N;

    for (int i = tid; i < S::TW_ELEMS; i += T) tw_s[i] = tw[i];
    for (int i = tid; i < S::TBL_ELEMS; i += T) tbl_s[i] = table[i];
    for (int i = tid; i < L * M && i < S::IC_OFF; i += T) taps_s[i] = taps[i];
    if constexpr (SIC) {
        static_assert(IPT == 1 || (S::TBL_TMEM && S::TMEM_KEEP_COLS >= IPT * 2 * M),
                      "the cancellation loop keeps one subcarrier per thread in registers, or the kept blocks in tensor memory");
        for (int i = tid; i < M && i < 32; i += T) taps_s[S::IC_OFF + i] = cscale(sic.ic_taps[i], inv_m); // 1/M of the IFFT folded in
    }
    // constellation: interference cancellation and the hard-decision output (mode 2)
    if constexpr (SIC || DEC)
        for (int i = tid; i < sic.n_points && i < S::MAX_POINTS; i += T) taps_s[S::PTS_OFF + i] = sic.points[i];
    // grid-constellation lookup table: behind the 16 floats the phase reduction of the SIC loop uses
    unsigned char* lut_s = reinterpret_cast<unsigned char*>(taps_s + S::RED_OFF + 16);
    if constexpr (SIC || DEC)
        if (tid < 64) lut_s[tid] = sic.grid.lut[tid];
    if (tid == 0) {
        mbar_init(bar_p, 1);
        mbar_init(bar_h, 1);
    }
    // table columns of this thread -> tensor memory (once per CTA)
    uint32_t tmem_base = 0, tmem_mine = 0;
    if constexpr (S::TBL_TMEM) {
        uint32_t* slot = reinterpret_cast<uint32_t*>(bar_p + 1);
        tmem_setup<S>(slot, table, tw, tid, tmem_base, tmem_mine);
    }
    __syncthreads();

    // sample rows n2 < PR of every frame of group gg -> P (one bulk copy per frame, or one per group)
    auto load_head = [&](int gg) {
        const int fhh = min(F, n_frames - gg * F);
        if (PR == M && fstride == (size_t)N) {
            const uint32_t bytes = (uint32_t)fhh * N * sizeof(cpx);
            mbar_expect_tx(bar_p, bytes);
            bulk_load(pre, in + (size_t)gg * F * N, bytes, bar_p);
        } else {
            constexpr uint32_t bytes = (uint32_t)PR * K * sizeof(cpx);
            mbar_expect_tx(bar_p, bytes * fhh);
            for (int f = 0; f < fhh; ++f)
                bulk_load(pre + (size_t)f * PR * K, in + ((size_t)gg * F + f) * fstride, bytes, bar_p);
        }
    };
    // sample rows n2 >= PR -> registers
    cpx xr[IPT][XR];
    auto load_rest = [&](int gg) {
        if constexpr (PR < M) {
            const int fhh = min(F, n_frames - gg * F);
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int it = tid + j * T;
                const int f = it / K, n1 = it - f * K;
                const cpx* src = in + ((size_t)gg * F + f) * fstride + n1;
#pragma unroll
                for (int r = 0; r < M - PR; ++r)
                    xr[j][r] = f < fhh ? ldg_stream(src + (size_t)(PR + r) * K) : cmake(0.f, 0.f);
            }
        }
    };

    int g = blockIdx.x;
    if (g < n_groups) {
        if (tid == 0) load_head(g);
        load_rest(g);
    }
    uint32_t phase = 0, phase_h = 0;
    STAGE_INIT();
    for (; g < n_groups; g += gridDim.x) {
        const int fh = min(F, n_frames - g * F);
        const int gn = g + gridDim.x;
        cpx v[IPT][M];
        // the channel of this group is needed ~10k cycles from now: into L2 with it
        if constexpr (EQ)
            if (tid == 0) bulk_prefetch_l2(eq + (size_t)g * F * N, (uint32_t)fh * N * sizeof(cpx));
        STAGE_MARK(15) // loop top (bulk store issue)
        mbar_wait(bar_p, phase);
        phase ^= 1;
        STAGE_MARK(16) // wait for the bulk load
        // ---- stage A': x[n1 + K*n2] -> registers (lanes = consecutive n1), M-point FFT over n2
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, n1 = it - f * K;
            const cpx* src = pre + (size_t)f * PR * K + n1;
#pragma unroll
            for (int n2 = 0; n2 < M; ++n2) v[j][n2] = n2 < PR ? src[n2 * K] : xr[j][n2 - PR < XR ? n2 - PR : 0];
        }
        __syncthreads(); // P consumed: refill it with the next group
        // (issued by the idle last warp during the row FFT where there is one, see the modulator)
        constexpr bool IDLE_WARP = S::TWO_PASS && S::ROWS * S::R2 + 32 <= T;
        if (!IDLE_WARP && tid == 0 && gn < n_groups) {
            fence_proxy_async();
            load_head(gn);
        }
        STAGE_MARK(17) // stage A' reads
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int n1 = it % K;
            rf::FFTN<M, -1>::run(v[j]);
            if constexpr (S::TBL_SMEM) {
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cmul(v[j][m], tbl_s[m * K + n1]);
            } else {
                float tf[2 * M]; // this item's table column out of tensor memory
                tmem_ld<2 * M>(tf, tmem_mine + j * 2 * M);
                tmem_wait_ld();
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cmul(v[j][m], cmake(tf[2 * m], tf[2 * m + 1]));
            }
        }
        // the previous group's bulk store must have finished reading R
        if ((tid & 31) == 0) bulk_wait_read(); // (each warp's lane 0 issued that warp's stores)
        __syncthreads();
        STAGE_MARK(18) // M-FFT + table, wait for the previous store
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, n1 = it - f * K;
            cpx* dst = buf + (size_t)f * M * RS + S::swz(n1);
#pragma unroll
            for (int m = 0; m < M; ++m) dst[m * RS] = v[j][m];
        }
        __syncthreads();
        STAGE_MARK(19) // row writes
        // ---- stage B: K-point forward FFT of every row (over n1)
        if (IDLE_WARP && tid == T - 32 && gn < n_groups) {
            fence_proxy_async();
            load_head(gn);
        }
        row_fft<S, -1>(buf, tw_s, tid, tmem_mine + S::TMEM_TBL_COLS);
        STAGE_MARK(20) // row FFT
        if (gn < n_groups) load_rest(gn); // tail rows of the next group: in flight during stage C'
        __syncthreads();
        STAGE_MARK(21) // barrier after row FFT
        // ---- stage C': column k of all rows = the subcarrier's M bins
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * T;
            const int f = it / K, k = it - f * K;
            const cpx* src = buf + (size_t)f * M * RS + k;
#pragma unroll
            for (int m = 0; m < M; ++m) v[j][m] = src[m * RS];
        }
        __syncthreads();
        STAGE_MARK(22) // stage C' reads
        // Equalisation with overlap 2 (every reference configuration): the division and the tap combine stay in registers.
        // The channel of the group arrives by ONE bulk copy (TMA) into the row buffer, which is free once the columns are in
        // registers, in exactly the [k][m] order its owners read it in; it was pulled into L2 at the top of the iteration,
        // so the copy is short.  The only neighbour a subcarrier needs is k-1, i.e. the previous lane (a shuffle) -- lane 0
        // takes it from a small per-warp hand-over array in the unused tail of the row buffer.
        constexpr bool EQ_FAST_OK = !S::TWO_PASS || S::BUF_ELEMS - F * N >= F * N / 32;
        if (EQ && L == 2 && EQ_FAST_OK) {
            const cpx* eqg = eq + (size_t)g * F * N;
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar_h, (uint32_t)fh * N * sizeof(cpx));
                bulk_load(buf, eqg, (uint32_t)fh * N * sizeof(cpx), bar_h);
            }
            mbar_wait(bar_h, phase_h);
            phase_h ^= 1;
            cpx* bnd = buf + F * N; // [item warp][M]: the equalised bins of every warp's last lane
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int it = tid + j * T;
                if (it < fh * K) { // (frames behind the batch end have no channel)
                    const cpx* hh = buf + (size_t)it * M;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const cpx h1 = hh[m];
                        const float rden = __fdividef(1.0f, h1.x * h1.x + h1.y * h1.y);
                        const cpx num = cmulc(v[j][m], h1); // y * conj(h) / |h|^2, volk_32fc_x2_divide_32fc
                        v[j][m] = cmake(num.x * rden, num.y * rden);
                    }
                }
                if constexpr (K >= 32) {
                    if ((tid & 31) == 31) {
                        cpx* b = bnd + (size_t)(it >> 5) * M;
#pragma unroll
                        for (int m = 0; m < M; ++m) b[m] = v[j][m];
                    }
                }
            }
            if constexpr (K >= 32) __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int it = tid + j * T;
                const int f = it / K, k = it - f * K;
                const int lane = tid & 31;
                // R_k[m] = taps[M+m] * Yeq[(k-1) M + m] + taps[m] * Yeq[k M + m]   (receiver_kernel_cc.cc:165-192, L = 2)
                if constexpr (K >= 32) {
                    // subcarrier k-1 is the previous lane; for lane 0 the last lane of another warp (k = 0: K-1 of the same frame)
                    const int src = k == 0 ? it + K - 1 : it - 1;
                    const cpx* b = bnd + (size_t)(src >> 5) * M;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        cpx pv = cmake(__shfl_up_sync(0xffffffffu, v[j][m].x, 1), __shfl_up_sync(0xffffffffu, v[j][m].y, 1));
                        if (lane == 0) pv = b[m];
                        v[j][m] = cadd(cmul(taps_s[M + m], pv), cmul(taps_s[m], v[j][m]));
                    }
                } else { // whole frames inside a warp: k-1 (wrapping inside the frame) is another lane
                    const int src = k == 0 ? lane + K - 1 : lane - 1;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const cpx pv = cmake(__shfl_sync(0xffffffffu, v[j][m].x, src), __shfl_sync(0xffffffffu, v[j][m].y, src));
                        v[j][m] = cadd(cmul(taps_s[M + m], pv), cmul(taps_s[m], v[j][m]));
                    }
                }
            }
            __syncthreads(); // the inverted channel and the hand-over array are dead: the row buffer may be reused
        } else if (EQ) {
            // any other overlap: Y[b*M+m] back to the linear [b][m] order, divide by the channel, combine the L parts
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                cpx* dst = buf + (size_t)(tid + j * T) * M;
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m] = v[j][m];
            }
            __syncthreads();
            const cpx* eqg = eq + (size_t)g * F * N;
            {
                // F*N/T = IPT*M bins per thread, consecutive lanes on consecutive bins: issue every
                // channel load before the first division so that their latencies overlap
                constexpr int PER = IPT * M;
                cpx hq[PER];
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int i = tid + q * T;
                    hq[q] = i < fh * N ? ldg_stream(eqg + i) : cmake(1.f, 0.f);
                }
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int i = tid + q * T;
                    const cpx y = buf[i], hh = hq[q];
                    const float rden = __fdividef(1.0f, hh.x * hh.x + hh.y * hh.y);
                    const cpx num = cmulc(y, hh); // y * conj(h) / |h|^2, volk_32fc_x2_divide_32fc
                    buf[i] = cmake(num.x * rden, num.y * rden);
                }
            }
            __syncthreads();
            const int h = L / 2;
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int it = tid + j * T;
                const int f = it / K, k = it - f * K;
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cmake(0.f, 0.f);
                for (int i = 0; i < L; ++i) {
                    int kk = k + i - h;
                    kk = kk < 0 ? kk + K : (kk >= K ? kk - K : kk);
                    const cpx* src = buf + ((size_t)f * K + kk) * M;
                    const cpx* tp = taps_s + ((i + h) % L) * M;
#pragma unroll
                    for (int m = 0; m < M; ++m) v[j][m] = cadd(v[j][m], cmul(tp[m], src[m]));
                }
            }
            __syncthreads();
        }
        if constexpr (SIC && IPT == 1) {
            // v[0] = R_k (kept frequency block of this thread's subcarrier); iterate decide -> re-modulate the
            // neighbours -> subtract -> back to time domain without leaving the SM
            const int f = tid / K, k = tid - f * K;
            const cpx* ic_s = taps_s + S::IC_OFF;
            const cpx* pts_s = taps_s + S::PTS_OFF;
            float* red_s = reinterpret_cast<float*>(taps_s + S::RED_OFF); // 64 floats
            const int cnt = sic.count[k];
            cpx y[M];
#pragma unroll
            for (int m = 0; m < M; ++m) y[m] = v[0][m];
            rf::FFTN<M, +1>::run(y);
#pragma unroll
            for (int m = 0; m < M; ++m) y[m] = cscale(y[m], inv_m);
            // the 1/M of every later transform_subcarriers_to_td is folded into the kept block and the taps (ic_s), once
#pragma unroll
            for (int m = 0; m < M; ++m) v[0][m] = cscale(v[0][m], inv_m);
            const float qa = cnt ? sic.qpsk_a : 0.f;
            for (int it = 0; it < sic.ic_iter; ++it) {
                cpx d[M];
                if (sic.qpsk_a > 0.f) {
                    // gr::digital::constellation_qpsk: idx = 2*(im > 0) + (re > 0), point = (+-a, +-a): no table lookup
#pragma unroll
                    for (int m = 0; m < M; ++m) d[m] = cmake(y[m].x > 0.f ? qa : -qa, y[m].y > 0.f ? qa : -qa);
                } else if (sic.rule != 1 && sic.grid.n_re > 0) {
                    // grid constellation (square / rectangular QAM): straight-line quantiser, no call in this loop
#pragma unroll
                    for (int m = 0; m < M; ++m)
                        d[m] = cnt ? pts_s[decide_symbol_grid_fast(y[m], sic.grid, lut_s)] : cmake(0.f, 0.f);
                } else {
#pragma unroll
                    for (int m = 0; m < M; ++m)
                        d[m] = cnt ? pts_s[decide_symbol_grid(y[m], pts_s, sic.n_points, sic.rule, sic.grid, lut_s)]
                                   : cmake(0.f, 0.f);
                }
                if (sic.phase_comp > 0 && it == 0) {
                    // calculate_phase_offset (:78-91): mean over the map of arg(decided) - arg(soft); deterministic
                    // tree: lanes of a frame -> per-warp partial -> fixed-order sum
                    float part = 0.f;
                    if (cnt) {
#pragma unroll
                        for (int m = 0; m < M; ++m) part += atan2f(d[m].y, d[m].x) - atan2f(y[m].y, y[m].x);
                        part *= (float)cnt;
                    }
                    constexpr int G = K < 32 ? K : 32;       // lanes of one frame inside a warp
                    constexpr int NP = K / G;                 // partials per frame
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if ((tid & (G - 1)) == 0) red_s[tid / G] = part;
                    __syncthreads();
                    float phi = 0.f;
                    for (int q = 0; q < NP; ++q) phi += red_s[f * NP + q];
                    phi *= sic.inv_map_total;
                    float sn, cs;
                    sincosf(phi, &sn, &cs);
                    const cpx rot = cmake(cs, sn);
#pragma unroll
                    for (int m = 0; m < M; ++m) v[0][m] = cmul(v[0][m], rot); // the kept block stays rotated (:61-71)
                }
                cpx* mine = buf + (size_t)tid * M;
#pragma unroll
                for (int m = 0; m < M; ++m) mine[m] = d[m];
                __syncthreads();
                const int kp = k == 0 ? K - 1 : k - 1, kn = k == K - 1 ? 0 : k + 1;
                const cpx* prev = buf + ((size_t)f * K + kp) * M;
                const cpx* next = buf + ((size_t)f * K + kn) * M;
#pragma unroll
                for (int m = 0; m < M; ++m) d[m] = cadd(prev[m], next[m]);
                __syncthreads();
                rf::FFTN<M, -1>::run(d);
#pragma unroll
                for (int m = 0; m < M; ++m) y[m] = csub(v[0][m], cmul(ic_s[m], d[m]));
                rf::FFTN<M, +1>::run(y);
            }
#pragma unroll
            for (int m = 0; m < M; ++m) v[0][m] = y[m];
        
        }
        if constexpr (SIC && IPT > 1) {
            // Two (or more) subcarriers per thread: the kept frequency blocks R_k / M wait in tensor memory, the soft symbols
            // never stay in registers across an iteration -- every new y is decided at once and only the DECISIONS (one byte
            // per symbol: constellation index, 255 = inactive subcarrier) are exchanged through shared memory, double
            // buffered in the tail of the row buffer.  lib/advanced_receiver_kernel_cc.cc:56-76, receiver_kernel_cc.cc:274-299.
            // A record = the M decisions of one subcarrier packed into W4 16-byte words: one 128-bit store by its owner, one
            // 128-bit load per neighbour (byte-wide accesses at a 15-byte stride would be 4-way bank conflicts each).
            constexpr int W4 = (M + 15) / 16;
            constexpr int EXB = F * K * W4 * 16;                        // bytes of one decision buffer
            constexpr int EX_OFF = S::BUF_ELEMS - (2 * EXB + 7) / 8;    // both buffers at the end of the row buffer
            static_assert(EX_OFF >= T * M && (EX_OFF % 2) == 0, "decision buffers must not overlap the output staging of item 0");
            uint4* ex0 = reinterpret_cast<uint4*>(buf + EX_OFF);
            const cpx* ic_s = taps_s + S::IC_OFF;
            const cpx* pts_s = taps_s + S::PTS_OFF;
            float* red_s = reinterpret_cast<float*>(taps_s + S::RED_OFF);
            const uint32_t keep = tmem_mine + S::TMEM_TBL_COLS + S::TMEM_TW_COLS;
            const float qa = sic.qpsk_a;
            const int n_pts = sic.n_points;
            auto point = [&](unsigned b) -> cpx {
                if (qa > 0.f) return b < 4u ? cmake((b & 1u) ? qa : -qa, (b & 2u) ? qa : -qa) : cmake(0.f, 0.f);
                return (int)b < n_pts ? pts_s[b] : cmake(0.f, 0.f);
            };
            auto decide = [&](cpx y) -> unsigned {
                if (qa > 0.f) return (unsigned)(2 * (y.y > 0.f) + (y.x > 0.f));
                if (sic.rule != 1 && sic.grid.n_re > 0) return (unsigned)decide_symbol_grid_fast(y, sic.grid, lut_s);
                return (unsigned)decide_symbol_grid(y, pts_s, n_pts, sic.rule, sic.grid, lut_s);
            };
            auto byte_of = [](const uint4 (&r)[W4], int m) -> unsigned {
                const uint4& q = r[m >> 4];
                const unsigned w = ((m >> 2) & 3) == 0 ? q.x : ((m >> 2) & 3) == 1 ? q.y : ((m >> 2) & 3) == 2 ? q.z : q.w;
                return (w >> (8 * (m & 3))) & 0xffu;
            };
            // decisions of one subcarrier -> its record in decision buffer `dstb`
            auto publish = [&](uint4* dstb, int it, const cpx (&y)[M], int cnt) {
                unsigned w[W4 * 4];
#pragma unroll
                for (int q = 0; q < W4 * 4; ++q) w[q] = 0xffffffffu; // 255 = no symbol (inactive subcarrier, padding)
                auto put = [&](int m, unsigned b) { w[m >> 2] = (w[m >> 2] & ~(0xffu << (8 * (m & 3)))) | (b << (8 * (m & 3))); };
                if (cnt) { // the decision rule is picked outside the unrolled loops: the common ones stay call-free
                    if (qa > 0.f) {
#pragma unroll
                        for (int m = 0; m < M; ++m) put(m, (unsigned)(2 * (y[m].y > 0.f) + (y[m].x > 0.f)));
                    } else if (sic.rule != 1 && sic.grid.n_re > 0) {
#pragma unroll
                        for (int m = 0; m < M; ++m) put(m, (unsigned)decide_symbol_grid_fast(y[m], sic.grid, lut_s));
                    } else {
#pragma unroll
                        for (int m = 0; m < M; ++m) put(m, decide(y[m]));
                    }
                }
#pragma unroll
                for (int q = 0; q < W4; ++q) dstb[(size_t)it * W4 + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            };
            // prologue: park R_k / M, first soft symbols, first decisions (+ the phase estimate's partial sums)
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int it = tid + j * T, k = it % K;
                const int cnt = sic.count[k];
                float kf[2 * M];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    v[j][m] = cscale(v[j][m], inv_m);
                    kf[2 * m] = v[j][m].x;
                    kf[2 * m + 1] = v[j][m].y;
                }
                tmem_st<2 * M>(keep + j * 2 * M, kf);
                rf::FFTN<M, +1>::run(v[j]); // y = IFFT_M(R_k) / M
                // (ic_iter >= 1: without iterations the host runs the plain receiver)
                publish(ex0, it, v[j], cnt);
                if (sic.phase_comp > 0 && cnt) {
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const cpx dd = point(decide(v[j][m]));
                        part += (float)cnt * (atan2f(dd.y, dd.x) - atan2f(v[j][m].y, v[j][m].x));
                    }
                }
            }
            tmem_wait_st();
            {
                if (sic.phase_comp > 0) {
                    // calculate_phase_offset (:78-91): mean over the map of arg(decided) - arg(soft); fixed-order tree.
                    // One frame per CTA pass here (F == 1): warp partials, then one sum
                    static_assert(F == 1, "phase compensation of the tensor-memory variant assumes one frame per CTA pass");
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if ((tid & 31) == 0) red_s[tid >> 5] = part;
                    __syncthreads();
                    float phi = 0.f;
                    for (int q = 0; q < T / 32; ++q) phi += red_s[q];
                    phi *= sic.inv_map_total;
                    float sn, cs;
                    sincosf(phi, &sn, &cs);
                    const cpx rot = cmake(cs, sn);
#pragma unroll
                    for (int j = 0; j < IPT; ++j) { // the kept blocks stay rotated (:61-71)
                        float kf[2 * M];
                        tmem_ld<2 * M>(kf, keep + j * 2 * M);
                        tmem_wait_ld();
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            const cpx r = cmul(cmake(kf[2 * m], kf[2 * m + 1]), rot);
                            kf[2 * m] = r.x;
                            kf[2 * m + 1] = r.y;
                        }
                        tmem_st<2 * M>(keep + j * 2 * M, kf);
                    }
                    tmem_wait_st();
                }
                __syncthreads();
                for (int iter = 0; iter < sic.ic_iter; ++iter) {
                    const bool last = iter == sic.ic_iter - 1;
                    const uint4* cur = ex0 + (size_t)(iter & 1) * (EXB / 16);
                    uint4* nxt = ex0 + (size_t)((iter + 1) & 1) * (EXB / 16);
#pragma unroll
                    for (int j = 0; j < IPT; ++j) {
                        const int it = tid + j * T, f = it / K, k = it - f * K;
                        const int kp = k == 0 ? K - 1 : k - 1, kn = k == K - 1 ? 0 : k + 1;
                        uint4 rp[W4], rn[W4];
#pragma unroll
                        for (int q = 0; q < W4; ++q) {
                            rp[q] = cur[((size_t)f * K + kp) * W4 + q];
                            rn[q] = cur[((size_t)f * K + kn) * W4 + q];
                        }
                        cpx d[M];
#pragma unroll
                        for (int m = 0; m < M; ++m) d[m] = cadd(point(byte_of(rp, m)), point(byte_of(rn, m)));
                        // the last iteration stages its results where the decision buffers live (item 0's slice lies in
                        // front of them): before the first overlapping slice is written every thread has read
                        if (last && j == 1) __syncthreads();
                        rf::FFTN<M, -1>::run(d);
                        float kf[2 * M];
                        tmem_ld<2 * M>(kf, keep + j * 2 * M);
                        tmem_wait_ld();
#pragma unroll
                        for (int m = 0; m < M; ++m) d[m] = csub(cmake(kf[2 * m], kf[2 * m + 1]), cmul(ic_s[m], d[m]));
                        rf::FFTN<M, +1>::run(d);
                        if (!last) {
                            publish(nxt, it, d, sic.count[k]);
                        } else {
                            cpx* dst = buf + (size_t)it * M; // output staging, linear [k][m] (bulk-stored below)
#pragma unroll
                            for (int m = 0; m < M; ++m) dst[m] = d[m];
                        }
                    }
                    if (!last) __syncthreads();
                }
            }
        }
        // output staging in the linear [k][m] order; item j covers the contiguous slice
        // [j*T*M, (j+1)*T*M), which leaves by bulk stores as soon as it is staged so that it
        // drains to HBM while the next item's M-point IFFT runs
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            if (!SIC && mode != 1) {
                rf::FFTN<M, +1>::run(v[j]);
#pragma unroll
                for (int m = 0; m < M; ++m) v[j][m] = cscale(v[j][m], inv_m);
            }
            if constexpr (DEC) {
                // hard decisions (symbols2bits' argmin / the constellation's decision rule): the soft symbols never
                // leave the SM, the frame goes out as one byte per symbol in the same [k][m] order
                unsigned char* dst = reinterpret_cast<unsigned char*>(buf) + (size_t)(tid + j * T) * M;
                const cpx* pts_s = taps_s + S::PTS_OFF;
                decide_block<M>(v[j], dst, pts_s, sic.n_points, sic.rule, sic.grid, lut_s);
            } else if constexpr (!(SIC && IPT > 1)) { // (that variant's cancellation loop has staged its results already)
                cpx* dst = buf + (size_t)(tid + j * T) * M;
#pragma unroll
                for (int m = 0; m < M; ++m) dst[m] = v[j][m];
            }
            if constexpr (!(SIC && IPT > 1)) {
                // every warp stores its own 32 records (a contiguous slice of the output) as soon as they are staged: no CTA
                // barrier in the output stage, and the first bytes leave a warp-skew earlier (measured against one store per
                // item, per group, and a deferred last store: experiments/README.md)
                fence_proxy_async();
                __syncwarp();
                if ((tid & 31) == 0) {
                    const int e0 = (j * T + (tid & ~31)) * M, e1 = min(e0 + 32 * M, fh * N);
                    if (e1 > e0) {
                        if constexpr (DEC)
                            bulk_store(reinterpret_cast<unsigned char*>(out) + (size_t)g * F * N + e0,
                                       reinterpret_cast<unsigned char*>(buf) + e0, (uint32_t)(e1 - e0));
                        else
                            bulk_store(out + (size_t)g * F * N + e0, buf + e0, (uint32_t)(e1 - e0) * sizeof(cpx));
                    }
                }
                continue;
            }
            // (the cancellation loop of the two-subcarrier shapes stages CTA-wide: one store per item)
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                const int e0 = j * T * M, e1 = min((j + 1) * T * M, fh * N);
                if (e1 > e0) {
                    if constexpr (DEC)
                        bulk_store(reinterpret_cast<unsigned char*>(out) + (size_t)g * F * N + e0,
                                   reinterpret_cast<unsigned char*>(buf) + e0, (uint32_t)(e1 - e0));
                    else
                        bulk_store(out + (size_t)g * F * N + e0, buf + e0, (uint32_t)(e1 - e0) * sizeof(cpx));
                }
            }
        }
        STAGE_MARK(23) // M-IFFT + output staging + store issue
    }
    if ((tid & 31) == 0) bulk_wait_all(); // every issuing lane waits for its own stores
    if constexpr (S::TBL_TMEM) {
        tmem_fence_before_sync();
        __syncthreads();
        if (tid < 32) tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------
// launchers + shape entries

template <class S>
static void launch_mod(cpx* out, const cpx* in, const cpx* table, const cpx* tw, int n_frames, int grid,
                       cudaStream_t s)
{
    fused_mod_kernel<S, false><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, table, tw, n_frames, TxArgs{});
}
template <class S>
static void launch_tx(cpx* out, const cpx* in, const cpx* table, const cpx* tw, int n_frames, int grid,
                      const TxArgs& tx, cudaStream_t s)
{
    fused_mod_kernel<S, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, table, tw, n_frames, tx);
}
template <class S>
static void launch_modc(cpx* out, const cpx* in, const cpx* table, const cpx* tw, int n_frames, int grid,
                        const TxArgs& tx, cudaStream_t s)
{
    fused_mod_kernel<S, false, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, table, tw, n_frames, tx);
}
template <class S>
static void launch_txc(cpx* out, const cpx* in, const cpx* table, const cpx* tw, int n_frames, int grid,
                       const TxArgs& tx, cudaStream_t s)
{
    fused_mod_kernel<S, true, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, table, tw, n_frames, tx);
}
// receiver with the hard-decision output (DEC): the constellation travels in SicArgs
template <class S>
static void launch_rxd(cpx* out, const cpx* in, const cpx* eq, const cpx* table, const cpx* tw, const cpx* taps, int L,
                       int n_frames, int grid, SicArgs sic, cudaStream_t s)
{
    if (eq)
        fused_rx_kernel<S, false, true, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, 0, n_frames, sic);
    else
        fused_rx_kernel<S, false, true, false><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, 0, n_frames, sic);
}
template <class S>
static void launch_rx(cpx* out, const cpx* in, const cpx* eq, const cpx* table, const cpx* tw, const cpx* taps, int L,
                      int mode, int n_frames, int grid, size_t in_stride, cudaStream_t s)
{
    SicArgs a{};
    a.in_stride = in_stride;
    if (eq)
        fused_rx_kernel<S, false, false, true><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, mode, n_frames, a);
    else
        fused_rx_kernel<S, false, false, false><<<grid, S::T, S::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, mode, n_frames, a);
}
// The cancellation loop runs on the shape itself when a thread owns one subcarrier, otherwise on a companion shape whose
// tensor memory has room for the kept blocks (row-FFT twiddles back in shared memory): SicShape<S>.
template <class S, bool ONE = S::IPT == 1>
struct SicShape {
    typedef S type;
};
template <class S>
struct SicShape<S, false> {
    typedef Shape<S::M, S::R1, S::R2, S::T, S::IPT, S::MINB, true, false, S::IPT * 2 * S::M> type;
};
template <class S>
constexpr bool sic_shape_ok()
{
    typedef typename SicShape<S>::type X;
    if (S::IPT == 1) return S::T / (S::K < 32 ? S::K : 32) <= 32; // partial sums of the phase estimate: 32-float scratch
    return X::TBL_TMEM && X::F == 1 && X::TMEM_USED <= 512 && X::TMEM_COLS * X::MINB <= 512 && X::T / 32 <= 32 &&
           X::BUF_ELEMS - 4 * X::F * X::K * ((X::M + 15) / 16) >= X::T * X::M; // two packed decision buffers behind item 0's staging
}
template <class S>
static void launch_sic(cpx* out, const cpx* in, const cpx* eq, const cpx* table, const cpx* tw, const cpx* taps, int L,
                       int n_frames, int grid, SicArgs sic, cudaStream_t s)
{
    typedef typename SicShape<S>::type X;
    if constexpr (sic_shape_ok<S>()) {
        if (eq)
            fused_rx_kernel<X, true, false, true><<<grid, X::T, X::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, 0, n_frames, sic);
        else
            fused_rx_kernel<X, true, false, false><<<grid, X::T, X::SMEM_BYTES, s>>>(out, in, eq, table, tw, taps, L, 0, n_frames, sic);
    }
}


template <class S>
static ShapeEntry make_entry(const char* mn, const char* rn, const char* tn)
{
    ShapeEntry e;
    e.M = S::M; e.K = S::K; e.R1 = S::R1; e.R2 = S::R2; e.T = S::T; e.F = S::F;
    e.smem = S::SMEM_BYTES;
    e.mod_name = mn;
    e.rx_name = rn;
    e.tx_name = tn;
    e.tx = &launch_tx<S>;
    e.tx_fn = (const void*)&fused_mod_kernel<S, true>;
    e.mod = &launch_mod<S>;
    e.rx = &launch_rx<S>;
    e.mod_fn = (const void*)&fused_mod_kernel<S, false>;
    e.rx_fn = (const void*)&fused_rx_kernel<S, false, false, false>;
    e.rx_eq_fn = (const void*)&fused_rx_kernel<S, false, false, true>;
    e.sic = nullptr;
    e.sic_fn = nullptr;
    e.sic_name = "none";
    e.modc = &launch_modc<S>;
    e.txc = &launch_txc<S>;
    e.rxd = &launch_rxd<S>;
    e.modc_fn = (const void*)&fused_mod_kernel<S, false, true>;
    e.txc_fn = (const void*)&fused_mod_kernel<S, true, true>;
    e.rxd_fn = (const void*)&fused_rx_kernel<S, false, true, false>;
    e.rxd_eq_fn = (const void*)&fused_rx_kernel<S, false, true, true>;
    e.modc_name = std::string(mn) + "+chunks";
    e.txc_name = std::string(tn) + "+chunks";
    e.rxd_name = std::string(rn) + "+decide";
    if constexpr (sic_shape_ok<S>()) {
        e.sic = &launch_sic<S>;
        e.sic_fn = (const void*)&fused_rx_kernel<typename SicShape<S>::type, true, false, false>;
        e.sic_eq_fn = (const void*)&fused_rx_kernel<typename SicShape<S>::type, true, false, true>;
        e.sic_smem = SicShape<S>::type::SMEM_BYTES;
    }
    return e;
}

#define GFDM_SHAPE(M, R1, R2, T, IPT, MINB)                                                            \
    make_entry<Shape<M, R1, R2, T, IPT, MINB>>("fused_mod_kernel<M=" #M ",K=" #R1 "x" #R2 ",T=" #T ">", \
                                                "fused_rx_kernel<M=" #M ",K=" #R1 "x" #R2 ",T=" #T ">",   \
                                                "fused_tx_chain_kernel<M=" #M ",K=" #R1 "x" #R2 ",T=" #T ">")

} // namespace gfdm
