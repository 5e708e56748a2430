// fused_modem.cu -- host side of the single-kernel GFDM modulator / receiver (kernels: fused_kernels.cuh, instantiated per
// shape in fused_shapes_*.cu): shape lookup, folded filter/twiddle tables, launch geometry.
#include "fused_shapes.h"

#include <cmath>
#include <cstring>
#include <string>

namespace gfdm {

static const std::vector<ShapeEntry>& shape_table()
{
    static const std::vector<ShapeEntry> t = [] {
        std::vector<ShapeEntry> all;
        for (auto part : { fused_shapes_baseline, fused_shapes_k16, fused_shapes_k32, fused_shapes_k64, fused_shapes_k128, fused_shapes_k256,
                           fused_shapes_k512, fused_shapes_k1024 }) {
            std::vector<ShapeEntry> p = part();
            all.insert(all.end(), p.begin(), p.end());
        }
        return all;
    }();
    return t;
}

struct FusedImpl {
    const ShapeEntry* e = nullptr; // single-pass shape, or
    TwoPass* tp = nullptr;         // two-pass kernels (frame larger than shared memory)
    int M = 0, K = 0, L = 0;
    bool eq_ok = false;        // rx: the equalising variant is available (single-pass shape, L*M <= 64)
    cpx* d_table = nullptr;    // tx: C_tx ; rx: C_rx (taps folded)
    cpx* d_table_eq = nullptr; // rx only: plain twiddle
    cpx* d_tw = nullptr;
    cpx* d_taps = nullptr;
    int mod_grid_cap = 0, rx_grid_cap = 0, sic_grid_cap = 0, tx_grid_cap = 0, modc_grid_cap = 0, txc_grid_cap = 0, rxd_grid_cap = 0;
    // interference cancellation (advanced receiver)
    cpx* d_ic = nullptr;
    cpx* d_points = nullptr;
    unsigned char* d_count = nullptr;
    SicArgs sic{};
    std::string sic_name;
};

static const ShapeEntry* find_shape(int M, int K)
{
    for (const ShapeEntry& e : shape_table())
        if (e.M == M && e.K == K) return &e;
    return nullptr;
}

int fused_grid_cap(const void* fn, int threads, size_t smem)
{
    int dev = 0, sms = 0, per_sm = 0;
    GFDM_CUDA_CHECK(cudaGetDevice(&dev));
    GFDM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GFDM_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GFDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) throw CudaError("fused kernel does not fit on this device");
    return sms * per_sm;
}

std::vector<cpx> make_row_twiddles(int R1, int R2)
{
    const int K = R1 * R2;
    std::vector<cpx> tw((size_t)K);
    for (int k1 = 0; k1 < R1; ++k1)
        for (int n0 = 0; n0 < R2; ++n0) {
            const double ph = -2.0 * M_PI * (double)((long)n0 * k1 % K) / (double)K;
            tw[(size_t)k1 * R2 + n0] = make_float2((float)std::cos(ph), (float)std::sin(ph));
        }
    return tw;
}

// sign = +1: modulator table (incl. 1/N and part_len); sign = -1: receiver table; with_taps = false: twiddle only
std::vector<std::complex<double>> make_fold_table_d(int M, int K, int L, const std::vector<std::complex<float>>& taps,
                                                    int sign, bool with_taps)
{
    const int N = M * K, h = L / 2;
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M;
    std::vector<std::complex<double>> t((size_t)N);
    // roots of unity once: a full-width receiver (L = K) sums K taps per table entry
    std::vector<std::complex<double>> wk((size_t)K);
    for (int e = 0; e < K; ++e) wk[e] = std::polar(1.0, sign * 2.0 * M_PI * (double)e / (double)K);
    for (int m = 0; m < M; ++m)
        for (int n1 = 0; n1 < K; ++n1) {
            std::complex<double> G(1.0, 0.0);
            if (with_taps) {
                G = 0.0;
                for (int i = 0; i < L; ++i) {
                    const std::complex<double> tp(taps[((i + h) % L) * M + m].real(), taps[((i + h) % L) * M + m].imag());
                    long e = ((long)(i - h) * n1) % K;
                    if (e < 0) e += K;
                    G += tp * wk[e];
                }
                if (sign > 0 && m >= part_len) G = 0.0;
            }
            const std::complex<double> w = std::polar(1.0, sign * 2.0 * M_PI * (double)((long)m * n1 % N) / (double)N);
            std::complex<double> c = G * w;
            if (sign > 0) c /= (double)N;
            t[(size_t)m * K + n1] = c;
        }
    return t;
}
std::vector<cpx> make_fold_table(int M, int K, int L, const std::vector<std::complex<float>>& taps, int sign,
                                 bool with_taps)
{
    const std::vector<std::complex<double>> d = make_fold_table_d(M, K, L, taps, sign, with_taps);
    std::vector<cpx> t(d.size());
    for (size_t i = 0; i < d.size(); ++i) t[i] = make_float2((float)d[i].real(), (float)d[i].imag());
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
    if (!e && L >= 1 && twopass_supported(M, K)) {
        TwoPass* tp = twopass_create_tx(M, K, L, taps);
        if (!tp) return;
        impl_ = new FusedImpl;
        impl_->tp = tp; impl_->M = M; impl_->K = K; impl_->L = L;
        return;
    }
    if (!e || L < 1) return;
    FusedImpl* p = new FusedImpl;
    p->e = e; p->M = M; p->K = K; p->L = L;
    try {
        p->mod_grid_cap = fused_grid_cap(e->mod_fn, e->T, e->smem);
        p->d_table = upload(make_fold_table(M, K, L, taps, +1, true));
        p->d_tw = upload(make_row_twiddles(e->R1, e->R2));
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
    if (!e && L >= 2 && twopass_supported(M, K)) {
        TwoPass* tp = twopass_create_rx(M, K, L, taps);
        if (!tp) return;
        impl_ = new FusedImpl;
        impl_->tp = tp; impl_->M = M; impl_->K = K; impl_->L = L;
        return;
    }
    if (!e || L < 2) return;
    FusedImpl* p = new FusedImpl;
    p->e = e; p->M = M; p->K = K; p->L = L;
    // Without equalisation any overlap (up to the full-width L = K of a zero-forcing receiver) costs nothing at run time:
    // the taps live in the folded table.  The equalising variant combines the L neighbouring blocks explicitly with the
    // taps held in shared memory (at most 64 of them).
    p->eq_ok = L * M <= 64;
    try {
        p->rx_grid_cap = fused_grid_cap(e->rx_fn, e->T, e->smem);
        fused_grid_cap(e->rx_eq_fn, e->T, e->smem); // opts the equalising kernel into its shared-memory size as well
        p->d_table = upload(make_fold_table(M, K, L, taps, -1, true));
        p->d_table_eq = upload(make_fold_table(M, K, L, taps, -1, false));
        p->d_tw = upload(make_row_twiddles(e->R1, e->R2));
        p->d_taps = upload(to_cpx(taps));
    } catch (...) {
        impl_ = p;
        destroy();
        throw;
    }
    impl_ = p;
}

bool FusedModem::supports_eq() const
{
    return impl_ && ((impl_->e != nullptr && impl_->eq_ok) || (impl_->tp != nullptr && twopass_supports_eq(impl_->tp)));
}
bool FusedModem::supports_stride() const { return impl_ && impl_->e != nullptr; }

int FusedModem::modulate(cpx* out, const cpx* in, size_t frames, cudaStream_t s)
{
    if (impl_->tp) return twopass_modulate(impl_->tp, out, in, frames, s);
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

bool FusedModem::supports_tx_chain(const TxArgs& tx) const
{
    // single-pass shapes only; the bulk copies of the compact symbol vectors need 16-byte granularity
    return impl_ && impl_->e && tx.n_in > 0 && tx.n_in % 2 == 0 && tx.n_ant >= 1 && tx.n_ant <= GFDM_TX_MAX_ANT;
}

int FusedModem::transmit(cpx* out, const cpx* in, const TxArgs& tx, size_t frames, cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    if (!impl_->tx_grid_cap) impl_->tx_grid_cap = fused_grid_cap(e->tx_fn, e->T, e->smem);
    int launches = 0;
    const size_t os = (size_t)tx.pre_pad + tx.P + (size_t)e->M * e->K + tx.cp + tx.cs + tx.post_pad;
    const size_t max_chunk = (size_t)1 << 20; // keep frame counts in int range
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->tx_grid_cap ? groups : impl_->tx_grid_cap;
        e->tx(out + f0 * os, in + f0 * (size_t)tx.n_in, impl_->d_table, impl_->d_tw, nf, grid, tx, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

const char* FusedModem::tx_name() const { return impl_ && impl_->e ? impl_->e->tx_name : "none"; }

// ---- chunk entries ---------------------------------------------------------------------------
// The bulk copies move whole 16-byte units: a frame of chunks (N bytes) and a store slice (T*M bytes) must be
// multiples of 16, which holds for every shape in the table; the constellation must fit the shared-memory slot.
bool FusedModem::supports_chunks(int n_points) const
{
    if (!impl_ || !impl_->e) return false;
    const ShapeEntry* e = impl_->e;
    return n_points >= 1 && n_points <= GFDM_FUSED_MAX_POINTS && (e->M * e->K) % 16 == 0 && (e->T * e->M) % 16 == 0;
}

int FusedModem::modulate_chunks(cpx* out, const unsigned char* chunks, const cpx* d_points, int n_points, size_t frames,
                                cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    if (!impl_->modc_grid_cap) impl_->modc_grid_cap = fused_grid_cap(e->modc_fn, e->T, e->smem);
    TxArgs tx;
    tx.points = d_points;
    tx.n_points = n_points;
    int launches = 0;
    const size_t N = (size_t)e->M * e->K, max_chunk = (size_t)1 << 20;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->modc_grid_cap ? groups : impl_->modc_grid_cap;
        e->modc(out + f0 * N, reinterpret_cast<const cpx*>(chunks + f0 * N), impl_->d_table, impl_->d_tw, nf, grid, tx, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

bool FusedModem::supports_tx_chain_chunks(const TxArgs& tx) const
{
    return impl_ && impl_->e && tx.n_in > 0 && tx.n_in % 16 == 0 && tx.n_ant >= 1 && tx.n_ant <= GFDM_TX_MAX_ANT &&
           tx.n_points >= 1 && tx.n_points <= GFDM_FUSED_MAX_POINTS &&
           (size_t)impl_->e->F * tx.n_in <= (size_t)impl_->e->F * impl_->e->M * impl_->e->K;
}

int FusedModem::transmit_chunks(cpx* out, const unsigned char* chunks, const TxArgs& tx, size_t frames, cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    if (!impl_->txc_grid_cap) impl_->txc_grid_cap = fused_grid_cap(e->txc_fn, e->T, e->smem);
    int launches = 0;
    const size_t os = (size_t)tx.pre_pad + tx.P + (size_t)e->M * e->K + tx.cp + tx.cs + tx.post_pad;
    const size_t max_chunk = (size_t)1 << 20;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->txc_grid_cap ? groups : impl_->txc_grid_cap;
        e->txc(out + f0 * os, reinterpret_cast<const cpx*>(chunks + f0 * (size_t)tx.n_in), impl_->d_table, impl_->d_tw, nf,
               grid, tx, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

int FusedModem::demodulate_decide(unsigned char* chunks_out, const cpx* in, const cpx* eq, const cpx* d_points, int n_points,
                                  int rule, const DecideGrid& dgrid, size_t frames, cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    if (!impl_->rxd_grid_cap) {
        impl_->rxd_grid_cap = fused_grid_cap(e->rxd_fn, e->T, e->smem);
        fused_grid_cap(e->rxd_eq_fn, e->T, e->smem);
    }
    int launches = 0;
    const size_t N = (size_t)e->M * e->K, max_chunk = (size_t)1 << 20;
    SicArgs a{};
    a.points = d_points;
    a.n_points = n_points;
    a.rule = rule;
    a.grid = dgrid;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->rxd_grid_cap ? groups : impl_->rxd_grid_cap;
        e->rxd(reinterpret_cast<cpx*>(chunks_out + f0 * N), in + f0 * N, eq ? eq + f0 * N : nullptr,
               eq ? impl_->d_table_eq : impl_->d_table, impl_->d_tw, impl_->d_taps, impl_->L, nf, grid, a, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

const char* FusedModem::modc_name() const { return impl_ && impl_->e ? impl_->e->modc_name.c_str() : "none"; }
const char* FusedModem::txc_name() const { return impl_ && impl_->e ? impl_->e->txc_name.c_str() : "none"; }
const char* FusedModem::rxd_name() const { return impl_ && impl_->e ? impl_->e->rxd_name.c_str() : "none"; }

int FusedModem::demodulate(cpx* out_td, cpx* out_fd, const cpx* in, const cpx* eq, size_t frames, cudaStream_t s, size_t in_stride)
{
    if (impl_->tp) {
        if (in_stride) throw std::invalid_argument("the two-pass receiver kernel takes packed frames only");
        int n = 0;
        if (out_td) n += twopass_demodulate(impl_->tp, out_td, in, eq, 0, frames, s);
        if (out_fd) n += twopass_demodulate(impl_->tp, out_fd, in, eq, 1, frames, s);
        return n;
    }
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
            e->rx(out + f0 * N, in + f0 * (in_stride ? in_stride : N), eq ? eq + f0 * N : nullptr,
                  eq ? impl_->d_table_eq : impl_->d_table, impl_->d_tw, impl_->d_taps, impl_->L, pass, nf, grid, in_stride, s);
            ++launches;
        }
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

bool FusedModem::init_sic(const std::vector<std::complex<float>>& ic_taps,
                          const std::vector<std::complex<float>>& points, int rule, const std::vector<int>& subcarrier_map)
{
    if (!impl_ || !impl_->e || !impl_->e->sic) return false;
    const ShapeEntry* e = impl_->e;
    if ((int)ic_taps.size() != e->M || e->M > 32 || points.empty() || points.size() > 64) return false;
    std::vector<unsigned char> count((size_t)e->K, 0);
    for (int k : subcarrier_map) {
        if (k < 0 || k >= e->K || count[k] == 255) return false;
        ++count[k];
    }
    impl_->sic_grid_cap = fused_grid_cap(e->sic_fn, e->T, e->sic_smem ? e->sic_smem : e->smem);
    fused_grid_cap(e->sic_eq_fn, e->T, e->sic_smem ? e->sic_smem : e->smem);
    impl_->d_ic = upload(to_cpx(ic_taps));
    impl_->d_points = upload(to_cpx(points));
    impl_->d_count = upload(count);
    impl_->sic.ic_taps = impl_->d_ic;
    impl_->sic.points = impl_->d_points;
    impl_->sic.count = impl_->d_count;
    impl_->sic.n_points = (int)points.size();
    impl_->sic.rule = rule;
    impl_->sic.grid = make_decide_grid(to_cpx(points)); // O(1) nearest-point decisions on grid constellations
    impl_->sic.qpsk_a = 0.f;
    if (rule == 1 && points.size() == 4) {
        const float a = points[3].real();
        const bool gr_qpsk = a > 0.f && points[3].imag() == a && points[0].real() == -a && points[0].imag() == -a &&
                             points[1].real() == a && points[1].imag() == -a && points[2].real() == -a && points[2].imag() == a;
        if (gr_qpsk) impl_->sic.qpsk_a = a;
    }
    impl_->sic.inv_map_total = subcarrier_map.empty() ? 0.f : 1.0f / (float)(subcarrier_map.size() * (size_t)e->M);
    impl_->sic_name = std::string(e->rx_name) + "+sic";
    return true;
}

bool FusedModem::sic_available() const { return impl_ && impl_->d_count != nullptr; }

int FusedModem::demodulate_sic(cpx* out, const cpx* in, const cpx* eq, size_t frames, int ic_iter, int phase_comp,
                               cudaStream_t s)
{
    const ShapeEntry* e = impl_->e;
    // no iterations: advanced_receiver_kernel_cc::generic_work is the plain receiver (perform_ic_iterations does nothing)
    if (ic_iter <= 0 && e->sic_smem) return demodulate(out, nullptr, in, eq, frames, s);
    int launches = 0;
    const size_t N = (size_t)e->M * e->K;
    const size_t max_chunk = (size_t)1 << 20;
    SicArgs a = impl_->sic;
    a.ic_iter = ic_iter;
    a.phase_comp = phase_comp;
    for (size_t f0 = 0; f0 < frames; f0 += max_chunk) {
        const int nf = (int)std::min(max_chunk, frames - f0);
        const int groups = (nf + e->F - 1) / e->F;
        const int grid = groups < impl_->sic_grid_cap ? groups : impl_->sic_grid_cap;
        e->sic(out + f0 * N, in + f0 * N, eq ? eq + f0 * N : nullptr, eq ? impl_->d_table_eq : impl_->d_table, impl_->d_tw,
               impl_->d_taps, impl_->L, nf, grid, a, s);
        ++launches;
    }
    GFDM_CUDA_CHECK(cudaGetLastError());
    return launches;
}

const char* FusedModem::sic_name() const { return sic_available() ? impl_->sic_name.c_str() : "none"; }
const char* FusedModem::mod_name() const { return !impl_ ? "none" : impl_->tp ? twopass_name(impl_->tp) : impl_->e->mod_name; }
const char* FusedModem::rx_name() const { return !impl_ ? "none" : impl_->tp ? twopass_name(impl_->tp) : impl_->e->rx_name; }


void FusedModem::destroy()
{
    if (!impl_) return;
    if (impl_->tp) twopass_destroy(impl_->tp);
    if (impl_->d_table) cudaFree(impl_->d_table);
    if (impl_->d_table_eq) cudaFree(impl_->d_table_eq);
    if (impl_->d_tw) cudaFree(impl_->d_tw);
    if (impl_->d_taps) cudaFree(impl_->d_taps);
    if (impl_->d_ic) cudaFree(impl_->d_ic);
    if (impl_->d_points) cudaFree(impl_->d_points);
    if (impl_->d_count) cudaFree(impl_->d_count);
    delete impl_;
    impl_ = nullptr;
}

} // namespace gfdm
