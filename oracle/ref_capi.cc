/*
 * oracle/ref_capi.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Exposes the UNMODIFIED gr-gfdm kernel classes (compiled from the sources
 * where they lie under /root/reference/lib, against the headers in
 * oracle/shim/) through the same C ABI as the product library
 * (include/gfdm_b200.h), so that parity tests can drive product and reference
 * through identical calls, and so that bench.py can time the reference's own
 * CPU implementation (`--impl reference`, `cpu_baseline.kind == "reference"`).
 *
 * Built by oracle/Makefile into oracle/_ref/libgfdm_ref.so.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 * Leaf arithmetic caveat: FFTW3/VOLK are not installable here, so FFT and
 * elementwise leaves are the scalar shims in oracle/shim/; every index
 * computation, scaling and quirk comes from the reference sources themselves.
 */
#include "../include/gfdm_b200.h"

#include <gfdm/add_cyclic_prefix_cc.h>
#include <gfdm/advanced_receiver_kernel_cc.h>
#include <gfdm/modulator_kernel_cc.h>
#include <gfdm/preamble_channel_estimator_cc.h>
#include <gfdm/receiver_kernel_cc.h>
#include <gfdm/resource_mapper_kernel_cc.h>
#include <gfdm/transmitter_kernel.h>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace gr::gfdm;
typedef std::complex<float> cf;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define REF_TRY try {
#define REF_CATCH                                              \
    }                                                          \
    catch (const std::invalid_argument& e)                     \
    {                                                          \
        return fail(GFDM_ERR_INVALID_ARGUMENT, e.what());      \
    }                                                          \
    catch (const std::exception& e)                            \
    {                                                          \
        return fail(GFDM_ERR_RUNTIME, e.what());               \
    }                                                          \
    return GFDM_OK;

#define REQUIRE_HOST(mem)                                                                 \
    if ((mem) != GFDM_MEM_HOST && (mem) != GFDM_MEM_HOST_ASYNC) /* a CPU library is trivially 'complete at return' */                                                           \
        return fail(GFDM_ERR_UNSUPPORTED, "oracle-ref is CPU only: GFDM_MEM_DEVICE is not supported");

static inline cf* C(gfdm_complex* p) { return reinterpret_cast<cf*>(p); }
static inline const cf* C(const gfdm_complex* p) { return reinterpret_cast<const cf*>(p); }
static std::vector<cf> vec(const gfdm_complex* p, int n)
{
    return std::vector<cf>(C(p), C(p) + (n > 0 ? n : 0));
}

struct gfdm_modulator { std::unique_ptr<modulator_kernel_cc> k; int M, K, L; };
struct gfdm_receiver { std::unique_ptr<receiver_kernel_cc> k; };
struct gfdm_advanced_receiver { std::unique_ptr<advanced_receiver_kernel_cc> k; };
struct gfdm_resource_mapper { std::unique_ptr<resource_mapper_kernel_cc> k; };
struct gfdm_cyclic_prefixer { std::unique_ptr<add_cyclic_prefix_cc> k; };
struct gfdm_channel_estimator { std::unique_ptr<preamble_channel_estimator_cc> k; int A; };
struct gfdm_transmitter { std::unique_ptr<transmitter_kernel> k; std::vector<cf> frame; int N; };
struct gfdm_fft {
    std::unique_ptr<gfdm_kernel_utils> u;
    std::vector<cf> in, out;
    fftwf_plan plan;
    int n;
};

extern "C" {

const char* gfdm_last_error(void) { return g_err.c_str(); }
/* used by gfdm_oracle_next.c (linked into this library) */
void gfdm_oracle_set_error(const char* msg) { g_err = msg; }
const char* gfdm_backend(void) { return "oracle-ref"; }
int gfdm_device_count(void) { return 0; }
int gfdm_set_device(int) { return GFDM_OK; }
int gfdm_set_stream(void*, void*) { return GFDM_OK; }
int gfdm_sync(void*) { return GFDM_OK; }
long long gfdm_launch_count(void*) { return 0; }
const char* gfdm_last_kernel(void*) { return "cpu"; }

int gfdm_calculate_signal_energy(float* energy, const gfdm_complex* in, int n)
{
    REF_TRY
    gfdm_kernel_utils u;
    *energy = u.calculate_signal_energy(C(in), n);
    REF_CATCH
}

int gfdm_fft_create(gfdm_fft** out, int fft_size, int forward)
{
    REF_TRY
    if (fft_size < 1) throw std::invalid_argument("fft_size MUST be positive");
    auto h = new gfdm_fft;
    h->u.reset(new gfdm_kernel_utils);
    h->n = fft_size;
    h->in.resize(fft_size);
    h->out.resize(fft_size);
    h->plan = h->u->initialize_fft(h->out.data(), h->in.data(), fft_size, forward != 0);
    *out = h;
    REF_CATCH
}
void gfdm_fft_destroy(gfdm_fft* h)
{
    if (!h) return;
    fftwf_destroy_plan(h->plan);
    delete h;
}
int gfdm_fft_execute_batch(gfdm_fft* h, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    for (int i = 0; i < n; ++i) {
        memcpy(h->in.data(), in + (size_t)i * h->n, sizeof(cf) * h->n);
        fftwf_execute(h->plan);
        memcpy(out + (size_t)i * h->n, h->out.data(), sizeof(cf) * h->n);
    }
    REF_CATCH
}

/* ---- modulator ---------------------------------------------------------- */
int gfdm_modulator_create(gfdm_modulator** out, int M, int K, int L, const gfdm_complex* taps, int n_taps)
{
    REF_TRY
    auto h = new gfdm_modulator;
    try {
        h->k.reset(new modulator_kernel_cc(M, K, L, vec(taps, n_taps)));
    } catch (...) {
        delete h;
        throw;
    }
    h->M = M; h->K = K; h->L = L;
    *out = h;
    REF_CATCH
}
void gfdm_modulator_destroy(gfdm_modulator* h) { delete h; }
int gfdm_modulator_block_size(const gfdm_modulator* h) { return h->k->block_size(); }
int gfdm_modulator_filter_taps(const gfdm_modulator* h, gfdm_complex* o)
{
    auto t = h->k->filter_taps();
    memcpy(o, t.data(), sizeof(cf) * t.size());
    return GFDM_OK;
}
int gfdm_modulator_work(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in)
{
    REF_TRY
    h->k->generic_work(C(out), C(in));
    REF_CATCH
}
int gfdm_modulator_work_batch(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i) h->k->generic_work(C(out) + i * bs, C(in) + i * bs);
    REF_CATCH
}

/* ---- receiver ----------------------------------------------------------- */
int gfdm_receiver_create(gfdm_receiver** out, int M, int K, int L, const gfdm_complex* taps, int n_taps)
{
    REF_TRY
    auto h = new gfdm_receiver;
    try {
        h->k.reset(new receiver_kernel_cc(M, K, L, vec(taps, n_taps)));
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    REF_CATCH
}
void gfdm_receiver_destroy(gfdm_receiver* h) { delete h; }
int gfdm_receiver_block_size(const gfdm_receiver* h) { return h->k->block_size(); }
int gfdm_receiver_timeslots(const gfdm_receiver* h) { return h->k->timeslots(); }
int gfdm_receiver_subcarriers(const gfdm_receiver* h) { return h->k->subcarriers(); }
int gfdm_receiver_overlap(const gfdm_receiver* h) { return h->k->overlap(); }
int gfdm_receiver_filter_taps(const gfdm_receiver* h, gfdm_complex* o)
{
    auto t = h->k->filter_taps();
    memcpy(o, t.data(), sizeof(cf) * t.size());
    return GFDM_OK;
}
int gfdm_receiver_ic_filter_taps(const gfdm_receiver* h, gfdm_complex* o)
{
    auto t = h->k->ic_filter_taps();
    memcpy(o, t.data(), sizeof(cf) * t.size());
    return GFDM_OK;
}
int gfdm_receiver_work_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                             const gfdm_complex* eq, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i) {
        if (eq)
            h->k->generic_work_equalize(C(out) + i * bs, C(in) + i * bs, C(eq) + i * bs);
        else
            h->k->generic_work(C(out) + i * bs, C(in) + i * bs);
    }
    REF_CATCH
}
int gfdm_receiver_work(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_receiver_work_batch(h, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_work_equalize(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_receiver_work_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_filter_downsample_batch(gfdm_receiver* h, gfdm_complex* out,
                                              const gfdm_complex* in, const gfdm_complex* eq, int n,
                                              int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i) {
        if (eq)
            h->k->fft_equalize_filter_downsample(C(out) + i * bs, C(in) + i * bs, C(eq) + i * bs);
        else
            h->k->fft_filter_downsample(C(out) + i * bs, C(in) + i * bs);
    }
    REF_CATCH
}
int gfdm_receiver_fft_filter_downsample(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_receiver_fft_filter_downsample_batch(h, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_equalize_filter_downsample(gfdm_receiver* h, gfdm_complex* out,
                                                 const gfdm_complex* in, const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_receiver_fft_filter_downsample_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_transform_subcarriers_to_td_batch(gfdm_receiver* h, gfdm_complex* out,
                                                    const gfdm_complex* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i) h->k->transform_subcarriers_to_td(C(out) + i * bs, C(in) + i * bs);
    REF_CATCH
}
int gfdm_receiver_transform_subcarriers_to_td(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_receiver_transform_subcarriers_to_td_batch(h, out, in, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_cancel_sc_interference_batch(gfdm_receiver* h, gfdm_complex* out,
                                               const gfdm_complex* td, const gfdm_complex* fd, int n,
                                               int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i)
        h->k->cancel_sc_interference(C(out) + i * bs, C(td) + i * bs, C(fd) + i * bs);
    REF_CATCH
}
int gfdm_receiver_cancel_sc_interference(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* td,
                                         const gfdm_complex* fd)
{
    return gfdm_receiver_cancel_sc_interference_batch(h, out, td, fd, 1, GFDM_MEM_HOST);
}

/* TEST-ONLY entry (not part of include/gfdm_b200.h): the reference's legacy 2-D interface run through its OWN
 * vector<vector<>> methods (lib/receiver_kernel_cc.cc:130-163,194-209,227-272), results serialised [k][m]:
 *   fd = filter_superposition(in); td = demodulate_subcarrier(fd); ic = remove_sc_interference(sc_symbols = td, fd).
 * tests/test_bindings.py compares the product's 2-D adapters (include/gfdm_b200.hpp) with these. */
extern "C" __attribute__((visibility("default"))) int gfdm_ref_receiver_legacy_2d(gfdm_receiver* h, gfdm_complex* fd_out,
                                                                                  gfdm_complex* td_out, gfdm_complex* ic_out,
                                                                                  const gfdm_complex* in)
{
    REF_TRY
    const int K = h->k->subcarriers(), M = h->k->timeslots();
    std::vector<std::vector<cf>> fd(K, std::vector<cf>(M)), td(K, std::vector<cf>(M));
    h->k->filter_superposition(fd, C(in));
    h->k->demodulate_subcarrier(td, fd);
    h->k->serialize_output(C(fd_out), fd);
    h->k->serialize_output(C(td_out), td);
    std::vector<std::vector<cf>> sym = td;
    h->k->remove_sc_interference(sym, fd);
    h->k->serialize_output(C(ic_out), sym);
    REF_CATCH
}

/* ---- advanced receiver -------------------------------------------------- */
int gfdm_advanced_receiver_create(gfdm_advanced_receiver** out, int M, int K, int L,
                                  const gfdm_complex* taps, int n_taps, const int* smap, int n_map,
                                  int ic_iter, const gfdm_constellation* c, int do_phase_comp)
{
    REF_TRY
    if (!c || c->n_points < 1 || !c->points)
        throw std::invalid_argument("constellation MUST hold at least one point");
    // ABI rules shared with gfdm_symbol_mapper_create: the sign rule indexes points[0..3]
    if (c->decision_rule != GFDM_DECISION_NEAREST && c->decision_rule != GFDM_DECISION_QPSK_SIGN)
        throw std::invalid_argument("unknown constellation decision rule!");
    if (c->decision_rule == GFDM_DECISION_QPSK_SIGN && c->n_points != 4)
        throw std::invalid_argument("the QPSK sign rule needs exactly 4 constellation points!");
    auto cst = std::make_shared<gr::digital::constellation>(vec(c->points, c->n_points),
                                                             c->decision_rule);
    auto h = new gfdm_advanced_receiver;
    try {
        h->k.reset(new advanced_receiver_kernel_cc(M, K, L, vec(taps, n_taps),
                                                   std::vector<int>(smap, smap + n_map), ic_iter,
                                                   cst, do_phase_comp));
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    REF_CATCH
}
void gfdm_advanced_receiver_destroy(gfdm_advanced_receiver* h) { delete h; }
int gfdm_advanced_receiver_block_size(const gfdm_advanced_receiver* h) { return h->k->block_size(); }
int gfdm_advanced_receiver_set_ic(gfdm_advanced_receiver* h, int v) { h->k->set_ic(v); return GFDM_OK; }
int gfdm_advanced_receiver_get_ic(const gfdm_advanced_receiver* h) { return h->k->get_ic(); }
int gfdm_advanced_receiver_set_phase_compensation(gfdm_advanced_receiver* h, int v)
{
    h->k->set_phase_compensation(v);
    return GFDM_OK;
}
int gfdm_advanced_receiver_get_phase_compensation(const gfdm_advanced_receiver* h)
{
    return h->k->get_phase_compensation();
}
int gfdm_advanced_receiver_work_batch(gfdm_advanced_receiver* h, gfdm_complex* out,
                                      const gfdm_complex* in, const gfdm_complex* eq, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size();
    for (int i = 0; i < n; ++i) {
        if (eq)
            h->k->generic_work_equalize(C(out) + i * bs, C(in) + i * bs, C(eq) + i * bs);
        else
            h->k->generic_work(C(out) + i * bs, C(in) + i * bs);
    }
    REF_CATCH
}
int gfdm_advanced_receiver_work(gfdm_advanced_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_advanced_receiver_work_batch(h, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_advanced_receiver_work_equalize(gfdm_advanced_receiver* h, gfdm_complex* out,
                                         const gfdm_complex* in, const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_advanced_receiver_work_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}

/* ---- resource mapper ---------------------------------------------------- */
int gfdm_resource_mapper_create(gfdm_resource_mapper** out, int M, int K, int A, const int* smap,
                                int n_map, int per_timeslot, int is_mapper)
{
    REF_TRY
    auto h = new gfdm_resource_mapper;
    try {
        h->k.reset(new resource_mapper_kernel_cc(M, K, A, std::vector<int>(smap, smap + n_map),
                                                 per_timeslot != 0, is_mapper != 0));
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    REF_CATCH
}
void gfdm_resource_mapper_destroy(gfdm_resource_mapper* h) { delete h; }
size_t gfdm_resource_mapper_frame_size(const gfdm_resource_mapper* h) { return h->k->frame_size(); }
size_t gfdm_resource_mapper_block_size(const gfdm_resource_mapper* h) { return h->k->block_size(); }
size_t gfdm_resource_mapper_input_vector_size(const gfdm_resource_mapper* h) { return h->k->input_vector_size(); }
size_t gfdm_resource_mapper_output_vector_size(const gfdm_resource_mapper* h) { return h->k->output_vector_size(); }
int gfdm_resource_mapper_map_to_resources(gfdm_resource_mapper* h, gfdm_complex* out,
                                          const gfdm_complex* in, size_t n)
{
    REF_TRY
    h->k->map_to_resources(C(out), C(in), n);
    REF_CATCH
}
int gfdm_resource_mapper_demap_from_resources(gfdm_resource_mapper* h, gfdm_complex* out,
                                              const gfdm_complex* in, size_t n)
{
    REF_TRY
    h->k->demap_from_resources(C(out), C(in), n);
    REF_CATCH
}
int gfdm_resource_mapper_map_to_resources_batch(gfdm_resource_mapper* h, gfdm_complex* out,
                                                const gfdm_complex* in, size_t sz, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    for (int i = 0; i < n; ++i)
        h->k->map_to_resources(C(out) + i * h->k->frame_size(), C(in) + i * sz, sz);
    REF_CATCH
}
int gfdm_resource_mapper_demap_from_resources_batch(gfdm_resource_mapper* h, gfdm_complex* out,
                                                    const gfdm_complex* in, size_t sz, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    // NOTE: demap_per_subcarrier writes one element past `sz` when one more
    // resource exists (lib/resource_mapper_kernel_cc.cc:155-159); demap into a
    // padded temporary so that the quirk cannot clobber the next frame.
    std::vector<cf> tmp(sz + 1);
    for (int i = 0; i < n; ++i) {
        h->k->demap_from_resources(tmp.data(), C(in) + i * h->k->frame_size(), sz);
        memcpy(C(out) + i * sz, tmp.data(), sizeof(cf) * sz);
    }
    REF_CATCH
}

/* ---- cyclic prefixer ---------------------------------------------------- */
int gfdm_cyclic_prefixer_create(gfdm_cyclic_prefixer** out, int block_len, int cp_len, int cs_len,
                                int ramp_len, const gfdm_complex* w, int n_w, int cyclic_shift)
{
    REF_TRY
    auto h = new gfdm_cyclic_prefixer;
    try {
        h->k.reset(new add_cyclic_prefix_cc(block_len, cp_len, cs_len, ramp_len, vec(w, n_w), cyclic_shift));
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    REF_CATCH
}
void gfdm_cyclic_prefixer_destroy(gfdm_cyclic_prefixer* h) { delete h; }
int gfdm_cyclic_prefixer_block_size(const gfdm_cyclic_prefixer* h) { return h->k->block_size(); }
int gfdm_cyclic_prefixer_frame_size(const gfdm_cyclic_prefixer* h) { return h->k->frame_size(); }
int gfdm_cyclic_prefixer_cyclic_shift(const gfdm_cyclic_prefixer* h) { return h->k->cyclic_shift(); }
int gfdm_cyclic_prefixer_work(gfdm_cyclic_prefixer* h, gfdm_complex* out, const gfdm_complex* in)
{
    REF_TRY
    h->k->generic_work(C(out), C(in));
    REF_CATCH
}
int gfdm_cyclic_prefixer_add_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                           const gfdm_complex* in, int shift)
{
    REF_TRY
    h->k->add_cyclic_prefix(C(out), C(in), shift);
    REF_CATCH
}
int gfdm_cyclic_prefixer_remove_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                              const gfdm_complex* in)
{
    REF_TRY
    h->k->remove_cyclic_prefix(C(out), C(in));
    REF_CATCH
}
int gfdm_cyclic_prefixer_add_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                                 const gfdm_complex* in, int shift, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size(), fs = h->k->frame_size();
    for (int i = 0; i < n; ++i) h->k->add_cyclic_prefix(C(out) + i * fs, C(in) + i * bs, shift);
    REF_CATCH
}
int gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                                    const gfdm_complex* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t bs = h->k->block_size(), fs = h->k->frame_size();
    for (int i = 0; i < n; ++i) h->k->remove_cyclic_prefix(C(out) + i * bs, C(in) + i * fs);
    REF_CATCH
}

/* ---- channel estimator -------------------------------------------------- */
int gfdm_channel_estimator_create(gfdm_channel_estimator** out, int M, int K, int A, int is_dc_free,
                                  int which, const gfdm_complex* preamble, int n_preamble)
{
    REF_TRY
    if (n_preamble < 2 * K)
        throw std::invalid_argument("preamble MUST hold at least 2 * fft_len samples");
    // odd A overflows the reference's own frame buffer in interpolate_frame (:238-274): rejected by the ABI
    if (A % 2) throw std::invalid_argument("active_subcarriers MUST be even (the reference writes past its frame buffer for odd values)");
    auto h = new gfdm_channel_estimator;
    try {
        h->k.reset(new preamble_channel_estimator_cc(M, K, A, is_dc_free != 0, which,
                                                     vec(preamble, n_preamble)));
    } catch (...) {
        delete h;
        throw;
    }
    h->A = A;
    *out = h;
    REF_CATCH
}
void gfdm_channel_estimator_destroy(gfdm_channel_estimator* h) { delete h; }
int gfdm_channel_estimator_fft_len(const gfdm_channel_estimator* h) { return h->k->fft_len(); }
int gfdm_channel_estimator_timeslots(const gfdm_channel_estimator* h) { return h->k->timeslots(); }
int gfdm_channel_estimator_frame_len(const gfdm_channel_estimator* h) { return h->k->frame_len(); }
int gfdm_channel_estimator_active_subcarriers(const gfdm_channel_estimator* h) { return h->k->active_subcarriers(); }
int gfdm_channel_estimator_is_dc_free(const gfdm_channel_estimator* h) { return h->k->is_dc_free(); }
int gfdm_channel_estimator_preamble_filter_taps(const gfdm_channel_estimator* h, float* o)
{
    auto t = h->k->preamble_filter_taps();
    memcpy(o, t.data(), sizeof(float) * t.size());
    return GFDM_OK;
}
int gfdm_channel_estimator_estimate_preamble_channel(gfdm_channel_estimator* h, gfdm_complex* o,
                                                     const gfdm_complex* rx)
{
    REF_TRY
    h->k->estimate_preamble_channel(C(o), C(rx));
    REF_CATCH
}
int gfdm_channel_estimator_filter_preamble_estimate(gfdm_channel_estimator* h, gfdm_complex* o,
                                                    const gfdm_complex* e)
{
    REF_TRY
    h->k->filter_preamble_estimate(C(o), C(e));
    REF_CATCH
}
int gfdm_channel_estimator_interpolate_frame(gfdm_channel_estimator* h, gfdm_complex* o,
                                             const gfdm_complex* e)
{
    REF_TRY
    h->k->interpolate_frame(C(o), C(e));
    REF_CATCH
}
int gfdm_channel_estimator_prepare_for_zf(gfdm_channel_estimator* h, gfdm_complex* o,
                                          const gfdm_complex* e)
{
    REF_TRY
    h->k->prepare_for_zf(C(o), C(e));
    REF_CATCH
}
int gfdm_channel_estimator_estimate_frame_batch(gfdm_channel_estimator* h, gfdm_complex* o,
                                                const gfdm_complex* rx, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t fl = h->k->frame_len(), pl = 2 * h->k->fft_len();
    for (int i = 0; i < n; ++i) h->k->estimate_frame(C(o) + i * fl, C(rx) + i * pl);
    REF_CATCH
}
int gfdm_channel_estimator_estimate_frame(gfdm_channel_estimator* h, gfdm_complex* o,
                                          const gfdm_complex* rx)
{
    return gfdm_channel_estimator_estimate_frame_batch(h, o, rx, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_estimate_snr_batch(gfdm_channel_estimator* h, float* snr, float* cnrs,
                                              const gfdm_complex* rx, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t pl = 2 * h->k->fft_len();
    std::vector<float> c;
    for (int i = 0; i < n; ++i) {
        snr[i] = h->k->estimate_snr(c, C(rx) + i * pl);
        if (cnrs) memcpy(cnrs + (size_t)i * h->A, c.data(), sizeof(float) * h->A);
    }
    REF_CATCH
}
int gfdm_channel_estimator_estimate_snr(gfdm_channel_estimator* h, float* snr, float* cnrs,
                                        const gfdm_complex* rx)
{
    return gfdm_channel_estimator_estimate_snr_batch(h, snr, cnrs, rx, 1, GFDM_MEM_HOST);
}

/* ---- transmitter -------------------------------------------------------- */
int gfdm_transmitter_create(gfdm_transmitter** out, int M, int K, int A, int cp, int cs, int ramp,
                            const int* smap, int n_map, int per_timeslot, int L,
                            const gfdm_complex* taps, int n_taps, const gfdm_complex* w, int n_w,
                            const int* shifts, int n_shifts, const gfdm_complex* const* preambles,
                            const int* preamble_sizes, int n_preambles)
{
    REF_TRY
    if (n_preambles < 1) // the reference reads preambles[0] unconditionally (transmitter_kernel.cc:54)
        throw std::invalid_argument("at least one preamble is required");
    std::vector<std::vector<cf>> pre;
    for (int i = 0; i < n_preambles; ++i) pre.push_back(vec(preambles[i], preamble_sizes[i]));
    auto h = new gfdm_transmitter;
    try {
        h->k.reset(new transmitter_kernel(M, K, A, cp, cs, ramp, std::vector<int>(smap, smap + n_map),
                                          per_timeslot != 0, L, vec(taps, n_taps), vec(w, n_w),
                                          std::vector<int>(shifts, shifts + n_shifts), pre));
    } catch (...) {
        delete h;
        throw;
    }
    h->N = M * K;
    h->frame.resize(h->N);
    *out = h;
    REF_CATCH
}
void gfdm_transmitter_destroy(gfdm_transmitter* h) { delete h; }
int gfdm_transmitter_input_vector_size(const gfdm_transmitter* h) { return h->k->input_vector_size(); }
int gfdm_transmitter_output_vector_size(const gfdm_transmitter* h) { return h->k->output_vector_size(); }
int gfdm_transmitter_n_cyclic_shifts(const gfdm_transmitter* h) { return (int)h->k->cyclic_shifts().size(); }
int gfdm_transmitter_cyclic_shifts(const gfdm_transmitter* h, int* o)
{
    const auto& s = h->k->cyclic_shifts();
    memcpy(o, s.data(), sizeof(int) * s.size());
    return GFDM_OK;
}
int gfdm_transmitter_work(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int n)
{
    REF_TRY
    h->k->generic_work(C(out), C(in), n);
    REF_CATCH
}
int gfdm_transmitter_modulate(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int n)
{
    REF_TRY
    h->k->modulate(C(out), C(in), n);
    REF_CATCH
}
int gfdm_transmitter_add_frame(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int shift)
{
    REF_TRY
    bool known = false;
    for (int s : h->k->cyclic_shifts()) known |= (s == shift);
    if (!known) throw std::invalid_argument("cyclic_shift has no preamble");
    h->k->add_frame(C(out), C(in), shift);
    REF_CATCH
}
int gfdm_transmitter_work_batch(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in,
                                int nin, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t os = h->k->output_vector_size();
    for (int i = 0; i < n; ++i) h->k->generic_work(C(out) + i * os, C(in) + (size_t)i * nin, nin);
    REF_CATCH
}
int gfdm_transmitter_set_chain_fusion(gfdm_transmitter*, int) { return GFDM_OK; }
int gfdm_transmitter_work_all_batch(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in,
                                    int nin, int n, int mem)
{
    REQUIRE_HOST(mem)
    REF_TRY
    const size_t os = h->k->output_vector_size();
    const auto& shifts = h->k->cyclic_shifts();
    for (int i = 0; i < n; ++i) { // loop of lib/transmitter_cc_impl.cc:165-177
        h->k->modulate(h->frame.data(), C(in) + (size_t)i * nin, nin);
        for (size_t a = 0; a < shifts.size(); ++a)
            h->k->add_frame(C(out) + (a * n + i) * os, h->frame.data(), shifts[a]);
    }
    REF_CATCH
}

} // extern "C"
