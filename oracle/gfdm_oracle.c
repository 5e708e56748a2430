/*
 * oracle/gfdm_oracle.c -- TEST INFRASTRUCTURE ONLY (the "port" oracle).
 *
 * Plain-C CPU restatement of gr-gfdm's GNU-Radio-free kernel layer, exporting
 * the same C ABI as the product library (include/gfdm_b200.h).  Every function
 * cites the reference file:line it follows.  Arithmetic model: arrays the
 * reference stores as complex<float> are stored as float here too; the leaf
 * DFT (FFTW in the reference) is evaluated in double precision and rounded to
 * float once; elementwise VOLK leaves are single-precision scalar loops.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *  (1) the golden vectors generated from the reference's own NumPy spec
 *      (tests/golden/ npz files, tests/golden/make_golden.py), and
 *  (2) oracle/_ref/libgfdm_ref.so = the unmodified reference C++ sources
 *      compiled against oracle/shim/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load this library; the product never routes through it.
 */
#define _GNU_SOURCE /* M_PI */
#include "../include/gfdm_b200.h"

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef gfdm_complex cf;
typedef double complex dc;

static _Thread_local char g_err[512] = "";

static int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

#define REQUIRE_HOST(mem)    \
    if ((mem) != GFDM_MEM_HOST && (mem) != GFDM_MEM_HOST_ASYNC) /* a CPU library is trivially 'complete at return' */ \
        return fail(GFDM_ERR_UNSUPPORTED, "oracle-port is CPU only: GFDM_MEM_DEVICE is not supported");

static inline cf cf_make(float re, float im) { cf r = { re, im }; return r; }
static inline cf cf_add(cf a, cf b) { return cf_make(a.re + b.re, a.im + b.im); }
static inline cf cf_sub(cf a, cf b) { return cf_make(a.re - b.re, a.im - b.im); }
/* volk_32fc_x2_multiply_32fc generic: (ar*br - ai*bi, ar*bi + ai*br) in float */
static inline cf cf_mul(cf a, cf b) { return cf_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }

/* ---- leaf DFT: double-precision mixed radix, unnormalised (FFTW semantics,
 *      lib/gfdm_kernel_utils.cc:45-49: fftwf_plan_dft_1d, FORWARD = exp(-j..)) */
typedef struct {
    int n;
    int sign;
    dc* tw;   /* tw[j] = exp(sign * 2 pi i j / n) */
    dc* a;    /* work buffers */
    dc* b;
} dft_plan;

static int smallest_factor(int n)
{
    if (n % 2 == 0) return 2;
    for (int p = 3; (long)p * p <= n; p += 2)
        if (n % p == 0) return p;
    return n;
}

static void dft_rec(const dft_plan* pl, int n, const dc* in, int stride, dc* out)
{
    if (n == 1) {
        out[0] = in[0];
        return;
    }
    const int p = smallest_factor(n);
    const int m = n / p;
    const int tws = pl->n / n;
    for (int r = 0; r < p; ++r) dft_rec(pl, m, in + (size_t)r * stride, stride * p, out + (size_t)r * m);
    dc t[p];
    dc acc[p];
    for (int k = 0; k < m; ++k) {
        for (int r = 0; r < p; ++r) t[r] = out[(size_t)r * m + k] * pl->tw[((long)r * k * tws) % pl->n];
        for (int q = 0; q < p; ++q) {
            dc s = t[0];
            for (int r = 1; r < p; ++r) s += t[r] * pl->tw[((long)r * q * m * tws) % pl->n];
            acc[q] = s;
        }
        for (int q = 0; q < p; ++q) out[(size_t)q * m + k] = acc[q];
    }
}

static int dft_init(dft_plan* pl, int n, int forward)
{
    pl->n = n;
    pl->sign = forward ? -1 : 1;
    pl->tw = (dc*)malloc(sizeof(dc) * (size_t)n);
    pl->a = (dc*)malloc(sizeof(dc) * (size_t)n);
    pl->b = (dc*)malloc(sizeof(dc) * (size_t)n);
    if (!pl->tw || !pl->a || !pl->b) return -1;
    for (int j = 0; j < n; ++j) {
        const double ph = pl->sign * 2.0 * M_PI * (double)j / (double)n;
        pl->tw[j] = cos(ph) + I * sin(ph);
    }
    return 0;
}

static void dft_free(dft_plan* pl)
{
    free(pl->tw);
    free(pl->a);
    free(pl->b);
    pl->tw = pl->a = pl->b = NULL;
}

/* out may alias in */
static void dft_exec(dft_plan* pl, cf* out, const cf* in)
{
    for (int i = 0; i < pl->n; ++i) pl->a[i] = (double)in[i].re + I * (double)in[i].im;
    dft_rec(pl, pl->n, pl->a, 1, pl->b);
    for (int i = 0; i < pl->n; ++i) out[i] = cf_make((float)creal(pl->b[i]), (float)cimag(pl->b[i]));
}

/* taps <- taps / sqrt(sum|taps|^2 / M): modulator_kernel_cc.cc:71-90, receiver_kernel_cc.cc:99-118.
 * The conjugate dot product accumulates in float; std::sqrt(std::abs(res) / n_timeslots)
 * is the float overload, `1. /` promotes to double, the result is cast to float. */
static void normalize_taps(cf* dst, const cf* src, int n, int n_timeslots)
{
    float er = 0.0f, ei = 0.0f;
    for (int i = 0; i < n; ++i) {
        er += src[i].re * src[i].re + src[i].im * src[i].im;
        ei += src[i].im * src[i].re - src[i].re * src[i].im;
    }
    const float absres = hypotf(er, ei);
    const float sf = (float)(1. / (double)sqrtf(absres / (float)n_timeslots)); /* std::sqrt(float) */
    const cf s = cf_make(sf, 0.0f);
    for (int i = 0; i < n; ++i) dst[i] = cf_mul(src[i], s);
}

/* ======================================================================== */
const char* gfdm_last_error(void) { return g_err; }
/* used by gfdm_oracle_next.c (linked into this library) */
void gfdm_oracle_set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
const char* gfdm_backend(void) { return "oracle-port"; }
int gfdm_device_count(void) { return 0; }
int gfdm_set_device(int device) { (void)device; return GFDM_OK; }
int gfdm_set_stream(void* h, void* s) { (void)h; (void)s; return GFDM_OK; }
int gfdm_sync(void* h) { (void)h; return GFDM_OK; }
long long gfdm_launch_count(void* h) { (void)h; return 0; }
const char* gfdm_last_kernel(void* h) { (void)h; return "cpu"; }

/* lib/gfdm_kernel_utils.cc:59-65 */
int gfdm_calculate_signal_energy(float* energy, const cf* in, int n)
{
    float e = 0.0f;
    for (int i = 0; i < n; ++i) e += in[i].re * in[i].re + in[i].im * in[i].im;
    *energy = e;
    return GFDM_OK;
}

struct gfdm_fft { dft_plan pl; };

int gfdm_fft_create(gfdm_fft** out, int fft_size, int forward)
{
    if (fft_size < 1) return fail(GFDM_ERR_INVALID_ARGUMENT, "fft_size MUST be positive");
    gfdm_fft* h = (gfdm_fft*)calloc(1, sizeof(*h));
    if (!h || dft_init(&h->pl, fft_size, forward)) return fail(GFDM_ERR_RUNTIME, "out of memory");
    *out = h;
    return GFDM_OK;
}
void gfdm_fft_destroy(gfdm_fft* h)
{
    if (!h) return;
    dft_free(&h->pl);
    free(h);
}
int gfdm_fft_execute_batch(gfdm_fft* h, cf* out, const cf* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i) dft_exec(&h->pl, out + (size_t)i * h->pl.n, in + (size_t)i * h->pl.n);
    return GFDM_OK;
}

/* ---- modulator_kernel_cc ------------------------------------------------ */
struct gfdm_modulator {
    int M, K, L, N;
    cf* taps;
    dft_plan sub_fft; /* M, forward */
    dft_plan ifft;    /* N, backward */
    cf* sub_out;      /* M */
    cf* ifft_in;      /* N */
    cf* ifft_out;     /* N */
};

/* lib/modulator_kernel_cc.cc:30-63 */
int gfdm_modulator_create(gfdm_modulator** out, int M, int K, int L, const cf* taps, int n_taps)
{
    if (n_taps != M * L) {
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "number of frequency taps(%d) MUST be equal to n_timeslots(%d) * overlap(%d) = %d!",
                 n_taps, M, L, M * L);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    if (M < 1 || K < 1 || L < 1) return fail(GFDM_ERR_INVALID_ARGUMENT, "timeslots, subcarriers and overlap MUST be positive");
    gfdm_modulator* h = (gfdm_modulator*)calloc(1, sizeof(*h));
    h->M = M; h->K = K; h->L = L; h->N = M * K;
    h->taps = (cf*)malloc(sizeof(cf) * (size_t)n_taps);
    normalize_taps(h->taps, taps, n_taps, M);
    dft_init(&h->sub_fft, M, 1);
    dft_init(&h->ifft, h->N, 0);
    h->sub_out = (cf*)malloc(sizeof(cf) * (size_t)M);
    h->ifft_in = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    h->ifft_out = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    *out = h;
    return GFDM_OK;
}
void gfdm_modulator_destroy(gfdm_modulator* h)
{
    if (!h) return;
    free(h->taps); free(h->sub_out); free(h->ifft_in); free(h->ifft_out);
    dft_free(&h->sub_fft); dft_free(&h->ifft);
    free(h);
}
int gfdm_modulator_block_size(const gfdm_modulator* h) { return h->N; }
int gfdm_modulator_filter_taps(const gfdm_modulator* h, cf* o)
{
    memcpy(o, h->taps, sizeof(cf) * (size_t)(h->M * h->L));
    return GFDM_OK;
}
/* lib/modulator_kernel_cc.cc:98-141 */
int gfdm_modulator_work(gfdm_modulator* h, cf* p_out, const cf* p_in)
{
    const int M = h->M, K = h->K, L = h->L, N = h->N;
    const int part_len = (M * L / 2 < M) ? M * L / 2 : M; /* :101 */
    memset(h->ifft_in, 0, sizeof(cf) * (size_t)N);         /* :104 */
    for (int k = 0; k < K; ++k) {
        dft_exec(&h->sub_fft, h->sub_out, p_in + (size_t)k * M); /* :109-110 */
        for (int i = 0; i < L; ++i) {
            const int src = ((i + L / 2) % L) * M;              /* :118 */
            const int tgt = ((k + i + K - (L / 2)) % K) * M;    /* :119-121 */
            for (int m = 0; m < part_len; ++m) {
                const cf f = cf_mul(h->sub_out[m], h->taps[src + m]); /* :123 */
                h->ifft_in[tgt + m] = cf_add(h->ifft_in[tgt + m], f); /* :128 */
            }
        }
    }
    dft_exec(&h->ifft, h->ifft_out, h->ifft_in); /* :137 */
    const cf s = cf_make((float)(1.0 / N), 0.0f);
    for (int n = 0; n < N; ++n) p_out[n] = cf_mul(h->ifft_out[n], s); /* :139-140 */
    return GFDM_OK;
}
int gfdm_modulator_work_batch(gfdm_modulator* h, cf* out, const cf* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i) gfdm_modulator_work(h, out + (size_t)i * h->N, in + (size_t)i * h->N);
    return GFDM_OK;
}

/* ---- receiver_kernel_cc ------------------------------------------------- */
struct gfdm_receiver {
    int M, K, L, N;
    cf* taps;
    cf* ic_taps;
    dft_plan in_fft;  /* N forward */
    dft_plan sc_ifft; /* M backward */
    dft_plan sc_fft;  /* M forward */
    cf* fft_out;      /* N */
    cf* equalized;    /* N */
    cf* sc_filtered;  /* N */
    cf* tmpM;         /* M */
};

/* lib/receiver_kernel_cc.cc:31-88 */
int gfdm_receiver_create(gfdm_receiver** out, int M, int K, int L, const cf* taps, int n_taps)
{
    if (n_taps != M * L) {
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "number of frequency taps(%d) MUST be equal to n_timeslots(%d) * overlap(%d) = %d!",
                 n_taps, M, L, M * L);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    if (L < 2) return fail(GFDM_ERR_INVALID_ARGUMENT, "overlap MUST be greater or equal 2");
    if (M < 1 || K < 1) return fail(GFDM_ERR_INVALID_ARGUMENT, "timeslots and subcarriers MUST be positive");
    gfdm_receiver* h = (gfdm_receiver*)calloc(1, sizeof(*h));
    h->M = M; h->K = K; h->L = L; h->N = M * K;
    h->taps = (cf*)malloc(sizeof(cf) * (size_t)n_taps);
    normalize_taps(h->taps, taps, n_taps, M);
    h->ic_taps = (cf*)malloc(sizeof(cf) * (size_t)M);
    for (int m = 0; m < M; ++m) h->ic_taps[m] = cf_mul(h->taps[m], h->taps[M * (L - 1) + m]); /* :56-63 */
    dft_init(&h->in_fft, h->N, 1);
    dft_init(&h->sc_ifft, M, 0);
    dft_init(&h->sc_fft, M, 1);
    h->fft_out = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    h->equalized = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    h->sc_filtered = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    h->tmpM = (cf*)malloc(sizeof(cf) * (size_t)M);
    *out = h;
    return GFDM_OK;
}
void gfdm_receiver_destroy(gfdm_receiver* h)
{
    if (!h) return;
    free(h->taps); free(h->ic_taps); free(h->fft_out); free(h->equalized); free(h->sc_filtered); free(h->tmpM);
    dft_free(&h->in_fft); dft_free(&h->sc_ifft); dft_free(&h->sc_fft);
    free(h);
}
int gfdm_receiver_block_size(const gfdm_receiver* h) { return h->N; }
int gfdm_receiver_timeslots(const gfdm_receiver* h) { return h->M; }
int gfdm_receiver_subcarriers(const gfdm_receiver* h) { return h->K; }
int gfdm_receiver_overlap(const gfdm_receiver* h) { return h->L; }
int gfdm_receiver_filter_taps(const gfdm_receiver* h, cf* o)
{
    memcpy(o, h->taps, sizeof(cf) * (size_t)(h->M * h->L));
    return GFDM_OK;
}
int gfdm_receiver_ic_filter_taps(const gfdm_receiver* h, cf* o)
{
    memcpy(o, h->ic_taps, sizeof(cf) * (size_t)h->M);
    return GFDM_OK;
}

/* lib/receiver_kernel_cc.cc:165-192 */
static void rx_filter_downsample_fd(gfdm_receiver* h, cf* p_out, const cf* p_in)
{
    const int M = h->M, K = h->K, L = h->L;
    memset(p_out, 0, sizeof(cf) * (size_t)h->N);
    for (int k = 0; k < K; ++k) {
        for (int i = 0; i < L; ++i) {
            const int src = ((k + i + K - (L / 2)) % K) * M; /* :175-177 */
            const int tgt = ((i + L / 2) % L) * M;           /* :178 */
            for (int m = 0; m < M; ++m) {
                const cf f = cf_mul(h->taps[tgt + m], p_in[src + m]);           /* :180-183 */
                p_out[(size_t)k * M + m] = cf_add(p_out[(size_t)k * M + m], f); /* :185-188 */
            }
        }
    }
}

/* lib/receiver_kernel_cc.cc:301-320; eq == NULL -> fft_filter_downsample */
static void rx_fft_filter(gfdm_receiver* h, cf* p_out, const cf* p_in, const cf* eq)
{
    dft_exec(&h->in_fft, h->fft_out, p_in); /* :304-305 / :313-314 */
    if (eq) {
        /* volk_32fc_x2_divide_32fc :315 -- a * conj(b) / |b|^2 */
        for (int n = 0; n < h->N; ++n) {
            const cf a = h->fft_out[n], b = eq[n];
            const cf num = cf_mul(a, cf_make(b.re, -b.im));
            const float den = b.re * b.re + b.im * b.im;
            h->equalized[n] = cf_make(num.re / den, num.im / den);
        }
        rx_filter_downsample_fd(h, p_out, h->equalized);
    } else {
        rx_filter_downsample_fd(h, p_out, h->fft_out);
    }
}

/* lib/receiver_kernel_cc.cc:211-225 */
static void rx_to_td(gfdm_receiver* h, cf* p_out, const cf* p_in)
{
    const cf s = cf_make((float)(1.0 / h->M), 0.0f);
    for (int k = 0; k < h->K; ++k) {
        dft_exec(&h->sc_ifft, h->tmpM, p_in + (size_t)k * h->M);
        for (int m = 0; m < h->M; ++m) p_out[(size_t)k * h->M + m] = cf_mul(h->tmpM[m], s);
    }
}

/* lib/receiver_kernel_cc.cc:274-299 */
static void rx_cancel(gfdm_receiver* h, cf* p_out, const cf* td, const cf* fd)
{
    const int M = h->M, K = h->K;
    for (int k = 0; k < K; ++k) {
        const int prev = (k - 1 + K) % K, next = (k + 1 + K) % K;
        for (int m = 0; m < M; ++m) h->tmpM[m] = cf_add(td[(size_t)prev * M + m], td[(size_t)next * M + m]);
        dft_exec(&h->sc_fft, h->tmpM, h->tmpM);
        for (int m = 0; m < M; ++m) {
            const cf f = cf_mul(h->ic_taps[m], h->tmpM[m]);
            p_out[(size_t)k * M + m] = cf_sub(fd[(size_t)k * M + m], f);
        }
    }
}

int gfdm_receiver_work_batch(gfdm_receiver* h, cf* out, const cf* in, const cf* eq, int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i) { /* :322-334 */
        rx_fft_filter(h, h->sc_filtered, in + (size_t)i * h->N, eq ? eq + (size_t)i * h->N : NULL);
        rx_to_td(h, out + (size_t)i * h->N, h->sc_filtered);
    }
    return GFDM_OK;
}
int gfdm_receiver_work(gfdm_receiver* h, cf* out, const cf* in)
{
    return gfdm_receiver_work_batch(h, out, in, NULL, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_work_equalize(gfdm_receiver* h, cf* out, const cf* in, const cf* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_receiver_work_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_filter_downsample_batch(gfdm_receiver* h, cf* out, const cf* in, const cf* eq,
                                              int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i)
        rx_fft_filter(h, out + (size_t)i * h->N, in + (size_t)i * h->N, eq ? eq + (size_t)i * h->N : NULL);
    return GFDM_OK;
}
int gfdm_receiver_fft_filter_downsample(gfdm_receiver* h, cf* out, const cf* in)
{
    return gfdm_receiver_fft_filter_downsample_batch(h, out, in, NULL, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_equalize_filter_downsample(gfdm_receiver* h, cf* out, const cf* in, const cf* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_receiver_fft_filter_downsample_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_transform_subcarriers_to_td_batch(gfdm_receiver* h, cf* out, const cf* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i) rx_to_td(h, out + (size_t)i * h->N, in + (size_t)i * h->N);
    return GFDM_OK;
}
int gfdm_receiver_transform_subcarriers_to_td(gfdm_receiver* h, cf* out, const cf* in)
{
    return gfdm_receiver_transform_subcarriers_to_td_batch(h, out, in, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_cancel_sc_interference_batch(gfdm_receiver* h, cf* out, const cf* td, const cf* fd,
                                               int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i)
        rx_cancel(h, out + (size_t)i * h->N, td + (size_t)i * h->N, fd + (size_t)i * h->N);
    return GFDM_OK;
}
int gfdm_receiver_cancel_sc_interference(gfdm_receiver* h, cf* out, const cf* td, const cf* fd)
{
    return gfdm_receiver_cancel_sc_interference_batch(h, out, td, fd, 1, GFDM_MEM_HOST);
}

/* ---- advanced_receiver_kernel_cc ---------------------------------------- */
struct gfdm_advanced_receiver {
    gfdm_receiver* rx;
    int* smap;
    int n_map;
    int ic_iter;
    int phase_comp;
    cf* points;
    int n_points;
    int rule;
    cf* freq_block;
    cf* ic_time;
    cf* ic_freq;
};

/* lib/advanced_receiver_kernel_cc.cc:32-52 */
int gfdm_advanced_receiver_create(gfdm_advanced_receiver** out, int M, int K, int L, const cf* taps,
                                  int n_taps, const int* smap, int n_map, int ic_iter,
                                  const gfdm_constellation* c, int do_phase_compensation)
{
    if (!c || c->n_points < 1 || !c->points)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "constellation MUST hold at least one point");
    /* ABI rules shared with gfdm_symbol_mapper_create: the sign rule indexes points[0..3] */
    if (c->decision_rule != GFDM_DECISION_NEAREST && c->decision_rule != GFDM_DECISION_QPSK_SIGN)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "unknown constellation decision rule!");
    if (c->decision_rule == GFDM_DECISION_QPSK_SIGN && c->n_points != 4)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "the QPSK sign rule needs exactly 4 constellation points!");
    for (int i = 0; i < n_map; ++i)
        if (smap[i] < 0 || smap[i] >= K)
            return fail(GFDM_ERR_INVALID_ARGUMENT, "subcarrier_map entries MUST lie in [0, subcarriers)");
    gfdm_receiver* rx = NULL;
    const int st = gfdm_receiver_create(&rx, M, K, L, taps, n_taps);
    if (st) return st;
    gfdm_advanced_receiver* h = (gfdm_advanced_receiver*)calloc(1, sizeof(*h));
    h->rx = rx;
    h->n_map = n_map;
    h->smap = (int*)malloc(sizeof(int) * (size_t)(n_map > 0 ? n_map : 1));
    memcpy(h->smap, smap, sizeof(int) * (size_t)n_map);
    h->ic_iter = ic_iter;
    h->phase_comp = do_phase_compensation;
    h->n_points = c->n_points;
    h->rule = c->decision_rule;
    h->points = (cf*)malloc(sizeof(cf) * (size_t)c->n_points);
    memcpy(h->points, c->points, sizeof(cf) * (size_t)c->n_points);
    h->freq_block = (cf*)malloc(sizeof(cf) * (size_t)rx->N);
    h->ic_time = (cf*)malloc(sizeof(cf) * (size_t)rx->N);
    h->ic_freq = (cf*)malloc(sizeof(cf) * (size_t)rx->N);
    *out = h;
    return GFDM_OK;
}
void gfdm_advanced_receiver_destroy(gfdm_advanced_receiver* h)
{
    if (!h) return;
    gfdm_receiver_destroy(h->rx);
    free(h->smap); free(h->points); free(h->freq_block); free(h->ic_time); free(h->ic_freq);
    free(h);
}
int gfdm_advanced_receiver_block_size(const gfdm_advanced_receiver* h) { return h->rx->N; }
int gfdm_advanced_receiver_set_ic(gfdm_advanced_receiver* h, int v) { h->ic_iter = v; return GFDM_OK; }
int gfdm_advanced_receiver_get_ic(const gfdm_advanced_receiver* h) { return h->ic_iter; }
int gfdm_advanced_receiver_set_phase_compensation(gfdm_advanced_receiver* h, int v) { h->phase_comp = v; return GFDM_OK; }
int gfdm_advanced_receiver_get_phase_compensation(const gfdm_advanced_receiver* h) { return h->phase_comp; }

/* stand-in for gr::digital::constellation::decision_maker (see include/gfdm_b200.h) */
static int decide(const gfdm_advanced_receiver* h, cf s)
{
    if (h->rule == GFDM_DECISION_QPSK_SIGN) return 2 * (s.im > 0) + (s.re > 0);
    int best = 0;
    float dmin = 0.0f;
    for (int i = 0; i < h->n_points; ++i) {
        const float dr = s.re - h->points[i].re, di = s.im - h->points[i].im;
        const float d = dr * dr + di * di;
        if (i == 0 || d < dmin) { dmin = d; best = i; }
    }
    return best;
}

/* lib/advanced_receiver_kernel_cc.cc:56-123 */
static void adv_ic_iterations(gfdm_advanced_receiver* h, cf* p_out, cf* freq_block)
{
    const int M = h->rx->M, N = h->rx->N;
    for (int j = 0; j < h->ic_iter; ++j) {
        /* map_symbols_to_constellation_points :109-123 */
        memset(h->ic_time, 0, sizeof(cf) * (size_t)N);
        for (int a = 0; a < h->n_map; ++a) {
            const int k = h->smap[a];
            for (int m = 0; m < M; ++m) h->ic_time[k * M + m] = h->points[decide(h, p_out[k * M + m])];
        }
        if (h->phase_comp > 0 && j == 0) {
            /* calculate_phase_offset :78-91 (serial float accumulation) */
            float ph = 0.0f;
            for (int a = 0; a < h->n_map; ++a) {
                const int k = h->smap[a];
                for (int m = 0; m < M; ++m) {
                    const int pos = k * M + m;
                    ph += atan2f(h->ic_time[pos].im, h->ic_time[pos].re) - atan2f(p_out[pos].im, p_out[pos].re);
                }
            }
            ph = ph / (float)((size_t)h->n_map * (size_t)M);
            /* std::polar(1.0f, ph); rotator with phase_inc = 1 :63-70 */
            const cf rot = cf_make(cosf(ph), sinf(ph));
            for (int n = 0; n < N; ++n) freq_block[n] = cf_mul(freq_block[n], rot);
        }
        rx_cancel(h->rx, h->ic_freq, h->ic_time, freq_block); /* :72-73 */
        rx_to_td(h->rx, p_out, h->ic_freq);                   /* :74 */
    }
}

int gfdm_advanced_receiver_work_batch(gfdm_advanced_receiver* h, cf* out, const cf* in, const cf* eq,
                                      int n, int mem)
{
    REQUIRE_HOST(mem)
    const size_t N = (size_t)h->rx->N;
    for (int i = 0; i < n; ++i) { /* :93-107 */
        rx_fft_filter(h->rx, h->freq_block, in + i * N, eq ? eq + i * N : NULL);
        rx_to_td(h->rx, out + i * N, h->freq_block);
        adv_ic_iterations(h, out + i * N, h->freq_block);
    }
    return GFDM_OK;
}
int gfdm_advanced_receiver_work(gfdm_advanced_receiver* h, cf* out, const cf* in)
{
    return gfdm_advanced_receiver_work_batch(h, out, in, NULL, 1, GFDM_MEM_HOST);
}
int gfdm_advanced_receiver_work_equalize(gfdm_advanced_receiver* h, cf* out, const cf* in, const cf* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_advanced_receiver_work_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}

/* ---- resource_mapper_kernel_cc ------------------------------------------ */
struct gfdm_resource_mapper {
    int M, K, A;
    size_t block_size, frame_size;
    int per_timeslot, is_mapper;
    int* smap; /* sorted */
};

static int cmp_int(const void* a, const void* b)
{
    const int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

/* lib/resource_mapper_kernel_cc.cc:30-70 */
int gfdm_resource_mapper_create(gfdm_resource_mapper** out, int M, int K, int A, const int* smap,
                                int n_map, int per_timeslot, int is_mapper)
{
    char msg[256];
    if (A > K) {
        snprintf(msg, sizeof(msg), "active_subcarriers(%d) MUST be smaller or equal to subcarriers(%d)!", A, K);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    if (n_map != A) {
        snprintf(msg, sizeof(msg), "number of subcarrier_map entries(%d) MUST be equal to active_subcarriers(%d)!", n_map, A);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    int* s = (int*)malloc(sizeof(int) * (size_t)(n_map > 0 ? n_map : 1));
    memcpy(s, smap, sizeof(int) * (size_t)n_map);
    qsort(s, (size_t)n_map, sizeof(int), cmp_int); /* :56 */
    for (int i = 1; i < n_map; ++i)
        if (s[i] == s[i - 1]) {
            free(s);
            return fail(GFDM_ERR_INVALID_ARGUMENT, "All entries in subcarrier_map MUST be unique!");
        }
    if (n_map > 0 && s[0] < 0) {
        free(s);
        return fail(GFDM_ERR_INVALID_ARGUMENT, "All subcarrier indices MUST be greater or equal to ZERO!");
    }
    /* The reference tests `> subcarriers` (:65), which lets index == subcarriers
     * through and then writes out of bounds.  That is undefined behaviour, not
     * an output; every implementation of this ABI rejects index >= subcarriers
     * with the reference's message. */
    if (n_map > 0 && s[n_map - 1] >= K) {
        free(s);
        return fail(GFDM_ERR_INVALID_ARGUMENT, "All subcarrier indices MUST be smaller or equal to subcarriers!");
    }
    gfdm_resource_mapper* h = (gfdm_resource_mapper*)calloc(1, sizeof(*h));
    h->M = M; h->K = K; h->A = A;
    h->block_size = (size_t)M * A;
    h->frame_size = (size_t)M * K;
    h->per_timeslot = per_timeslot;
    h->is_mapper = is_mapper;
    h->smap = s;
    *out = h;
    return GFDM_OK;
}
void gfdm_resource_mapper_destroy(gfdm_resource_mapper* h)
{
    if (!h) return;
    free(h->smap);
    free(h);
}
size_t gfdm_resource_mapper_frame_size(const gfdm_resource_mapper* h) { return h->frame_size; }
size_t gfdm_resource_mapper_block_size(const gfdm_resource_mapper* h) { return h->block_size; }
size_t gfdm_resource_mapper_input_vector_size(const gfdm_resource_mapper* h) { return h->is_mapper ? h->block_size : h->frame_size; }
size_t gfdm_resource_mapper_output_vector_size(const gfdm_resource_mapper* h) { return h->is_mapper ? h->frame_size : h->block_size; }

/* lib/resource_mapper_kernel_cc.cc:74-89, 108-134 */
int gfdm_resource_mapper_map_to_resources(gfdm_resource_mapper* h, cf* p_out, const cf* p_in, size_t n)
{
    if (n > h->block_size) {
        char msg[256];
        snprintf(msg, sizeof(msg), "input vector size(%zu) MUST not exceed active_subcarriers * timeslots(%zu)!", n, h->block_size);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    memset(p_out, 0, sizeof(cf) * h->frame_size);
    size_t ctr = 0;
    if (h->per_timeslot) {
        for (int t = 0; t < h->M; ++t)
            for (int a = 0; a < h->A; ++a) {
                p_out[(size_t)h->M * h->smap[a] + t] = ctr < n ? p_in[ctr] : cf_make(0.0f, 0.0f);
                ctr++;
            }
    } else {
        for (int a = 0; a < h->A; ++a)
            for (int t = 0; t < h->M; ++t) {
                p_out[(size_t)h->M * h->smap[a] + t] = ctr < n ? p_in[ctr] : cf_make(0.0f, 0.0f);
                ctr++;
            }
    }
    return GFDM_OK;
}
/* lib/resource_mapper_kernel_cc.cc:91-106, 136-162.  The per-subcarrier branch
 * returns only after the counter EXCEEDS noutput_size (:155-159), i.e. it
 * writes element [noutput_size] too when one more resource exists; callers
 * that pass a full block never see it.  Restated faithfully here. */
int gfdm_resource_mapper_demap_from_resources(gfdm_resource_mapper* h, cf* p_out, const cf* p_in, size_t n)
{
    if (n > h->block_size) {
        char msg[256];
        snprintf(msg, sizeof(msg), "output vector size(%zu) MUST not exceed active_subcarriers * timeslots(%zu)!", n, h->block_size);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    memset(p_out, 0, sizeof(cf) * n);
    if (h->per_timeslot) {
        for (size_t i = 0; i < n; ++i) {
            const size_t t = i / (size_t)h->A;
            const int s = h->smap[i % (size_t)h->A];
            p_out[i] = p_in[(size_t)h->M * s + t];
        }
    } else {
        size_t ctr = 0;
        for (int a = 0; a < h->A; ++a)
            for (int t = 0; t < h->M; ++t) {
                p_out[ctr] = p_in[(size_t)h->M * h->smap[a] + t];
                ctr++;
                if (ctr > n) return GFDM_OK;
            }
    }
    return GFDM_OK;
}
int gfdm_resource_mapper_map_to_resources_batch(gfdm_resource_mapper* h, cf* out, const cf* in, size_t sz,
                                                int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i) {
        const int st = gfdm_resource_mapper_map_to_resources(h, out + (size_t)i * h->frame_size, in + (size_t)i * sz, sz);
        if (st) return st;
    }
    return GFDM_OK;
}
int gfdm_resource_mapper_demap_from_resources_batch(gfdm_resource_mapper* h, cf* out, const cf* in,
                                                    size_t sz, int n, int mem)
{
    REQUIRE_HOST(mem)
    cf* tmp = (cf*)malloc(sizeof(cf) * (sz + 1)); /* keeps the one-past write inside the frame */
    for (int i = 0; i < n; ++i) {
        const int st = gfdm_resource_mapper_demap_from_resources(h, tmp, in + (size_t)i * h->frame_size, sz);
        if (st) { free(tmp); return st; }
        memcpy(out + (size_t)i * sz, tmp, sizeof(cf) * sz);
    }
    free(tmp);
    return GFDM_OK;
}

/* ---- add_cyclic_prefix_cc ----------------------------------------------- */
struct gfdm_cyclic_prefixer {
    int block_len, cp_len, cs_len, ramp_len, cyclic_shift;
    cf* front;
    cf* back;
};

/* lib/add_cyclic_prefix_cc.cc:30-57 */
int gfdm_cyclic_prefixer_create(gfdm_cyclic_prefixer** out, int block_len, int cp_len, int cs_len,
                                int ramp_len, const cf* w, int n_w, int cyclic_shift)
{
    const int window_len = block_len + cp_len + cs_len;
    if (n_w != window_len && n_w != 2 * ramp_len) {
        char msg[256];
        snprintf(msg, sizeof(msg), "number of window taps(%d) MUST be equal to 2*ramp_len(%d) OR block_len+cp_len (%d)!",
                 n_w, 2 * ramp_len, window_len);
        return fail(GFDM_ERR_INVALID_ARGUMENT, msg);
    }
    if (block_len < 1 || cp_len < 0 || cs_len < 0 || ramp_len < 0 || ramp_len > n_w)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "block_len MUST be positive; cp_len, cs_len, ramp_len MUST NOT be negative");
    gfdm_cyclic_prefixer* h = (gfdm_cyclic_prefixer*)calloc(1, sizeof(*h));
    h->block_len = block_len; h->cp_len = cp_len; h->cs_len = cs_len; h->ramp_len = ramp_len;
    h->cyclic_shift = cyclic_shift;
    h->front = (cf*)malloc(sizeof(cf) * (size_t)(ramp_len > 0 ? ramp_len : 1));
    h->back = (cf*)malloc(sizeof(cf) * (size_t)(ramp_len > 0 ? ramp_len : 1));
    memcpy(h->front, w, sizeof(cf) * (size_t)ramp_len);                  /* :52-53 */
    memcpy(h->back, w + (n_w - ramp_len), sizeof(cf) * (size_t)ramp_len); /* :54-56 */
    *out = h;
    return GFDM_OK;
}
void gfdm_cyclic_prefixer_destroy(gfdm_cyclic_prefixer* h)
{
    if (!h) return;
    free(h->front); free(h->back);
    free(h);
}
int gfdm_cyclic_prefixer_block_size(const gfdm_cyclic_prefixer* h) { return h->block_len; }
int gfdm_cyclic_prefixer_frame_size(const gfdm_cyclic_prefixer* h) { return h->block_len + h->cp_len + h->cs_len; }
int gfdm_cyclic_prefixer_cyclic_shift(const gfdm_cyclic_prefixer* h) { return h->cyclic_shift; }

/* lib/add_cyclic_prefix_cc.cc:67-98 */
int gfdm_cyclic_prefixer_add_cyclic_prefix(gfdm_cyclic_prefixer* h, cf* out, const cf* in, int shift)
{
    const int N = h->block_len;
    if (shift < 0 || shift > h->cs_len || h->cp_len + shift > N)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "cyclic_shift MUST lie in [0, cs_len] and cp_len + cyclic_shift MUST NOT exceed block_len");
    const int cp_start = N - h->cp_len - shift;      /* :82 */
    const int shifted_cp_len = h->cp_len + shift;    /* :83 */
    memcpy(out, in + cp_start, sizeof(cf) * (size_t)shifted_cp_len);
    memcpy(out + shifted_cp_len, in, sizeof(cf) * (size_t)N);
    const int shifted_cs_len = h->cs_len - shift;    /* :87 */
    memcpy(out + shifted_cp_len + N, in, sizeof(cf) * (size_t)shifted_cs_len);
    if (h->ramp_len > 0) {                           /* :73-75, 92-98 */
        const int tail = N + h->cp_len + h->cs_len - h->ramp_len;
        for (int i = 0; i < h->ramp_len; ++i) out[i] = cf_mul(out[i], h->front[i]);
        for (int i = 0; i < h->ramp_len; ++i) out[tail + i] = cf_mul(out[tail + i], h->back[i]);
    }
    return GFDM_OK;
}
int gfdm_cyclic_prefixer_work(gfdm_cyclic_prefixer* h, cf* out, const cf* in)
{
    return gfdm_cyclic_prefixer_add_cyclic_prefix(h, out, in, h->cyclic_shift); /* :61-64 */
}
/* lib/add_cyclic_prefix_cc.cc:100-104 */
int gfdm_cyclic_prefixer_remove_cyclic_prefix(gfdm_cyclic_prefixer* h, cf* out, const cf* in)
{
    memcpy(out, in + h->cp_len, sizeof(cf) * (size_t)h->block_len);
    return GFDM_OK;
}
int gfdm_cyclic_prefixer_add_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, cf* out, const cf* in, int shift,
                                                 int n, int mem)
{
    REQUIRE_HOST(mem)
    const size_t fs = (size_t)gfdm_cyclic_prefixer_frame_size(h);
    for (int i = 0; i < n; ++i) {
        const int st = gfdm_cyclic_prefixer_add_cyclic_prefix(h, out + i * fs, in + (size_t)i * h->block_len, shift);
        if (st) return st;
    }
    return GFDM_OK;
}
int gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, cf* out, const cf* in, int n, int mem)
{
    REQUIRE_HOST(mem)
    const size_t fs = (size_t)gfdm_cyclic_prefixer_frame_size(h);
    for (int i = 0; i < n; ++i) gfdm_cyclic_prefixer_remove_cyclic_prefix(h, out + (size_t)i * h->block_len, in + i * fs);
    return GFDM_OK;
}

/* ---- preamble_channel_estimator_cc -------------------------------------- */
struct gfdm_channel_estimator {
    int M, K, A, dc_free, which;
    dft_plan fftK;  /* K forward */
    dft_plan fft2K; /* 2K forward */
    cf* inv0;
    cf* inv1;
    float g[9];
    cf* tmpK;
    cf* tmpK2;
    cf* inter;    /* A + 9 + dc */
    cf* pre_est;  /* K */
    cf* filt;     /* A + dc */
    cf* snr_out;  /* 2K */
};

/* lib/preamble_channel_estimator_cc.cc:111-119 */
static void est_init_inv(gfdm_channel_estimator* h, cf* dst, const cf* part)
{
    dft_exec(&h->fftK, h->tmpK, part);
    for (int i = 0; i < h->K; ++i) {
        /* std::complex<float> division gfdm_complex(0.5, 0.0) / x */
        const float complex q = (0.5f + 0.0f * I) / (h->tmpK[i].re + h->tmpK[i].im * I);
        dst[i] = cf_make(crealf(q), cimagf(q));
    }
}

/* lib/preamble_channel_estimator_cc.cc:34-100 */
int gfdm_channel_estimator_create(gfdm_channel_estimator** out, int M, int K, int A, int is_dc_free,
                                  int which, const cf* preamble, int n_preamble)
{
    if (M < 1 || K < 2 || A < 2 || A > K)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "timeslots MUST be positive and 2 <= active_subcarriers <= fft_len");
    if (n_preamble < 2 * K) return fail(GFDM_ERR_INVALID_ARGUMENT, "preamble MUST hold at least 2 * fft_len samples");
    /* odd A: interpolate_frame (:238-274) writes up to index N + M/2 - 1 of an N-long frame when dc free, and
       filter_preamble_estimate / estimate_snr read unfilled entries; the ABI rejects it on every backend */
    if (A % 2) return fail(GFDM_ERR_INVALID_ARGUMENT, "active_subcarriers MUST be even (the reference writes past its frame buffer for odd values)");
    if (A + (is_dc_free ? 1 : 0) > K)
        return fail(GFDM_ERR_INVALID_ARGUMENT, "active_subcarriers (+1 if dc free) MUST NOT exceed fft_len");
    gfdm_channel_estimator* h = (gfdm_channel_estimator*)calloc(1, sizeof(*h));
    h->M = M; h->K = K; h->A = A; h->dc_free = is_dc_free ? 1 : 0; h->which = which;
    dft_init(&h->fftK, K, 1);
    dft_init(&h->fft2K, 2 * K, 1);
    h->inv0 = (cf*)malloc(sizeof(cf) * (size_t)K);
    h->inv1 = (cf*)malloc(sizeof(cf) * (size_t)K);
    h->tmpK = (cf*)malloc(sizeof(cf) * (size_t)K);
    h->tmpK2 = (cf*)malloc(sizeof(cf) * (size_t)K);
    h->inter = (cf*)malloc(sizeof(cf) * (size_t)(A + 9 + 1));
    h->pre_est = (cf*)malloc(sizeof(cf) * (size_t)K);
    h->filt = (cf*)malloc(sizeof(cf) * (size_t)(A + 1));
    h->snr_out = (cf*)malloc(sizeof(cf) * (size_t)(2 * K));
    est_init_inv(h, h->inv0, preamble);
    est_init_inv(h, h->inv1, preamble + K);
    /* initialize_gaussian_filter(sigma_sq = 1, 9 taps) :86-100 */
    float s = 0.0f;
    for (int i = 0; i < 9; ++i) {
        const float val = powf((float)(i - 9 / 2), 2.0f) / 1.0f;
        h->g[i] = expf(-0.5f * val);
        s += h->g[i];
    }
    for (int i = 0; i < 9; ++i) h->g[i] = h->g[i] / s;
    *out = h;
    return GFDM_OK;
}
void gfdm_channel_estimator_destroy(gfdm_channel_estimator* h)
{
    if (!h) return;
    dft_free(&h->fftK); dft_free(&h->fft2K);
    free(h->inv0); free(h->inv1); free(h->tmpK); free(h->tmpK2); free(h->inter); free(h->pre_est);
    free(h->filt); free(h->snr_out);
    free(h);
}
int gfdm_channel_estimator_fft_len(const gfdm_channel_estimator* h) { return h->K; }
int gfdm_channel_estimator_timeslots(const gfdm_channel_estimator* h) { return h->M; }
int gfdm_channel_estimator_frame_len(const gfdm_channel_estimator* h) { return h->M * h->K; }
int gfdm_channel_estimator_active_subcarriers(const gfdm_channel_estimator* h) { return h->A; }
int gfdm_channel_estimator_is_dc_free(const gfdm_channel_estimator* h) { return h->dc_free; }
int gfdm_channel_estimator_preamble_filter_taps(const gfdm_channel_estimator* h, float* o)
{
    memcpy(o, h->g, sizeof(h->g));
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:121-143 */
int gfdm_channel_estimator_estimate_preamble_channel(gfdm_channel_estimator* h, cf* o, const cf* rx)
{
    const int K = h->K;
    dft_exec(&h->fftK, h->tmpK, rx);
    for (int i = 0; i < K; ++i) h->tmpK2[i] = cf_mul(h->tmpK[i], h->inv0[i]);
    dft_exec(&h->fftK, h->tmpK, rx + K);
    for (int i = 0; i < K; ++i) o[i] = cf_add(cf_mul(h->tmpK[i], h->inv1[i]), h->tmpK2[i]);
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:145-185 */
int gfdm_channel_estimator_filter_preamble_estimate(gfdm_channel_estimator* h, cf* filtered, const cf* est)
{
    const int K = h->K, A = h->A, G = 9;
    cf* fi = h->inter;
    for (int i = 0; i < G / 2; ++i) fi[i] = est[K - A / 2];
    for (int i = 0; i < A / 2; ++i) fi[i + G / 2] = est[i + K - A / 2];
    int offset = 0;
    if (h->dc_free) {
        /* (estimate[fft_len - 1] + estimate[1]) / gfdm_complex(2.0f, 0.0f) */
        const float complex q = ((est[K - 1].re + est[1].re) + (est[K - 1].im + est[1].im) * I) / (2.0f + 0.0f * I);
        fi[G / 2 + A / 2] = cf_make(crealf(q), cimagf(q));
        offset = 1;
    }
    for (int i = 0; i < A / 2; ++i) fi[i + offset + G / 2 + A / 2] = est[offset + i];
    for (int i = A / 2; i < A / 2 + G / 2; ++i) fi[i + offset + G / 2 + A / 2] = est[offset + A / 2 - 1];
    const int n_taps = A + offset;
    for (int i = 0; i < n_taps; ++i) { /* volk_32fc_32f_dot_prod_32fc :179-184 */
        float re = 0.0f, im = 0.0f;
        for (int t = 0; t < G; ++t) {
            re += fi[i + t].re * h->g[t];
            im += fi[i + t].im * h->g[t];
        }
        filtered[i] = cf_make(re, im);
    }
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:238-274 */
int gfdm_channel_estimator_interpolate_frame(gfdm_channel_estimator* h, cf* fe, const cf* est)
{
    const int M = h->M, K = h->K, A = h->A;
    const int n_est = A + (h->dc_free ? 1 : 0);
    const int center = K * M / 2;
    const int dead = K - A;
    const cf step = cf_make(1.0f / (float)M, 0.0f);
    for (int i = center; i < center + M * dead / 2; ++i) fe[i] = est[0];
    for (int i = M * A / 2; i < center; ++i) fe[i] = est[n_est - 1];
    for (int i = 0; i < n_est / 2; ++i) {
        const cf inc = cf_mul(cf_sub(est[i + 1], est[i]), step);
        cf factor = est[i];
        for (int j = 0; j < M; ++j) {
            fe[center + M * dead / 2 + i * M + j] = factor;
            factor = cf_add(factor, inc);
        }
    }
    for (int i = n_est / 2; i < n_est - 1; ++i) {
        const int off = (i - n_est / 2) * M;
        const cf inc = cf_mul(cf_sub(est[i + 1], est[i]), step);
        cf factor = est[i];
        for (int j = 0; j < M; ++j) {
            fe[off + j] = factor;
            factor = cf_add(factor, inc);
        }
    }
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:276-282: conj(1 / frame_estimate) */
int gfdm_channel_estimator_prepare_for_zf(gfdm_channel_estimator* h, cf* o, const cf* fe)
{
    for (int i = 0; i < h->M * h->K; ++i) {
        const cf b = fe[i];
        const float den = b.re * b.re + b.im * b.im;
        o[i] = cf_make(b.re / den, b.im / den); /* conj((1 * conj(b)) / |b|^2) */
    }
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:285-294 */
int gfdm_channel_estimator_estimate_frame(gfdm_channel_estimator* h, cf* fe, const cf* rx)
{
    gfdm_channel_estimator_estimate_preamble_channel(h, h->pre_est, rx);
    gfdm_channel_estimator_filter_preamble_estimate(h, h->filt, h->pre_est);
    return gfdm_channel_estimator_interpolate_frame(h, fe, h->filt);
}
int gfdm_channel_estimator_estimate_frame_batch(gfdm_channel_estimator* h, cf* fe, const cf* rx, int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i)
        gfdm_channel_estimator_estimate_frame(h, fe + (size_t)i * h->M * h->K, rx + (size_t)i * 2 * h->K);
    return GFDM_OK;
}
/* lib/preamble_channel_estimator_cc.cc:187-235 */
int gfdm_channel_estimator_estimate_snr(gfdm_channel_estimator* h, float* snr_lin, float* cnrs, const cf* rx)
{
    const int K = h->K, A = h->A;
    float* c = cnrs ? cnrs : (float*)malloc(sizeof(float) * (size_t)A);
    dft_exec(&h->fft2K, h->snr_out, rx);
    float se_sum = 0.0f, ne_sum = 0.0f;
    const unsigned active_half = (unsigned)A / 2;
    const unsigned offset = h->dc_free ? 1 : 0;
    for (unsigned i = 0; i < active_half; ++i) {
        const unsigned pos = 2 * (i + offset);
        const float se = h->snr_out[pos].re * h->snr_out[pos].re + h->snr_out[pos].im * h->snr_out[pos].im;
        const float ne = h->snr_out[pos + 1].re * h->snr_out[pos + 1].re + h->snr_out[pos + 1].im * h->snr_out[pos + 1].im;
        c[i] = se;
        se_sum += se;
        ne_sum += ne;
    }
    const unsigned unused_half = (unsigned)(K - A) / 2;
    const unsigned low_offset = unused_half + (unsigned)K / 2;
    for (unsigned i = 0; i < active_half; ++i) {
        const unsigned pos = 2 * (i + low_offset);
        const float se = h->snr_out[pos].re * h->snr_out[pos].re + h->snr_out[pos].im * h->snr_out[pos].im;
        const float ne = h->snr_out[pos + 1].re * h->snr_out[pos + 1].re + h->snr_out[pos + 1].im * h->snr_out[pos + 1].im;
        c[active_half + i] = se;
        se_sum += se;
        ne_sum += ne;
    }
    const float snr = (se_sum - ne_sum) / ne_sum;
    const float scale = snr / (se_sum / (float)A);
    for (int i = 0; i < A; ++i) c[i] = c[i] * scale;
    *snr_lin = snr;
    if (!cnrs) free(c);
    return GFDM_OK;
}
int gfdm_channel_estimator_estimate_snr_batch(gfdm_channel_estimator* h, float* snr, float* cnrs, const cf* rx,
                                              int n, int mem)
{
    REQUIRE_HOST(mem)
    for (int i = 0; i < n; ++i)
        gfdm_channel_estimator_estimate_snr(h, snr + i, cnrs ? cnrs + (size_t)i * h->A : NULL, rx + (size_t)i * 2 * h->K);
    return GFDM_OK;
}

/* ---- transmitter_kernel ------------------------------------------------- */
struct gfdm_transmitter {
    gfdm_resource_mapper* mapper;
    gfdm_modulator* mod;
    gfdm_cyclic_prefixer* cp;
    int n_shifts;
    int* shifts;
    int preamble_size;
    cf* preambles; /* [n_shifts][preamble_size] */
    cf* mapped;
    cf* frame;
    int N;
};

/* lib/transmitter_kernel.cc:34-74 */
int gfdm_transmitter_create(gfdm_transmitter** out, int M, int K, int A, int cp, int cs, int ramp,
                            const int* smap, int n_map, int per_timeslot, int L, const cf* taps, int n_taps,
                            const cf* w, int n_w, const int* shifts, int n_shifts,
                            const cf* const* preambles, const int* preamble_sizes, int n_preambles)
{
    if (n_preambles < 1) return fail(GFDM_ERR_INVALID_ARGUMENT, "at least one preamble is required");
    gfdm_resource_mapper* mp = NULL;
    gfdm_modulator* md = NULL;
    gfdm_cyclic_prefixer* cpx = NULL;
    int st = gfdm_resource_mapper_create(&mp, M, K, A, smap, n_map, per_timeslot, 1);
    if (!st) st = gfdm_modulator_create(&md, M, K, L, taps, n_taps);
    if (!st) st = gfdm_cyclic_prefixer_create(&cpx, M * K, cp, cs, ramp, w, n_w, 0);
    if (!st && n_shifts != n_preambles)
        st = fail(GFDM_ERR_INVALID_ARGUMENT, "Number of cyclic shifts and number of preambles do not match!");
    for (int i = 0; !st && i < n_preambles; ++i)
        if (preamble_sizes[i] != preamble_sizes[0])
            st = fail(GFDM_ERR_INVALID_ARGUMENT, "All preambles must have equal size!");
    if (st) {
        gfdm_resource_mapper_destroy(mp);
        gfdm_modulator_destroy(md);
        gfdm_cyclic_prefixer_destroy(cpx);
        return st;
    }
    gfdm_transmitter* h = (gfdm_transmitter*)calloc(1, sizeof(*h));
    h->mapper = mp; h->mod = md; h->cp = cpx;
    h->n_shifts = n_shifts;
    h->N = M * K;
    h->shifts = (int*)malloc(sizeof(int) * (size_t)n_shifts);
    memcpy(h->shifts, shifts, sizeof(int) * (size_t)n_shifts);
    h->preamble_size = preamble_sizes[0];
    h->preambles = (cf*)malloc(sizeof(cf) * (size_t)n_shifts * (size_t)(h->preamble_size > 0 ? h->preamble_size : 1));
    for (int i = 0; i < n_shifts; ++i)
        memcpy(h->preambles + (size_t)i * h->preamble_size, preambles[i], sizeof(cf) * (size_t)h->preamble_size);
    h->mapped = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    h->frame = (cf*)malloc(sizeof(cf) * (size_t)h->N);
    *out = h;
    return GFDM_OK;
}
void gfdm_transmitter_destroy(gfdm_transmitter* h)
{
    if (!h) return;
    gfdm_resource_mapper_destroy(h->mapper);
    gfdm_modulator_destroy(h->mod);
    gfdm_cyclic_prefixer_destroy(h->cp);
    free(h->shifts); free(h->preambles); free(h->mapped); free(h->frame);
    free(h);
}
int gfdm_transmitter_input_vector_size(const gfdm_transmitter* h) { return (int)gfdm_resource_mapper_input_vector_size(h->mapper); }
int gfdm_transmitter_output_vector_size(const gfdm_transmitter* h) { return gfdm_cyclic_prefixer_frame_size(h->cp) + h->preamble_size; }
int gfdm_transmitter_n_cyclic_shifts(const gfdm_transmitter* h) { return h->n_shifts; }
int gfdm_transmitter_cyclic_shifts(const gfdm_transmitter* h, int* o)
{
    memcpy(o, h->shifts, sizeof(int) * (size_t)h->n_shifts);
    return GFDM_OK;
}
/* lib/transmitter_kernel.cc:78-84 */
int gfdm_transmitter_modulate(gfdm_transmitter* h, cf* out, const cf* in, int n)
{
    if (n < 0) return fail(GFDM_ERR_INVALID_ARGUMENT, "ninput_size MUST NOT be negative");
    const int st = gfdm_resource_mapper_map_to_resources(h->mapper, h->mapped, in, (size_t)n);
    if (st) return st;
    return gfdm_modulator_work(h->mod, out, h->mapped);
}
/* lib/transmitter_kernel.cc:86-98.  The reference's unordered_map lookup of an
 * unknown shift yields an empty preamble and reads out of bounds; here (and in
 * the product) an unknown shift is an invalid argument.  With duplicate shift
 * values the map keeps the FIRST preamble (emplace does not overwrite, :67-71). */
int gfdm_transmitter_add_frame(gfdm_transmitter* h, cf* out, const cf* in, int shift)
{
    int idx = -1;
    for (int i = 0; i < h->n_shifts && idx < 0; ++i)
        if (h->shifts[i] == shift) idx = i;
    if (idx < 0) return fail(GFDM_ERR_INVALID_ARGUMENT, "cyclic_shift has no preamble");
    memcpy(out, h->preambles + (size_t)idx * h->preamble_size, sizeof(cf) * (size_t)h->preamble_size);
    return gfdm_cyclic_prefixer_add_cyclic_prefix(h->cp, out + h->preamble_size, in, shift);
}
/* lib/transmitter_kernel.cc:101-107 */
int gfdm_transmitter_work(gfdm_transmitter* h, cf* out, const cf* in, int n)
{
    const int st = gfdm_transmitter_modulate(h, h->frame, in, n);
    if (st) return st;
    return gfdm_transmitter_add_frame(h, out, h->frame, h->shifts[0]);
}
int gfdm_transmitter_work_batch(gfdm_transmitter* h, cf* out, const cf* in, int nin, int n, int mem)
{
    REQUIRE_HOST(mem)
    const size_t os = (size_t)gfdm_transmitter_output_vector_size(h);
    for (int i = 0; i < n; ++i) {
        const int st = gfdm_transmitter_work(h, out + i * os, in + (size_t)i * nin, nin);
        if (st) return st;
    }
    return GFDM_OK;
}
int gfdm_transmitter_set_chain_fusion(gfdm_transmitter* h, int on) { (void)h; (void)on; return GFDM_OK; }
/* the per-frame loop of lib/transmitter_cc_impl.cc:165-177 */
int gfdm_transmitter_work_all_batch(gfdm_transmitter* h, cf* out, const cf* in, int nin, int n, int mem)
{
    REQUIRE_HOST(mem)
    const size_t os = (size_t)gfdm_transmitter_output_vector_size(h);
    for (int i = 0; i < n; ++i) {
        int st = gfdm_transmitter_modulate(h, h->frame, in + (size_t)i * nin, nin);
        for (int a = 0; !st && a < h->n_shifts; ++a)
            st = gfdm_transmitter_add_frame(h, out + ((size_t)a * n + i) * os, h->frame, h->shifts[a]);
        if (st) return st;
    }
    return GFDM_OK;
}
