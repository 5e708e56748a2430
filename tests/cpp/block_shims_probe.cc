// block_shims_probe.cc -- drives every work() of include/gfdm_b200_blocks.hpp the way the GNU Radio scheduler would
// (tests/stub_gnuradio/ stands in for GNU Radio) and checks it frame by frame against the kernel classes' own per-frame
// calls -- the loops the reference blocks run (lib/*_impl.cc).  `--dry`: no device needed, constructor validation only.
#include <gfdm_b200_blocks.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

using namespace gr::gfdm;
typedef std::complex<float> cf;

static std::mt19937 rng(7);
static std::vector<cf> crand(size_t n)
{
    std::normal_distribution<float> d(0.f, 1.f);
    std::vector<cf> v(n);
    for (auto& x : v) x = cf(d(rng), d(rng));
    return v;
}
static bool same(const cf* a, const cf* b, size_t n) { return std::memcmp(a, b, sizeof(cf) * n) == 0; }
#define CHECK(cond, what)                         \
    do {                                          \
        if (!(cond)) {                            \
            printf("FAIL %s (line %d)\n", what, __LINE__); \
            return 1;                             \
        }                                         \
    } while (0)

struct qpsk_like { // what gr::digital::constellation_sptr offers to the shim: points()
    std::vector<cf> pts;
    const std::vector<cf>& points() const { return pts; }
};

int main(int argc, char** argv)
{
    const int M = 9, K = 64, L = 2, A = 52, cp = 16, cs = 8, ramp = 8, N = M * K;
    // RRC-like taps: any L*M taps do, the kernels are tap-agnostic
    std::vector<cf> taps(L * M);
    for (int i = 0; i < L * M; ++i) taps[i] = cf(std::cos(0.3f * i) + 1.5f, 0.1f * std::sin(0.7f * i));
    if (argc > 1 && std::string(argv[1]) == "--dry") {
        bool threw = false;
        try {
            b200::simple_modulator_cc::make(M, K, L, std::vector<cf>(3));
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        CHECK(threw, "simple_modulator_cc ctor must forward the kernel's std::invalid_argument");
        threw = false;
        try {
            b200::short_burst_shaper::make(-1, 0, cf(1, 0));
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        CHECK(threw, "short_burst_shaper ctor must reject negative padding");
        printf("OK dry\n");
        return 0;
    }
    const int F = 6;
    // ---- simple_modulator_cc / simple_receiver_cc
    auto mod = b200::simple_modulator_cc::make(M, K, L, taps);
    CHECK(mod->output_multiple() == N, "modulator output multiple");
    std::vector<cf> d = crand((size_t)F * N), x((size_t)F * N), x1(N);
    {
        gr_vector_const_void_star in{ d.data() };
        gr_vector_void_star out{ x.data() };
        CHECK(mod->work(F * N, in, out) == F * N, "modulator work return");
        modulator_kernel_cc k(M, K, L, taps);
        for (int f = 0; f < F; ++f) {
            k.generic_work(x1.data(), d.data() + (size_t)f * N);
            CHECK(same(x1.data(), x.data() + (size_t)f * N, N), "simple_modulator_cc vs generic_work per frame");
        }
    }
    std::vector<cf> rxt(taps.size());
    for (size_t i = 0; i < taps.size(); ++i) rxt[i] = std::conj(taps[i]);
    auto rx = b200::simple_receiver_cc::make(M, K, L, rxt);
    std::vector<cf> y((size_t)F * N), y1(N);
    {
        gr_vector_const_void_star in{ x.data() };
        gr_vector_void_star out{ y.data() };
        CHECK(rx->work(F * N, in, out) == F * N, "receiver work return");
        receiver_kernel_cc k(M, K, L, rxt);
        for (int f = 0; f < F; ++f) {
            k.generic_work(y1.data(), x.data() + (size_t)f * N);
            CHECK(same(y1.data(), y.data() + (size_t)f * N, N), "simple_receiver_cc vs generic_work per frame");
        }
    }
    // ---- advanced_receiver_sb_cc: two inputs, the channel input advances frame by frame; tags of input 1 are forwarded
    std::vector<int> smap;
    for (int i = 1; i <= A / 2; ++i) smap.push_back(i);
    for (int i = K - A / 2; i < K; ++i) smap.push_back(i);
    {
        auto cst = std::make_shared<qpsk_like>();
        cst->pts = constellation::qpsk().points;
        auto adv = b200::advanced_receiver_sb_cc::make(M, K, L, 2, rxt, cst, smap, 0);
        std::vector<cf> eq((size_t)F * N);
        for (size_t i = 0; i < eq.size(); ++i) eq[i] = cf(1.0f + 0.01f * (float)(i % 7), 0.02f * (float)(i % 5));
        gr::tag_t t;
        t.offset = N + 3;
        t.key = pmt::intern("frame_start");
        t.value = pmt::from_long(42);
        adv->test_input_tags(1).push_back(t);
        gr_vector_const_void_star in{ x.data(), eq.data() };
        gr_vector_void_star out{ y.data() };
        CHECK(adv->work(F * N, in, out) == F * N, "advanced receiver work return");
        CHECK(adv->test_output_tags(0).size() == 1 && adv->test_output_tags(0)[0].offset == (uint64_t)N + 3, "tag forwarding");
        advanced_receiver_kernel_cc k(M, K, L, rxt, smap, 2, constellation::qpsk(), 0);
        for (int f = 0; f < F; ++f) {
            k.generic_work_equalize(y1.data(), x.data() + (size_t)f * N, eq.data() + (size_t)f * N);
            CHECK(same(y1.data(), y.data() + (size_t)f * N, N), "advanced_receiver_sb_cc vs generic_work_equalize per frame");
        }
        gr_vector_const_void_star in1{ x.data() };
        CHECK(adv->work(2 * N, in1, out) == 2 * N, "advanced receiver, one input");
        k.generic_work(y1.data(), x.data() + N);
        CHECK(same(y1.data(), y.data() + N, N), "advanced_receiver_sb_cc vs generic_work (no channel input)");
    }
    // ---- transmitter_cc: two antennas (cyclic shifts), length tags
    {
        const int W = N + cp + cs, P = 2 * K + cp + ramp;
        std::vector<cf> window(W, cf(1, 0));
        for (int i = 0; i < ramp; ++i) {
            window[i] = cf((float)(i + 1) / (ramp + 1), 0);
            window[W - 1 - i] = window[i];
        }
        std::vector<int> shifts{ 0, 4 };
        std::vector<std::vector<cf>> pre{ crand(P), crand(P) };
        auto tx = b200::transmitter_cc::make(M, K, A, cp, cs, ramp, smap, true, L, taps, window, shifts, pre, "packet_len");
        const int is = tx->kernel().input_vector_size(), os = tx->kernel().output_vector_size();
        CHECK(os == P + W && is == A * M, "transmitter sizes");
        CHECK(tx->fixed_rate_noutput_to_ninput(3 * os) == 3 * is, "transmitter rate");
        std::vector<cf> sym = crand((size_t)F * is), o0((size_t)F * os), o1((size_t)F * os), fr(N), one(os);
        gr::tag_t t;
        t.offset = 0;
        t.key = pmt::string_to_symbol("packet_len");
        t.value = pmt::from_long(is);
        tx->test_input_tags(0).push_back(t);
        gr_vector_int nin{ F * is };
        gr_vector_const_void_star in{ sym.data() };
        gr_vector_void_star out{ o0.data(), o1.data() };
        CHECK(tx->general_work(F * os + 5, nin, in, out) == F * os, "transmitter general_work return");
        CHECK(tx->consumed() == F * is, "transmitter consume");
        CHECK(tx->test_input_tags(0).empty(), "length tag removed from the input");
        CHECK(tx->test_output_tags(0).size() == (size_t)F && tx->test_output_tags(1).size() == (size_t)F &&
                  pmt::to_long(tx->test_output_tags(1)[2].value) == os && tx->test_output_tags(1)[2].offset == (uint64_t)2 * os,
              "length tags on both outputs");
        transmitter_kernel k(M, K, A, cp, cs, ramp, smap, true, L, taps, window, shifts, pre);
        for (int f = 0; f < F; ++f) {
            k.modulate(fr.data(), sym.data() + (size_t)f * is, is);
            for (int a = 0; a < 2; ++a) {
                k.add_frame(one.data(), fr.data(), shifts[a]);
                const cf* got = (a ? o1.data() : o0.data()) + (size_t)f * os;
                double err = 0, nrm = 0; // the one-kernel chain and the staged per-frame calls agree to fp32 rounding
                for (int i = 0; i < os; ++i) { err += std::norm(got[i] - one[i]); nrm += std::norm(one[i]); }
                CHECK(err <= 1e-10 * nrm, "transmitter_cc vs modulate + add_frame per antenna");
            }
        }
    }
    // ---- channel_estimator_cc: estimate + SNR tags
    {
        std::vector<cf> core = crand(2 * K);
        for (int i = 0; i < K; ++i) core[K + i] = core[i];
        auto est = b200::channel_estimator_cc::make(M, K, A, true, 1, core);
        std::vector<cf> rxp = crand((size_t)F * 2 * K), h((size_t)F * N), h1(N);
        gr_vector_int nin{ F * 2 * K };
        gr_vector_const_void_star in{ rxp.data() };
        gr_vector_void_star out{ h.data() };
        CHECK(est->general_work(F * N, nin, in, out) == F * N && est->consumed() == F * 2 * K, "estimator general_work");
        preamble_channel_estimator_cc k(M, K, A, true, 1, core);
        for (int f = 0; f < F; ++f) {
            k.estimate_frame(h1.data(), rxp.data() + (size_t)f * 2 * K);
            CHECK(same(h1.data(), h.data() + (size_t)f * N, N), "channel_estimator_cc vs estimate_frame per frame");
        }
        CHECK(est->test_output_tags(0).size() == (size_t)2 * F, "two tags per frame");
        std::vector<float> cnrs;
        const float snr = k.estimate_snr(cnrs, rxp.data() + (size_t)2 * 2 * K);
        const gr::tag_t& ts = est->test_output_tags(0)[4];
        CHECK(pmt::symbol_to_string(ts.key) == "snr_lin" && ts.offset == (uint64_t)2 * N && std::fabs(pmt::to_double(ts.value) - snr) <= 1e-4f * std::fabs(snr), "snr tag");
        CHECK(pmt::f32vector_elements(est->test_output_tags(0)[5].value).size() == (size_t)A, "cnr tag");
    }
    // ---- resource_mapper_cc / resource_demapper_cc / cyclic_prefixer_cc
    {
        auto mp = b200::resource_mapper_cc::make(M, K, A, smap, true);
        auto dm = b200::resource_demapper_cc::make(M, K, A, smap, true);
        std::vector<cf> s = crand((size_t)F * A * M), g((size_t)F * N), s2((size_t)F * A * M);
        gr_vector_int nin{ F * A * M };
        gr_vector_const_void_star in{ s.data() };
        gr_vector_void_star out{ g.data() };
        CHECK(mp->general_work(F * N, nin, in, out) == F * N && mp->consumed() == F * A * M, "mapper general_work");
        gr_vector_int nin2{ F * N };
        gr_vector_const_void_star in2{ g.data() };
        gr_vector_void_star out2{ s2.data() };
        CHECK(dm->general_work(F * A * M, nin2, in2, out2) == F * A * M, "demapper general_work");
        CHECK(same(s.data(), s2.data(), s.size()), "map -> demap is the identity");
        const int W = N + cp + cs;
        std::vector<cf> window(2 * ramp, cf(0.5f, 0));
        auto pf = b200::cyclic_prefixer_cc::make(N, cp, cs, ramp, window, 0);
        std::vector<cf> fr((size_t)F * W), f1(W);
        gr_vector_int nin3{ F * N };
        gr_vector_void_star out3{ fr.data() };
        CHECK(pf->general_work(F * W, nin3, in2, out3) == F * W && pf->consumed() == F * N, "prefixer general_work");
        add_cyclic_prefix_cc k(N, cp, cs, ramp, window, 0);
        for (int f = 0; f < F; ++f) {
            k.generic_work(f1.data(), g.data() + (size_t)f * N);
            CHECK(same(f1.data(), fr.data() + (size_t)f * W, W), "cyclic_prefixer_cc vs generic_work per frame");
        }
    }
    // ---- short_burst_shaper
    {
        const cf sc(0.5f, -0.25f);
        auto sh = b200::short_burst_shaper::make(11, 5, sc, 2, "packet_len");
        std::vector<cf> a = crand(100), b = crand(100), oa(116), ob(116);
        gr_vector_int nin{ 100, 100 };
        CHECK(sh->calculate_output_stream_length(nin) == 116, "shaper output length");
        gr_vector_const_void_star in{ a.data(), b.data() };
        gr_vector_void_star out{ oa.data(), ob.data() };
        CHECK(sh->work(116, nin, in, out) == 116, "shaper work return");
        for (int i = 0; i < 116; ++i) {
            const cf want = (i < 11 || i >= 111) ? cf(0, 0) : cf(b[i - 11].real() * sc.real() - b[i - 11].imag() * sc.imag(),
                                                                  b[i - 11].real() * sc.imag() + b[i - 11].imag() * sc.real());
            CHECK(ob[i] == want, "short_burst_shaper padding + scale");
        }
    }
    printf("OK blocks\n");
    return 0;
}
