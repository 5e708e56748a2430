// Minimal stand-in for GNU Radio's <pmt/pmt.h> (TEST INFRASTRUCTURE: GNU Radio cannot be installed in this image).
// Only what include/gfdm_b200_blocks.hpp touches: symbols, longs, floats and f32 vectors as opaque shared values.
#ifndef STUB_PMT_H
#define STUB_PMT_H
#include <memory>
#include <string>
#include <vector>
namespace pmt {
struct pmt_base {
    enum kind_t { SYMBOL, LONG, FLOAT, F32VECTOR } kind;
    std::string sym;
    long l = 0;
    double d = 0.0;
    std::vector<float> f32;
};
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t string_to_symbol(const std::string& s) { auto p = std::make_shared<pmt_base>(); p->kind = pmt_base::SYMBOL; p->sym = s; return p; }
inline pmt_t intern(const std::string& s) { return string_to_symbol(s); }
inline pmt_t mp(const std::string& s) { return string_to_symbol(s); }
inline pmt_t from_long(long v) { auto p = std::make_shared<pmt_base>(); p->kind = pmt_base::LONG; p->l = v; return p; }
inline pmt_t from_float(double v) { auto p = std::make_shared<pmt_base>(); p->kind = pmt_base::FLOAT; p->d = v; return p; }
inline pmt_t init_f32vector(size_t, const std::vector<float>& v) { auto p = std::make_shared<pmt_base>(); p->kind = pmt_base::F32VECTOR; p->f32 = v; return p; }
inline bool eqv(const pmt_t& a, const pmt_t& b) { return a && b && a->kind == b->kind && a->sym == b->sym && a->l == b->l; }
inline std::string symbol_to_string(const pmt_t& p) { return p->sym; }
inline long to_long(const pmt_t& p) { return p->l; }
inline double to_double(const pmt_t& p) { return p->d; }
inline std::vector<float> f32vector_elements(const pmt_t& p) { return p->f32; }
} // namespace pmt
#endif
