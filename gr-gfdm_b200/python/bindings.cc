// bindings.cc -- pybind11 module `gfdm_python` retargeted onto the B200 engine.
//
// Same module name, class names, method names, argument order, return shapes and
// error text as gr-gfdm's python/bindings (python_bindings.cc:55-70 registers the
// five kernel classes):
//   Modulator                  python/bindings/modulator_python.cc:34-59
//   Demodulator                python/bindings/demodulator_python.cc:35-205
//   Cyclic_prefixer            python/bindings/cyclic_prefix_python.cc:34-94
//   Resource_mapper            python/bindings/resource_mapper_python.cc:34-85
//   Preamble_channel_estimator python/bindings/preamble_channel_estimator_python.cc:34-99
// so the reference's python/qa_python_bindings.py runs against it unchanged apart from
// the import line.  New here: Advanced_receiver, Transmitter, `*_batch` methods taking
// 2-D arrays [n_frames][size] (GIL released while the GPU works).
//
// The classes bound are the C++ host layer of include/gfdm_b200.hpp, which sits on the
// C ABI of libgfdm_b200.so; there is no other backend.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <gfdm_b200.hpp>

namespace py = pybind11;
using namespace gr::gfdm;
typedef std::complex<float> cf;
typedef py::array_t<cf, py::array::c_style | py::array::forcecast> carray;

namespace {

const cf* in_1d(const py::buffer_info& b, py::ssize_t want, const char* what)
{
    if (b.ndim != 1) throw std::runtime_error("Only ONE-dimensional vectors allowed!");
    if (b.size != want)
        throw std::runtime_error("Input vector size(" + std::to_string(b.size) + ") MUST be equal to " + what + "(" +
                                 std::to_string(want) + ")!");
    return static_cast<const cf*>(b.ptr);
}

// [n_frames][want] (a 1-D array of exactly `want` elements counts as one frame)
py::ssize_t frames_2d(const py::buffer_info& b, py::ssize_t want, const char* what)
{
    if (b.ndim == 1 && b.size == want) return 1;
    if (b.ndim != 2) throw std::runtime_error("Only TWO-dimensional arrays [n_frames][size] allowed!");
    if (b.shape[1] != want)
        throw std::runtime_error("Frame size(" + std::to_string(b.shape[1]) + ") MUST be equal to " + what + "(" +
                                 std::to_string(want) + ")!");
    return b.shape[0];
}

py::array_t<cf> out_2d(py::ssize_t n, py::ssize_t size)
{
    return py::array_t<cf>(std::vector<py::ssize_t>{ n, size });
}

template <class F>
void nogil(F&& f)
{
    py::gil_scoped_release release;
    f();
}

} // namespace

PYBIND11_MODULE(gfdm_python, m)
{
    m.doc() = "gr-gfdm kernel bindings on the B200 CUDA engine (libgfdm_b200)";
    m.def("backend", []() { return std::string(gfdm_backend()); });
    m.def("device_count", []() { return gfdm_device_count(); });
    m.def("set_device", [](int d) { detail::check(gfdm_set_device(d)); });

    // ------------------------------------------------------------------ Modulator
    py::class_<modulator_kernel_cc>(m, "Modulator")
        .def(py::init<int, int, int, std::vector<cf>>())
        .def("block_size", &modulator_kernel_cc::block_size)
        .def("filter_taps", &modulator_kernel_cc::filter_taps)
        .def("modulate",
             [](modulator_kernel_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, self.block_size(), "Modulator.block_size");
                 auto result = py::array_t<cf>(inb.size);
                 self.generic_work(static_cast<cf*>(result.request().ptr), in);
                 return result;
             })
        .def("modulate_batch", [](modulator_kernel_cc& self, const carray array) {
            py::buffer_info inb = array.request();
            const py::ssize_t n = frames_2d(inb, self.block_size(), "Modulator.block_size");
            auto result = out_2d(n, self.block_size());
            cf* out = static_cast<cf*>(result.request().ptr);
            nogil([&] { self.generic_work_batch(out, static_cast<const cf*>(inb.ptr), (int)n); });
            return result;
        });

    // ------------------------------------------------------------------ Demodulator
    auto unary = [](void (receiver_kernel_cc::*fn)(cf*, const cf*)) {
        return [fn](receiver_kernel_cc& self, const carray array) {
            py::buffer_info inb = array.request();
            const cf* in = in_1d(inb, self.block_size(), "Modulator.block_size");
            auto result = py::array_t<cf>(inb.size);
            (self.*fn)(static_cast<cf*>(result.request().ptr), in);
            return result;
        };
    };
    auto binary = [](void (receiver_kernel_cc::*fn)(cf*, const cf*, const cf*)) {
        return [fn](receiver_kernel_cc& self, const carray array, const carray eq_arr) {
            py::buffer_info inb = array.request();
            py::buffer_info eqb = eq_arr.request();
            const cf* in = in_1d(inb, self.block_size(), "Modulator.block_size");
            const cf* eq = in_1d(eqb, self.block_size(), "Modulator.block_size");
            auto result = py::array_t<cf>(inb.size);
            (self.*fn)(static_cast<cf*>(result.request().ptr), in, eq);
            return result;
        };
    };
    py::class_<receiver_kernel_cc>(m, "Demodulator")
        .def(py::init<int, int, int, std::vector<cf>>())
        .def("timeslots", &receiver_kernel_cc::timeslots)
        .def("subcarriers", &receiver_kernel_cc::subcarriers)
        .def("overlap", &receiver_kernel_cc::overlap)
        .def("block_size", &receiver_kernel_cc::block_size)
        .def("filter_taps", &receiver_kernel_cc::filter_taps)
        .def("ic_filter_taps", &receiver_kernel_cc::ic_filter_taps)
        .def("demodulate", unary(&receiver_kernel_cc::generic_work))
        .def("fft_filter_downsample", unary(&receiver_kernel_cc::fft_filter_downsample))
        .def("transform_subcarriers_to_td", unary(&receiver_kernel_cc::transform_subcarriers_to_td))
        .def("demodulate_equalize", binary(&receiver_kernel_cc::generic_work_equalize))
        .def("fft_equalize_filter_downsample", binary(&receiver_kernel_cc::fft_equalize_filter_downsample))
        .def("cancel_sc_interference", binary(&receiver_kernel_cc::cancel_sc_interference))
        .def("demodulate_batch",
             [](receiver_kernel_cc& self, const carray array, py::object eq_arr) {
                 py::buffer_info inb = array.request();
                 const py::ssize_t n = frames_2d(inb, self.block_size(), "Demodulator.block_size");
                 const cf* eq = nullptr;
                 carray eq_keep;
                 if (!eq_arr.is_none()) {
                     eq_keep = eq_arr.cast<carray>();
                     py::buffer_info eqb = eq_keep.request();
                     if (frames_2d(eqb, self.block_size(), "Demodulator.block_size") != n)
                         throw std::runtime_error("Equalizer array MUST hold one row per frame!");
                     eq = static_cast<const cf*>(eqb.ptr);
                 }
                 auto result = out_2d(n, self.block_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] { self.generic_work_batch(out, static_cast<const cf*>(inb.ptr), eq, (int)n); });
                 return result;
             },
             py::arg("array"), py::arg("eq_arr") = py::none());

    // ------------------------------------------------------------------ Advanced_receiver (new)
    py::class_<advanced_receiver_kernel_cc>(m, "Advanced_receiver")
        .def(py::init([](int timeslots, int subcarriers, int overlap, std::vector<cf> taps, std::vector<int> smap,
                         int ic_iter, std::vector<cf> points, int decision_rule, int do_phase_compensation) {
                 constellation c{ std::move(points), decision_rule };
                 if (c.points.empty()) c = constellation::qpsk();
                 return new advanced_receiver_kernel_cc(timeslots, subcarriers, overlap, std::move(taps),
                                                        std::move(smap), ic_iter, c, do_phase_compensation);
             }),
             py::arg("timeslots"), py::arg("subcarriers"), py::arg("overlap"), py::arg("frequency_taps"),
             py::arg("subcarrier_map"), py::arg("ic_iter"), py::arg("constellation_points") = std::vector<cf>(),
             py::arg("decision_rule") = (int)GFDM_DECISION_QPSK_SIGN, py::arg("do_phase_compensation") = 0)
        .def("block_size", &advanced_receiver_kernel_cc::block_size)
        .def("set_ic", &advanced_receiver_kernel_cc::set_ic)
        .def("get_ic", &advanced_receiver_kernel_cc::get_ic)
        .def("set_phase_compensation", &advanced_receiver_kernel_cc::set_phase_compensation)
        .def("get_phase_compensation", &advanced_receiver_kernel_cc::get_phase_compensation)
        .def("demodulate",
             [](advanced_receiver_kernel_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, self.block_size(), "Advanced_receiver.block_size");
                 auto result = py::array_t<cf>(inb.size);
                 self.generic_work(static_cast<cf*>(result.request().ptr), in);
                 return result;
             })
        .def("demodulate_equalize",
             [](advanced_receiver_kernel_cc& self, const carray array, const carray eq_arr) {
                 py::buffer_info inb = array.request();
                 py::buffer_info eqb = eq_arr.request();
                 const cf* in = in_1d(inb, self.block_size(), "Advanced_receiver.block_size");
                 const cf* eq = in_1d(eqb, self.block_size(), "Advanced_receiver.block_size");
                 auto result = py::array_t<cf>(inb.size);
                 self.generic_work_equalize(static_cast<cf*>(result.request().ptr), in, eq);
                 return result;
             })
        .def("demodulate_batch",
             [](advanced_receiver_kernel_cc& self, const carray array, py::object eq_arr) {
                 py::buffer_info inb = array.request();
                 const py::ssize_t n = frames_2d(inb, self.block_size(), "Advanced_receiver.block_size");
                 const cf* eq = nullptr;
                 carray eq_keep;
                 if (!eq_arr.is_none()) {
                     eq_keep = eq_arr.cast<carray>();
                     py::buffer_info eqb = eq_keep.request();
                     if (frames_2d(eqb, self.block_size(), "Advanced_receiver.block_size") != n)
                         throw std::runtime_error("Equalizer array MUST hold one row per frame!");
                     eq = static_cast<const cf*>(eqb.ptr);
                 }
                 auto result = out_2d(n, self.block_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] { self.generic_work_batch(out, static_cast<const cf*>(inb.ptr), eq, (int)n); });
                 return result;
             },
             py::arg("array"), py::arg("eq_arr") = py::none());

    // ------------------------------------------------------------------ Cyclic_prefixer
    py::class_<add_cyclic_prefix_cc>(m, "Cyclic_prefixer")
        .def(py::init<int, int, int, int, std::vector<cf>, int>(), py::arg("block_len"), py::arg("cp_len"),
             py::arg("cs_len"), py::arg("ramp_len"), py::arg("window_taps"), py::arg("cyclic_shift") = 0)
        .def("block_size", &add_cyclic_prefix_cc::block_size)
        .def("frame_size", &add_cyclic_prefix_cc::frame_size)
        .def("cyclic_shift", &add_cyclic_prefix_cc::cyclic_shift)
        .def("add_cyclic_prefix",
             [](add_cyclic_prefix_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, self.block_size(), "Cyclic_prefixer.block_size");
                 auto result = py::array_t<cf>(self.frame_size());
                 self.generic_work(static_cast<cf*>(result.request().ptr), in);
                 return result;
             })
        .def("remove_cyclic_prefix",
             [](add_cyclic_prefix_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, self.frame_size(), "Cyclic_prefixer.frame_size");
                 auto result = py::array_t<cf>(self.block_size());
                 self.remove_cyclic_prefix(static_cast<cf*>(result.request().ptr), in);
                 return result;
             })
        .def("add_cyclic_prefix_batch",
             [](add_cyclic_prefix_cc& self, const carray array, py::object shift) {
                 py::buffer_info inb = array.request();
                 const py::ssize_t n = frames_2d(inb, self.block_size(), "Cyclic_prefixer.block_size");
                 const int s = shift.is_none() ? self.cyclic_shift() : shift.cast<int>();
                 auto result = out_2d(n, self.frame_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] { self.add_cyclic_prefix_batch(out, static_cast<const cf*>(inb.ptr), s, (int)n); });
                 return result;
             },
             py::arg("array"), py::arg("cyclic_shift") = py::none())
        .def("remove_cyclic_prefix_batch", [](add_cyclic_prefix_cc& self, const carray array) {
            py::buffer_info inb = array.request();
            const py::ssize_t n = frames_2d(inb, self.frame_size(), "Cyclic_prefixer.frame_size");
            auto result = out_2d(n, self.block_size());
            cf* out = static_cast<cf*>(result.request().ptr);
            nogil([&] { self.remove_cyclic_prefix_batch(out, static_cast<const cf*>(inb.ptr), (int)n); });
            return result;
        });

    // ------------------------------------------------------------------ Resource_mapper
    py::class_<resource_mapper_kernel_cc>(m, "Resource_mapper")
        .def(py::init<int, int, int, std::vector<int>, bool>(), py::arg("timeslots"), py::arg("subcarriers"),
             py::arg("active_subcarriers"), py::arg("subcarrier_map"), py::arg("per_timeslot") = true)
        .def("block_size", &resource_mapper_kernel_cc::block_size)
        .def("frame_size", &resource_mapper_kernel_cc::frame_size)
        .def("map_to_resources",
             [](resource_mapper_kernel_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, (py::ssize_t)self.block_size(), "Resource_mapper.block_size");
                 auto result = py::array_t<cf>((py::ssize_t)self.frame_size());
                 self.map_to_resources(static_cast<cf*>(result.request().ptr), in, (size_t)inb.size);
                 return result;
             })
        .def("demap_from_resources",
             [](resource_mapper_kernel_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, (py::ssize_t)self.frame_size(), "Resource_mapper.frame_size");
                 auto result = py::array_t<cf>((py::ssize_t)self.block_size());
                 self.demap_from_resources(static_cast<cf*>(result.request().ptr), in, self.block_size());
                 return result;
             })
        .def("map_to_resources_batch",
             [](resource_mapper_kernel_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const py::ssize_t n = frames_2d(inb, (py::ssize_t)self.block_size(), "Resource_mapper.block_size");
                 auto result = out_2d(n, (py::ssize_t)self.frame_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] {
                     self.map_to_resources_batch(out, static_cast<const cf*>(inb.ptr), self.block_size(), (int)n);
                 });
                 return result;
             })
        .def("demap_from_resources_batch", [](resource_mapper_kernel_cc& self, const carray array) {
            py::buffer_info inb = array.request();
            const py::ssize_t n = frames_2d(inb, (py::ssize_t)self.frame_size(), "Resource_mapper.frame_size");
            auto result = out_2d(n, (py::ssize_t)self.block_size());
            cf* out = static_cast<cf*>(result.request().ptr);
            nogil([&] {
                self.demap_from_resources_batch(out, static_cast<const cf*>(inb.ptr), self.block_size(), (int)n);
            });
            return result;
        });

    // ------------------------------------------------------------------ Preamble_channel_estimator
    py::class_<preamble_channel_estimator_cc>(m, "Preamble_channel_estimator")
        .def(py::init<int, int, int, bool, int, std::vector<cf>>(), py::arg("timeslots"), py::arg("subcarriers"),
             py::arg("active_subcarriers"), py::arg("is_dc_free"), py::arg("which_estimator"), py::arg("preamble"))
        .def("timeslots", &preamble_channel_estimator_cc::timeslots)
        .def("subcarriers", &preamble_channel_estimator_cc::fft_len)
        .def("active_subcarriers", &preamble_channel_estimator_cc::active_subcarriers)
        .def("frame_len", &preamble_channel_estimator_cc::frame_len)
        .def("is_dc_free", &preamble_channel_estimator_cc::is_dc_free)
        .def("preamble_filter_taps", &preamble_channel_estimator_cc::preamble_filter_taps)
        .def("estimate_frame",
             [](preamble_channel_estimator_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, 2 * self.fft_len(), "2 * subcarriers");
                 auto result = py::array_t<cf>(self.frame_len());
                 py::buffer_info resb = result.request();
                 std::fill_n(static_cast<cf*>(resb.ptr), resb.size, cf(0.f, 0.f)); // bins the kernel never writes
                 self.estimate_frame(static_cast<cf*>(resb.ptr), in);
                 return result;
             })
        .def("estimate_snr",
             [](preamble_channel_estimator_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, 2 * self.fft_len(), "2 * subcarriers");
                 std::vector<float> cnrs;
                 return self.estimate_snr(cnrs, in);
             })
        .def("estimate_snr_cnrs",
             [](preamble_channel_estimator_cc& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const cf* in = in_1d(inb, 2 * self.fft_len(), "2 * subcarriers");
                 std::vector<float> cnrs;
                 const float snr = self.estimate_snr(cnrs, in);
                 return py::make_tuple(snr, py::array_t<float>((py::ssize_t)cnrs.size(), cnrs.data()));
             })
        .def("estimate_frame_batch", [](preamble_channel_estimator_cc& self, const carray array) {
            py::buffer_info inb = array.request();
            const py::ssize_t n = frames_2d(inb, 2 * self.fft_len(), "2 * subcarriers");
            auto result = out_2d(n, self.frame_len());
            py::buffer_info resb = result.request();
            std::fill_n(static_cast<cf*>(resb.ptr), resb.size, cf(0.f, 0.f));
            cf* out = static_cast<cf*>(resb.ptr);
            nogil([&] { self.estimate_frame_batch(out, static_cast<const cf*>(inb.ptr), (int)n); });
            return result;
        });

    // ------------------------------------------------------------------ Transmitter (new)
    py::class_<transmitter_kernel>(m, "Transmitter")
        .def(py::init<int, int, int, int, int, int, std::vector<int>, bool, int, std::vector<cf>, std::vector<cf>,
                      std::vector<int>, std::vector<std::vector<cf>>>(),
             py::arg("timeslots"), py::arg("subcarriers"), py::arg("active_subcarriers"), py::arg("cp_len"),
             py::arg("cs_len"), py::arg("ramp_len"), py::arg("subcarrier_map"), py::arg("per_timeslot"),
             py::arg("overlap"), py::arg("frequency_taps"), py::arg("window_taps"), py::arg("cyclic_shifts"),
             py::arg("preambles"))
        .def("input_vector_size", &transmitter_kernel::input_vector_size)
        .def("output_vector_size", &transmitter_kernel::output_vector_size)
        .def("cyclic_shifts", &transmitter_kernel::cyclic_shifts)
        .def("generic_work",
             [](transmitter_kernel& self, const carray array) {
                 py::buffer_info inb = array.request();
                 if (inb.ndim != 1) throw std::runtime_error("Only ONE-dimensional vectors allowed!");
                 auto result = py::array_t<cf>(self.output_vector_size());
                 self.generic_work(static_cast<cf*>(result.request().ptr), static_cast<const cf*>(inb.ptr),
                                   (int)inb.size);
                 return result;
             })
        .def("generic_work_batch",
             [](transmitter_kernel& self, const carray array) {
                 py::buffer_info inb = array.request();
                 const py::ssize_t n = frames_2d(inb, self.input_vector_size(), "Transmitter.input_vector_size");
                 auto result = out_2d(n, self.output_vector_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] {
                     self.generic_work_batch(out, static_cast<const cf*>(inb.ptr), self.input_vector_size(), (int)n);
                 });
                 return result;
             })
        .def("generic_work_all_batch", [](transmitter_kernel& self, const carray array) {
            py::buffer_info inb = array.request();
            const py::ssize_t n = frames_2d(inb, self.input_vector_size(), "Transmitter.input_vector_size");
            const py::ssize_t na = (py::ssize_t)self.cyclic_shifts().size();
            auto result = py::array_t<cf>(std::vector<py::ssize_t>{ na, n, (py::ssize_t)self.output_vector_size() });
            cf* out = static_cast<cf*>(result.request().ptr);
            nogil([&] {
                self.generic_work_all_batch(out, static_cast<const cf*>(inb.ptr), self.input_vector_size(), (int)n);
            });
            return result;
        });

    // ------------------------------------------------------------------ rows either side of the path
    typedef py::array_t<unsigned char, py::array::c_style | py::array::forcecast> barray;
    py::class_<symbol_mapper>(m, "Symbol_mapper")
        .def(py::init([](std::vector<cf> points, int decision_rule) {
                 constellation c{ std::move(points), decision_rule };
                 if (c.points.empty()) c = constellation::qpsk();
                 return new symbol_mapper(c);
             }),
             py::arg("constellation_points") = std::vector<cf>(), py::arg("decision_rule") = (int)GFDM_DECISION_NEAREST)
        .def("n_points", &symbol_mapper::n_points)
        .def("bits_per_symbol", &symbol_mapper::bits_per_symbol)
        .def("map_chunks",
             [](symbol_mapper& self, const barray chunks) {
                 py::buffer_info b = chunks.request();
                 auto result = py::array_t<cf>(b.size);
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] { self.map_chunks(out, static_cast<const unsigned char*>(b.ptr), (size_t)b.size); });
                 return result;
             })
        .def("decide",
             [](symbol_mapper& self, const carray symbols) {
                 py::buffer_info b = symbols.request();
                 auto result = py::array_t<unsigned char>(b.size);
                 unsigned char* out = static_cast<unsigned char*>(result.request().ptr);
                 nogil([&] { self.decide(out, static_cast<const cf*>(b.ptr), (size_t)b.size); });
                 return result;
             })
        .def("bits2symbols",
             [](symbol_mapper& self, const barray bits) {
                 py::buffer_info b = bits.request();
                 const int bps = self.bits_per_symbol() > 0 ? self.bits_per_symbol() : 1;
                 if (b.size % bps) throw std::runtime_error("number of bits MUST be a multiple of bits_per_symbol!");
                 auto result = py::array_t<cf>(b.size / bps);
                 self.bits2symbols(static_cast<cf*>(result.request().ptr), static_cast<const unsigned char*>(b.ptr),
                                   (size_t)(b.size / bps));
                 return result;
             })
        .def("symbols2bits",
             [](symbol_mapper& self, const carray symbols) {
                 py::buffer_info b = symbols.request();
                 const int bps = self.bits_per_symbol() > 0 ? self.bits_per_symbol() : 1;
                 auto result = py::array_t<unsigned char>(b.size * bps);
                 self.symbols2bits(static_cast<unsigned char*>(result.request().ptr), static_cast<const cf*>(b.ptr),
                                   (size_t)b.size);
                 return result;
             })
        .def("modulate_chunks_batch",
             [](symbol_mapper& self, modulator_kernel_cc& mod, const barray chunks) {
                 py::buffer_info b = chunks.request();
                 if (b.ndim != 2 || b.shape[1] != mod.block_size())
                     throw std::runtime_error("Only TWO-dimensional arrays [n_frames][block_size] allowed!");
                 auto result = out_2d(b.shape[0], mod.block_size());
                 cf* out = static_cast<cf*>(result.request().ptr);
                 nogil([&] { self.modulate_chunks(mod, out, static_cast<const unsigned char*>(b.ptr), (int)b.shape[0]); });
                 return result;
             })
        .def("demodulate_decide_batch", [](symbol_mapper& self, receiver_kernel_cc& rx, const carray array) {
            py::buffer_info b = array.request();
            const py::ssize_t n = frames_2d(b, rx.block_size(), "Demodulator.block_size");
            auto result = py::array_t<unsigned char>(std::vector<py::ssize_t>{ n, (py::ssize_t)rx.block_size() });
            unsigned char* out = static_cast<unsigned char*>(result.request().ptr);
            nogil([&] { self.demodulate_decide(rx, out, static_cast<const cf*>(b.ptr), nullptr, (int)n); });
            return result;
        });

    py::class_<remove_prefix>(m, "Remove_prefix")
        .def(py::init<int, int, int>(), py::arg("frame_len"), py::arg("block_len"), py::arg("offset"))
        .def("work_batch", [](remove_prefix& self, const carray array) {
            py::buffer_info b = array.request();
            const py::ssize_t n = frames_2d(b, self.frame_len(), "Remove_prefix.frame_len");
            auto result = out_2d(n, self.block_len());
            cf* out = static_cast<cf*>(result.request().ptr);
            nogil([&] { self.work_batch(out, static_cast<const cf*>(b.ptr), (int)n); });
            return result;
        });

    py::class_<extract_burst>(m, "Extract_burst")
        .def(py::init<int, int, bool>(), py::arg("burst_len"), py::arg("tag_backoff"),
             py::arg("activate_cfo_correction") = false)
        .def("activate_cfo_compensation", &extract_burst::activate_cfo_compensation)
        .def("work",
             [](extract_burst& self, const carray stream, std::vector<long long> burst_starts, std::vector<float> scale_factors,
                std::vector<cf> phase_rotations) {
                 py::buffer_info b = stream.request();
                 const py::ssize_t mb = (py::ssize_t)burst_starts.size();
                 auto bursts = out_2d(mb > 0 ? mb : 1, self.burst_len());
                 extract_burst::result r = self.work(static_cast<cf*>(bursts.request().ptr), (int)mb,
                                                     static_cast<const cf*>(b.ptr), (long long)b.size, burst_starts,
                                                     scale_factors, phase_rotations);
                 return py::make_tuple(bursts, r.n_produced, r.n_consumed);
             },
             py::arg("stream"), py::arg("burst_starts"), py::arg("scale_factors") = std::vector<float>(),
             py::arg("phase_rotations") = std::vector<cf>());
}
