// api.cu -- the C ABI of include/gfdm_b200.h on top of the CUDA kernels.
//
// Host-side mirror of gr-gfdm's kernel classes: constructor validation with the
// reference's conditions and messages, tap renormalisation, precomputed tables,
// per-handle stream + scratch, host<->device staging for HOST pointers.  There is
// no CPU compute path in this file: every data-path entry launches CUDA kernels
// and fails with GFDM_ERR_CUDA when no device is usable.
#include "../../include/gfdm_b200.h"

#include "engine.h"
#include "fused.h"

#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges cost a predicted branch unless a profiler is attached

#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sstream>

using namespace gfdm;
typedef std::complex<float> cf;

static thread_local std::string g_err;
// Device selection is PER HOST THREAD (one thread per GPU may create and drive its own handles concurrently);
// a thread that never called gfdm_set_device inherits the last selection made anywhere in the process.
static thread_local int t_device = -1;
static std::atomic<int> g_default_device{ 0 };
static int current_device_choice() { return t_device >= 0 ? t_device : g_default_device.load(std::memory_order_relaxed); }

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

// one NVTX range per C-ABI entry (tracing hooks), named after the entry point
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define API_TRY                  \
    NvtxRange nvtx_range_(__func__); \
    try {
#define API_CATCH                                                                     \
    }                                                                                 \
    catch (const std::invalid_argument& e) { return fail(GFDM_ERR_INVALID_ARGUMENT, e.what()); } \
    catch (const CudaError& e) { return fail(GFDM_ERR_CUDA, e.what()); }              \
    catch (const std::exception& e) { return fail(GFDM_ERR_RUNTIME, e.what()); }      \
    return GFDM_OK;

static const uint32_t HANDLE_MAGIC = 0x47464442u; // "GFDB"

// Common prefix of every handle (gfdm_set_stream / gfdm_sync accept any of them).
struct HandleBase {
    uint32_t magic = HANDLE_MAGIC;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool opened = false; // open() ran: device resources may exist (the *_destroy functions select the device only then)
    long long launches = 0;
    const char* last_kernel = "none";
    DeviceBuf stage_in, stage_in2, stage_out; // HOST-pointer staging
    DeviceBuf work_a, work_b, work_c;         // pipeline scratch
    DeviceBuf work_fmt;                       // complex64 side of an sc16 batch
    // pipelined HOST batches (host_pipeline below): two chunk slots, copy engines on their own streams
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[2] = { nullptr, nullptr }, ev_run[2] = { nullptr, nullptr }, ev_out[2] = { nullptr, nullptr };
    DeviceBuf slot_in0[2], slot_in1[2], slot_out[2];

    void open()
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0)
            throw CudaError(std::string("gfdm_b200: no usable CUDA device (") + cudaGetErrorString(e) +
                            "); this library has no CPU fallback");
        device = current_device_choice();
        if (device >= n) throw CudaError("gfdm_b200: the selected device does not exist");
        GFDM_CUDA_CHECK(cudaSetDevice(device));
        opened = true;
        GFDM_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    void close()
    {
        stage_in.release(); stage_in2.release(); stage_out.release();
        work_a.release(); work_b.release(); work_c.release(); work_fmt.release();
        for (int i = 0; i < 2; ++i) {
            slot_in0[i].release(); slot_in1[i].release(); slot_out[i].release();
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_run[i]) cudaEventDestroy(ev_run[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
            ev_in[i] = ev_run[i] = ev_out[i] = nullptr;
        }
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        s_h2d = s_d2h = nullptr;
        if (own_stream && stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
    void use() { GFDM_CUDA_CHECK(cudaSetDevice(device)); }
    void sync()
    {
        if (s_d2h) GFDM_CUDA_CHECK(cudaStreamSynchronize(s_d2h)); // device->host copies of a GFDM_MEM_HOST_ASYNC batch
        GFDM_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
};

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static std::vector<cf> vec(const gfdm_complex* p, int n)
{
    return std::vector<cf>(reinterpret_cast<const cf*>(p), reinterpret_cast<const cf*>(p) + (n > 0 ? n : 0));
}
static cpx* dev_upload(const std::vector<cf>& v)
{
    std::vector<cpx> t(v.size());
    for (size_t i = 0; i < v.size(); ++i) t[i] = make_float2(v[i].real(), v[i].imag());
    return upload(t);
}

// taps <- taps / sqrt(sum|taps|^2 / M)   (lib/modulator_kernel_cc.cc:71-90, lib/receiver_kernel_cc.cc:99-118)
static std::vector<cf> normalize_taps(const std::vector<cf>& taps, int n_timeslots)
{
    cf res(0.f, 0.f);
    for (const cf& t : taps) res += t * std::conj(t);
    const cf sf = cf(1. / std::sqrt(std::abs(res) / n_timeslots), 0.0f);
    std::vector<cf> out(taps.size());
    for (size_t i = 0; i < taps.size(); ++i)
        out[i] = cf(taps[i].real() * sf.real() - taps[i].imag() * sf.imag(),
                    taps[i].real() * sf.imag() + taps[i].imag() * sf.real());
    return out;
}

static void check_taps(size_t n_taps, int M, int L)
{
    if ((int)n_taps != M * L) {
        std::stringstream s;
        s << "number of frequency taps(" << n_taps << ") MUST be equal to n_timeslots(" << M << ") * overlap(" << L
          << ") = " << M * L << "!";
        throw std::invalid_argument(s.str());
    }
}

// Run `body(d_out, d_in..., n)` either directly on device pointers or through staging.
// in_sizes / out_size are per-frame element counts (cpx unless noted).
struct Staging {
    HandleBase* h;
    int mem;
    bool async = false; // GFDM_MEM_HOST_ASYNC: the host pipeline returns without waiting (gfdm_sync completes it)
    explicit Staging(HandleBase* hb, int m, bool allow_async = false) : h(hb), mem(m)
    {
        if (m == GFDM_MEM_HOST_ASYNC) {
            if (!allow_async) throw std::invalid_argument("GFDM_MEM_HOST_ASYNC is accepted by the pipelined batch entries only");
            mem = GFDM_MEM_HOST;
            async = true;
        } else if (m != GFDM_MEM_HOST && m != GFDM_MEM_DEVICE) {
            throw std::invalid_argument("mem MUST be GFDM_MEM_HOST, GFDM_MEM_DEVICE or GFDM_MEM_HOST_ASYNC");
        }
    }
    const cpx* in(const gfdm_complex* p, size_t elems, DeviceBuf& buf)
    {
        if (!p) return nullptr;
        if (mem == GFDM_MEM_DEVICE) return reinterpret_cast<const cpx*>(p);
        buf.ensure(elems * sizeof(cpx));
        GFDM_CUDA_CHECK(cudaMemcpyAsync(buf.p, p, elems * sizeof(cpx), cudaMemcpyHostToDevice, h->stream));
        return buf.as<cpx>();
    }
    cpx* out(gfdm_complex* p, size_t elems, DeviceBuf& buf)
    {
        if (mem == GFDM_MEM_DEVICE) return reinterpret_cast<cpx*>(p);
        buf.ensure(elems * sizeof(cpx));
        return buf.as<cpx>();
    }
    void finish(gfdm_complex* p, size_t elems, DeviceBuf& buf)
    {
        if (mem == GFDM_MEM_DEVICE) return;
        GFDM_CUDA_CHECK(cudaMemcpyAsync(p, buf.p, elems * sizeof(cpx), cudaMemcpyDeviceToHost, h->stream));
        h->sync();
    }
};

// frames per staging chunk for HOST batches (bounds device staging memory)
static size_t chunk_frames(size_t bytes_per_frame, size_t n)
{
    const size_t budget = (size_t)256 << 20;
    size_t c = budget / (bytes_per_frame ? bytes_per_frame : 1);
    if (c < 1) c = 1;
    return c < n ? c : n;
}

// Pipelined HOST batch: the batch is cut into chunks of `chunk` frames; the host->device copy of chunk
// i+1 (stream s_h2d), the kernels of chunk i (the handle's stream) and the device->host copy of chunk
// i-1 (stream s_d2h) run concurrently, so both PCIe directions stay busy.  run(dout, d0, d1, f0, nf)
// enqueues the kernels of one chunk on h->stream.  Synchronous for the caller, like generic_work -- unless `async`
// (GFDM_MEM_HOST_ASYNC): then the call returns with everything enqueued and gfdm_sync() completes it; a following
// call on the same handle is ordered behind it by the slot events.
// seed_out: the staged output starts as the caller's data (kernels that leave elements untouched).
// Sizes are BYTES per frame (the symbol side of a frame may be one byte per symbol, see the chunk entries).
template <class Run>
static void host_pipeline_bytes(HandleBase* h, size_t n, size_t chunk, const void* in0v, size_t in0_sz, const void* in1v,
                                size_t in1_sz, void* outv, size_t out_sz, bool seed_out, bool async, Run run)
{
    if (!n) return;
    if (chunk < 1) chunk = 1;
    const unsigned char* in0 = static_cast<const unsigned char*>(in0v);
    const unsigned char* in1 = static_cast<const unsigned char*>(in1v);
    unsigned char* out = static_cast<unsigned char*>(outv);
    if (!h->s_h2d) {
        GFDM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        GFDM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GFDM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            GFDM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_run[i], cudaEventDisableTiming));
            GFDM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
    }
    const size_t c0 = std::min(chunk, n);
    for (int i = 0; i < 2; ++i) {
        if (in0 && in0_sz) h->slot_in0[i].ensure(c0 * in0_sz);
        if (in1 && in1_sz) h->slot_in1[i].ensure(c0 * in1_sz);
        h->slot_out[i].ensure(std::max<size_t>(c0 * out_sz, 16));
    }
    // work queued earlier on the handle's stream may still use the slots (a DEVICE call followed by a HOST call)
    GFDM_CUDA_CHECK(cudaEventRecord(h->ev_run[0], h->stream));
    GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_run[0], 0));
    size_t idx = 0;
    for (size_t f0 = 0; f0 < n; f0 += chunk, ++idx) {
        const size_t nf = std::min(chunk, n - f0);
        const int sl = (int)(idx & 1);
        void* d0 = nullptr;
        void* d1 = nullptr;
        void* dout = h->slot_out[sl].p;
        // inputs of this slot are free once the kernels of chunk idx-2 have run (chunks of an earlier call: ev_run[0] above)
        if (idx >= 2) GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_run[sl], 0));
        if (in0 && in0_sz) {
            d0 = h->slot_in0[sl].p;
            GFDM_CUDA_CHECK(cudaMemcpyAsync(d0, in0 + f0 * in0_sz, nf * in0_sz, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (in1 && in1_sz) {
            d1 = h->slot_in1[sl].p;
            GFDM_CUDA_CHECK(cudaMemcpyAsync(d1, in1 + f0 * in1_sz, nf * in1_sz, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (seed_out) {
            // the output slot is free once chunk idx-2 (or the last chunk of an earlier asynchronous call that used this
            // slot) has been copied back; waiting on a never-recorded event is a no-op
            GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_out[sl], 0));
            GFDM_CUDA_CHECK(cudaMemcpyAsync(dout, out + f0 * out_sz, nf * out_sz, cudaMemcpyHostToDevice, h->s_h2d));
        }
        GFDM_CUDA_CHECK(cudaEventRecord(h->ev_in[sl], h->s_h2d));
        GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_in[sl], 0));
        GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_out[sl], 0));
        run(dout, d0, d1, f0, nf);
        GFDM_CUDA_CHECK(cudaEventRecord(h->ev_run[sl], h->stream));
        GFDM_CUDA_CHECK(cudaStreamWaitEvent(h->s_d2h, h->ev_run[sl], 0));
        if (out_sz)
            GFDM_CUDA_CHECK(cudaMemcpyAsync(out + f0 * out_sz, dout, nf * out_sz, cudaMemcpyDeviceToHost, h->s_d2h));
        GFDM_CUDA_CHECK(cudaEventRecord(h->ev_out[sl], h->s_d2h));
    }
    if (async) return;
    GFDM_CUDA_CHECK(cudaStreamSynchronize(h->s_d2h));
    GFDM_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}
// the same with every array complex64 and sizes in elements per frame
template <class Run>
static void host_pipeline(HandleBase* h, size_t n, size_t chunk, const gfdm_complex* in0, size_t in0_sz,
                          const gfdm_complex* in1, size_t in1_sz, gfdm_complex* out, size_t out_sz, bool seed_out,
                          bool async, Run run)
{
    host_pipeline_bytes(h, n, chunk, in0, in0_sz * sizeof(cpx), in1, in1_sz * sizeof(cpx), out, out_sz * sizeof(cpx), seed_out, async,
                        [&](void* dout, const void* d0, const void* d1, size_t f0, size_t nf) {
                            run(static_cast<cpx*>(dout), static_cast<const cpx*>(d0), static_cast<const cpx*>(d1), f0, nf);
                        });
}

// frames per pipeline chunk: 32 MB of host traffic per chunk keeps fill/drain below a millisecond
static size_t pipe_chunk(size_t bytes_per_frame, size_t n)
{
    // GFDM_PIPE_CHUNK_MB overrides the chunk size (tools/e2e_sweep.py measured 8..128 MB on B200)
    static const size_t override_mb = [] {
        const char* e = std::getenv("GFDM_PIPE_CHUNK_MB");
        const long v = e ? std::atol(e) : 0;
        return (size_t)(v > 0 && v <= 1024 ? v : 0);
    }();
    const size_t budget = (override_mb ? override_mb : (size_t)32) << 20;
    size_t c = budget / (bytes_per_frame ? bytes_per_frame : 1);
    if (c < 1) c = 1;
    if (c >= n) return n;
    // equal chunks: a short last chunk would leave one copy engine idle for most of a pipeline step
    const size_t k = (n + c - 1) / c;
    return (n + k - 1) / k;
}

/* ======================================================================== */
/* library-wide                                                              */
extern "C" {

const char* gfdm_last_error(void) { return g_err.c_str(); }
const char* gfdm_backend(void) { return "cuda-sm_100a"; }
int gfdm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int gfdm_set_device(int device)
{
    API_TRY
    int n = 0;
    GFDM_CUDA_CHECK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) throw std::invalid_argument("device index out of range");
    t_device = device;
    g_default_device.store(device, std::memory_order_relaxed);
    API_CATCH
}
static HandleBase* base(void* handle)
{
    HandleBase* b = reinterpret_cast<HandleBase*>(handle);
    if (!b || b->magic != HANDLE_MAGIC) throw std::invalid_argument("not a gfdm_b200 handle");
    return b;
}
int gfdm_set_stream(void* handle, void* cuda_stream)
{
    API_TRY
    HandleBase* b = base(handle);
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    b->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    b->own_stream = false;
    API_CATCH
}
int gfdm_sync(void* handle)
{
    API_TRY
    HandleBase* b = base(handle);
    b->use();
    b->sync();
    API_CATCH
}
long long gfdm_launch_count(void* handle)
{
    HandleBase* b = reinterpret_cast<HandleBase*>(handle);
    return (b && b->magic == HANDLE_MAGIC) ? b->launches : -1;
}
const char* gfdm_last_kernel(void* handle)
{
    HandleBase* b = reinterpret_cast<HandleBase*>(handle);
    return (b && b->magic == HANDLE_MAGIC) ? b->last_kernel : "invalid";
}

int gfdm_calculate_signal_energy(float* energy, const gfdm_complex* in, int n)
{
    API_TRY
    if (n < 0) throw std::invalid_argument("n MUST NOT be negative");
    HandleBase hb;
    hb.open();
    try {
        hb.stage_in.ensure(sizeof(cpx) * (size_t)(n > 0 ? n : 1));
        hb.stage_out.ensure(sizeof(float));
        GFDM_CUDA_CHECK(cudaMemcpyAsync(hb.stage_in.p, in, sizeof(cpx) * (size_t)n, cudaMemcpyHostToDevice, hb.stream));
        launch_energy(hb.stage_out.as<float>(), hb.stage_in.as<cpx>(), (size_t)n, hb.stream);
        GFDM_CUDA_CHECK(cudaMemcpyAsync(energy, hb.stage_out.p, sizeof(float), cudaMemcpyDeviceToHost, hb.stream));
        hb.sync();
    } catch (...) {
        hb.close();
        throw;
    }
    hb.close();
    API_CATCH
}

/* ---- FFT engine --------------------------------------------------------- */
struct gfdm_fft : HandleBase {
    FftPlan plan;
    bool forward = true;
};
int gfdm_fft_create(gfdm_fft** out, int fft_size, int forward)
{
    API_TRY
    if (fft_size < 1) throw std::invalid_argument("fft_size MUST be positive");
    std::unique_ptr<gfdm_fft, void (*)(gfdm_fft*)> h(new gfdm_fft, gfdm_fft_destroy); // a throwing step releases stream + device memory
    h->open();
    h->plan.init(fft_size);
    h->forward = forward != 0;
    *out = h.release();
    API_CATCH
}
void gfdm_fft_destroy(gfdm_fft* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->plan.destroy();
    h->close();
    delete h;
}
int gfdm_fft_execute_batch(gfdm_fft* h, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_transforms MUST NOT be negative");
    Staging st(h, mem);
    const size_t el = (size_t)n * h->plan.n;
    const cpx* di = st.in(in, el, h->stage_in);
    cpx* dout = st.out(out, el, h->stage_out);
    h->work_a.ensure(el * sizeof(cpx));
    h->launches += fft_exec(h->plan, dout, di, h->work_a.as<cpx>(), (size_t)n, !h->forward, 1.0f, h->stream);
    h->last_kernel = "stockham_pass";
    st.finish(out, el, h->stage_out);
    API_CATCH
}

/* ---- modulator_kernel_cc ------------------------------------------------ */
struct gfdm_modulator : HandleBase {
    int M = 0, K = 0, L = 0, N = 0;
    std::vector<cf> taps; // normalised
    cpx* d_taps = nullptr;
    FftPlan fft_m, fft_n;
    FusedModem fused; // fused single-kernel path when (M, K, L) has one
};

static void modulator_run(gfdm_modulator* h, cpx* out, const cpx* in, size_t frames)
{
    if (!frames) return;
    if (h->fused.available() && aligned16(in) && aligned16(out)) { // cp.async.bulk / 128-bit accesses need 16-byte aligned addresses
        h->launches += h->fused.modulate(out, in, frames, h->stream);
        h->last_kernel = h->fused.mod_name();
        return;
    }
    if (h->fft_m.n == 0) h->fft_m.init(h->M);
    if (h->fft_n.n == 0) h->fft_n.init(h->N);
    if (generic_smem_supported(h->M, h->K, h->L, h->fft_m, h->fft_n)) { // frame resident in shared memory: one kernel, 16 N bytes
        h->launches += launch_generic_smem_mod(out, in, h->M, h->K, h->L, h->fft_m, h->fft_n, h->d_taps, frames, h->stream);
        h->last_kernel = "generic_smem_mod_kernel";
        return;
    }
    const size_t el = frames * (size_t)h->N;
    h->work_a.ensure(el * sizeof(cpx));
    h->work_b.ensure(el * sizeof(cpx));
    cpx* A = h->work_a.as<cpx>();
    cpx* B = h->work_b.as<cpx>();
    h->launches += fft_exec(h->fft_m, A, in, B, frames * h->K, false, 1.0f, h->stream); // D_k = FFT_M(d_k)
    launch_mod_filter(B, A, h->d_taps, h->M, h->K, h->L, frames, h->stream);            // X
    h->launches += 1;
    h->launches += fft_exec(h->fft_n, out, B, A, frames, true, (float)(1.0 / h->N), h->stream); // IFFT_N / N
    h->last_kernel = "generic:fft_m+mod_filter+ifft_n";
}

int gfdm_modulator_create(gfdm_modulator** out, int M, int K, int L, const gfdm_complex* taps, int n_taps)
{
    API_TRY
    check_taps((size_t)(n_taps > 0 ? n_taps : 0), M, L);
    if (M < 1 || K < 1 || L < 1) throw std::invalid_argument("timeslots, subcarriers and overlap MUST be positive");
    std::unique_ptr<gfdm_modulator, void (*)(gfdm_modulator*)> h(new gfdm_modulator, gfdm_modulator_destroy); // a throwing step releases stream + device memory
    h->M = M; h->K = K; h->L = L; h->N = M * K;
    h->taps = normalize_taps(vec(taps, n_taps), M);
    h->open();
    h->d_taps = dev_upload(h->taps);
    h->fused.init_tx(M, K, L, h->taps);
    *out = h.release();
    API_CATCH
}
void gfdm_modulator_destroy(gfdm_modulator* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    if (h->d_taps) cudaFree(h->d_taps);
    h->fft_m.destroy();
    h->fft_n.destroy();
    h->fused.destroy();
    h->close();
    delete h;
}
int gfdm_modulator_block_size(const gfdm_modulator* h) { return h->N; }
int gfdm_modulator_filter_taps(const gfdm_modulator* h, gfdm_complex* o)
{
    memcpy(o, h->taps.data(), sizeof(cf) * h->taps.size());
    return GFDM_OK;
}
int gfdm_modulator_work_batch(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem, true);
    if (mem == GFDM_MEM_DEVICE) {
        modulator_run(h, reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in), (size_t)n);
    } else {
        host_pipeline(h, (size_t)n, pipe_chunk(2 * sizeof(cpx) * h->N, (size_t)n), in, h->N, nullptr, 0, out, h->N, false,
                      st.async, [&](cpx* dout, const cpx* d0, const cpx*, size_t, size_t nf) { modulator_run(h, dout, d0, nf); });
    }
    API_CATCH
}
int gfdm_modulator_work(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_modulator_work_batch(h, out, in, 1, GFDM_MEM_HOST);
}

/* ---- receiver_kernel_cc ------------------------------------------------- */
struct gfdm_receiver : HandleBase {
    int M = 0, K = 0, L = 0, N = 0;
    std::vector<cf> taps, ic_taps;
    cpx* d_taps = nullptr;
    cpx* d_ic = nullptr;
    FftPlan fft_m, fft_n;
    FusedModem fused;
};

static void receiver_init(gfdm_receiver* h, int M, int K, int L, const std::vector<cf>& taps_in)
{
    check_taps(taps_in.size(), M, L);
    if (L < 2) throw std::invalid_argument("overlap MUST be greater or equal 2");
    if (M < 1 || K < 1) throw std::invalid_argument("timeslots and subcarriers MUST be positive");
    h->M = M; h->K = K; h->L = L; h->N = M * K;
    h->taps = normalize_taps(taps_in, M);
    h->ic_taps.resize(M);
    for (int m = 0; m < M; ++m) { // lib/receiver_kernel_cc.cc:56-63
        const cf a = h->taps[m], b = h->taps[M * (L - 1) + m];
        h->ic_taps[m] = cf(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
    }
    h->open();
    h->d_taps = dev_upload(h->taps);
    h->d_ic = dev_upload(h->ic_taps);
    h->fused.init_rx(M, K, L, h->taps, h->ic_taps);
    h->fft_m.init(M);
}
static void receiver_free(gfdm_receiver* h)
{
    if (h->opened) cudaSetDevice(h->device);
    if (h->d_taps) cudaFree(h->d_taps);
    if (h->d_ic) cudaFree(h->d_ic);
    h->fft_m.destroy();
    h->fft_n.destroy();
    h->fused.destroy();
    h->close();
}

// fft_filter_downsample / fft_equalize_filter_downsample (:301-320) -> R
static void receiver_fd(gfdm_receiver* h, cpx* R, const cpx* in, const cpx* eq, size_t frames)
{
    if (!frames) return;
    if (h->fused.available() && (!eq || h->fused.supports_eq()) && aligned16(R) && aligned16(in) && aligned16(eq)) {
        h->launches += h->fused.demodulate(nullptr, R, in, eq, frames, h->stream);
        h->last_kernel = h->fused.rx_name();
        return;
    }
    if (h->fft_n.n == 0) h->fft_n.init(h->N);
    if (generic_smem_supported(h->M, h->K, h->L, h->fft_m, h->fft_n)) {
        h->launches += launch_generic_smem_rx(R, in, eq, 1, h->M, h->K, h->L, h->fft_m, h->fft_n, h->d_taps, frames, h->stream);
        h->last_kernel = "generic_smem_rx_kernel";
        return;
    }
    const size_t el = frames * (size_t)h->N;
    h->work_a.ensure(el * sizeof(cpx));
    h->work_b.ensure(el * sizeof(cpx));
    cpx* A = h->work_a.as<cpx>();
    cpx* B = h->work_b.as<cpx>();
    h->launches += fft_exec(h->fft_n, A, in, B, frames, false, 1.0f, h->stream); // Y = FFT_N(in)
    if (eq) {
        launch_eq_divide(A, A, eq, el, h->stream);
        h->launches += 1;
    }
    launch_rx_filter(R, A, h->d_taps, h->M, h->K, h->L, frames, h->stream);
    h->launches += 1;
    h->last_kernel = "generic:fft_n+rx_filter";
}
// transform_subcarriers_to_td (:211-225); out != in
static void receiver_td(gfdm_receiver* h, cpx* out, const cpx* R, size_t frames)
{
    if (!frames) return;
    const size_t el = frames * (size_t)h->N;
    h->work_c.ensure(el * sizeof(cpx));
    h->launches += fft_exec(h->fft_m, out, R, h->work_c.as<cpx>(), frames * h->K, true, (float)(1.0 / h->M), h->stream);
}
// both results of the receiver for the cancellation loop: R (kept frequency blocks) and y = IFFT_M(R)/M.  With a
// fused kernel that is two launches of it (the second transform is cheaper there than a staged M-point pass over HBM)
static void receiver_fd_td(gfdm_receiver* h, cpx* R, cpx* y, const cpx* in, const cpx* eq, size_t frames)
{
    if (!frames) return;
    if (h->fused.available() && (!eq || h->fused.supports_eq()) && aligned16(R) && aligned16(y) && aligned16(in) &&
        aligned16(eq)) {
        h->launches += h->fused.demodulate(y, R, in, eq, frames, h->stream);
        return;
    }
    receiver_fd(h, R, in, eq, frames);
    receiver_td(h, y, R, frames);
}
// generic_work[_equalize] (:322-334)
static void receiver_run(gfdm_receiver* h, cpx* out, const cpx* in, const cpx* eq, size_t frames)
{
    if (!frames) return;
    if (h->fused.available() && (!eq || h->fused.supports_eq()) && aligned16(out) && aligned16(in) && aligned16(eq)) {
        h->launches += h->fused.demodulate(out, nullptr, in, eq, frames, h->stream);
        h->last_kernel = h->fused.rx_name();
        return;
    }
    if (h->fft_n.n == 0) h->fft_n.init(h->N);
    if (generic_smem_supported(h->M, h->K, h->L, h->fft_m, h->fft_n)) {
        h->launches += launch_generic_smem_rx(out, in, eq, 0, h->M, h->K, h->L, h->fft_m, h->fft_n, h->d_taps, frames, h->stream);
        h->last_kernel = "generic_smem_rx_kernel";
        return;
    }
    const size_t el = frames * (size_t)h->N;
    h->work_c.ensure(el * sizeof(cpx));
    cpx* R = h->work_c.as<cpx>();
    receiver_fd(h, R, in, eq, frames); // uses work_a / work_b
    // M-point IFFTs: scratch = work_a (free again after receiver_fd)
    h->launches += fft_exec(h->fft_m, out, R, h->work_a.as<cpx>(), frames * h->K, true, (float)(1.0 / h->M), h->stream);
    h->last_kernel = "generic:fft_n+rx_filter+ifft_m";
}
// cancel_sc_interference (:274-299); out may not alias inputs
static void receiver_cancel(gfdm_receiver* h, cpx* out, const cpx* td, const cpx* fd, size_t frames)
{
    if (!frames) return;
    const size_t el = frames * (size_t)h->N;
    h->work_a.ensure(el * sizeof(cpx));
    h->work_b.ensure(el * sizeof(cpx));
    cpx* A = h->work_a.as<cpx>();
    cpx* B = h->work_b.as<cpx>();
    launch_neighbor_sum(A, td, h->M, h->K, frames, h->stream);
    h->launches += 1;
    h->launches += fft_exec(h->fft_m, B, A, out, frames * h->K, false, 1.0f, h->stream); // out as scratch
    launch_ic_subtract(out, fd, B, h->d_ic, h->M, h->K, frames, h->stream);
    h->launches += 1;
    h->last_kernel = "generic:neighbor_sum+fft_m+ic_subtract";
}

int gfdm_receiver_create(gfdm_receiver** out, int M, int K, int L, const gfdm_complex* taps, int n_taps)
{
    API_TRY
    std::unique_ptr<gfdm_receiver, void (*)(gfdm_receiver*)> h(new gfdm_receiver, gfdm_receiver_destroy); // a throwing step releases stream + device memory
    receiver_init(h.get(), M, K, L, vec(taps, n_taps));
    *out = h.release();
    API_CATCH
}
void gfdm_receiver_destroy(gfdm_receiver* h)
{
    if (!h) return;
    receiver_free(h);
    delete h;
}
int gfdm_receiver_block_size(const gfdm_receiver* h) { return h->N; }
int gfdm_receiver_timeslots(const gfdm_receiver* h) { return h->M; }
int gfdm_receiver_subcarriers(const gfdm_receiver* h) { return h->K; }
int gfdm_receiver_overlap(const gfdm_receiver* h) { return h->L; }
int gfdm_receiver_filter_taps(const gfdm_receiver* h, gfdm_complex* o)
{
    memcpy(o, h->taps.data(), sizeof(cf) * h->taps.size());
    return GFDM_OK;
}
int gfdm_receiver_ic_filter_taps(const gfdm_receiver* h, gfdm_complex* o)
{
    memcpy(o, h->ic_taps.data(), sizeof(cf) * h->ic_taps.size());
    return GFDM_OK;
}

enum RxOp { RX_WORK, RX_FD, RX_TD, RX_CANCEL };
static int receiver_batch(gfdm_receiver* h, RxOp op, gfdm_complex* out, const gfdm_complex* in0,
                          const gfdm_complex* in1, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (op == RX_CANCEL && !in1) throw std::invalid_argument("fd_in MUST NOT be NULL");
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](cpx* dout, const cpx* d0, const cpx* d1, size_t, size_t nf) {
        switch (op) {
        case RX_WORK: receiver_run(h, dout, d0, d1, nf); break;
        case RX_FD: receiver_fd(h, dout, d0, d1, nf); break;
        case RX_TD:
            receiver_td(h, dout, d0, nf);
            h->last_kernel = "generic:ifft_m";
            break;
        case RX_CANCEL: receiver_cancel(h, dout, d0, d1, nf); break;
        }
    };
    if (mem == GFDM_MEM_DEVICE)
        run(reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in0), reinterpret_cast<const cpx*>(in1), 0, (size_t)n);
    else
        host_pipeline(h, (size_t)n, pipe_chunk((in1 ? 3 : 2) * sizeof(cpx) * N, (size_t)n), in0, N, in1, N, out, N, false, st.async, run);
    API_CATCH
}
int gfdm_receiver_work_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* eq,
                             int n, int mem)
{
    return receiver_batch(h, RX_WORK, out, in, eq, n, mem);
}
int gfdm_receiver_work(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return receiver_batch(h, RX_WORK, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_work_equalize(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return receiver_batch(h, RX_WORK, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_filter_downsample_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                              const gfdm_complex* eq, int n, int mem)
{
    return receiver_batch(h, RX_FD, out, in, eq, n, mem);
}
int gfdm_receiver_fft_filter_downsample(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return receiver_batch(h, RX_FD, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_fft_equalize_filter_downsample(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                                 const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return receiver_batch(h, RX_FD, out, in, eq, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_transform_subcarriers_to_td_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                                    int n, int mem)
{
    return receiver_batch(h, RX_TD, out, in, nullptr, n, mem);
}
int gfdm_receiver_transform_subcarriers_to_td(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return receiver_batch(h, RX_TD, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_receiver_cancel_sc_interference_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* td,
                                               const gfdm_complex* fd, int n, int mem)
{
    return receiver_batch(h, RX_CANCEL, out, td, fd, n, mem);
}
int gfdm_receiver_cancel_sc_interference(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* td,
                                         const gfdm_complex* fd)
{
    return receiver_batch(h, RX_CANCEL, out, td, fd, 1, GFDM_MEM_HOST);
}

/* ---- advanced_receiver_kernel_cc ---------------------------------------- */
struct gfdm_advanced_receiver : gfdm_receiver {
    std::vector<int> smap;
    int ic_iter = 0;
    int phase_comp = 0;
    int n_points = 0, rule = 0;
    int* d_smap = nullptr;
    unsigned char* d_active = nullptr;
    cpx* d_points = nullptr;
    DecideGrid grid; // O(1) nearest-point decisions when the constellation is a uniform rectangular grid
    DeviceBuf freq_block, ic_time, ic_freq;
};

// generic_work[_equalize] + perform_ic_iterations, lib/advanced_receiver_kernel_cc.cc:56-107
static void advanced_run(gfdm_advanced_receiver* h, cpx* out, const cpx* in, const cpx* eq, size_t frames)
{
    if (!frames) return;
    if (h->fused.sic_available() && (!eq || h->fused.supports_eq()) && h->ic_iter >= 0 && aligned16(out) && aligned16(in) &&
        aligned16(eq)) {
        h->launches += h->fused.demodulate_sic(out, in, eq, frames, h->ic_iter, h->phase_comp, h->stream);
        h->last_kernel = h->fused.sic_name();
        return;
    }
    const size_t el = frames * (size_t)h->N;
    h->freq_block.ensure(el * sizeof(cpx));
    h->ic_time.ensure(el * sizeof(cpx));
    h->ic_freq.ensure(el * sizeof(cpx));
    cpx* FB = h->freq_block.as<cpx>();
    cpx* IT = h->ic_time.as<cpx>();
    cpx* IF = h->ic_freq.as<cpx>();
    const bool one_kernel_iter = sic_iter_supported(h->M, h->K, h->n_points, frames);
    const bool staged_first = !one_kernel_iter || (h->phase_comp > 0 && h->ic_iter > 0);
    {
        // soft symbols go where the first consumer reads them (see the ping-pong below)
        const int rest0 = std::max(0, h->ic_iter);
        cpx* y0 = (staged_first || rest0 % 2 == 0) ? out : IT;
        receiver_fd_td(h, FB, y0, in, eq, frames);
    }
    // iterations that need no phase estimate run as ONE kernel each, ping-ponging between `out` and IT; the first
    // buffer is chosen so that the last iteration lands in `out`
    int j0 = 0;
    if (staged_first) {
        const int staged = one_kernel_iter ? 1 : h->ic_iter; // with phase compensation only iteration 0 is staged
        for (int j = 0; j < staged; ++j) {
            launch_decide(IT, out, h->d_active, h->d_points, h->n_points, h->rule, h->M, h->K, frames, h->stream);
            h->launches += 1;
            if (h->phase_comp > 0 && j == 0) {
                launch_phase_rotate(FB, IT, out, h->d_smap, (int)h->smap.size(), h->M, h->K, frames, h->stream);
                h->launches += 1;
            }
            receiver_cancel(h, IF, IT, FB, frames);
            receiver_td(h, out, IF, frames);
        }
        j0 = staged;
    }
    if (one_kernel_iter) {
        const int rest = std::max(0, h->ic_iter - j0);
        cpx* cur = (rest % 2 == 0) ? out : IT;
        cpx* nxt = (rest % 2 == 0) ? IT : out;
        if (j0 != 0 && cur != out) GFDM_CUDA_CHECK(cudaMemcpyAsync(cur, out, el * sizeof(cpx), cudaMemcpyDeviceToDevice, h->stream));
        for (int j = 0; j < rest; ++j) {
            launch_sic_iter(nxt, cur, FB, h->d_ic, h->d_active, h->d_points, h->n_points, h->rule, h->grid, h->M, h->K, frames, h->stream);
            h->launches += 1;
            std::swap(cur, nxt);
        }
        h->last_kernel = "receiver -> sic_iter_kernel per iteration";
        return;
    }
    h->last_kernel = "generic:advanced_receiver";
}

int gfdm_advanced_receiver_create(gfdm_advanced_receiver** out, int M, int K, int L, const gfdm_complex* taps,
                                  int n_taps, const int* smap, int n_map, int ic_iter, const gfdm_constellation* c,
                                  int do_phase_compensation)
{
    API_TRY
    if (!c || c->n_points < 1 || !c->points) throw std::invalid_argument("constellation MUST hold at least one point");
    // same rules as gfdm_symbol_mapper_create: the sign rule indexes points[0..3]
    if (c->decision_rule != GFDM_DECISION_NEAREST && c->decision_rule != GFDM_DECISION_QPSK_SIGN)
        throw std::invalid_argument("unknown constellation decision rule!");
    if (c->decision_rule == GFDM_DECISION_QPSK_SIGN && c->n_points != 4)
        throw std::invalid_argument("the QPSK sign rule needs exactly 4 constellation points!");
    if (n_map < 0) throw std::invalid_argument("subcarrier_map size MUST NOT be negative");
    for (int i = 0; i < n_map; ++i)
        if (smap[i] < 0 || smap[i] >= K) throw std::invalid_argument("subcarrier_map entries MUST lie in [0, subcarriers)");
    std::unique_ptr<gfdm_advanced_receiver, void (*)(gfdm_advanced_receiver*)> h(new gfdm_advanced_receiver, gfdm_advanced_receiver_destroy); // a throwing step releases stream + device memory
    h->smap.assign(smap, smap + n_map);
    h->ic_iter = ic_iter;
    h->phase_comp = do_phase_compensation;
    h->n_points = c->n_points;
    h->rule = c->decision_rule;
    receiver_init(h.get(), M, K, L, vec(taps, n_taps));
    std::vector<unsigned char> active(K, 0);
    for (int k : h->smap) active[k] = 1;
    h->d_active = upload(active);
    h->d_smap = upload(h->smap);
    h->d_points = dev_upload(vec(c->points, c->n_points));
    {
        std::vector<cpx> p((size_t)c->n_points);
        for (int i = 0; i < c->n_points; ++i) p[i] = make_float2(c->points[i].re, c->points[i].im);
        h->grid = make_decide_grid(p);
    }
    if (h->fused.available()) h->fused.init_sic(h->ic_taps, vec(c->points, c->n_points), h->rule, h->smap);
    *out = h.release();
    API_CATCH
}
void gfdm_advanced_receiver_destroy(gfdm_advanced_receiver* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    if (h->d_smap) cudaFree(h->d_smap);
    if (h->d_active) cudaFree(h->d_active);
    if (h->d_points) cudaFree(h->d_points);
    h->freq_block.release(); h->ic_time.release(); h->ic_freq.release();
    receiver_free(h);
    delete h;
}
int gfdm_advanced_receiver_block_size(const gfdm_advanced_receiver* h) { return h->N; }
int gfdm_advanced_receiver_set_ic(gfdm_advanced_receiver* h, int v) { h->ic_iter = v; return GFDM_OK; }
int gfdm_advanced_receiver_get_ic(const gfdm_advanced_receiver* h) { return h->ic_iter; }
int gfdm_advanced_receiver_set_phase_compensation(gfdm_advanced_receiver* h, int v) { h->phase_comp = v; return GFDM_OK; }
int gfdm_advanced_receiver_get_phase_compensation(const gfdm_advanced_receiver* h) { return h->phase_comp; }
int gfdm_advanced_receiver_work_batch(gfdm_advanced_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                      const gfdm_complex* eq, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](cpx* dout, const cpx* d0, const cpx* d1, size_t, size_t nf) { advanced_run(h, dout, d0, d1, nf); };
    if (mem == GFDM_MEM_DEVICE)
        run(reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in), reinterpret_cast<const cpx*>(eq), 0, (size_t)n);
    else
        host_pipeline(h, (size_t)n, pipe_chunk((eq ? 3 : 2) * sizeof(cpx) * N, (size_t)n), in, N, eq, N, out, N, false, st.async, run);
    API_CATCH
}
int gfdm_advanced_receiver_work(gfdm_advanced_receiver* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_advanced_receiver_work_batch(h, out, in, nullptr, 1, GFDM_MEM_HOST);
}
int gfdm_advanced_receiver_work_equalize(gfdm_advanced_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                         const gfdm_complex* eq)
{
    if (!eq) return fail(GFDM_ERR_INVALID_ARGUMENT, "f_eq_in MUST NOT be NULL");
    return gfdm_advanced_receiver_work_batch(h, out, in, eq, 1, GFDM_MEM_HOST);
}

/* ---- resource_mapper_kernel_cc ------------------------------------------ */
struct MapperCore {
    int M = 0, K = 0, A = 0;
    size_t block_size = 0, frame_size = 0;
    bool per_timeslot = true, is_mapper = true;
    std::vector<int> smap; // sorted
    int* d_smap = nullptr;
    int* d_inv = nullptr;
    // lib/resource_mapper_kernel_cc.cc:30-70
    void validate(int M_, int K_, int A_, const int* map, int n_map, bool pts, bool mapper)
    {
        if (A_ > K_)
            throw std::invalid_argument("active_subcarriers(" + std::to_string(A_) +
                                        ") MUST be smaller or equal to subcarriers(" + std::to_string(K_) + ")!");
        if (n_map != A_)
            throw std::invalid_argument("number of subcarrier_map entries(" + std::to_string(n_map) +
                                        ") MUST be equal to active_subcarriers(" + std::to_string(A_) + ")!");
        if (M_ < 1 || K_ < 1) throw std::invalid_argument("timeslots and subcarriers MUST be positive");
        smap.assign(map, map + n_map);
        std::sort(smap.begin(), smap.end());
        if (std::adjacent_find(smap.begin(), smap.end()) != smap.end())
            throw std::invalid_argument("All entries in subcarrier_map MUST be unique!");
        if (!smap.empty() && smap.front() < 0)
            throw std::invalid_argument("All subcarrier indices MUST be greater or equal to ZERO!");
        // the reference tests `> subcarriers` (:65) and would then index out of bounds for == subcarriers
        if (!smap.empty() && smap.back() >= K_)
            throw std::invalid_argument("All subcarrier indices MUST be smaller or equal to subcarriers!");
        M = M_; K = K_; A = A_;
        block_size = (size_t)M * A;
        frame_size = (size_t)M * K;
        per_timeslot = pts;
        is_mapper = mapper;
    }
    void to_device()
    {
        std::vector<int> inv(K, -1);
        for (int a = 0; a < A; ++a) inv[smap[a]] = a;
        d_smap = upload(smap);
        d_inv = upload(inv);
    }
    void destroy()
    {
        if (d_smap) cudaFree(d_smap);
        if (d_inv) cudaFree(d_inv);
        d_smap = d_inv = nullptr;
    }
    void check_map_size(size_t n) const
    {
        if (n > block_size)
            throw std::invalid_argument("input vector size(" + std::to_string(n) +
                                        ") MUST not exceed active_subcarriers * timeslots(" +
                                        std::to_string(block_size) + ")!");
    }
    void check_demap_size(size_t n) const
    {
        if (n > block_size)
            throw std::invalid_argument("output vector size(" + std::to_string(n) +
                                        ") MUST not exceed active_subcarriers * timeslots(" +
                                        std::to_string(block_size) + ")!");
    }
};
struct gfdm_resource_mapper : HandleBase {
    MapperCore c;
};
int gfdm_resource_mapper_create(gfdm_resource_mapper** out, int M, int K, int A, const int* smap, int n_map,
                                int per_timeslot, int is_mapper)
{
    API_TRY
    std::unique_ptr<gfdm_resource_mapper, void (*)(gfdm_resource_mapper*)> h(new gfdm_resource_mapper, gfdm_resource_mapper_destroy); // a throwing step releases stream + device memory
    h->c.validate(M, K, A, smap, n_map, per_timeslot != 0, is_mapper != 0);
    h->open();
    h->c.to_device();
    *out = h.release();
    API_CATCH
}
void gfdm_resource_mapper_destroy(gfdm_resource_mapper* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->c.destroy();
    h->close();
    delete h;
}
size_t gfdm_resource_mapper_frame_size(const gfdm_resource_mapper* h) { return h->c.frame_size; }
size_t gfdm_resource_mapper_block_size(const gfdm_resource_mapper* h) { return h->c.block_size; }
size_t gfdm_resource_mapper_input_vector_size(const gfdm_resource_mapper* h) { return h->c.is_mapper ? h->c.block_size : h->c.frame_size; }
size_t gfdm_resource_mapper_output_vector_size(const gfdm_resource_mapper* h) { return h->c.is_mapper ? h->c.frame_size : h->c.block_size; }
int gfdm_resource_mapper_map_to_resources_batch(gfdm_resource_mapper* h, gfdm_complex* out, const gfdm_complex* in,
                                                size_t sz, int n, int mem)
{
    API_TRY
    h->use();
    h->c.check_map_size(sz);
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem);
    const size_t fs = h->c.frame_size;
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (fs + sz), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = sz ? st.in(in + f0 * sz, nf * sz, h->stage_in) : nullptr;
        cpx* dout = st.out(out + f0 * fs, nf * fs, h->stage_out);
        launch_map(dout, di, h->c.d_inv, h->c.M, h->c.K, h->c.A, h->c.per_timeslot, sz, sz, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "map_kernel";
        st.finish(out + f0 * fs, nf * fs, h->stage_out);
    }
    API_CATCH
}
int gfdm_resource_mapper_map_to_resources(gfdm_resource_mapper* h, gfdm_complex* out, const gfdm_complex* in, size_t n)
{
    return gfdm_resource_mapper_map_to_resources_batch(h, out, in, n, 1, GFDM_MEM_HOST);
}
int gfdm_resource_mapper_demap_from_resources_batch(gfdm_resource_mapper* h, gfdm_complex* out,
                                                    const gfdm_complex* in, size_t sz, int n, int mem)
{
    API_TRY
    h->use();
    h->c.check_demap_size(sz);
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (sz == 0) return GFDM_OK;
    Staging st(h, mem);
    const size_t fs = h->c.frame_size;
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (fs + sz), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = st.in(in + f0 * fs, nf * fs, h->stage_in);
        cpx* dout = st.out(out + f0 * sz, nf * sz, h->stage_out);
        launch_demap(dout, di, h->c.d_smap, h->c.M, h->c.K, h->c.A, h->c.per_timeslot, sz, sz, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "demap_kernel";
        st.finish(out + f0 * sz, nf * sz, h->stage_out);
    }
    API_CATCH
}
int gfdm_resource_mapper_demap_from_resources(gfdm_resource_mapper* h, gfdm_complex* out, const gfdm_complex* in,
                                              size_t n)
{
    return gfdm_resource_mapper_demap_from_resources_batch(h, out, in, n, 1, GFDM_MEM_HOST);
}

/* ---- add_cyclic_prefix_cc ----------------------------------------------- */
struct PrefixCore {
    int block_len = 0, cp_len = 0, cs_len = 0, ramp_len = 0, cyclic_shift = 0;
    std::vector<cf> front, back;
    cpx* d_front = nullptr;
    cpx* d_back = nullptr;
    // lib/add_cyclic_prefix_cc.cc:30-57
    void validate(int bl, int cp, int cs, int ramp, const std::vector<cf>& w, int shift)
    {
        const int window_len = bl + cp + cs;
        if (w.size() != (size_t)window_len && w.size() != (size_t)(2 * ramp)) {
            std::stringstream s;
            s << "number of window taps(" << w.size() << ") MUST be equal to 2*ramp_len(" << 2 * ramp
              << ") OR block_len+cp_len (" << window_len << ")!";
            throw std::invalid_argument(s.str());
        }
        if (bl < 1 || cp < 0 || cs < 0 || ramp < 0 || (size_t)ramp > w.size())
            throw std::invalid_argument("block_len MUST be positive; cp_len, cs_len, ramp_len MUST NOT be negative");
        block_len = bl; cp_len = cp; cs_len = cs; ramp_len = ramp; cyclic_shift = shift;
        front.assign(w.begin(), w.begin() + ramp);
        back.assign(w.end() - ramp, w.end());
    }
    void to_device()
    {
        d_front = dev_upload(front);
        d_back = dev_upload(back);
    }
    void destroy()
    {
        if (d_front) cudaFree(d_front);
        if (d_back) cudaFree(d_back);
        d_front = d_back = nullptr;
    }
    int frame_size() const { return block_len + cp_len + cs_len; }
    void check_shift(int s) const
    {
        if (s < 0 || s > cs_len || cp_len + s > block_len)
            throw std::invalid_argument("cyclic_shift MUST lie in [0, cs_len] and cp_len + cyclic_shift MUST NOT exceed block_len");
    }
};
struct gfdm_cyclic_prefixer : HandleBase {
    PrefixCore c;
};
int gfdm_cyclic_prefixer_create(gfdm_cyclic_prefixer** out, int block_len, int cp_len, int cs_len, int ramp_len,
                                const gfdm_complex* w, int n_w, int cyclic_shift)
{
    API_TRY
    std::unique_ptr<gfdm_cyclic_prefixer, void (*)(gfdm_cyclic_prefixer*)> h(new gfdm_cyclic_prefixer, gfdm_cyclic_prefixer_destroy); // a throwing step releases stream + device memory
    h->c.validate(block_len, cp_len, cs_len, ramp_len, vec(w, n_w), cyclic_shift);
    h->open();
    h->c.to_device();
    *out = h.release();
    API_CATCH
}
void gfdm_cyclic_prefixer_destroy(gfdm_cyclic_prefixer* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->c.destroy();
    h->close();
    delete h;
}
int gfdm_cyclic_prefixer_block_size(const gfdm_cyclic_prefixer* h) { return h->c.block_len; }
int gfdm_cyclic_prefixer_frame_size(const gfdm_cyclic_prefixer* h) { return h->c.frame_size(); }
int gfdm_cyclic_prefixer_cyclic_shift(const gfdm_cyclic_prefixer* h) { return h->c.cyclic_shift; }
int gfdm_cyclic_prefixer_add_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, gfdm_complex* out, const gfdm_complex* in,
                                                 int shift, int n, int mem)
{
    API_TRY
    h->use();
    h->c.check_shift(shift);
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem);
    const size_t N = h->c.block_len, W = h->c.frame_size();
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (N + W), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = st.in(in + f0 * N, nf * N, h->stage_in);
        cpx* dout = st.out(out + f0 * W, nf * W, h->stage_out);
        launch_add_cp(dout, di, h->c.block_len, h->c.cp_len, h->c.cs_len, h->c.ramp_len, h->c.d_front, h->c.d_back,
                      shift, W, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "add_cp_kernel";
        st.finish(out + f0 * W, nf * W, h->stage_out);
    }
    API_CATCH
}
int gfdm_cyclic_prefixer_add_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out, const gfdm_complex* in, int shift)
{
    return gfdm_cyclic_prefixer_add_cyclic_prefix_batch(h, out, in, shift, 1, GFDM_MEM_HOST);
}
int gfdm_cyclic_prefixer_work(gfdm_cyclic_prefixer* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_cyclic_prefixer_add_cyclic_prefix_batch(h, out, in, h->c.cyclic_shift, 1, GFDM_MEM_HOST);
}
int gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                                    const gfdm_complex* in, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem);
    const size_t N = h->c.block_len, W = h->c.frame_size();
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (N + W), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = st.in(in + f0 * W, nf * W, h->stage_in);
        cpx* dout = st.out(out + f0 * N, nf * N, h->stage_out);
        launch_remove_cp(dout, di, h->c.block_len, h->c.cp_len, h->c.cs_len, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "remove_cp_kernel";
        st.finish(out + f0 * N, nf * N, h->stage_out);
    }
    API_CATCH
}
int gfdm_cyclic_prefixer_remove_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out, const gfdm_complex* in)
{
    return gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(h, out, in, 1, GFDM_MEM_HOST);
}

/* ---- preamble_channel_estimator_cc -------------------------------------- */
struct gfdm_channel_estimator : HandleBase {
    int M = 0, K = 0, A = 0, dc_free = 0, which = 0;
    float g[9];
    FftPlan fft_k, fft_2k;
    cpx* d_inv0 = nullptr;
    cpx* d_inv1 = nullptr;
    float* d_g = nullptr;
    DeviceBuf fbuf, hbuf, filt, fscratch;
    int n_est() const { return A + (dc_free ? 1 : 0); }
};

int gfdm_channel_estimator_create(gfdm_channel_estimator** out, int M, int K, int A, int is_dc_free, int which,
                                  const gfdm_complex* preamble, int n_preamble)
{
    API_TRY
    if (M < 1 || K < 2 || A < 2 || A > K)
        throw std::invalid_argument("timeslots MUST be positive and 2 <= active_subcarriers <= fft_len");
    if (n_preamble < 2 * K) throw std::invalid_argument("preamble MUST hold at least 2 * fft_len samples");
    // odd A: the reference's interpolate_frame (:238-274) writes up to index N + M/2 - 1 of the N-long frame when dc free,
    // filter_preamble_estimate correlates against an unfilled entry and estimate_snr leaves cnrs[A-1] unset
    if (A % 2) throw std::invalid_argument("active_subcarriers MUST be even (the reference writes past its frame buffer for odd values)");
    if (A + (is_dc_free ? 1 : 0) > K)
        throw std::invalid_argument("active_subcarriers (+1 if dc free) MUST NOT exceed fft_len");
    std::unique_ptr<gfdm_channel_estimator, void (*)(gfdm_channel_estimator*)> h(new gfdm_channel_estimator, gfdm_channel_estimator_destroy); // a throwing step releases stream + device memory
    h->M = M; h->K = K; h->A = A; h->dc_free = is_dc_free ? 1 : 0; h->which = which;
    // initialize_gaussian_filter(sigma_sq = 1, 9 taps), lib/preamble_channel_estimator_cc.cc:86-100
    float s = 0.0f;
    for (int i = 0; i < 9; ++i) {
        const float val = std::pow(float(i - (9 / 2)), 2.0f) / 1.0f;
        h->g[i] = std::exp(-0.5f * val);
        s += h->g[i];
    }
    for (int i = 0; i < 9; ++i) h->g[i] = h->g[i] / s;
    h->open();
    h->fft_k.init(K);
    h->fft_2k.init(2 * K);
    h->d_g = upload(std::vector<float>(h->g, h->g + 9));
    // inv_ref_h = 0.5 / FFT_K(preamble half h)  (:111-119): FFT on the device engine, division on the host
    std::vector<cf> pre = vec(preamble, 2 * K), F(2 * K);
    h->fbuf.ensure(sizeof(cpx) * 2 * K);
    h->hbuf.ensure(sizeof(cpx) * 2 * K);
    h->fscratch.ensure(sizeof(cpx) * 2 * K);
    GFDM_CUDA_CHECK(cudaMemcpyAsync(h->hbuf.p, pre.data(), sizeof(cpx) * 2 * K, cudaMemcpyHostToDevice, h->stream));
    h->launches += fft_exec(h->fft_k, h->fbuf.as<cpx>(), h->hbuf.as<cpx>(), h->fscratch.as<cpx>(), 2, false, 1.0f, h->stream);
    GFDM_CUDA_CHECK(cudaMemcpyAsync(F.data(), h->fbuf.p, sizeof(cpx) * 2 * K, cudaMemcpyDeviceToHost, h->stream));
    h->sync();
    std::vector<cf> inv(2 * K);
    for (int i = 0; i < 2 * K; ++i) inv[i] = cf(0.5f, 0.0f) / F[i];
    h->d_inv0 = dev_upload(std::vector<cf>(inv.begin(), inv.begin() + K));
    h->d_inv1 = dev_upload(std::vector<cf>(inv.begin() + K, inv.end()));
    *out = h.release();
    API_CATCH
}
void gfdm_channel_estimator_destroy(gfdm_channel_estimator* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->fft_k.destroy(); h->fft_2k.destroy();
    if (h->d_inv0) cudaFree(h->d_inv0);
    if (h->d_inv1) cudaFree(h->d_inv1);
    if (h->d_g) cudaFree(h->d_g);
    h->fbuf.release(); h->hbuf.release(); h->filt.release(); h->fscratch.release();
    h->close();
    delete h;
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

// rx [frames][2K] -> H [frames][K]
static void est_preamble_channel(gfdm_channel_estimator* h, cpx* H, const cpx* rx, size_t frames)
{
    const size_t el = frames * 2 * (size_t)h->K;
    h->fbuf.ensure(el * sizeof(cpx));
    h->fscratch.ensure(el * sizeof(cpx));
    h->launches += fft_exec(h->fft_k, h->fbuf.as<cpx>(), rx, h->fscratch.as<cpx>(), 2 * frames, false, 1.0f, h->stream);
    launch_est_combine(H, h->fbuf.as<cpx>(), h->d_inv0, h->d_inv1, h->K, frames, h->stream);
    h->launches += 1;
}
static void est_frame(gfdm_channel_estimator* h, cpx* fe, const cpx* rx, size_t frames)
{
    if (!frames) return;
    if (est_fused_supported(h->K, h->A, h->dc_free)) {
        launch_est_fused(fe, rx, h->fft_k.d_tw, h->d_inv0, h->d_inv1, h->d_g, h->M, h->K, h->A, h->dc_free, frames, h->stream);
        h->launches += 1;
        h->last_kernel = "est_fused_kernel";
        return;
    }
    h->hbuf.ensure(frames * (size_t)h->K * sizeof(cpx));
    h->filt.ensure(frames * (size_t)h->n_est() * sizeof(cpx));
    est_preamble_channel(h, h->hbuf.as<cpx>(), rx, frames);
    launch_est_filter(h->filt.as<cpx>(), h->hbuf.as<cpx>(), h->d_g, h->K, h->A, h->dc_free, frames, h->stream);
    launch_est_interp(fe, h->filt.as<cpx>(), h->M, h->K, h->A, h->dc_free, frames, h->stream);
    h->launches += 2;
    h->last_kernel = "fft_k+est_combine+est_filter+est_interp";
}

enum EstOp { EST_CHANNEL, EST_FILTER, EST_INTERP, EST_ZF, EST_FRAME };
static int estimator_batch(gfdm_channel_estimator* h, EstOp op, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem);
    const size_t K = h->K, N = (size_t)h->M * h->K, ne = h->n_est();
    size_t in_sz = 0, out_sz = 0;
    switch (op) {
    case EST_CHANNEL: in_sz = 2 * K; out_sz = K; break;
    case EST_FILTER: in_sz = K; out_sz = ne; break;
    case EST_INTERP: in_sz = ne; out_sz = N; break;
    case EST_ZF: in_sz = N; out_sz = N; break;
    case EST_FRAME: in_sz = 2 * K; out_sz = N; break;
    }
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (in_sz + out_sz), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = st.in(in + f0 * in_sz, nf * in_sz, h->stage_in);
        cpx* dout = st.out(out + f0 * out_sz, nf * out_sz, h->stage_out);
        // interpolation leaves some bins untouched (reference quirk): seed the staged output with the caller's data
        if (mem == GFDM_MEM_HOST && (op == EST_INTERP || op == EST_FRAME))
            GFDM_CUDA_CHECK(cudaMemcpyAsync(dout, out + f0 * out_sz, nf * out_sz * sizeof(cpx), cudaMemcpyHostToDevice, h->stream));
        switch (op) {
        case EST_CHANNEL:
            est_preamble_channel(h, dout, di, nf);
            h->last_kernel = "fft_k+est_combine";
            break;
        case EST_FILTER:
            launch_est_filter(dout, di, h->d_g, h->K, h->A, h->dc_free, nf, h->stream);
            h->launches += 1;
            h->last_kernel = "est_filter_kernel";
            break;
        case EST_INTERP:
            launch_est_interp(dout, di, h->M, h->K, h->A, h->dc_free, nf, h->stream);
            h->launches += 1;
            h->last_kernel = "est_interp_kernel";
            break;
        case EST_ZF:
            launch_zf_prepare(dout, di, nf * N, h->stream);
            h->launches += 1;
            h->last_kernel = "zf_prepare_kernel";
            break;
        case EST_FRAME: est_frame(h, dout, di, nf); break;
        }
        st.finish(out + f0 * out_sz, nf * out_sz, h->stage_out);
    }
    API_CATCH
}
int gfdm_channel_estimator_estimate_preamble_channel(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* rx)
{
    return estimator_batch(h, EST_CHANNEL, o, rx, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_filter_preamble_estimate(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* e)
{
    return estimator_batch(h, EST_FILTER, o, e, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_interpolate_frame(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* e)
{
    return estimator_batch(h, EST_INTERP, o, e, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_prepare_for_zf(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* e)
{
    return estimator_batch(h, EST_ZF, o, e, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_estimate_frame(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* rx)
{
    return estimator_batch(h, EST_FRAME, o, rx, 1, GFDM_MEM_HOST);
}
int gfdm_channel_estimator_estimate_frame_batch(gfdm_channel_estimator* h, gfdm_complex* o, const gfdm_complex* rx,
                                                int n, int mem)
{
    return estimator_batch(h, EST_FRAME, o, rx, n, mem);
}
int gfdm_channel_estimator_estimate_snr_batch(gfdm_channel_estimator* h, float* snr, float* cnrs,
                                              const gfdm_complex* rx, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (n == 0) return GFDM_OK;
    Staging st(h, mem);
    const size_t nf = (size_t)n, el = nf * 2 * (size_t)h->K;
    const cpx* di = st.in(rx, el, h->stage_in);
    h->fbuf.ensure(el * sizeof(cpx));
    h->fscratch.ensure(el * sizeof(cpx));
    h->launches += fft_exec(h->fft_2k, h->fbuf.as<cpx>(), di, h->fscratch.as<cpx>(), nf, false, 1.0f, h->stream);
    float* d_snr = snr;
    float* d_cnr = cnrs;
    if (mem == GFDM_MEM_HOST) {
        h->stage_out.ensure(sizeof(float) * nf * (size_t)(1 + h->A));
        d_snr = h->stage_out.as<float>();
        d_cnr = cnrs ? d_snr + nf : nullptr;
    }
    launch_est_snr(d_snr, d_cnr, h->fbuf.as<cpx>(), h->K, h->A, h->dc_free, nf, h->stream);
    h->launches += 1;
    h->last_kernel = "fft_2k+est_snr_kernel";
    if (mem == GFDM_MEM_HOST) {
        GFDM_CUDA_CHECK(cudaMemcpyAsync(snr, d_snr, sizeof(float) * nf, cudaMemcpyDeviceToHost, h->stream));
        if (cnrs) GFDM_CUDA_CHECK(cudaMemcpyAsync(cnrs, d_cnr, sizeof(float) * nf * h->A, cudaMemcpyDeviceToHost, h->stream));
        h->sync();
    }
    API_CATCH
}
int gfdm_channel_estimator_estimate_snr(gfdm_channel_estimator* h, float* snr, float* cnrs, const gfdm_complex* rx)
{
    return gfdm_channel_estimator_estimate_snr_batch(h, snr, cnrs, rx, 1, GFDM_MEM_HOST);
}

/* ---- transmitter_kernel ------------------------------------------------- */
struct gfdm_transmitter : HandleBase {
    MapperCore map;
    PrefixCore pre;
    int M = 0, K = 0, L = 0, N = 0;
    std::vector<cf> taps;
    cpx* d_taps = nullptr;
    FftPlan fft_m, fft_n;
    FusedModem fused;
    std::vector<int> shifts;
    int preamble_size = 0;
    cpx* d_preambles = nullptr; // [n_shifts][preamble_size]
    bool force_staged = false;  // tests: run the chain as separate kernels (mapper, modulator, preamble, prefixer)
    DeviceBuf mapped, frame;
    int out_size() const { return pre.frame_size() + preamble_size; }
    int shift_index(int s) const
    {
        for (size_t i = 0; i < shifts.size(); ++i)
            if (shifts[i] == s) return (int)i; // first entry wins, as unordered_map::emplace (transmitter_kernel.cc:67-71)
        throw std::invalid_argument("cyclic_shift has no preamble");
    }
};

// map + modulate (lib/transmitter_kernel.cc:78-84): in [frames][nin] -> blk [frames][N]
static void tx_modulate(gfdm_transmitter* h, cpx* blk, const cpx* in, size_t nin, size_t frames)
{
    if (!frames) return;
    const size_t el = frames * (size_t)h->N;
    h->mapped.ensure(el * sizeof(cpx));
    cpx* mp = h->mapped.as<cpx>();
    launch_map(mp, in, h->map.d_inv, h->M, h->K, h->map.A, h->map.per_timeslot, nin, nin, frames, h->stream);
    h->launches += 1;
    if (h->fused.available()) { // `mapped` is our own 256-byte aligned buffer
        h->launches += h->fused.modulate(blk, mp, frames, h->stream);
        return;
    }
    if (h->fft_m.n == 0) h->fft_m.init(h->M);
    if (h->fft_n.n == 0) h->fft_n.init(h->N);
    if (generic_smem_supported(h->M, h->K, h->L, h->fft_m, h->fft_n)) {
        h->launches += launch_generic_smem_mod(blk, mp, h->M, h->K, h->L, h->fft_m, h->fft_n, h->d_taps, frames, h->stream);
        return;
    }
    h->work_a.ensure(el * sizeof(cpx));
    h->work_b.ensure(el * sizeof(cpx));
    cpx* A = h->work_a.as<cpx>();
    cpx* B = h->work_b.as<cpx>();
    h->launches += fft_exec(h->fft_m, A, mp, B, frames * h->K, false, 1.0f, h->stream);
    launch_mod_filter(B, A, h->d_taps, h->M, h->K, h->L, frames, h->stream);
    h->launches += 1;
    h->launches += fft_exec(h->fft_n, blk, B, A, frames, true, (float)(1.0 / h->N), h->stream);
}
// preamble + CP frame for one shift (lib/transmitter_kernel.cc:86-98): out [frames][out_size]
static void tx_add_frame(gfdm_transmitter* h, cpx* out, const cpx* blk, int shift, size_t frames)
{
    if (!frames) return;
    const int idx = h->shift_index(shift);
    h->pre.check_shift(shift);
    const size_t os = h->out_size();
    launch_copy_rows(out, h->d_preambles + (size_t)idx * h->preamble_size, h->preamble_size, os, frames, h->stream);
    launch_add_cp(out + h->preamble_size, blk, h->N, h->pre.cp_len, h->pre.cs_len, h->pre.ramp_len, h->pre.d_front,
                  h->pre.d_back, shift, os, frames, h->stream);
    h->launches += 2;
}

// Shaper parameters as the kernels see them (gfdm_burst_shaper below owns them)
struct ShaperArgs {
    int pre = 0, post = 0;
    cpx scale = make_float2(1.f, 0.f);
};

// transmitter_kernel::generic_work for nf frames on DEVICE pointers, n_ant antennas (cyclic shifts): antenna a at
// out + a*ant_stride, one row per frame.  One kernel for the whole chain when the shape has a shared-memory resident
// modulator, separate kernels otherwise.  sh != nullptr: every row passes through the short_burst_shaper as well
// (row = pre + out_size + post) -- inside the same kernel on the fused path.
static void tx_chain_device(gfdm_transmitter* h, cpx* out, size_t ant_stride, const cpx* di, size_t in_sz, size_t nf, size_t n_ant,
                            const ShaperArgs* sh)
{
    if (!nf) return;
    const size_t N = h->N, os = h->out_size();
    TxArgs ta;
    ta.inv_map = h->map.d_inv; ta.front = h->pre.d_front; ta.back = h->pre.d_back; ta.preambles = h->d_preambles;
    ta.A = h->map.A; ta.per_timeslot = h->map.per_timeslot ? 1 : 0; ta.n_in = (int)in_sz;
    ta.cp = h->pre.cp_len; ta.cs = h->pre.cs_len; ta.ramp = h->pre.ramp_len; ta.P = h->preamble_size;
    ta.n_ant = (int)n_ant;
    ta.ant_stride = ant_stride;
    for (size_t a = 0; a < n_ant && a < (size_t)GFDM_TX_MAX_ANT; ++a) {
        ta.shift[a] = h->shifts[a];
        ta.pre_idx[a] = h->shift_index(h->shifts[a]);
    }
    if (sh) {
        ta.shaped = 1; ta.pre_pad = sh->pre; ta.post_pad = sh->post; ta.sc_re = sh->scale.x; ta.sc_im = sh->scale.y;
    }
    if (!h->force_staged && h->fused.available() && h->fused.supports_tx_chain(ta) && aligned16(di)) {
        h->launches += h->fused.transmit(out, di, ta, nf, h->stream);
        h->last_kernel = h->fused.tx_name();
        return;
    }
    h->frame.ensure(nf * N * sizeof(cpx));
    tx_modulate(h, h->frame.as<cpx>(), di, in_sz, nf);
    for (size_t a = 0; a < n_ant; ++a) {
        cpx* o = out + a * ant_stride;
        if (!sh) {
            tx_add_frame(h, o, h->frame.as<cpx>(), h->shifts[a], nf);
        } else {
            h->work_c.ensure(nf * os * sizeof(cpx));
            tx_add_frame(h, h->work_c.as<cpx>(), h->frame.as<cpx>(), h->shifts[a], nf);
            launch_burst_shape(o, h->work_c.as<cpx>(), (int)os, sh->pre, sh->post, sh->scale, nf, h->stream);
            h->launches += 1;
        }
    }
    h->last_kernel = h->fused.available() ? (sh ? "map+fused_mod+copy_rows+add_cp+burst_shape" : "map+fused_mod+copy_rows+add_cp")
                                          : "generic:transmitter";
}

int gfdm_transmitter_create(gfdm_transmitter** out, int M, int K, int A, int cp, int cs, int ramp, const int* smap,
                            int n_map, int per_timeslot, int L, const gfdm_complex* taps, int n_taps,
                            const gfdm_complex* w, int n_w, const int* shifts, int n_shifts,
                            const gfdm_complex* const* preambles, const int* preamble_sizes, int n_preambles)
{
    API_TRY
    if (n_preambles < 1) throw std::invalid_argument("at least one preamble is required");
    std::unique_ptr<gfdm_transmitter, void (*)(gfdm_transmitter*)> h(new gfdm_transmitter, gfdm_transmitter_destroy); // a throwing step releases stream + device memory
    // member construction order of the reference: mapper, modulator, prefixer (transmitter_kernel.cc:47-52)
    h->map.validate(M, K, A, smap, n_map, per_timeslot != 0, true);
    check_taps((size_t)(n_taps > 0 ? n_taps : 0), M, L);
    if (L < 1) throw std::invalid_argument("overlap MUST be positive");
    h->pre.validate(M * K, cp, cs, ramp, vec(w, n_w), 0);
    if (n_shifts != n_preambles)
        throw std::invalid_argument("Number of cyclic shifts and number of preambles do not match!");
    for (int i = 0; i < n_preambles; ++i)
        if (preamble_sizes[i] != preamble_sizes[0]) throw std::invalid_argument("All preambles must have equal size!");
    h->M = M; h->K = K; h->L = L; h->N = M * K;
    h->taps = normalize_taps(vec(taps, n_taps), M);
    h->shifts.assign(shifts, shifts + n_shifts);
    h->preamble_size = preamble_sizes[0];
    h->open();
    h->map.to_device();
    h->pre.to_device();
    h->d_taps = dev_upload(h->taps);
    h->fused.init_tx(M, K, L, h->taps);
    std::vector<cf> all;
    for (int i = 0; i < n_preambles; ++i) {
        std::vector<cf> p = vec(preambles[i], preamble_sizes[i]);
        all.insert(all.end(), p.begin(), p.end());
    }
    h->d_preambles = dev_upload(all);
    *out = h.release();
    API_CATCH
}
void gfdm_transmitter_destroy(gfdm_transmitter* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->map.destroy();
    h->pre.destroy();
    if (h->d_taps) cudaFree(h->d_taps);
    if (h->d_preambles) cudaFree(h->d_preambles);
    h->fft_m.destroy(); h->fft_n.destroy();
    h->fused.destroy();
    h->mapped.release(); h->frame.release();
    h->close();
    delete h;
}
int gfdm_transmitter_input_vector_size(const gfdm_transmitter* h) { return (int)h->map.block_size; }
int gfdm_transmitter_output_vector_size(const gfdm_transmitter* h) { return h->out_size(); }
int gfdm_transmitter_n_cyclic_shifts(const gfdm_transmitter* h) { return (int)h->shifts.size(); }
int gfdm_transmitter_cyclic_shifts(const gfdm_transmitter* h, int* o)
{
    memcpy(o, h->shifts.data(), sizeof(int) * h->shifts.size());
    return GFDM_OK;
}

int gfdm_transmitter_set_chain_fusion(gfdm_transmitter* h, int on)
{
    h->force_staged = on == 0;
    return GFDM_OK;
}

enum TxOp { TX_WORK, TX_WORK_ALL, TX_MODULATE, TX_ADD_FRAME };
static int transmitter_batch(gfdm_transmitter* h, TxOp op, gfdm_complex* out, const gfdm_complex* in, int nin,
                             int shift, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (op != TX_ADD_FRAME) {
        if (nin < 0) throw std::invalid_argument("ninput_size MUST NOT be negative");
        h->map.check_map_size((size_t)nin);
    } else {
        h->shift_index(shift);
        h->pre.check_shift(shift);
    }
    if (op == TX_WORK || op == TX_WORK_ALL)
        for (int s : h->shifts) h->pre.check_shift(s);
    Staging st(h, mem);
    const size_t N = h->N, os = h->out_size(), ns = h->shifts.size();
    const size_t in_sz = op == TX_ADD_FRAME ? N : (size_t)nin;
    const size_t out_sz = op == TX_MODULATE ? N : os;
    const size_t n_ant = op == TX_WORK_ALL ? ns : 1;
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * (in_sz + out_sz * n_ant + 3 * N), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = in_sz ? st.in(in + f0 * in_sz, nf * in_sz, h->stage_in) : nullptr;
        // HOST staging of the all-antenna form: the chunk is laid out [ant][nf][os] on the device
        cpx* dout = mem == GFDM_MEM_DEVICE ? reinterpret_cast<cpx*>(out) : st.out(out, nf * out_sz * n_ant, h->stage_out);
        switch (op) {
        case TX_MODULATE:
            tx_modulate(h, mem == GFDM_MEM_DEVICE ? dout + f0 * N : dout, di, in_sz, nf);
            h->last_kernel = h->fused.available() ? h->fused.mod_name() : "generic:map+fft_m+mod_filter+ifft_n";
            break;
        case TX_ADD_FRAME:
            tx_add_frame(h, mem == GFDM_MEM_DEVICE ? dout + f0 * os : dout, di, shift, nf);
            h->last_kernel = "copy_rows+add_cp_kernel";
            break;
        case TX_WORK:
        case TX_WORK_ALL:
            tx_chain_device(h, mem == GFDM_MEM_DEVICE ? dout + f0 * os : dout, (mem == GFDM_MEM_DEVICE ? (size_t)n : nf) * os, di,
                            in_sz, nf, n_ant, nullptr);
            break;
        }
        if (mem == GFDM_MEM_HOST) {
            for (size_t a = 0; a < n_ant; ++a)
                GFDM_CUDA_CHECK(cudaMemcpyAsync(out + (a * (size_t)n + f0) * out_sz, dout + a * nf * out_sz,
                                                nf * out_sz * sizeof(cpx), cudaMemcpyDeviceToHost, h->stream));
            h->sync();
        }
    }
    API_CATCH
}
int gfdm_transmitter_work(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int n)
{
    return transmitter_batch(h, TX_WORK, out, in, n, 0, 1, GFDM_MEM_HOST);
}
int gfdm_transmitter_modulate(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int n)
{
    return transmitter_batch(h, TX_MODULATE, out, in, n, 0, 1, GFDM_MEM_HOST);
}
int gfdm_transmitter_add_frame(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int shift)
{
    return transmitter_batch(h, TX_ADD_FRAME, out, in, 0, shift, 1, GFDM_MEM_HOST);
}
int gfdm_transmitter_work_batch(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int nin, int n, int mem)
{
    return transmitter_batch(h, TX_WORK, out, in, nin, 0, n, mem);
}
int gfdm_transmitter_work_all_batch(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in, int nin, int n,
                                    int mem)
{
    return transmitter_batch(h, TX_WORK_ALL, out, in, nin, 0, n, mem);
}

/* ======================================================================== */
/* rows either side of the hot path     */

/* ---- remove_prefix_cc: lib/remove_prefix_cc_impl.cc:84-115 ---------------- */
struct gfdm_remove_prefix : HandleBase {
    int frame_len = 0, block_len = 0, offset = 0;
};
int gfdm_remove_prefix_create(gfdm_remove_prefix** out, int frame_len, int block_len, int offset)
{
    API_TRY
    // the block does not validate (:44-61) and would read past the frame; the ABI rejects it
    if (frame_len < 1 || block_len < 1 || offset < 0 || offset + block_len > frame_len)
        throw std::invalid_argument("remove_prefix: offset + block_len MUST NOT exceed frame_len!");
    std::unique_ptr<gfdm_remove_prefix, void (*)(gfdm_remove_prefix*)> h(new gfdm_remove_prefix, gfdm_remove_prefix_destroy); // a throwing step releases stream + device memory
    h->frame_len = frame_len; h->block_len = block_len; h->offset = offset;
    h->open();
    *out = h.release();
    API_CATCH
}
void gfdm_remove_prefix_destroy(gfdm_remove_prefix* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->close();
    delete h;
}
int gfdm_remove_prefix_work_batch(gfdm_remove_prefix* h, gfdm_complex* out, const gfdm_complex* in, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    Staging st(h, mem, true);
    // the strided copy of remove_cyclic_prefix with cp = offset, cs = whatever follows the block
    const int cs = h->frame_len - h->block_len - h->offset;
    auto run = [&](cpx* dout, const cpx* d0, const cpx*, size_t, size_t nf) {
        launch_remove_cp(dout, d0, h->block_len, h->offset, cs, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "remove_cp_kernel";
    };
    if (mem == GFDM_MEM_DEVICE)
        run(reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in), nullptr, 0, (size_t)n);
    else
        host_pipeline(h, (size_t)n, pipe_chunk(sizeof(cpx) * ((size_t)h->frame_len + h->block_len), (size_t)n), in,
                      (size_t)h->frame_len, nullptr, 0, out, (size_t)h->block_len, false, st.async, run);
    API_CATCH
}

/* ---- extract_burst_cc: lib/extract_burst_cc_impl.cc:43-242 ----------------- */
struct gfdm_extract_burst : HandleBase {
    int burst_len = 0, tag_backoff = 0;
    bool cfo = false;
    // burst descriptors: two pinned host slots + device copies, so that a DEVICE call never waits for the GPU
    // (slot i is rewritten only after the copy that read it two calls ago has completed)
    BurstDesc* h_desc[2] = { nullptr, nullptr };
    size_t h_cap[2] = { 0, 0 };
    DeviceBuf d_desc[2];
    cudaEvent_t ev_copied[2] = { nullptr, nullptr };
    int slot = 0;
};
int gfdm_extract_burst_create(gfdm_extract_burst** out, int burst_len, int tag_backoff, int activate_cfo_correction)
{
    API_TRY
    if (burst_len < 1) throw std::invalid_argument("extract_burst: burst_len MUST be positive!");
    // A negative back-off moves the burst BEHIND its tag; the block's admission test (:152) looks at the tag only and the
    // reference then reads past its input window.  The ABI rejects the configuration (deliberate deviation).
    if (tag_backoff < 0) throw std::invalid_argument("extract_burst: tag_backoff MUST NOT be negative!");
    std::unique_ptr<gfdm_extract_burst, void (*)(gfdm_extract_burst*)> h(new gfdm_extract_burst, gfdm_extract_burst_destroy); // a throwing step releases stream + device memory
    h->burst_len = burst_len; h->tag_backoff = tag_backoff; h->cfo = activate_cfo_correction != 0;
    h->open();
    *out = h.release();
    API_CATCH
}
void gfdm_extract_burst_destroy(gfdm_extract_burst* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    for (int i = 0; i < 2; ++i) {
        h->d_desc[i].release();
        if (h->h_desc[i]) cudaFreeHost(h->h_desc[i]);
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    }
    h->close();
    delete h;
}
int gfdm_extract_burst_activate_cfo_compensation(gfdm_extract_burst* h, int on)
{
    h->cfo = on != 0;
    return GFDM_OK;
}
int gfdm_extract_burst_work(gfdm_extract_burst* h, gfdm_complex* out, int max_bursts, const gfdm_complex* in,
                            long long n_in, const long long* burst_starts, const float* scale_factors,
                            const gfdm_complex* phase_rotations, int n_tags, int* n_produced, long long* n_consumed,
                            int mem)
{
    API_TRY
    h->use();
    if (max_bursts < 0 || n_in < 0 || n_tags < 0) throw std::invalid_argument("sizes MUST NOT be negative");
    for (int i = 1; i < n_tags; ++i)
        if (burst_starts[i] < burst_starts[i - 1]) throw std::invalid_argument("extract_burst: burst_starts MUST be sorted!");
    Staging st(h, mem);
    // the control flow of general_work (:117-242) runs on the host over the tag arrays and yields one
    // descriptor per produced burst; the samples are touched by the kernel only
    const long long BL = h->burst_len, noutput_items = (long long)max_bursts * BL, avail_items = n_in;
    long long consumed_items = avail_items, produced_items = 0;
    const int sl = h->slot;
    h->slot ^= 1;
    if (!h->ev_copied[sl]) GFDM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_copied[sl], cudaEventDisableTiming));
    GFDM_CUDA_CHECK(cudaEventSynchronize(h->ev_copied[sl])); // the copy that last read this slot (no-op when never recorded)
    const size_t need = (size_t)std::min<long long>(n_tags, max_bursts);
    if (need > h->h_cap[sl]) {
        if (h->h_desc[sl]) cudaFreeHost(h->h_desc[sl]);
        h->h_desc[sl] = nullptr;
        h->h_cap[sl] = 0;
        GFDM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&h->h_desc[sl]), sizeof(BurstDesc) * need));
        h->h_cap[sl] = need;
    }
    BurstDesc* desc = h->h_desc[sl];
    int nb = 0;
    for (int t = 0; t < n_tags; ++t) {
        const long long burst_start = burst_starts[t];
        if (avail_items - burst_start >= BL && produced_items + BL <= noutput_items) {
            BurstDesc d{};
            d.start = burst_start - h->tag_backoff;
            d.scale = scale_factors ? scale_factors[t] : 1.0f;
            d.pr_re = phase_rotations ? phase_rotations[t].re : 1.0f; // get_phase_rotation (:88-96) runs on the device
            d.pr_im = phase_rotations ? phase_rotations[t].im : 0.0f;
            desc[nb++] = d;
            produced_items += BL;
            consumed_items = burst_start + BL;
        } else {
            consumed_items = burst_start > 0 ? burst_start : 0;
            break;
        }
    }
    if (nb > 0) {
        h->d_desc[sl].ensure(sizeof(BurstDesc) * (size_t)nb);
        GFDM_CUDA_CHECK(cudaMemcpyAsync(h->d_desc[sl].p, desc, sizeof(BurstDesc) * (size_t)nb, cudaMemcpyHostToDevice, h->stream));
        GFDM_CUDA_CHECK(cudaEventRecord(h->ev_copied[sl], h->stream));
        const cpx* di = st.in(in, (size_t)n_in, h->stage_in);
        cpx* dout = st.out(out, (size_t)nb * BL, h->stage_out);
        h->launches += launch_extract_burst(dout, di, h->d_desc[sl].as<BurstDesc>(), h->burst_len, h->cfo, nb, n_in, h->stream);
        h->last_kernel = "extract_burst_kernel";
        st.finish(out, (size_t)nb * BL, h->stage_out);
    }
    if (n_produced) *n_produced = nb;
    if (n_consumed) *n_consumed = consumed_items;
    API_CATCH
}

/* ---- symbol mapping: python/pygfdm/symbolmapping.py:27-47, utils.py:47-51 -- */
struct gfdm_symbol_mapper : HandleBase {
    std::vector<cf> points;
    int rule = 0, bits = 0;
    cpx* d_points = nullptr;
    DecideGrid grid; // O(1) decisions when the constellation is a uniform rectangular grid
};
int gfdm_symbol_mapper_create(gfdm_symbol_mapper** out, const gfdm_constellation* c)
{
    API_TRY
    if (!c || !c->points || c->n_points < 1 || c->n_points > 256)
        throw std::invalid_argument("constellation MUST have between 1 and 256 points!");
    if (c->decision_rule != GFDM_DECISION_NEAREST && c->decision_rule != GFDM_DECISION_QPSK_SIGN)
        throw std::invalid_argument("unknown constellation decision rule!");
    if (c->decision_rule == GFDM_DECISION_QPSK_SIGN && c->n_points != 4)
        throw std::invalid_argument("the QPSK sign rule needs exactly 4 constellation points!");
    std::unique_ptr<gfdm_symbol_mapper, void (*)(gfdm_symbol_mapper*)> h(new gfdm_symbol_mapper, gfdm_symbol_mapper_destroy); // a throwing step releases stream + device memory
    h->points = vec(c->points, c->n_points);
    h->rule = c->decision_rule;
    for (int b = 0; b <= 8; ++b)
        if ((1 << b) == c->n_points) h->bits = b;
    {
        std::vector<cpx> p(h->points.size());
        for (size_t i = 0; i < p.size(); ++i) p[i] = make_float2(h->points[i].real(), h->points[i].imag());
        h->grid = make_decide_grid(p);
    }
    h->open();
    h->d_points = dev_upload(h->points);
    *out = h.release();
    API_CATCH
}
void gfdm_symbol_mapper_destroy(gfdm_symbol_mapper* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    if (h->d_points) cudaFree(h->d_points);
    h->close();
    delete h;
}
int gfdm_symbol_mapper_n_points(const gfdm_symbol_mapper* h) { return (int)h->points.size(); }
int gfdm_symbol_mapper_bits_per_symbol(const gfdm_symbol_mapper* h) { return h->bits; }
int gfdm_symbol_mapper_points(const gfdm_symbol_mapper* h, gfdm_complex* o)
{
    memcpy(o, h->points.data(), sizeof(cf) * h->points.size());
    return GFDM_OK;
}
int gfdm_symbol_mapper_decision_rule(const gfdm_symbol_mapper* h) { return h->rule; }

extern "C++" {
// run(dout, din, n) on device pointers; HOST arrays are staged in chunks of <= 64 Mi symbols
template <class Run>
static void symbol_batch(gfdm_symbol_mapper* h, void* out, size_t out_b, const void* in, size_t in_b, size_t n, int mem,
                         Run run)
{
    Staging st(h, mem);
    if (mem == GFDM_MEM_DEVICE) {
        run(out, in, n);
        return;
    }
    const size_t c = std::min<size_t>(n, (size_t)64 << 20);
    for (size_t i0 = 0; i0 < n; i0 += c) {
        const size_t ni = std::min(c, n - i0);
        h->stage_in.ensure(ni * in_b);
        h->stage_out.ensure(ni * out_b);
        GFDM_CUDA_CHECK(cudaMemcpyAsync(h->stage_in.p, static_cast<const unsigned char*>(in) + i0 * in_b, ni * in_b,
                                        cudaMemcpyHostToDevice, h->stream));
        run(h->stage_out.p, h->stage_in.p, ni);
        GFDM_CUDA_CHECK(cudaMemcpyAsync(static_cast<unsigned char*>(out) + i0 * out_b, h->stage_out.p, ni * out_b,
                                        cudaMemcpyDeviceToHost, h->stream));
        h->sync();
    }
}
} // extern "C++"
int gfdm_symbol_mapper_map_chunks_batch(gfdm_symbol_mapper* h, gfdm_complex* out, const unsigned char* chunks, size_t n,
                                        int mem)
{
    API_TRY
    h->use();
    symbol_batch(h, out, sizeof(cpx), chunks, 1, n, mem, [&](void* o, const void* i, size_t ni) {
        launch_map_chunks(static_cast<cpx*>(o), static_cast<const unsigned char*>(i), h->d_points, (int)h->points.size(), ni,
                          h->stream);
        h->launches += ni ? 1 : 0;
        h->last_kernel = "map_chunks_kernel";
    });
    API_CATCH
}
int gfdm_symbol_mapper_decide_batch(gfdm_symbol_mapper* h, unsigned char* chunks, const gfdm_complex* in, size_t n, int mem)
{
    API_TRY
    h->use();
    symbol_batch(h, chunks, 1, in, sizeof(cpx), n, mem, [&](void* o, const void* i, size_t ni) {
        launch_decide_chunks(static_cast<unsigned char*>(o), static_cast<const cpx*>(i), h->d_points, (int)h->points.size(),
                             h->rule, h->grid, ni, h->stream);
        h->launches += ni ? 1 : 0;
        h->last_kernel = "decide_chunks_kernel";
    });
    API_CATCH
}
int gfdm_symbol_mapper_bits2symbols_batch(gfdm_symbol_mapper* h, gfdm_complex* out, const unsigned char* bits, size_t n,
                                          int mem)
{
    API_TRY
    h->use();
    if (h->bits < 1) throw std::invalid_argument("bits2symbols: the constellation size MUST be a power of two >= 2!");
    symbol_batch(h, out, sizeof(cpx), bits, (size_t)h->bits, n, mem, [&](void* o, const void* i, size_t ni) {
        launch_bits2symbols(static_cast<cpx*>(o), static_cast<const unsigned char*>(i), h->d_points, (int)h->points.size(),
                            h->bits, ni, h->stream);
        h->launches += ni ? 1 : 0;
        h->last_kernel = "bits2symbols_kernel";
    });
    API_CATCH
}
int gfdm_symbol_mapper_symbols2bits_batch(gfdm_symbol_mapper* h, unsigned char* bits, const gfdm_complex* in, size_t n,
                                          int mem)
{
    API_TRY
    h->use();
    if (h->bits < 1) throw std::invalid_argument("symbols2bits: the constellation size MUST be a power of two >= 2!");
    symbol_batch(h, bits, (size_t)h->bits, in, sizeof(cpx), n, mem, [&](void* o, const void* i, size_t ni) {
        launch_symbols2bits(static_cast<unsigned char*>(o), static_cast<const cpx*>(i), h->d_points, (int)h->points.size(),
                            h->rule, h->bits, h->grid, ni, h->stream);
        h->launches += ni ? 1 : 0;
        h->last_kernel = "symbols2bits_kernel";
    });
    API_CATCH
}

/* ---- chunk entries of the path kernels -------------------------------------- */
// The constellation lives on the symbol mapper's device; handles of one chain share a device.
static void check_same_device(const HandleBase* a, const gfdm_symbol_mapper* sm)
{
    if (!sm || sm->magic != HANDLE_MAGIC) throw std::invalid_argument("symbol mapper MUST NOT be NULL");
    if (a->device != sm->device) throw std::invalid_argument("the symbol mapper lives on another device");
}

static void modulator_run_chunks(gfdm_modulator* h, const gfdm_symbol_mapper* sm, cpx* out, const unsigned char* chunks,
                                 size_t frames)
{
    if (!frames) return;
    const int np = (int)sm->points.size();
    if (h->fused.available() && h->fused.supports_chunks(np) && aligned16(chunks) && aligned16(out)) {
        h->launches += h->fused.modulate_chunks(out, chunks, sm->d_points, np, frames, h->stream);
        h->last_kernel = h->fused.modc_name();
        return;
    }
    // shapes without a byte-input kernel: lookup kernel into scratch, then the symbol path
    const size_t el = frames * (size_t)h->N;
    h->work_c.ensure(el * sizeof(cpx));
    launch_map_chunks(h->work_c.as<cpx>(), chunks, sm->d_points, np, el, h->stream);
    h->launches += 1;
    modulator_run(h, out, h->work_c.as<cpx>(), frames);
}
int gfdm_modulator_work_chunks_batch(gfdm_modulator* h, const gfdm_symbol_mapper* sm, gfdm_complex* out,
                                     const unsigned char* chunks, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    check_same_device(h, sm);
    Staging st(h, mem, true);
    if (mem == GFDM_MEM_DEVICE) {
        modulator_run_chunks(h, sm, reinterpret_cast<cpx*>(out), chunks, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk((sizeof(cpx) + 1) * h->N, (size_t)n), chunks, (size_t)h->N, nullptr, 0, out,
                            sizeof(cpx) * h->N, false, st.async, [&](void* dout, const void* d0, const void*, size_t, size_t nf) {
                                modulator_run_chunks(h, sm, static_cast<cpx*>(dout), static_cast<const unsigned char*>(d0), nf);
                            });
    }
    API_CATCH
}

static void receiver_run_decide(gfdm_receiver* h, const gfdm_symbol_mapper* sm, unsigned char* chunks_out, const cpx* in,
                                const cpx* eq, size_t frames)
{
    if (!frames) return;
    const int np = (int)sm->points.size();
    if (h->fused.available() && (!eq || h->fused.supports_eq()) && h->fused.supports_chunks(np) && aligned16(in) &&
        aligned16(chunks_out) && (!eq || aligned16(eq))) {
        h->launches += h->fused.demodulate_decide(chunks_out, in, eq, sm->d_points, np, sm->rule, sm->grid, frames, h->stream);
        h->last_kernel = h->fused.rxd_name();
        return;
    }
    const size_t el = frames * (size_t)h->N;
    h->work_c.ensure(el * sizeof(cpx));
    receiver_run(h, h->work_c.as<cpx>(), in, eq, frames);
    launch_decide_chunks(chunks_out, h->work_c.as<cpx>(), sm->d_points, np, sm->rule, sm->grid, el, h->stream);
    h->launches += 1;
}
int gfdm_receiver_work_decide_batch(gfdm_receiver* h, const gfdm_symbol_mapper* sm, unsigned char* chunks_out,
                                    const gfdm_complex* in, const gfdm_complex* eq, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    check_same_device(h, sm);
    Staging st(h, mem, true);
    const size_t N = h->N;
    if (mem == GFDM_MEM_DEVICE) {
        receiver_run_decide(h, sm, chunks_out, reinterpret_cast<const cpx*>(in), reinterpret_cast<const cpx*>(eq), (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(((eq ? 2 : 1) * sizeof(cpx) + 1) * N, (size_t)n), in, sizeof(cpx) * N, eq,
                            sizeof(cpx) * N, chunks_out, N, false, st.async,
                            [&](void* dout, const void* d0, const void* d1, size_t, size_t nf) {
                                receiver_run_decide(h, sm, static_cast<unsigned char*>(dout), static_cast<const cpx*>(d0),
                                                    static_cast<const cpx*>(d1), nf);
                            });
    }
    API_CATCH
}

int gfdm_transmitter_work_chunks_batch(gfdm_transmitter* h, const gfdm_symbol_mapper* sm, gfdm_complex* out,
                                       const unsigned char* chunks, int nin, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (nin < 0) throw std::invalid_argument("ninput_size MUST NOT be negative");
    h->map.check_map_size((size_t)nin);
    check_same_device(h, sm);
    for (int s : h->shifts) h->pre.check_shift(s);
    Staging st(h, mem, true);
    const size_t os = h->out_size(), N = h->N;
    const int np = (int)sm->points.size();
    auto run = [&](cpx* dout, const unsigned char* dch, size_t nf) {
        TxArgs ta;
        ta.inv_map = h->map.d_inv; ta.front = h->pre.d_front; ta.back = h->pre.d_back; ta.preambles = h->d_preambles;
        ta.A = h->map.A; ta.per_timeslot = h->map.per_timeslot ? 1 : 0; ta.n_in = nin;
        ta.cp = h->pre.cp_len; ta.cs = h->pre.cs_len; ta.ramp = h->pre.ramp_len; ta.P = h->preamble_size;
        ta.n_ant = 1;
        ta.ant_stride = nf * os;
        ta.shift[0] = h->shifts[0];
        ta.pre_idx[0] = h->shift_index(h->shifts[0]);
        ta.points = sm->d_points;
        ta.n_points = np;
        if (!h->force_staged && h->fused.available() && h->fused.supports_tx_chain_chunks(ta) && aligned16(dch)) {
            h->launches += h->fused.transmit_chunks(dout, dch, ta, nf, h->stream);
            h->last_kernel = h->fused.txc_name();
            return;
        }
        // lookup kernel, then the symbol chain (fused or staged, as gfdm_transmitter_work_batch would run it)
        const size_t el = nf * (size_t)nin;
        h->work_c.ensure(std::max<size_t>(el, 1) * sizeof(cpx));
        launch_map_chunks(h->work_c.as<cpx>(), dch, sm->d_points, np, el, h->stream);
        h->launches += 1;
        const cpx* di = h->work_c.as<cpx>();
        if (!h->force_staged && h->fused.available() && h->fused.supports_tx_chain(ta)) {
            h->launches += h->fused.transmit(dout, di, ta, nf, h->stream);
            h->last_kernel = h->fused.tx_name();
            return;
        }
        h->frame.ensure(nf * N * sizeof(cpx));
        tx_modulate(h, h->frame.as<cpx>(), di, (size_t)nin, nf);
        tx_add_frame(h, dout, h->frame.as<cpx>(), h->shifts[0], nf);
        h->last_kernel = h->fused.available() ? "map_chunks+map+fused_mod+copy_rows+add_cp" : "generic:transmitter";
    };
    if (mem == GFDM_MEM_DEVICE) {
        run(reinterpret_cast<cpx*>(out), chunks, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(sizeof(cpx) * os + (size_t)nin, (size_t)n), chunks, (size_t)nin, nullptr, 0,
                            out, sizeof(cpx) * os, false, st.async, [&](void* dout, const void* d0, const void*, size_t, size_t nf) {
                                run(static_cast<cpx*>(dout), static_cast<const unsigned char*>(d0), nf);
                            });
    }
    API_CATCH
}

int gfdm_resource_mapper_demap_chunks_batch(gfdm_resource_mapper* h, unsigned char* out, const unsigned char* in, size_t sz,
                                            int n, int mem)
{
    API_TRY
    h->use();
    h->c.check_demap_size(sz);
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (sz == 0) return GFDM_OK;
    Staging st(h, mem, true);
    const size_t fs = h->c.frame_size;
    auto run = [&](void* dout, const void* d0, const void*, size_t, size_t nf) {
        launch_demap_chunks(static_cast<unsigned char*>(dout), static_cast<const unsigned char*>(d0), h->c.d_smap, h->c.M,
                            h->c.K, h->c.A, h->c.per_timeslot, sz, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "demap_chunks_kernel";
    };
    if (mem == GFDM_MEM_DEVICE)
        run(out, in, nullptr, 0, (size_t)n);
    else
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(fs + sz, (size_t)n), in, fs, nullptr, 0, out, sz, false, st.async, run);
    API_CATCH
}

/* ---- receiver reading frames in place (remove_prefix_cc fused into its loads) ---- */
int gfdm_receiver_work_strided_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in, const gfdm_complex* eq,
                                     size_t in_stride, size_t in_offset, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    const size_t N = h->N;
    if (in_stride < in_offset + N) throw std::invalid_argument("in_stride MUST hold in_offset + block_size samples");
    Staging st(h, mem, true);
    auto run = [&](cpx* dout, const cpx* d0, const cpx* d1, size_t, size_t nf) {
        if (!nf) return;
        const cpx* first = d0 + in_offset;
        // cp.async.bulk needs 16-byte aligned frame starts: an even stride behind an aligned first frame
        if (h->fused.available() && h->fused.supports_stride() && (!d1 || h->fused.supports_eq()) && in_stride % 2 == 0 &&
            aligned16(first) && aligned16(dout) && aligned16(d1)) {
            h->launches += h->fused.demodulate(dout, nullptr, first, d1, nf, h->stream, in_stride == N ? 0 : in_stride);
            h->last_kernel = h->fused.rx_name();
            return;
        }
        h->work_fmt.ensure(nf * N * sizeof(cpx));
        launch_remove_cp(h->work_fmt.as<cpx>(), d0, (int)N, (int)in_offset, (int)(in_stride - N - in_offset), nf, h->stream);
        h->launches += 1;
        receiver_run(h, dout, h->work_fmt.as<cpx>(), d1, nf);
    };
    if (mem == GFDM_MEM_DEVICE)
        run(reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in), reinterpret_cast<const cpx*>(eq), 0, (size_t)n);
    else
        host_pipeline(h, (size_t)n, pipe_chunk(sizeof(cpx) * (in_stride + (eq ? 2 : 1) * N), (size_t)n), in, in_stride, eq, N, out, N, false,
                      st.async, run);
    API_CATCH
}

/* ---- short_burst_shaper: lib/short_burst_shaper_impl.cc:57-84, 161-182 ------ */
struct gfdm_burst_shaper : HandleBase {
    ShaperArgs a;
};
int gfdm_burst_shaper_create(gfdm_burst_shaper** out, int pre_padding, int post_padding, float scale_re, float scale_im)
{
    API_TRY
    if (pre_padding < 0) throw std::invalid_argument("Pre-padding length MUST be >= 0!");   // :77-79
    if (post_padding < 0) throw std::invalid_argument("Post-padding length MUST be >= 0!"); // :80-82
    std::unique_ptr<gfdm_burst_shaper, void (*)(gfdm_burst_shaper*)> h(new gfdm_burst_shaper, gfdm_burst_shaper_destroy);
    h->a.pre = pre_padding; h->a.post = post_padding; h->a.scale = make_float2(scale_re, scale_im);
    h->open();
    *out = h.release();
    API_CATCH
}
void gfdm_burst_shaper_destroy(gfdm_burst_shaper* h)
{
    if (!h) return;
    if (h->opened) cudaSetDevice(h->device);
    h->close();
    delete h;
}
int gfdm_burst_shaper_pre_padding(const gfdm_burst_shaper* h) { return h->a.pre; }
int gfdm_burst_shaper_post_padding(const gfdm_burst_shaper* h) { return h->a.post; }
int gfdm_burst_shaper_work_batch(gfdm_burst_shaper* h, gfdm_complex* out, const gfdm_complex* in, int burst_len, int n, int mem)
{
    API_TRY
    h->use();
    if (n < 0 || burst_len < 0) throw std::invalid_argument("burst_len and n_bursts MUST NOT be negative");
    Staging st(h, mem, true);
    const size_t row = (size_t)h->a.pre + burst_len + h->a.post;
    auto run = [&](cpx* dout, const cpx* d0, const cpx*, size_t, size_t nf) {
        launch_burst_shape(dout, d0, burst_len, h->a.pre, h->a.post, h->a.scale, nf, h->stream);
        h->launches += 1;
        h->last_kernel = "burst_shape_kernel";
    };
    if (mem == GFDM_MEM_DEVICE)
        run(reinterpret_cast<cpx*>(out), reinterpret_cast<const cpx*>(in), nullptr, 0, (size_t)n);
    else
        host_pipeline(h, (size_t)n, pipe_chunk(sizeof(cpx) * (row + burst_len), (size_t)n), in, (size_t)burst_len, nullptr, 0, out, row,
                      false, st.async, run);
    API_CATCH
}
int gfdm_transmitter_work_shaped_batch(gfdm_transmitter* h, const gfdm_burst_shaper* sh, gfdm_complex* out, const gfdm_complex* in,
                                       int nin, int n, int all_antennas, int mem)
{
    API_TRY
    h->use();
    if (!sh || sh->magic != HANDLE_MAGIC) throw std::invalid_argument("burst shaper MUST NOT be NULL");
    if (sh->device != h->device) throw std::invalid_argument("the burst shaper lives on another device");
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (nin < 0) throw std::invalid_argument("ninput_size MUST NOT be negative");
    h->map.check_map_size((size_t)nin);
    for (int s : h->shifts) h->pre.check_shift(s);
    Staging st(h, mem);
    const size_t row = (size_t)sh->a.pre + h->out_size() + sh->a.post, n_ant = all_antennas ? h->shifts.size() : 1;
    const size_t c = mem == GFDM_MEM_DEVICE ? (size_t)n : chunk_frames(sizeof(cpx) * ((size_t)nin + row * n_ant + 4 * (size_t)h->N), (size_t)n);
    for (size_t f0 = 0; f0 < (size_t)n; f0 += c) {
        const size_t nf = std::min(c, (size_t)n - f0);
        const cpx* di = nin ? st.in(in + f0 * (size_t)nin, nf * (size_t)nin, h->stage_in) : nullptr;
        cpx* dout = mem == GFDM_MEM_DEVICE ? reinterpret_cast<cpx*>(out) : st.out(out, nf * row * n_ant, h->stage_out);
        tx_chain_device(h, mem == GFDM_MEM_DEVICE ? dout + f0 * row : dout, (mem == GFDM_MEM_DEVICE ? (size_t)n : nf) * row, di,
                        (size_t)nin, nf, n_ant, &sh->a);
        if (mem == GFDM_MEM_HOST) {
            for (size_t a = 0; a < n_ant; ++a)
                GFDM_CUDA_CHECK(cudaMemcpyAsync(out + (a * (size_t)n + f0) * row, dout + a * nf * row, nf * row * sizeof(cpx),
                                                cudaMemcpyDeviceToHost, h->stream));
            h->sync();
        }
    }
    API_CATCH
}

/* ---- sc16 sample format on the host side of a batch (include/gfdm_b200.h) ---- */
static void check_sc16(float scale, const void* iq, int n)
{
    if (n < 0) throw std::invalid_argument("n_frames MUST NOT be negative");
    if (!(scale > 0.f) || !std::isfinite(scale)) throw std::invalid_argument("sc16 scale MUST be positive and finite");
    if (reinterpret_cast<uintptr_t>(iq) & 3u) throw std::invalid_argument("sc16 arrays MUST be aligned to a whole I/Q pair");
}
// device-pointer form: the 16-byte / 8-byte vector accesses of the conversion kernels
static void check_sc16_device(const void* iq)
{
    if (reinterpret_cast<uintptr_t>(iq) & 7u) throw std::invalid_argument("sc16 device arrays MUST be 8-byte aligned");
}
int gfdm_modulator_work_batch_sc16(gfdm_modulator* h, short* out, const gfdm_complex* in, float scale, int n, int mem)
{
    API_TRY
    h->use();
    check_sc16(scale, out, n);
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](void* dout, const void* d0, const void*, size_t, size_t nf) {
        h->work_fmt.ensure(nf * N * sizeof(cpx));
        modulator_run(h, h->work_fmt.as<cpx>(), static_cast<const cpx*>(d0), nf);
        launch_cf32_to_sc16(static_cast<short*>(dout), h->work_fmt.as<cpx>(), scale, nf * N, h->stream);
        h->launches += nf ? 1 : 0;
    };
    if (mem == GFDM_MEM_DEVICE) {
        check_sc16_device(out);
        run(out, in, nullptr, 0, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk((sizeof(cpx) + 4) * N, (size_t)n), in, sizeof(cpx) * N, nullptr, 0, out, 4 * N,
                            false, st.async, run);
    }
    API_CATCH
}
int gfdm_modulator_work_chunks_batch_sc16(gfdm_modulator* h, const gfdm_symbol_mapper* sm, short* out, const unsigned char* chunks,
                                          float scale, int n, int mem)
{
    API_TRY
    h->use();
    check_sc16(scale, out, n);
    check_same_device(h, sm);
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](void* dout, const void* d0, const void*, size_t, size_t nf) {
        h->work_fmt.ensure(nf * N * sizeof(cpx));
        modulator_run_chunks(h, sm, h->work_fmt.as<cpx>(), static_cast<const unsigned char*>(d0), nf);
        launch_cf32_to_sc16(static_cast<short*>(dout), h->work_fmt.as<cpx>(), scale, nf * N, h->stream);
        h->launches += nf ? 1 : 0;
    };
    if (mem == GFDM_MEM_DEVICE) {
        check_sc16_device(out);
        run(out, chunks, nullptr, 0, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(5 * N, (size_t)n), chunks, N, nullptr, 0, out, 4 * N, false, st.async, run);
    }
    API_CATCH
}
int gfdm_receiver_work_batch_sc16(gfdm_receiver* h, gfdm_complex* out, const short* in, const gfdm_complex* eq, float scale, int n,
                                  int mem)
{
    API_TRY
    h->use();
    check_sc16(scale, in, n);
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](void* dout, const void* d0, const void* d1, size_t, size_t nf) {
        h->work_fmt.ensure(nf * N * sizeof(cpx));
        launch_sc16_to_cf32(h->work_fmt.as<cpx>(), static_cast<const short*>(d0), scale, nf * N, h->stream);
        h->launches += nf ? 1 : 0;
        receiver_run(h, static_cast<cpx*>(dout), h->work_fmt.as<cpx>(), static_cast<const cpx*>(d1), nf);
    };
    if (mem == GFDM_MEM_DEVICE) {
        check_sc16_device(in);
        run(out, in, eq, 0, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(((eq ? 2 : 1) * sizeof(cpx) + 4) * N, (size_t)n), in, 4 * N, eq, sizeof(cpx) * N,
                            out, sizeof(cpx) * N, false, st.async, run);
    }
    API_CATCH
}
int gfdm_receiver_work_decide_batch_sc16(gfdm_receiver* h, const gfdm_symbol_mapper* sm, unsigned char* chunks_out, const short* in,
                                         const gfdm_complex* eq, float scale, int n, int mem)
{
    API_TRY
    h->use();
    check_sc16(scale, in, n);
    check_same_device(h, sm);
    Staging st(h, mem, true);
    const size_t N = h->N;
    auto run = [&](void* dout, const void* d0, const void* d1, size_t, size_t nf) {
        h->work_fmt.ensure(nf * N * sizeof(cpx));
        launch_sc16_to_cf32(h->work_fmt.as<cpx>(), static_cast<const short*>(d0), scale, nf * N, h->stream);
        h->launches += nf ? 1 : 0;
        receiver_run_decide(h, sm, static_cast<unsigned char*>(dout), h->work_fmt.as<cpx>(), static_cast<const cpx*>(d1), nf);
    };
    if (mem == GFDM_MEM_DEVICE) {
        check_sc16_device(in);
        run(chunks_out, in, eq, 0, (size_t)n);
    } else {
        host_pipeline_bytes(h, (size_t)n, pipe_chunk(((eq ? sizeof(cpx) : 0) + 5) * N, (size_t)n), in, 4 * N, eq, sizeof(cpx) * N,
                            chunks_out, N, false, st.async, run);
    }
    API_CATCH
}

} // extern "C"
