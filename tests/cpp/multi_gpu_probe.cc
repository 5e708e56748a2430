// multi_gpu_probe.cc -- gr::gfdm::multi_gpu (include/gfdm_b200.hpp): one worker thread per device, each with its own
// kernel objects, a HOST batch split contiguously over them; the result must equal the single-handle result bit for
// bit.  usage: multi_gpu_probe <n_workers> <n_frames>   (devices = worker index mod device count)
#include <gfdm_b200.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace gr::gfdm;
typedef std::complex<float> cf;

struct modem { // what one worker owns
    modulator_kernel_cc mod;
    receiver_kernel_cc dem;
    modem(int M, int K, int L, const std::vector<cf>& tx, const std::vector<cf>& rx) : mod(M, K, L, tx), dem(M, K, L, rx) {}
};

int main(int argc, char** argv)
{
    const int workers = argc > 1 ? atoi(argv[1]) : 2, frames = argc > 2 ? atoi(argv[2]) : 257;
    const int M = 15, K = 256, L = 2, N = M * K;
    const int n_dev = gfdm_device_count();
    if (n_dev < 1) { printf("FAIL no device\n"); return 1; }
    std::vector<cf> tx(L * M), rx(L * M);
    for (int i = 0; i < L * M; ++i) { tx[i] = cf(1.0f + 0.05f * i, 0.01f * i); rx[i] = std::conj(tx[i]); }
    std::mt19937 rng(3);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<cf> d((size_t)frames * N), x1(d.size()), y1(d.size()), x2(d.size()), y2(d.size());
    for (auto& v : d) v = cf(nd(rng), nd(rng));
    {
        modem one(M, K, L, tx, rx);
        one.mod.generic_work_batch(x1.data(), d.data(), frames);
        one.dem.generic_work_batch(y1.data(), x1.data(), nullptr, frames);
    }
    std::vector<int> devices;
    for (int i = 0; i < workers; ++i) devices.push_back(i % n_dev);
    multi_gpu<modem> pool(devices, [&]() { return std::unique_ptr<modem>(new modem(M, K, L, tx, rx)); });
    const auto t0 = std::chrono::steady_clock::now();
    pool.for_each_shard((size_t)frames, [&](modem& m, size_t f0, size_t nf) {
        m.mod.generic_work_batch(x2.data() + f0 * N, d.data() + f0 * N, (int)nf);
        m.dem.generic_work_batch(y2.data() + f0 * N, x2.data() + f0 * N, nullptr, (int)nf);
    });
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (std::memcmp(x1.data(), x2.data(), sizeof(cf) * x1.size()) || std::memcmp(y1.data(), y2.data(), sizeof(cf) * y1.size())) {
        printf("FAIL sharded result differs from the single-handle result\n");
        return 1;
    }
    // a worker's exception reaches the caller
    bool threw = false;
    try {
        pool.for_each_shard(4, [&](modem&, size_t, size_t) { throw std::runtime_error("boom"); });
    } catch (const std::runtime_error&) {
        threw = true;
    }
    if (!threw) { printf("FAIL exception not propagated\n"); return 1; }
    // and a bad device index fails at construction
    threw = false;
    try {
        multi_gpu<modem> bad({ n_dev + 7 }, [&]() { return std::unique_ptr<modem>(new modem(M, K, L, tx, rx)); });
    } catch (const std::exception&) {
        threw = true;
    }
    if (!threw) { printf("FAIL bad device accepted\n"); return 1; }
    printf("OK %d workers on %d device(s), %d frames, %.1f ms\n", workers, n_dev, frames, dt * 1e3);
    return 0;
}
