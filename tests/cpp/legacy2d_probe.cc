// legacy2d_probe.cc -- runs the legacy 2-D interface of receiver_kernel_cc (filter_superposition, demodulate_subcarrier,
// serialize_output, vectorize_2d, remove_sc_interference; lib/receiver_kernel_cc.cc:130-163,194-209,227-272) through
// include/gfdm_b200.hpp on whatever library exports the C ABI, and dumps fd | td | ic (serialised [k][m]).
// usage: legacy2d_probe M K L taps.bin x.bin out.bin
#include <gfdm/receiver_kernel_cc.h>

#include <cstdio>
#include <cstdlib>

using namespace gr::gfdm;
typedef std::complex<float> cf;

int main(int argc, char** argv)
{
    if (argc < 7) return 2;
    const int M = atoi(argv[1]), K = atoi(argv[2]), L = atoi(argv[3]), N = M * K;
    std::vector<cf> taps((size_t)L * M), x(N);
    FILE* f = fopen(argv[4], "rb");
    if (!f || fread(taps.data(), sizeof(cf), taps.size(), f) != taps.size()) return 2;
    fclose(f);
    f = fopen(argv[5], "rb");
    if (!f || fread(x.data(), sizeof(cf), x.size(), f) != x.size()) return 2;
    fclose(f);
    receiver_kernel_cc dem(M, K, L, taps);
    receiver_kernel_cc::matrix fd(K, std::vector<cf>(M)), td(K, std::vector<cf>(M));
    dem.filter_superposition(fd, x.data());
    dem.demodulate_subcarrier(td, fd);
    std::vector<cf> out((size_t)3 * N);
    dem.serialize_output(out.data(), fd);
    dem.serialize_output(out.data() + N, td);
    // round trip of the two converters, then the interference step on a copy of the symbols
    receiver_kernel_cc::matrix sym(K, std::vector<cf>(M));
    dem.vectorize_2d(sym, out.data() + N);
    if (sym != td) { printf("FAIL vectorize_2d(serialize_output(td)) != td\n"); return 3; }
    dem.remove_sc_interference(sym, fd);
    dem.serialize_output(out.data() + 2 * N, sym);
    f = fopen(argv[6], "wb");
    fwrite(out.data(), sizeof(cf), out.size(), f);
    fclose(f);
    printf("OK\n");
    return 0;
}
