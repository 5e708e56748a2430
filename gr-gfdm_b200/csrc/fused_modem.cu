// fused_modem.cu -- placeholder: no fused shapes registered yet (filled in next).
#include "fused.h"
namespace gfdm {
struct FusedImpl {};
void FusedModem::init_tx(int, int, int, const std::vector<std::complex<float>>&) { impl_ = nullptr; }
void FusedModem::init_rx(int, int, int, const std::vector<std::complex<float>>&,
                         const std::vector<std::complex<float>>&) { impl_ = nullptr; }
int FusedModem::modulate(cpx*, const cpx*, size_t, cudaStream_t) { return 0; }
int FusedModem::demodulate(cpx*, cpx*, const cpx*, const cpx*, size_t, cudaStream_t) { return 0; }
const char* FusedModem::mod_name() const { return "none"; }
const char* FusedModem::rx_name() const { return "none"; }
void FusedModem::destroy() { impl_ = nullptr; }
} // namespace gfdm
