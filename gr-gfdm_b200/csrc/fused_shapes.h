// fused_shapes.h -- host-visible description of one compiled fused-kernel shape (launchers, names, shared-memory size)
// and the argument block of the interference-cancelling / hard-decision receiver variants.  The kernels themselves are
// in fused_kernels.cuh and are instantiated per shape in fused_shapes_*.cu (several translation units, compiled in
// parallel); fused_modem.cu only sees this header.
#pragma once
#include "fused.h"

#include <string>
#include <vector>

namespace gfdm {

// successive interference cancellation resident in the receiver kernel
// (lib/advanced_receiver_kernel_cc.cc:56-123, lib/receiver_kernel_cc.cc:274-299)
struct SicArgs {
    const cpx* ic_taps;          // [M]
    const cpx* points;           // constellation points
    const unsigned char* count;  // [K] multiplicity of subcarrier k in the subcarrier map (0 = inactive)
    int n_points, rule, ic_iter, phase_comp;
    float inv_map_total;         // 1 / (map.size() * M)
    float qpsk_a;                // > 0: the constellation is gr::digital's QPSK (+-a +-ja in its index order): decisions by sign
    DecideGrid grid;             // hard-decision output (DEC): O(1) decisions on grid constellations
    size_t in_stride;            // elements between the sample frames of two consecutive frames (0: contiguous, = N)
};

typedef void (*mod_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, int, int, cudaStream_t);
typedef void (*tx_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, int, int, const TxArgs&, cudaStream_t);
typedef void (*rx_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, const cpx*, const cpx*, int, int, int, int, size_t,
                            cudaStream_t);
typedef void (*sic_launch_t)(cpx*, const cpx*, const cpx*, const cpx*, const cpx*, const cpx*, int, int, int,
                             SicArgs, cudaStream_t);

struct ShapeEntry {
    int M, K, R1, R2, T, F;
    size_t smem;
    const char* mod_name;
    const char* rx_name;
    const char* tx_name;
    tx_launch_t tx;
    const void* tx_fn;
    mod_launch_t mod;
    rx_launch_t rx;
    sic_launch_t sic; // null when the shape keeps more than one subcarrier per thread
    const void* mod_fn;
    const void* rx_fn;
    const void* sic_fn;
    // the equalising variants are kernels of their own (same launch geometry and shared memory)
    const void* rx_eq_fn = nullptr;
    const void* rxd_eq_fn = nullptr;
    const void* sic_eq_fn = nullptr;
    const char* sic_name;
    size_t sic_smem = 0; // shared memory of the cancellation variant (a companion shape when IPT > 1)
    // chunk entries: modulator / transmitter chain with byte input, receiver with hard-decision output
    tx_launch_t modc, txc;
    sic_launch_t rxd;
    const void* modc_fn;
    const void* txc_fn;
    const void* rxd_fn;
    std::string modc_name, txc_name, rxd_name;
};

// shape lists of the translation units fused_shapes_*.cu (tools/shape_chooser.py picks the parameters)
std::vector<ShapeEntry> fused_shapes_baseline(); // the four BASELINE.json shapes
std::vector<ShapeEntry> fused_shapes_k16(); // K = 4, 8, 16
std::vector<ShapeEntry> fused_shapes_k32();
std::vector<ShapeEntry> fused_shapes_k64();
std::vector<ShapeEntry> fused_shapes_k128();
std::vector<ShapeEntry> fused_shapes_k256();
std::vector<ShapeEntry> fused_shapes_k512();
std::vector<ShapeEntry> fused_shapes_k1024();

} // namespace gfdm
