// fused_shapes_baseline.cu -- instantiations of the fused kernels (fused_kernels.cuh) for the four BASELINE.json shapes
// (hand-tuned in round 1; the other groups take their parameters from tools/shape_chooser.py).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_baseline()
{
    return {
        GFDM_SHAPE(5, 16, 1, 256, 1, 3),   // K=16   (BASELINE config 1)
        GFDM_SHAPE(9, 8, 8, 256, 1, 3),    // K=64   (config 2)
        GFDM_SHAPE(15, 16, 16, 256, 1, 2), // K=256  (config 4)
        GFDM_SHAPE(15, 32, 32, 512, 2, 1), // K=1024 (config 3, headline)
    };
}

#ifdef GFDM_PROFILE_STAGES
extern "C" __attribute__((visibility("default"))) int gfdm_debug_stage_cycles(unsigned long long* out32, int reset)
{
    if (out32 && cudaMemcpyFromSymbol(out32, g_stage_cycles, sizeof(g_stage_cycles)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[32] = { 0 };
        if (cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

} // namespace gfdm
