// fused_shapes_k1024.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k1024()
{
    return {
        GFDM_SHAPE(3, 32, 32, 512, 2, 1), // K=1024: 1 frame(s) per pass, table in smem, PR=3, 84288 B smem, regs 116/128
        GFDM_SHAPE(5, 32, 32, 512, 2, 1), // K=1024: 1 frame(s) per pass, table in smem, PR=5, 133952 B smem, regs 116/128
        GFDM_SHAPE(7, 32, 32, 512, 2, 1), // K=1024: 1 frame(s) per pass, table in smem, PR=7, 183616 B smem, regs 116/128
        GFDM_SHAPE(9, 32, 32, 512, 2, 1), // K=1024: 1 frame(s) per pass, table in tmem, PR=9, 151360 B smem, regs 116/128
    };
}

} // namespace gfdm
