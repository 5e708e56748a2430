// fused_shapes_k64.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k64()
{
    return {
        GFDM_SHAPE(3, 8, 8, 256, 1, 3), // K=64: 4 frame(s) per pass, table in smem, PR=3, 16704 B smem, regs 80/85
        GFDM_SHAPE(5, 8, 8, 256, 1, 3), // K=64: 4 frame(s) per pass, table in smem, PR=5, 26432 B smem, regs 80/85
        GFDM_SHAPE(7, 8, 8, 256, 1, 3), // K=64: 4 frame(s) per pass, table in smem, PR=7, 36160 B smem, regs 80/85
        GFDM_SHAPE(15, 8, 8, 128, 1, 4), // K=64: 2 frame(s) per pass, table in smem, PR=15, 42432 B smem, regs 88/128
        GFDM_SHAPE(21, 8, 8, 128, 1, 4), // K=64: 2 frame(s) per pass, table in tmem, PR=21, 47296 B smem, regs 100/128
        GFDM_SHAPE(16, 8, 8, 128, 1, 4), // K=64: 2 frame(s) per pass, table in smem, PR=16, 45120 B smem, regs 90/128
    };
}

} // namespace gfdm
