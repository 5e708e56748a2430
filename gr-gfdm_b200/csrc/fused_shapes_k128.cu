// fused_shapes_k128.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k128()
{
    return {
        GFDM_SHAPE(3, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in smem, PR=3, 12032 B smem, regs 96/128
        GFDM_SHAPE(5, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in smem, PR=5, 18304 B smem, regs 96/128
        GFDM_SHAPE(7, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in smem, PR=7, 24576 B smem, regs 96/128
        GFDM_SHAPE(9, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in smem, PR=9, 30848 B smem, regs 96/128
        GFDM_SHAPE(15, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in smem, PR=15, 49664 B smem, regs 96/128
        GFDM_SHAPE(21, 8, 16, 128, 1, 4), // K=128: 1 frame(s) per pass, table in tmem, PR=21, 45952 B smem, regs 100/128
    };
}

} // namespace gfdm
