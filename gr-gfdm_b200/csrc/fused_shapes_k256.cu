// fused_shapes_k256.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k256()
{
    return {
        GFDM_SHAPE(3, 16, 16, 256, 1, 2), // K=256: 1 frame(s) per pass, table in smem, PR=3, 22464 B smem, regs 96/128
        GFDM_SHAPE(5, 16, 16, 256, 1, 2), // K=256: 1 frame(s) per pass, table in smem, PR=5, 35008 B smem, regs 96/128
        GFDM_SHAPE(7, 16, 16, 256, 1, 2), // K=256: 1 frame(s) per pass, table in smem, PR=7, 47552 B smem, regs 96/128
        GFDM_SHAPE(9, 16, 16, 256, 1, 2), // K=256: 1 frame(s) per pass, table in smem, PR=9, 60096 B smem, regs 96/128
        GFDM_SHAPE(21, 16, 16, 256, 1, 2), // K=256: 1 frame(s) per pass, table in tmem, PR=21, 90304 B smem, regs 100/128
    };
}

} // namespace gfdm
