// fused_shapes_k32.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k32()
{
    return {
        GFDM_SHAPE(3, 4, 8, 256, 1, 3), // K=32: 8 frame(s) per pass, table in smem, PR=3, 15680 B smem, regs 80/85
        GFDM_SHAPE(5, 4, 8, 256, 1, 3), // K=32: 8 frame(s) per pass, table in smem, PR=5, 24896 B smem, regs 80/85
        GFDM_SHAPE(7, 4, 8, 256, 1, 3), // K=32: 8 frame(s) per pass, table in smem, PR=7, 34112 B smem, regs 80/85
        GFDM_SHAPE(9, 4, 8, 256, 1, 3), // K=32: 8 frame(s) per pass, table in smem, PR=9, 43328 B smem, regs 80/85
        GFDM_SHAPE(15, 4, 8, 128, 1, 4), // K=32: 4 frame(s) per pass, table in smem, PR=15, 38336 B smem, regs 88/128
        GFDM_SHAPE(21, 4, 8, 128, 1, 4), // K=32: 4 frame(s) per pass, table in smem, PR=21, 52928 B smem, regs 100/128
        GFDM_SHAPE(19, 4, 8, 128, 1, 4), // K=32: 4 frame(s) per pass, table in smem, PR=19, 48064 B smem, regs 96/128
    };
}

} // namespace gfdm
