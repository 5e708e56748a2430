// fused_shapes_k512.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k512()
{
    return {
        GFDM_SHAPE(3, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in smem, PR=3, 42944 B smem, regs 128/128
        GFDM_SHAPE(5, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in smem, PR=5, 67776 B smem, regs 128/128
        GFDM_SHAPE(7, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in smem, PR=7, 92608 B smem, regs 128/128
        GFDM_SHAPE(9, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in smem, PR=9, 117440 B smem, regs 128/128
        GFDM_SHAPE(15, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in smem, PR=15, 191936 B smem, regs 128/128
        GFDM_SHAPE(21, 16, 32, 512, 1, 1), // K=512: 1 frame(s) per pass, table in tmem, PR=21, 176320 B smem, regs 128/128
    };
}

} // namespace gfdm
