// fused_shapes_k16.cu -- instantiations of the fused kernels (fused_kernels.cuh) for one group of shapes; parameters from
// tools/shape_chooser.py (one subcarrier per thread where possible, then the largest resident thread count whose shared
// memory, tensor memory and registers fit).
#include "fused_kernels.cuh"

namespace gfdm {

std::vector<ShapeEntry> fused_shapes_k16()
{
    return {
        GFDM_SHAPE(3, 4, 4, 256, 1, 3), // K=16: 16 frame(s) per pass, table in smem, PR=3, 15936 B smem, regs 72/85
        GFDM_SHAPE(7, 4, 4, 256, 1, 3), // K=16: 16 frame(s) per pass, table in smem, PR=7, 34880 B smem, regs 72/85
        GFDM_SHAPE(9, 4, 4, 256, 1, 3), // K=16: 16 frame(s) per pass, table in smem, PR=9, 44352 B smem, regs 76/85
        GFDM_SHAPE(15, 4, 4, 128, 1, 4), // K=16: 8 frame(s) per pass, table in smem, PR=15, 38208 B smem, regs 88/128
        GFDM_SHAPE(21, 4, 4, 128, 1, 4), // K=16: 8 frame(s) per pass, table in smem, PR=21, 52800 B smem, regs 100/128
        GFDM_SHAPE(8, 4, 4, 256, 1, 3), // K=16: 16 frame(s) per pass, table in smem, PR=8, 39616 B smem, regs 74/85
        GFDM_SHAPE(16, 4, 1, 128, 1, 4), // K=4: 32 frame(s) per pass, table in smem, PR=16, 38976 B smem, regs 90/128
        GFDM_SHAPE(7, 4, 2, 256, 1, 3), // K=8: 32 frame(s) per pass, table in smem, PR=7, 37952 B smem, regs 72/85
    };
}

} // namespace gfdm
