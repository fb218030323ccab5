// instantiations of the fused slab kernel (slab_fused_kernel.cuh) for cubes: all three passes use the tile shape the
// column tables use for that length
#include "slab_fused_kernel.cuh"

namespace fftb200 {

#define FUSED(T_, PREC_, L_, R_, W_)                                                                      \
    { fft_slab_fused_kernel<T_, L_, R_, W_>, PREC_, L_, R_, W_, TileTraits<T_, L_, R_, W_, V_CC>::THREADS, \
      TileTraits<T_, L_, R_, W_, V_CC>::SMEM_BYTES }

static const SlabFusedKernelInfo k_fused[] = {
    FUSED(double, 1, 128, 8, 8),  FUSED(double, 1, 256, 8, 8),   FUSED(double, 1, 512, 8, 8),   FUSED(double, 1, 1024, 16, 8),
    FUSED(float, 0, 128, 16, 16), FUSED(float, 0, 256, 16, 16),  FUSED(float, 0, 512, 16, 16),  FUSED(float, 0, 1024, 16, 16),
};

const SlabFusedKernelInfo *find_slab_fused_kernel(int prec, int L) {
    for (const SlabFusedKernelInfo &k : k_fused)
        if (k.prec == prec && k.L == L) return &k;
    return nullptr;
}

}  // namespace fftb200
