// instantiations of the fused two-axis pass (tile_kernel.cuh: fft_fused_ab_kernel) for equal-shaped
// passes: the same (L, R, W) rows the tile tables use for these lengths
#include "tile_registry.h"

namespace fftb200 {

#define FROW(T_, L_, R_, W_)                                                                                  \
    { fft_fused_ab_kernel<T_, L_, R_, W_, L_, R_, W_>, L_, L_, TileTraits<T_, L_, R_, W_, V_RR>::THREADS,        \
      TileTraits<T_, L_, R_, W_, V_RR>::SMEM_BYTES, TileTraits<T_, L_, R_, W_, V_RR>::MIN_CTAS, R_, W_, R_, W_ }

static const FusedKernelInfo k_fused_f64[] = {FROW(double, 128, 8, 8), FROW(double, 256, 8, 8), FROW(double, 512, 8, 8)};
static const FusedKernelInfo k_fused_f32[] = {FROW(float, 128, 16, 16), FROW(float, 256, 16, 16), FROW(float, 512, 16, 16)};

const FusedKernelInfo *find_fused_kernel(int prec, int LA, int LB) {
    const FusedKernelInfo *t = prec ? k_fused_f64 : k_fused_f32;
    const int n = 3;
    for (int i = 0; i < n; ++i)
        if (t[i].LA == LA && t[i].LB == LB) return &t[i];
    return nullptr;
}

}  // namespace fftb200
