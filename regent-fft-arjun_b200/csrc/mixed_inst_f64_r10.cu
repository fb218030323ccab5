// mixed-radix kernels, double, radices up to 10 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<double, 10>(bool rowmap, int io) {
    if (!rowmap) return fft_mixed_kernel<double, false, 10, MIXED_C2C>;
    if (io == MIXED_R2C) return fft_mixed_kernel<double, true, 10, MIXED_R2C>;
    if (io == MIXED_C2R) return fft_mixed_kernel<double, true, 10, MIXED_C2R>;
    return fft_mixed_kernel<double, true, 10, MIXED_C2C>;
}
}  // namespace fftb200
