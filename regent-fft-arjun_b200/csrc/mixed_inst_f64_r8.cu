// mixed-radix kernels, double, radices up to 8 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<double, 8>(bool rowmap) {
    return rowmap ? fft_mixed_kernel<double, true, 8> : fft_mixed_kernel<double, false, 8>;
}
}  // namespace fftb200
