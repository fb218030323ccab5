// mixed-radix kernels, double, radices up to 10 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<double, 10>(bool rowmap) {
    return rowmap ? fft_mixed_kernel<double, true, 10> : fft_mixed_kernel<double, false, 10>;
}
}  // namespace fftb200
