// mixed-radix kernels, float, radices up to 10 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<float, 10>(bool rowmap) {
    return rowmap ? fft_mixed_kernel<float, true, 10> : fft_mixed_kernel<float, false, 10>;
}
}  // namespace fftb200
