// mixed-radix kernels, float, radices up to 8 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<float, 8>(bool rowmap) {
    return rowmap ? fft_mixed_kernel<float, true, 8> : fft_mixed_kernel<float, false, 8>;
}
}  // namespace fftb200
