// mixed-radix kernels, float, radices up to 16 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<float, 16>(bool rowmap) {
    return rowmap ? fft_mixed_kernel<float, true, 16> : fft_mixed_kernel<float, false, 16>;
}
}  // namespace fftb200
