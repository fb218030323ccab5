// mixed-radix kernels, float, radices up to 16 (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {
template <> MixedKernelFn mixed_kernel_inst<float, 16>(bool rowmap, int io) {
    if (!rowmap) return io == MIXED_TW ? fft_mixed_kernel<float, false, 16, MIXED_TW> : fft_mixed_kernel<float, false, 16, MIXED_C2C>;
    switch (io) {
        case MIXED_R2C: return fft_mixed_kernel<float, true, 16, MIXED_R2C>;
        case MIXED_C2R: return fft_mixed_kernel<float, true, 16, MIXED_C2R>;
        case MIXED_R2C_HALF: return fft_mixed_kernel<float, true, 16, MIXED_R2C_HALF>;
        case MIXED_C2R_HALF: return fft_mixed_kernel<float, true, 16, MIXED_C2R_HALF>;
        case MIXED_RC: return fft_mixed_kernel<float, true, 16, MIXED_RC>;
        default: return fft_mixed_kernel<float, true, 16, MIXED_C2C>;
    }
}
}  // namespace fftb200
