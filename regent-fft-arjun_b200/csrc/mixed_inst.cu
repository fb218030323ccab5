// mixed_inst.cu — the instantiations of the mixed-radix shared-memory kernel (mixed_kernel.cuh)
#include "mixed_kernel.cuh"

namespace fftb200 {

template <typename T> static MixedKernelFn pick(bool rowmap, int maxr) {
    if (maxr <= 8) return rowmap ? fft_mixed_kernel<T, true, 8> : fft_mixed_kernel<T, false, 8>;
    if (maxr <= 10) return rowmap ? fft_mixed_kernel<T, true, 10> : fft_mixed_kernel<T, false, 10>;
    return rowmap ? fft_mixed_kernel<T, true, 16> : fft_mixed_kernel<T, false, 16>;
}

MixedKernelFn mixed_kernel(int prec, bool rowmap, int maxr) { return prec ? pick<double>(rowmap, maxr) : pick<float>(rowmap, maxr); }

}  // namespace fftb200
