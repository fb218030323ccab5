// mixed_inst.cu — lookup of the mixed-radix shared-memory kernels (mixed_kernel.cuh); the instantiations live in
// mixed_inst_f32_r*.cu / mixed_inst_f64_r*.cu (one translation unit per precision and largest radix)
#include "mixed_kernel.cuh"

namespace fftb200 {

template <typename T> static MixedKernelFn pick(bool rowmap, int maxr, int io) {
    if (maxr <= 8) return mixed_kernel_inst<T, 8>(rowmap, io);
    if (maxr <= 10) return mixed_kernel_inst<T, 10>(rowmap, io);
    return mixed_kernel_inst<T, 16>(rowmap, io);
}

MixedKernelFn mixed_kernel(int prec, bool rowmap, int maxr, int io) {
    if (rowmap ? io == MIXED_TW : (io != MIXED_C2C && io != MIXED_TW)) return nullptr;
    return prec ? pick<double>(rowmap, maxr, io) : pick<float>(rowmap, maxr, io);
}

}  // namespace fftb200
