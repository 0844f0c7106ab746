// Instantiations of the decode GEMV for the other vector lengths the reference dispatches
// (csrc/quant_gemv.cu:209-234: 2, 4, 6, 10, 12, 16), one token per pass.
#include "gemv_kernel.cuh"

namespace vptq_b200 {

template <typename T, int V>
static GemvKernelFn pickv(bool main_smem, bool res) {
  if (main_smem) return res ? gemv_kernel<T, V, 1, true, true> : gemv_kernel<T, V, 1, true, false>;
  return res ? gemv_kernel<T, V, 1, false, true> : gemv_kernel<T, V, 1, false, false>;
}

template <typename T>
static GemvKernelFn pickv_v(int v, bool main_smem, bool res) {
  switch (v) {
    case 2: return pickv<T, 2>(main_smem, res);
    case 4: return pickv<T, 4>(main_smem, res);
    case 6: return pickv<T, 6>(main_smem, res);
    case 10: return pickv<T, 10>(main_smem, res);
    case 12: return pickv<T, 12>(main_smem, res);
    case 16: return pickv<T, 16>(main_smem, res);
    default: return nullptr;
  }
}

GemvKernelFn gemv_kernel_vx(int dtype, int v, bool main_smem, bool res) {
  if (dtype == VPTQ_FP16) return pickv_v<__half>(v, main_smem, res);
  if (dtype == VPTQ_BF16) return pickv_v<__nv_bfloat16>(v, main_smem, res);
  return nullptr;
}

}  // namespace vptq_b200
