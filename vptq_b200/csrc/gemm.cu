// Prefill path placeholder: replaced by the tcgen05 kernel (gemm_tcgen05.cu) once it lands.
#include "kernels.h"

namespace vptq_b200 {

size_t gemm_workspace_bytes(const vptq_linear_desc&, int) { return 0; }

int gemm_launch(const vptq_linear_desc&, const void*, int64_t, void*, int64_t, int, void*, size_t, uint32_t,
                cudaStream_t) {
  set_error("quant_gemm (prefill, tcgen05) is not built into this library yet");
  return VPTQ_ERR_UNSUPPORTED;
}

}  // namespace vptq_b200
