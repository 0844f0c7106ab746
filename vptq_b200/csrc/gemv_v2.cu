// The reference's second GEMV op (`quant_gemv_v2`, csrc/quant_gemv_v2.cu:25-180 +
// csrc/kernels/quant_gemv_v2.cuh:15-184) with UNPACKED indices:
//   indices u16 [Ro][I], residual_indices u8|u16 [Ro][I], one codebook, no perm, no outliers,
//   y[t][r*v+e] = sum_c x[t][c] * (scale[c] * (C[idx[r][c]][e] + R[ridx[r][c]][e]) + sbias[c]) + bias
// (layout pinned by the reference's tests/test_quant_gemv.py:86-105).
//
// The op is not reachable from VQuantLinear.forward; it is kept for surface
// parity.  One warp owns one index row; lanes stride over columns so the u16/u8 index loads are
// coalesced; codebooks are gathered through L1/L2; x' = x * scale lives in shared memory.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace vptq_b200 {

namespace {

struct V2Params {
  const void* x;
  void* y;
  const uint16_t* idx;
  const void* ridx;
  int ridx_bytes;
  const void* cent;
  const void* rcent;
  const void* scale;
  const void* sbias;
  const void* bias;
  int I, O, Ro, tokens;
};

template <typename T, int V>
__global__ void __launch_bounds__(256) gemv_v2_kernel(const __grid_constant__ V2Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* sx = reinterpret_cast<float*>(smem);  // [I]
  __shared__ float s_red[8];
  __shared__ float s_cbias;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int t = blockIdx.y;
  const T* x = reinterpret_cast<const T*>(p.x) + int64_t(t) * p.I;
  const T* scale = reinterpret_cast<const T*>(p.scale);
  const T* sbias = reinterpret_cast<const T*>(p.sbias);
  float bs = 0.f;
  for (int c = tid; c < p.I; c += blockDim.x) {
    const float xv = DT<T>::to_float(x[c]);
    sx[c] = scale ? xv * DT<T>::to_float(scale[c]) : xv;
    if (sbias) bs = fmaf(xv, DT<T>::to_float(sbias[c]), bs);
  }
  bs = warp_sum(bs);
  if (lane == 0) s_red[warp] = bs;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += s_red[w];
    s_cbias = v;
  }
  __syncthreads();

  const uint64_t pol = policy_evict_last();
  const T* cent = reinterpret_cast<const T*>(p.cent);
  const T* rcent = reinterpret_cast<const T*>(p.rcent);
  for (int r = blockIdx.x * nwarps + warp; r < p.Ro; r += gridDim.x * nwarps) {
    const uint16_t* irow = p.idx + int64_t(r) * p.I;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int c = lane; c < p.I; c += 32) {
      const uint32_t mi = irow[c];
      uint32_t cw[V / 2];
      ldg_entry<V>(cw, cent + size_t(mi) * V, pol);
      float w[V];
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        const float2 f2 = DT<T>::unpack2(cw[i]);
        w[2 * i] = f2.x, w[2 * i + 1] = f2.y;
      }
      if (rcent) {
        const uint32_t ri = p.ridx_bytes == 1 ? uint32_t(reinterpret_cast<const uint8_t*>(p.ridx)[int64_t(r) * p.I + c])
                                              : uint32_t(reinterpret_cast<const uint16_t*>(p.ridx)[int64_t(r) * p.I + c]);
        uint32_t rw[V / 2];
        ldg_entry<V>(rw, rcent + size_t(ri) * V, pol);
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
          const float2 f2 = DT<T>::unpack2(rw[i]);
          w[2 * i] += f2.x, w[2 * i + 1] += f2.y;
        }
      }
      const float xv = sx[c];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = fmaf(xv, w[e], acc[e]);
    }
    float mine = 0.f;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float v = warp_sum(acc[e]);
      if (lane == e) mine = v;
    }
    const int o = r * V + lane;
    if (lane < V && o < p.O) {
      const float bv = p.bias ? DT<T>::to_float(reinterpret_cast<const T*>(p.bias)[o]) : 0.f;
      reinterpret_cast<T*>(p.y)[int64_t(t) * p.O + o] = DT<T>::from_float(mine + s_cbias + bv);
    }
  }
}

template <typename T>
int launch(const V2Params& p, int v, int sms, cudaStream_t stream) {
  const int warps = 8;
  dim3 block(warps * 32), grid(unsigned(std::min((p.Ro + warps - 1) / warps, sms * 2)), unsigned(p.tokens));
  const size_t smem = size_t(p.I) * 4;
  switch (v) {
#define VPTQ_CASE(VV)                                                                                        \
  case VV:                                                                                                   \
    if (smem > 48 * 1024)                                                                                    \
      cudaFuncSetAttribute(gemv_v2_kernel<T, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));   \
    gemv_v2_kernel<T, VV><<<grid, block, smem, stream>>>(p);                                                 \
    break;
    VPTQ_CASE(4) VPTQ_CASE(8) VPTQ_CASE(16)
#undef VPTQ_CASE
    default:
      // the reference supports exactly these (csrc/quant_gemv_v2.cu:61-63)
      set_error("quant_gemv_v2: vector_len %d not supported (4, 8, 16)", v);
      return VPTQ_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("quant_gemv_v2 launch: %s", cudaGetErrorString(e));
    return VPTQ_ERR_CUDA;
  }
  return 0;
}

}  // namespace

size_t gemv_v2_workspace_bytes(int, int, int, int) { return 0; }

int gemv_v2_launch(const GemvV2Args& a, void*, size_t, uint32_t, cudaStream_t stream) {
  const DeviceInfo* dev = device_info();
  if (!dev) return VPTQ_ERR_CUDA;
  if (a.in_features > 16384) {
    set_error("quant_gemv_v2: in_features %d > 16384", a.in_features);
    return VPTQ_ERR_UNSUPPORTED;
  }
  V2Params p{};
  p.x = a.x, p.y = a.y, p.idx = a.indices;
  p.ridx = a.num_res_centroids > 0 ? a.residual_indices : nullptr;
  p.ridx_bytes = a.res_index_bytes;
  p.cent = a.centroids;
  p.rcent = a.num_res_centroids > 0 ? a.residual_centroids : nullptr;
  p.scale = a.scale_weights, p.sbias = a.scale_bias, p.bias = a.bias;
  p.I = a.in_features, p.O = a.out_features, p.tokens = a.tokens;
  p.Ro = (a.out_features + a.vector_len - 1) / a.vector_len;
  return a.dtype == VPTQ_FP16 ? launch<__half>(p, a.vector_len, dev->sm_count, stream)
                              : launch<__nv_bfloat16>(p, a.vector_len, dev->sm_count, stream);
}

}  // namespace vptq_b200
