// Prefill path: y[T][O] = x[T][I] * W[O][I]^T + bias for many tokens, on the 5th-generation tensor
// cores (tcgen05.mma, accumulator in TMEM, operands staged in shared memory by TMA).
//
// Replaces the reference's tokens >= 3 branch (vptq/ops/quant_gemm.py:231-275: the `dequant` CUDA
// op followed by torch F.linear / cuBLAS).  The algebra is rearranged so that the tensor-core
// operand is the RAW quantised weight (no per-column scale / bias / permutation inside the GEMM):
//
//   y[t][o] = sum_c x'[t][c] * Wq[o][c]  +  rowbias[t]  +  bias[o]
//   x'[t][c]   = x[t][perm[c]] * scale[perm[c]]                 (prep kernel, one pass over x)
//   rowbias[t] = sum_f x[t][f] * wbias[f]                        (same kernel, fp32)
//   Wq[o][c]   = C[idx[o/v][c]][o%v] + R[ridx[o/v][c]][o%v]      (outlier columns from their codebook)
//
// Three kernels on the caller's stream:
//   1. prefill_prep_x     x -> x' (16 bit, quantised column order) and rowbias (fp32)
//   2. dequant (dequant.cu, quantised-order mode)  packed indices -> Wq tile source [O][Ipad]
//   3. gemm_tn_tcgen05    persistent, warp-specialised: TMA producer / single-thread MMA issuer /
//                         4 epilogue warps; 128x256x64 tiles, 4-stage smem ring (48 KB per stage), two
//                         fp32 accumulators in TMEM (2 x 256 columns) so the epilogue of one tile
//                         (adds rowbias[t] + bias[o], writes 16-bit y) overlaps the MMAs of the next.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace vptq_b200 {

// implemented in dequant.cu
int dequant_quant_order_launch(const vptq_linear_desc& d, void* wq_out, int64_t ld, cudaStream_t stream);

namespace {

constexpr int BM = 128, BN = 256, BK = 64;  // CTA tile: tokens x outputs x reduction
constexpr int UMMA_K = 16;                  // K per tcgen05.mma for 16-bit inputs
constexpr int STAGES = 4;
constexpr int TILE_A_BYTES = BM * BK * 2, TILE_B_BYTES = BN * BK * 2;
constexpr int ACC_STAGES = 2;               // accumulators in flight: epilogue of tile i overlaps the MMAs of tile i+1
constexpr int TMEM_COLS = ACC_STAGES * BN;  // fp32 accumulators: 128 lanes x (2 x 256) columns = all of TMEM
constexpr int GEMM_THREADS = 192;           // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int GEMM_SMEM = STAGES * (TILE_A_BYTES + TILE_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;

// ---------------------------------------------------------------------------------------------
// x' = x[perm] * scale[perm]  and  rowbias = x . wbias
// ---------------------------------------------------------------------------------------------
// One CTA per token: the row of x is staged in shared memory with coalesced 16-byte loads, then
// gathered from there (2-byte shared-memory reads) while perm / scale stream in coalesced, and x' is
// written with 16-byte stores.  `scale_q` is the optional quantised-order copy scale[perm[c]]
// (vptq_linear_desc.weight_scale_q); without it scale is gathered through perm.
template <typename T>
__global__ void __launch_bounds__(256) prefill_prep_x(const T* __restrict__ x, int64_t x_stride,
                                                      const uint16_t* __restrict__ perm,
                                                      const T* __restrict__ scale, const T* __restrict__ scale_q,
                                                      const T* __restrict__ wbias, T* __restrict__ xq,
                                                      int64_t xq_stride, float* __restrict__ rowbias, int I) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  T* sx = reinterpret_cast<T*>(prep_smem);
  const int t = blockIdx.x, tid = threadIdx.x;
  const T* xr = x + int64_t(t) * x_stride;
  T* out = xq + int64_t(t) * xq_stride;
  float bs = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) && ((I & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(wbias) & 15u) == 0);
  if (vec) {
    for (int i = tid * 8; i < I; i += 256 * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
      *reinterpret_cast<uint4*>(sx + i) = v;
      if (wbias) {
        const uint4 w = *reinterpret_cast<const uint4*>(wbias + i);
        const uint32_t* vp = reinterpret_cast<const uint32_t*>(&v);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(&w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = DT<T>::unpack2(vp[k]), c = DT<T>::unpack2(wp[k]);
          bs = fmaf(a.x, c.x, fmaf(a.y, c.y, bs));
        }
      }
    }
  } else {
    for (int i = tid; i < I; i += 256) {
      sx[i] = xr[i];
      if (wbias) bs = fmaf(DT<T>::to_float(xr[i]), DT<T>::to_float(wbias[i]), bs);
    }
  }
  __syncthreads();
  const bool vec_out = ((I & 7) == 0) && (!perm || (reinterpret_cast<uintptr_t>(perm) & 15u) == 0) &&
                       (!scale_q || (reinterpret_cast<uintptr_t>(scale_q) & 15u) == 0) && ((xq_stride & 7) == 0);
  if (vec_out) {
    for (int c = tid * 8; c < I; c += 256 * 8) {
      uint32_t pc[8];
      if (perm) {
        const uint4 pv = *reinterpret_cast<const uint4*>(perm + c);
        const uint32_t* pp = reinterpret_cast<const uint32_t*>(&pv);
#pragma unroll
        for (int k = 0; k < 4; ++k) pc[2 * k] = pp[k] & 0xffffu, pc[2 * k + 1] = pp[k] >> 16;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) pc[k] = uint32_t(c + k);
      }
      float sc[8];
      if (scale_q) {
        const uint4 sv = *reinterpret_cast<const uint4*>(scale_q + c);
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f2 = DT<T>::unpack2(sp[k]);
          sc[2 * k] = f2.x, sc[2 * k + 1] = f2.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) sc[k] = scale ? DT<T>::to_float(scale[pc[k]]) : 1.f;
      }
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k] = DT<T>::pack2(DT<T>::to_float(sx[pc[2 * k]]) * sc[2 * k], DT<T>::to_float(sx[pc[2 * k + 1]]) * sc[2 * k + 1]);
      *reinterpret_cast<uint4*>(out + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    for (int c = tid; c < I; c += 256) {
      const int f = perm ? int(perm[c]) : c;
      const float sc = scale_q ? DT<T>::to_float(scale_q[c]) : (scale ? DT<T>::to_float(scale[f]) : 1.f);
      out[c] = DT<T>::from_float(DT<T>::to_float(sx[f]) * sc);
    }
  }
  for (int c = I + tid; c < xq_stride; c += 256) out[c] = DT<T>::from_float(0.f);  // K padding
  __shared__ float red[8];
  bs = warp_sum(bs);
  if ((tid & 31) == 0) red[tid >> 5] = bs;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    rowbias[t] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMA PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (lane i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major tile whose rows are 128 bytes (64 x 16 bit), 128B swizzle
// (8-row x 128-byte atoms, 1024 bytes apart), sm_100 descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);        // bits [0,14)  start address >> 4
  d |= uint64_t(0) << 16;                        // bits [16,30) leading-dim byte offset (unused: K-major, swizzled)
  d |= uint64_t(1024 >> 4) << 32;                // bits [32,46) stride-dim byte offset: next 8-row group
  d |= uint64_t(1) << 46;                        // bits [46,48) descriptor version (Blackwell)
  d |= uint64_t(2) << 61;                        // bits [61,64) layout: SWIZZLE_128B
  return d;
}

// instruction descriptor, kind::f16: fp32 accumulate, A and B both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc(int is_bf16, int m, int n) {
  return (1u << 4)                        // D format: f32
         | (uint32_t(is_bf16) << 7)       // A format: 0 f16, 1 bf16
         | (uint32_t(is_bf16) << 10)      // B format
         | (0u << 15) | (0u << 16)        // A, B major: K
         | (uint32_t(n >> 3) << 17)       // N / 8
         | (uint32_t(m >> 4) << 24);      // M / 16
}

struct GemmParams {
  const void* bias;       // [O] 16 bit or nullptr
  const float* rowbias;   // [T]
  void* y;
  int64_t y_stride;
  int T, O, K;            // K = padded reduction length (multiple of 8)
  int is_bf16;
};

template <typename T>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + STAGES * TILE_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (TILE_A_BYTES + TILE_B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;        // MMA -> epilogue: accumulator a is complete
  uint64_t* acc_empty = acc_full + ACC_STAGES;  // epilogue -> MMA: accumulator a has been drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + BK - 1) / BK;
  // persistent CTAs walk the tile list with stride gridDim.x; tokens (m) vary fastest so that the
  // CTAs running at the same time share the same weight (B) tiles in L2
  const int tiles_m = (p.T + BM - 1) / BM, tiles_n = (p.O + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < ACC_STAGES; ++a) mbar_init(&acc_full[a], 1), mbar_init(&acc_empty[a], 4);  // 4 epilogue warps
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one lane) =====
    if (lane == 0) {
      int it = 0;  // running k-block counter across tiles: slot = it % STAGES, round = it / STAGES
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);  // slot free (passes immediately in round 0)
          mbar_arrive_expect_tx(&full[s], TILE_A_BYTES + TILE_B_BYTES);
          tma_load_2d(sa + s * TILE_A_BYTES, &map_a, kb * BK, m0, &full[s]);
          tma_load_2d(sb + s * TILE_B_BYTES, &map_b, kb * BK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane issues for the whole CTA) =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(p.is_bf16, BM, BN);
      int it = 0, nt = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
        const int a = nt % ACC_STAGES;
        mbar_wait(&acc_empty[a], ((nt / ACC_STAGES) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(a * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);  // TMA bytes have landed
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sa + s * TILE_A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sb + s * TILE_B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in 16-byte units
            umma_f16_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees this smem slot once the MMAs above have read it
        }
        umma_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+rowbias +bias) -> 16-bit y =====
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const T* bias = reinterpret_cast<const T*>(p.bias);
    int nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
      const int a = nt % ACC_STAGES;
      mbar_wait(&acc_full[a], (nt / ACC_STAGES) & 1);
      tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
      const float rb = (p.rowbias && m < p.T) ? p.rowbias[m] : 0.f;
      T* yrow = reinterpret_cast<T*>(p.y) + int64_t(m) * p.y_stride;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(a * BN + c0), r);
        tmem_ld_wait();
        if (m < p.T) {
          const int nb = n0 + c0;
          if (nb + 32 <= p.O && (p.y_stride & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int n = nb + j + 2 * i;
                const float b0 = bias ? DT<T>::to_float(bias[n]) : 0.f, b1 = bias ? DT<T>::to_float(bias[n + 1]) : 0.f;
                w[i] = DT<T>::pack2(__uint_as_float(r[j + 2 * i]) + rb + b0, __uint_as_float(r[j + 2 * i + 1]) + rb + b1);
              }
              *reinterpret_cast<uint4*>(yrow + nb + j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          } else {
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              if (n < p.O) yrow[n] = DT<T>::from_float(__uint_as_float(r[j]) + rb + (bias ? DT<T>::to_float(bias[n]) : 0.f));
            }
          }
        }
      }
      // this warp is done reading accumulator a: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// row-major [rows][cols] 16-bit matrix, row pitch `ld` elements; box = 64 columns x box_rows rows, 128B swizzle
int make_map(CUtensorMap* m, int is_bf16, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) {
    set_error("quant_gemm: cuTensorMapEncodeTiled not available from the driver");
    return VPTQ_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {BK, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("quant_gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", int(r), (long long)rows,
              (long long)cols, (long long)ld);
    return VPTQ_ERR_CUDA;
  }
  return 0;
}

struct GemmWorkspace {
  size_t off_rowbias, off_xq, off_wq, total;
  int64_t kpad;
};
GemmWorkspace gemm_layout(const vptq_linear_desc& d, int tokens) {
  GemmWorkspace w;
  w.kpad = int64_t(align_up(size_t(d.in_features), 64));  // whole BK blocks: no partially filled swizzle rows
  size_t off = kZeroRegionBytes;                         // the zero-at-rest region stays untouched
  w.off_rowbias = off;
  off += align_up(size_t(tokens) * 4, 1024);
  w.off_xq = off;
  off += align_up(size_t(tokens) * w.kpad * 2, 1024);
  w.off_wq = off;
  off += align_up(size_t(d.out_features) * w.kpad * 2, 1024);
  w.total = off;
  return w;
}

}  // namespace

size_t gemm_workspace_bytes(const vptq_linear_desc& d, int tokens) { return gemm_layout(d, tokens).total; }

int gemm_launch(const vptq_linear_desc& d, const void* x, int64_t x_stride, void* y, int64_t y_stride, int tokens,
                void* workspace, size_t workspace_bytes, uint32_t /*flags*/, cudaStream_t stream) {
  const GemmWorkspace w = gemm_layout(d, tokens);
  if (!workspace || workspace_bytes < w.total) {
    set_error("quant_gemm: workspace %zu bytes < required %zu", workspace_bytes, w.total);
    return VPTQ_ERR_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* rowbias = reinterpret_cast<float*>(ws + w.off_rowbias);
  void* xq = ws + w.off_xq;
  void* wq = ws + w.off_wq;
  const int is_bf16 = d.dtype == VPTQ_BF16;

  // Prep-free path: W in ORIGINAL column order with scale and weight_bias folded in (what the reference's dequant
  // returns, csrc/kernels/dequant.cuh:9-115) through the 8-columns-per-thread dequant, and x handed to TMA as it
  // is -- no x' pass over the tokens, no rowbias.  Needs whole BK blocks (in_features % 64 == 0) and TMA-able x rows.
  const bool direct = std::getenv("VPTQ_B200_GEMM_PREP") == nullptr && int64_t(d.in_features) == w.kpad &&
                      (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (x_stride % 8) == 0 &&
                      dequant_orig_fast_ok(d, ws + kZeroRegionBytes + align_up(size_t(d.in_features) * 2, 1024), w.kpad);
  if (direct) {
    void* wo = ws + kZeroRegionBytes + align_up(size_t(d.in_features) * 2, 1024);  // behind the inverse permutation
    if (int rc = dequant_launch(d, wo, workspace, workspace_bytes, stream, w.kpad)) return rc;
    CUtensorMap map_a, map_b;
    if (int rc = make_map(&map_a, is_bf16, x, tokens, w.kpad, x_stride, BM)) return rc;
    if (int rc = make_map(&map_b, is_bf16, wo, d.out_features, w.kpad, w.kpad, BN)) return rc;
    GemmParams p{};
    p.bias = d.bias, p.rowbias = nullptr, p.y = y, p.y_stride = y_stride;
    p.T = tokens, p.O = d.out_features, p.K = int(w.kpad), p.is_bf16 = is_bf16;
    const DeviceInfo* dev = device_info();
    if (!dev) return VPTQ_ERR_CUDA;
    const int ntiles = ((d.out_features + BN - 1) / BN) * ((tokens + BM - 1) / BM);
    dim3 grid(unsigned(std::min(ntiles, dev->sm_count)));
    if (is_bf16) {
      if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(gemm_tn_tcgen05<__nv_bfloat16>), GEMM_SMEM)) return rc;
      gemm_tn_tcgen05<__nv_bfloat16><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(map_a, map_b, p);
    } else {
      if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(gemm_tn_tcgen05<__half>), GEMM_SMEM)) return rc;
      gemm_tn_tcgen05<__half><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(map_a, map_b, p);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("quant_gemm launch: %s", cudaGetErrorString(e));
      return VPTQ_ERR_CUDA;
    }
    return 0;
  }

  // 1. x' and rowbias
  const size_t prep_smem = align_up(size_t(d.in_features) * 2, 16);
  {
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(prefill_prep_x<__half>), 132 * 1024)) return rc;
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(prefill_prep_x<__nv_bfloat16>), 132 * 1024)) return rc;
  }
  if (is_bf16)
    prefill_prep_x<__nv_bfloat16><<<tokens, 256, prep_smem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), x_stride, d.perm,
        reinterpret_cast<const __nv_bfloat16*>(d.weight_scale), reinterpret_cast<const __nv_bfloat16*>(d.weight_scale_q),
        reinterpret_cast<const __nv_bfloat16*>(d.weight_bias), reinterpret_cast<__nv_bfloat16*>(xq), w.kpad, rowbias,
        d.in_features);
  else
    prefill_prep_x<__half><<<tokens, 256, prep_smem, stream>>>(
        reinterpret_cast<const __half*>(x), x_stride, d.perm, reinterpret_cast<const __half*>(d.weight_scale),
        reinterpret_cast<const __half*>(d.weight_scale_q), reinterpret_cast<const __half*>(d.weight_bias),
        reinterpret_cast<__half*>(xq), w.kpad, rowbias, d.in_features);
  // 2. Wq in quantised column order (no scale / bias / perm)
  if (int rc = dequant_quant_order_launch(d, wq, w.kpad, stream)) return rc;
  // 3. tensor-core GEMM
  CUtensorMap map_a, map_b;
  if (int rc = make_map(&map_a, is_bf16, xq, tokens, w.kpad, w.kpad, BM)) return rc;
  if (int rc = make_map(&map_b, is_bf16, wq, d.out_features, w.kpad, w.kpad, BN)) return rc;
  GemmParams p{};
  p.bias = d.bias, p.rowbias = rowbias, p.y = y, p.y_stride = y_stride;
  p.T = tokens, p.O = d.out_features, p.K = int(w.kpad), p.is_bf16 = is_bf16;
  const DeviceInfo* dev = device_info();
  if (!dev) return VPTQ_ERR_CUDA;
  const int ntiles = ((d.out_features + BN - 1) / BN) * ((tokens + BM - 1) / BM);
  dim3 grid(unsigned(std::min(ntiles, dev->sm_count)));  // persistent: one CTA per SM
  cudaError_t e;
  if (is_bf16) {
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(gemm_tn_tcgen05<__nv_bfloat16>), GEMM_SMEM)) return rc;
    gemm_tn_tcgen05<__nv_bfloat16><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(map_a, map_b, p);
  } else {
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(gemm_tn_tcgen05<__half>), GEMM_SMEM)) return rc;
    gemm_tn_tcgen05<__half><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(map_a, map_b, p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("quant_gemm launch: %s", cudaGetErrorString(e));
    return VPTQ_ERR_CUDA;
  }
  return 0;
}

}  // namespace vptq_b200
