// Decode GEMV (one token) for layers whose main codebook is too large for one SM's shared memory
// (K = 65536 entries of 16 bytes = 1 MiB), sliced-codebook variant.
//
// The generic kernel (gemv_kernel.cuh) gathers such a codebook through L1/L2 and is bound by the
// L1TEX tag stage at ~1.1 gathers per clock per SM (tools/gather_microbench.cu) -- about 14 % of the
// HBM roofline whatever else is done.  Shared memory sustains 3-6 random 16-byte reads per clock, so
// this kernel turns every gather into a shared-memory access:
//   * the codebook is cut into NS = K/8192 slices of 128 KiB; a thread-block cluster of NS CTAs
//     covers a contiguous range of index rows and CTA s stages slice s ONCE (1-D TMA bulk copies);
//   * at load time the indices of every row are re-bucketed by slice (vptq_linear_desc::sliced_*,
//     built by vptq_b200.native): list (s, r) holds, for the fields of row r that fall into slice s,
//     the 13 low index bits, the quantised column and the residual index.  Because a sum does not
//     care about the order of its terms, the builder also orders every list so that 8 consecutive
//     entries touch 8 different 16-byte bank groups: the 128-bit gathers are (nearly) conflict-free;
//   * the lists (s, r0..r1) of a CTA are contiguous in memory; the CTA cuts that range into 16 equal
//     runs of 32-entry steps, one per warp, whatever the row boundaries are.  A warp streams its run
//     through a private TMA ring in fixed-size stages, looks x'[column] up in a shared-memory copy
//     of the whole x' row, accumulates in fp32 registers and, whenever its run crosses into the next
//     row, parks the finished row piece in shared memory (pieces are summed in warp order);
//   * the NS partial sums of a row meet in the shared memory of ONE CTA of the cluster (row i of
//     the range is owned by CTA i mod NS): st.async + mbarrier complete_tx, as in the generic
//     kernel, but with the epilogue spread over all CTAs.  The owner sums the slices in order
//     (deterministic) and writes y.
// Mathematics and reference citations: gemv_kernel.cuh (the reference's kernel is
// csrc/kernels/quant_gemv.cuh:11-186; nothing of its structure is used here).
#include <algorithm>
#include <cstring>
#include <type_traits>

#include "gemv_kernel.cuh"

namespace vptq_b200 {

namespace {

constexpr int kSliceEntries = 8192;             // main-codebook entries per slice
constexpr int kSliceBytes = kSliceEntries * 16;  // 128 KiB
constexpr int kSlicedWarps = 16, kSlicedThreads = kSlicedWarps * 32;
constexpr int kSPS = 4;         // steps (32 entries each) per ring stage
constexpr int kColBatch = 16;     // columns per thread whose loads are in flight together (x' prologue)
constexpr int kMaxRowsCta = 256;  // rows of one cluster (bounds the offset / partial-sum tables)

struct SlicedLayer {
  const uint8_t* stream;    // step records: 32 entry words (+ 32 residual-index bytes)
  const uint32_t* offsets;  // [NS*Ro + 1], in steps
  const void* centroids;
  const void* res_centroids;
  const uint16_t* perm;
  const void* scale_q;  // quantised column order (or nullptr: no scale / bias)
  const void* wbias_q;
  const void* bias;
  void* y;
  int Cq, O, Ro, Kr;
  int ncl;        // clusters (CTA groups) working on this layer
  int part_row0;  // global-reduce variant: first row of this layer in the partial-sum table
};

struct SlicedParams {
  int n, ns;                                  // fused layers, slices (= cluster size)
  uint32_t grid_begin[kMaxFusedLayers + 1];   // layer l owns blocks [grid_begin[l], grid_begin[l+1])
  const void* x;
  SlicedLayer layer[kMaxFusedLayers];
  // shared-memory carve-up (bytes)
  uint32_t off_bars, off_offs, off_red, off_slice, off_res, off_x, off_recv, off_wsum, off_ring;
  int res_rep, stages;
  uint32_t stage_bytes;
  unsigned long long* prof;  // developer aid: %globaltimer stamps of the first / last CTA (or nullptr)
  // global-reduce variant (CLUSTER = false): the NS CTAs of a group are ordinary CTAs; their per-row sums
  // meet in `part` [row][slice][8] and the last CTA of the group to arrive (counter) writes y
  float* part;
  uint32_t* counters;  // one per group (blockIdx.x / ns), zero at rest
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}

// sm_100 mixed-precision scalar ops (SASS FHADD / FHFMA): a 16-bit operand is widened inside the
// instruction, so c + r costs one conversion and one add per element instead of two and one; both
// forms round exactly like convert-then-FADD, i.e. like the generic kernel's fma_entry.
template <typename T>
__device__ __forceinline__ float add_f32_16(uint16_t a, float c);
template <>
__device__ __forceinline__ float add_f32_16<__half>(uint16_t a, float c) {
  float r;
  asm("add.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(a), "f"(c));
  return r;
}
template <>
__device__ __forceinline__ float add_f32_16<__nv_bfloat16>(uint16_t a, float c) {
  float r;
  asm("add.f32.bf16 %0, %1, %2;" : "=f"(r) : "h"(a), "f"(c));
  return r;
}

// acc[e] += xv * (c[e] + r[e]), fp32
template <typename T, bool RES>
__device__ __forceinline__ void fma_entry8(float (&acc)[8], float xv, const uint32_t (&cw)[4], const uint32_t (&rw)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (RES) {
      const float2 r = DT<T>::unpack2(rw[i]);
      acc[2 * i] = fmaf(xv, add_f32_16<T>(uint16_t(cw[i] & 0xffffu), r.x), acc[2 * i]);
      acc[2 * i + 1] = fmaf(xv, add_f32_16<T>(uint16_t(cw[i] >> 16), r.y), acc[2 * i + 1]);
    } else {
      const float2 c = DT<T>::unpack2(cw[i]);
      acc[2 * i] = fmaf(xv, c.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(xv, c.y, acc[2 * i + 1]);
    }
  }
}

template <typename T, bool RES, bool CLUSTER>
__global__ void __launch_bounds__(kSlicedThreads, 1) gemv_sliced_kernel(const __grid_constant__ SlicedParams mp) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t REC = RES ? 160u : 128u;  // bytes per step record
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxFusedLayers; ++i)
    if (i < mp.n && blockIdx.x >= mp.grid_begin[i]) l = i;
  const SlicedLayer& L = mp.layer[l];
  const uint32_t bidx = blockIdx.x - mp.grid_begin[l];
  const int ns = mp.ns;
  const int s = int(bidx % uint32_t(ns));  // == %cluster_ctarank (grid_begin[] are multiples of ns)
  const int q = int(bidx / uint32_t(ns));
  const int r0 = int(int64_t(L.Ro) * q / L.ncl), r1 = int(int64_t(L.Ro) * (q + 1) / L.ncl);
  const int nrows = r1 - r0;                    // index rows of this cluster
  const int nown_max = (nrows + ns - 1) / ns;   // rows owned (reduced + written) per CTA, at most
  const int nown = nrows > s ? (nrows - s + ns - 1) / ns : 0;
  const int Cq = L.Cq, stages = mp.stages;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + mp.off_bars);
  uint32_t* s_offs = reinterpret_cast<uint32_t*>(smem + mp.off_offs);  // [nrows + 1]
  float* s_red = reinterpret_cast<float*>(smem + mp.off_red);           // [warps + 1]
  uint8_t* s_slice = smem + mp.off_slice;
  uint8_t* s_res = smem + mp.off_res;
  float* s_x = reinterpret_cast<float*>(smem + mp.off_x);               // [Cq + 1], slot Cq stays 0
  uint32_t* s_xw = reinterpret_cast<uint32_t*>(smem + mp.off_x);        // before x arrives: perm | scale << 16
  float* s_recv = reinterpret_cast<float*>(smem + mp.off_recv);         // [ns][nown_max][8]
  float* s_wsum = reinterpret_cast<float*>(smem + mp.off_wsum);         // [nrows + warps][8] row pieces
  uint8_t* ring = smem + mp.off_ring + size_t(warp) * stages * mp.stage_bytes;
  uint64_t* slice_bar = &bars[0];
  uint64_t* recv_bar = &bars[1];
  uint64_t* full = &bars[2 + warp * stages];

  // phase stamps (ns) of thread 0 of the first and of the last CTA: tools/profile_gemv.py --phases
  auto stamp = [&](int slot) {
    if (mp.prof && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      mp.prof[(blockIdx.x == 0 ? 0 : 16) + slot] = t;
    }
  };
  stamp(0);
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();

  // -------- barriers; the slice copy leaves at once (its issuer needs no CTA barrier) ----------------
  if (tid < 2 + kSlicedWarps * stages) {
    mbar_init(&bars[tid], 1);
    fence_mbar_init();
    if (tid == 0) {
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(slice_bar, uint32_t(kSliceBytes));
      const uint8_t* src = reinterpret_cast<const uint8_t*>(L.centroids) + size_t(s) * kSliceBytes;
      for (uint32_t off = 0; off < uint32_t(kSliceBytes); off += 32768u)
        tma_bulk_g2s(s_slice + off, src + off, 32768u, slice_bar, pol_keep);
    }
    // owner side of the reduction: every slice delivers 8 floats per owned row with st.async
    if (CLUSTER && tid == 1 && nown > 0) mbar_arrive_expect_tx(recv_bar, uint32_t(ns * nown * 8) * 4u);
  }
  // residual codebook (<= 256 entries of 16 bytes): one entry per thread, stored res_rep times below
  uint4 res_entry = make_uint4(0u, 0u, 0u, 0u);
  if constexpr (RES) {
    if (tid < L.Kr) res_entry = ldg_nc_v4(reinterpret_cast<const uint8_t*>(L.res_centroids) + tid * 16, pol_keep);
  }
  // -------- x-independent column data: list offsets of this CTA's rows; perm and scale of every
  // quantised column, parked in the x' array as (perm | scale bits << 16) until x arrives ---------
  for (int i = tid; i <= nrows; i += kSlicedThreads) s_offs[i] = L.offsets[size_t(s) * L.Ro + r0 + i];
  // row pieces of warps whose run is empty (fewer steps than warps) are read as zero by the epilogue
  for (int i = tid; i < (nrows + kSlicedWarps) * 8; i += kSlicedThreads) s_wsum[i] = 0.f;
  {
    const T* scale_q = reinterpret_cast<const T*>(L.scale_q);
    const T one = DT<T>::from_float(1.f);
    for (int c0 = tid; c0 < Cq; c0 += kColBatch * kSlicedThreads) {
      uint32_t pc[kColBatch];
      T sc[kColBatch];
#pragma unroll
      for (int k = 0; k < kColBatch; ++k) {
        const int c = c0 + k * kSlicedThreads;
        pc[k] = c < Cq ? (L.perm ? uint32_t(L.perm[c]) : uint32_t(c)) : 0u;
        sc[k] = (c < Cq && scale_q) ? scale_q[c] : one;
      }
#pragma unroll
      for (int k = 0; k < kColBatch; ++k) {
        const int c = c0 + k * kSlicedThreads;
        if (c < Cq) s_xw[c] = pc[k] | (uint32_t(*reinterpret_cast<const uint16_t*>(&sc[k])) << 16);
      }
    }
  }
  __syncthreads();
  stamp(1);
  // "this CTA runs and its barriers exist"; waited (acquire) right before the first st.async
  if constexpr (CLUSTER) cluster_arrive_relaxed();
  pdl_launch_dependents();

  // -------- this warp's run of steps: an equal share of the CTA's contiguous step range -------------
  const int T0 = int(s_offs[0]), TT = int(s_offs[nrows]) - T0;  // first step, number of steps of the CTA
  const int t_begin = T0 + int(int64_t(TT) * warp / kSlicedWarps);
  const int t_end = T0 + int(int64_t(TT) * (warp + 1) / kSlicedWarps);
  const int nstage = (t_end - t_begin + kSPS - 1) / kSPS;
  // warp-collective: start the copy of stage qi of the run into ring slot `slot` (= qi mod stages)
  auto issue = [&](int qi, int slot) {
    if (lane == 0) {  // (the slot was only READ through the generic proxy before)
      const int t = t_begin + qi * kSPS;
      const uint32_t bytes = uint32_t(min(kSPS, t_end - t)) * REC;
      mbar_arrive_expect_tx(&full[slot], bytes);
      tma_bulk_g2s(ring + size_t(slot) * mp.stage_bytes, L.stream + size_t(t) * REC, bytes, &full[slot], pol_stream);
    }
  };
  for (int qi = 0; qi < min(stages, nstage); ++qi) issue(qi, qi);

  // -------- residual codebook: copy k of entry i sits at 16-byte slot i*rep + k and lane L reads copy
  // L mod rep, so that (rep = 8) the 8 lanes of a quarter-warp always hit 8 different bank groups -----
  if constexpr (RES) {
    if (tid < L.Kr) {
      const uint32_t dst = smem_u32(s_res);
      for (int c = 0; c < mp.res_rep; ++c) sts_v4(dst + uint32_t(tid * mp.res_rep + c) * 16u, res_entry);
    }
  }

  stamp(2);
  // -------- x arrives from the previous kernel: x'[c] = x[perm c] * scale[perm c] ------------------
  pdl_wait_prior_grid();
  stamp(3);
  {
    const T* x = reinterpret_cast<const T*>(mp.x);
    const T* wbias_q = (s == 0) ? reinterpret_cast<const T*>(L.wbias_q) : nullptr;
    float bs = 0.f;  // slice 0 also forms sum_c x[perm c] * wbias[perm c], the weight_bias term of every row
    for (int c0 = tid; c0 < Cq; c0 += kColBatch * kSlicedThreads) {
      uint32_t w[kColBatch];
      T xv[kColBatch], wb[kColBatch];
#pragma unroll
      for (int k = 0; k < kColBatch; ++k) {
        const int c = c0 + k * kSlicedThreads;
        w[k] = c < Cq ? s_xw[c] : 0u;  // own slots only: no barrier since they were written
      }
#pragma unroll
      for (int k = 0; k < kColBatch; ++k) {
        const int c = c0 + k * kSlicedThreads;
        xv[k] = x[w[k] & 0xffffu];
        wb[k] = (wbias_q && c < Cq) ? wbias_q[c] : DT<T>::from_float(0.f);
      }
#pragma unroll
      for (int k = 0; k < kColBatch; ++k) {
        const int c = c0 + k * kSlicedThreads;
        if (c < Cq) {
          const uint16_t sb = uint16_t(w[k] >> 16);
          const float xf = DT<T>::to_float(xv[k]);
          s_x[c] = xf * DT<T>::to_float(*reinterpret_cast<const T*>(&sb));
          bs = fmaf(xf, DT<T>::to_float(wb[k]), bs);
        }
      }
    }
    if (tid == 0) s_x[Cq] = 0.f;  // the column of null entries
    if (s == 0) {
      const float v = warp_sum(bs);
      if (lane == 0) s_red[warp] = v;
    }
  }
  __syncthreads();
  if (s == 0 && tid == 0) {
    float v = 0.f;
    for (int w = 0; w < kSlicedWarps; ++w) v += s_red[w];
    s_red[kSlicedWarps] = v;  // read after the post-loop barrier
  }
  stamp(4);
  mbar_wait(slice_bar, 0);
  stamp(5);

  // -------- main loop ---------------------------------------------------------------------------
  if (nstage > 0) {
    const uint32_t slice_base = smem_u32(s_slice);
    const uint32_t res_lane = smem_u32(s_res) + uint32_t(lane & (mp.res_rep - 1)) * 16u;
    const uint32_t res_stride = 16u * uint32_t(mp.res_rep);
    const uint32_t x_base = smem_u32(s_x);
    // the row this run starts in: the last i with s_offs[i] <= t_begin (rows may be empty)
    int row;
    {
      int lo = 0, hi = nrows - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (int(s_offs[mid]) <= t_begin) lo = mid;
        else hi = mid - 1;
      }
      row = lo;
    }
    int row_end = int(s_offs[row + 1]);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // a piece (the part of a row that lies in this run) is complete: park its 8 sums.  Pieces are
    // numbered row + warp, which is unique and increasing along the CTA's step range.
    auto flush = [&]() {
      const float mine = warp_reduce_to_lane<8>(acc, lane);
      if (lane < 8) s_wsum[(row + warp) * 8 + lane] = mine;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    };
    int t = t_begin, slot = 0;
    uint32_t par = 0;
    // one ring stage: FULL = all kSPS steps present (every stage but the last of the run)
    auto stage_body = [&](auto full_tag, int cnt) {
      constexpr bool FULL = decltype(full_tag)::value;
      const uint32_t st = smem_u32(ring + size_t(slot) * mp.stage_bytes);
      uint32_t ent[kSPS], rix[kSPS];
      uint32_t cw[kSPS][4], rw[kSPS][4];
      float xv[kSPS];
#pragma unroll
      for (int j = 0; j < kSPS; ++j) {
        if (FULL || j < cnt) {
          ent[j] = lds_u32(st + uint32_t(j) * REC + uint32_t(lane) * 4u);
          if constexpr (RES) rix[j] = lds_u8(st + uint32_t(j) * REC + 128u + uint32_t(lane));
        }
      }
#pragma unroll
      for (int j = 0; j < kSPS; ++j) {
        if (FULL || j < cnt) {
          lds_entry<8>(cw[j], slice_base + (ent[j] & 0x1fffu) * 16u);
          if constexpr (RES) lds_entry<8>(rw[j], res_lane + rix[j] * res_stride);
          xv[j] = lds_f32(x_base + (ent[j] >> 16) * 4u);
        }
      }
#pragma unroll
      for (int j = 0; j < kSPS; ++j) {
        if (FULL || j < cnt) {
          while (t == row_end) {  // the run crosses into the next row (possibly over empty rows)
            flush();
            ++row;
            row_end = int(s_offs[row + 1]);
          }
          fma_entry8<T, RES>(acc, xv[j], cw[j], rw[j]);
          ++t;
        }
      }
    };
    for (int qi = 0; qi < nstage; ++qi) {
      const int cnt = min(kSPS, t_end - t);
      mbar_wait(&full[slot], par);
      if (cnt == kSPS) stage_body(std::true_type{}, cnt);
      else stage_body(std::false_type{}, cnt);
      __syncwarp();  // every lane has read its words of the stage: refill it
      if (qi + stages < nstage) issue(qi + stages, slot);
      if (++slot == stages) slot = 0, par ^= 1u;
    }
    flush();
  }
  stamp(6);  // warp 0 finished its run
  __syncthreads();
  stamp(7);

  if constexpr (CLUSTER) {
    // -------- reduction over the slices: row i of the range goes to CTA i mod ns ---------------------
    cluster_wait();  // every CTA of the cluster runs and has armed its barrier
    stamp(8);
    {
      const float cbias = (s == 0) ? s_red[kSlicedWarps] : 0.f;
      const uint32_t recv0 = smem_u32(s_recv), bar0 = smem_u32(recv_bar);
      for (int i = tid; i < nrows * 8; i += kSlicedThreads) {
        const int row = i >> 3, e = i & 7;
        float v = cbias;
        const int a = int(s_offs[row]) - T0, b = int(s_offs[row + 1]) - T0;  // the row's steps, relative to the CTA
        if (b > a) {
          // warp of step u: the w with floor(TT w / 16) <= u < floor(TT (w+1) / 16)
          const int fw = int((int64_t(a + 1) * kSlicedWarps - 1) / TT), lw = int((int64_t(b) * kSlicedWarps - 1) / TT);
          for (int w = fw; w <= lw; ++w) v += s_wsum[(row + w) * 8 + e];
        }
        const uint32_t owner = uint32_t(row % ns), j = uint32_t(row / ns);
        st_async_f32(mapa_shared(recv0 + ((uint32_t(s) * nown_max + j) * 8u + e) * 4u, owner), v,
                     mapa_shared(bar0, owner));
      }
    }
    if (nown > 0) {
      mbar_wait(recv_bar, 0);  // ns * nown * 8 partial sums have landed
      stamp(9);
      const T* bias = reinterpret_cast<const T*>(L.bias);
      T* y = reinterpret_cast<T*>(L.y);
      // one thread per output value: the 8 values of a row leave as 8 adjacent 2-byte stores
      for (int i = tid; i < nown * 8; i += kSlicedThreads) {
        const int j = i >> 3, e = i & 7;
        const int o = (r0 + j * ns + s) * 8 + e;
        if (o < L.O) {
          float v = bias ? DT<T>::to_float(bias[o]) : 0.f;
          for (int sl = 0; sl < ns; ++sl) v += s_recv[(sl * nown_max + j) * 8 + e];
          y[o] = DT<T>::from_float(v);
        }
      }
    }
  } else {
    // -------- reduction over the slices through global memory (L2): every CTA of the group parks its
    // per-row sums, the last one to arrive adds the NS slices in order (deterministic) and writes y ---
    const float cbias = (s == 0) ? s_red[kSlicedWarps] : 0.f;
    float* part = mp.part + size_t(L.part_row0 + r0) * ns * 8;
    for (int i = tid; i < nrows * 8; i += kSlicedThreads) {
      const int row = i >> 3, e = i & 7;
      float v = cbias;
      const int a = int(s_offs[row]) - T0, b = int(s_offs[row + 1]) - T0;
      if (b > a) {
        const int fw = int((int64_t(a + 1) * kSlicedWarps - 1) / TT), lw = int((int64_t(b) * kSlicedWarps - 1) / TT);
        for (int w = fw; w <= lw; ++w) v += s_wsum[(row + w) * 8 + e];
      }
      part[(size_t(row) * ns + s) * 8 + e] = v;
    }
    __threadfence();
    __syncthreads();
    uint32_t* flag = reinterpret_cast<uint32_t*>(s_red) + kSlicedWarps + 1;
    if (tid == 0) {
      uint32_t* ctr = mp.counters + blockIdx.x / uint32_t(ns);
      const uint32_t prev = atomicAdd(ctr, 1u);
      *flag = prev == uint32_t(ns - 1) ? 1u : 0u;
      if (prev == uint32_t(ns - 1)) *ctr = 0u;  // leave the counter zeroed for the next launch
    }
    __syncthreads();
    stamp(8);
    if (*flag) {
      __threadfence();
      const T* bias = reinterpret_cast<const T*>(L.bias);
      T* y = reinterpret_cast<T*>(L.y);
      for (int i = tid; i < nrows * 8; i += kSlicedThreads) {
        const int row = i >> 3, e = i & 7;
        const int o = (r0 + row) * 8 + e;
        if (o < L.O) {
          float v = bias ? DT<T>::to_float(bias[o]) : 0.f;
          for (int sl = 0; sl < ns; ++sl) v += ldg_cg_f32(&part[(size_t(row) * ns + sl) * 8 + e]);
          y[o] = DT<T>::from_float(v);
        }
      }
    }
  }
  stamp(10);
}

using SlicedKernelFn = void (*)(const SlicedParams);

template <typename T>
SlicedKernelFn pick_sliced_t(bool res, bool cluster) {
  if (cluster) return res ? gemv_sliced_kernel<T, true, true> : gemv_sliced_kernel<T, false, true>;
  return res ? gemv_sliced_kernel<T, true, false> : gemv_sliced_kernel<T, false, false>;
}
SlicedKernelFn pick_sliced(int dtype, bool res, bool cluster) {
  if (dtype == VPTQ_FP16) return pick_sliced_t<__half>(res, cluster);
  if (dtype == VPTQ_BF16) return pick_sliced_t<__nv_bfloat16>(res, cluster);
  return nullptr;
}

}  // namespace

bool gemv_sliced_eligible(const vptq_linear_desc& d) {
  if (!d.sliced_stream || !d.sliced_offsets) return false;
  const bool outl = d.outlier_size > 0 && d.outlier_indices != nullptr;
  if (d.vector_len != 8 || d.num_codebooks != 1 || outl) return false;
  if (d.num_centroids < 2 * kSliceEntries || d.num_centroids % kSliceEntries) return false;
  if (d.num_centroids / kSliceEntries > 8) return false;
  if (d.num_res_centroids > 256) return false;
  if (d.weight_scale && (!d.weight_scale_q || !d.weight_bias_q)) return false;
  if (d.in_features >= 65535) return false;  // column + null column in 16 bits
  if ((reinterpret_cast<uintptr_t>(d.sliced_stream) & 15u) || (reinterpret_cast<uintptr_t>(d.sliced_offsets) & 3u))
    return false;
  return true;
}

size_t gemv_sliced_workspace_bytes(const vptq_linear_desc& d) {
  if (!gemv_sliced_eligible(d)) return 0;
  return kCounterRegionBytes + size_t((d.out_features + 7) / 8) * (d.num_centroids / kSliceEntries) * 32;
}

int gemv_sliced_launch(int n, const vptq_linear_desc* const* descs, const void* x, void* const* ys, uint32_t flags,
                       cudaStream_t stream, void* workspace, size_t workspace_bytes) {
  const DeviceInfo* dev = device_info();
  if (!dev) return VPTQ_ERR_CUDA;
  if (n < 1 || n > kMaxFusedLayers) {
    set_error("gemv_sliced: 1..%d layers", kMaxFusedLayers);
    return VPTQ_ERR_UNSUPPORTED;
  }
  const vptq_linear_desc& d0 = *descs[0];
  const bool res = d0.num_res_centroids > 0;
  for (int l = 0; l < n; ++l) {
    const vptq_linear_desc& d = *descs[l];
    if (!gemv_sliced_eligible(d) || d.dtype != d0.dtype || d.in_features != d0.in_features ||
        d.num_centroids != d0.num_centroids || (d.num_res_centroids > 0) != res) {
      set_error("gemv_sliced: layer %d is not eligible / does not match layer 0", l);
      return VPTQ_ERR_UNSUPPORTED;
    }
  }
  const int ns = d0.num_centroids / kSliceEntries;
  const int Cq = d0.in_features;
  // Reduction variant.  Default: thread-block clusters (st.async into the owner's shared memory).
  // VPTQ_B200_GEMV_TUNE="sliced=2" (experimental, needs a workspace): independent CTAs + a last-arriver
  // reduction through global memory, which is not tied to the 15 co-resident 8-CTA clusters.
  size_t part_rows = 0;
  for (int l = 0; l < n; ++l) part_rows += size_t((descs[l]->out_features + 7) / 8);
  const size_t ws_need = kCounterRegionBytes + part_rows * ns * 32;
  const bool cluster = !(gemv_tune_sliced() == 2 && workspace && workspace_bytes >= ws_need);
  SlicedKernelFn fn = pick_sliced(d0.dtype, res, cluster);
  if (!fn) return VPTQ_ERR_UNSUPPORTED;
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(fn), dev->smem_optin)) return rc;

  // ---- clusters that can run at once, shared out to the layers in proportion to their rows (all
  // lists of a launch have the same expected length): minimise the largest rows-per-cluster ---------
  int avail = dev->sm_count / ns;
  if (cluster) {
    const int nmax = max_active_clusters(reinterpret_cast<const void*>(fn), ns, kSlicedThreads, 220 * 1024, dev->smem_optin);
    if (nmax > 0) avail = std::min(avail, nmax);
  }
  int Ro[kMaxFusedLayers], share[kMaxFusedLayers];
  for (int l = 0; l < n; ++l) Ro[l] = (descs[l]->out_features + 7) / 8;
  if (avail < n) {
    set_error("gemv_sliced: %d layers but only %d co-resident clusters of %d CTAs", n, avail, ns);
    return VPTQ_ERR_UNSUPPORTED;
  }
  {
    // the smallest rows-per-cluster bound R that the available clusters can honour, then the spare
    // clusters go to whichever layer is currently the slowest
    int R = 1, used = 0;
    for (;; ++R) {
      used = 0;
      for (int l = 0; l < n; ++l) used += (Ro[l] + R - 1) / R;
      if (used <= avail) break;
    }
    for (int l = 0; l < n; ++l) share[l] = (Ro[l] + R - 1) / R;
    while (used < avail) {
      int worst = -1;
      double worst_rows = 0;
      for (int l = 0; l < n; ++l) {
        const double rows = double(Ro[l]) / share[l];
        if (share[l] < Ro[l] && rows > worst_rows) worst = l, worst_rows = rows;
      }
      if (worst < 0) break;
      ++share[worst], ++used;
    }
  }
  int max_rows = 0, max_kr = 0;
  for (int l = 0; l < n; ++l) {
    max_rows = std::max(max_rows, (Ro[l] + share[l] - 1) / share[l]);
    max_kr = std::max(max_kr, descs[l]->num_res_centroids > 0 ? descs[l]->num_res_centroids : 0);
  }
  if (max_rows > kMaxRowsCta) {
    set_error("gemv_sliced: %d rows per cluster exceed %d", max_rows, kMaxRowsCta);
    return VPTQ_ERR_UNSUPPORTED;
  }

  // ---- shared-memory carve-up --------------------------------------------------------------------
  SlicedParams mp{};
  const size_t limit = size_t(dev->smem_optin);
  auto carve = [&](int res_rep, int stages) -> size_t {
    size_t off = 0;
    mp.off_bars = uint32_t(off), off += align_up(size_t(2 + kSlicedWarps * stages) * 8, 128);
    mp.off_offs = uint32_t(off), off += align_up(size_t(max_rows + 1) * 4, 128);
    mp.off_red = uint32_t(off), off += 128;
    mp.off_slice = uint32_t(off), off += kSliceBytes;
    mp.off_res = uint32_t(off), off += align_up(size_t(max_kr) * 16 * res_rep, 128);
    mp.off_x = uint32_t(off), off += align_up(size_t(Cq + 1) * 4, 128);
    mp.off_recv = uint32_t(off), off += align_up(size_t(ns) * ((max_rows + ns - 1) / ns) * 32, 128);
    mp.off_wsum = uint32_t(off), off += align_up(size_t(max_rows + kSlicedWarps) * 32, 128);
    mp.stage_bytes = uint32_t(kSPS * (res ? 160 : 128));
    mp.off_ring = uint32_t(off), off += size_t(kSlicedWarps) * stages * mp.stage_bytes;
    mp.res_rep = res_rep, mp.stages = stages;
    return off;
  };
  // what to shed, in order, until the layout fits next to the 128 KiB slice and the x' row
  struct Shape { int rep, stages; };
  const Shape shapes[] = {{8, 4}, {8, 3}, {4, 3}, {8, 2}, {4, 2}, {2, 2}, {1, 2}};
  size_t need = 0;
  bool placed = false;
  for (const Shape& sh : shapes) {
    need = carve(res ? sh.rep : 1, sh.stages);
    if (need <= limit) {
      placed = true;
      break;
    }
  }
  if (!placed) {
    set_error("gemv_sliced: no shared-memory layout fits (%zu bytes needed, in_features %d)", need, Cq);
    return VPTQ_ERR_UNSUPPORTED;
  }

  mp.n = n, mp.ns = ns, mp.x = x, mp.prof = gemv_profile_buffer();
  uint32_t begin = 0;
  int part_row0 = 0;
  if (!cluster) {
    mp.counters = reinterpret_cast<uint32_t*>(workspace);
    mp.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCounterRegionBytes);
  }
  for (int l = 0; l < n; ++l) {
    const vptq_linear_desc& d = *descs[l];
    SlicedLayer& L = mp.layer[l];
    L.part_row0 = part_row0;
    part_row0 += Ro[l];
    L.stream = reinterpret_cast<const uint8_t*>(d.sliced_stream), L.offsets = d.sliced_offsets;
    L.centroids = d.centroids, L.res_centroids = d.res_centroids;
    L.perm = d.perm, L.scale_q = d.weight_scale ? d.weight_scale_q : nullptr;
    L.wbias_q = d.weight_scale ? d.weight_bias_q : nullptr;
    L.bias = d.bias, L.y = ys[l];
    L.Cq = Cq, L.O = d.out_features, L.Ro = Ro[l], L.Kr = d.num_res_centroids > 0 ? d.num_res_centroids : 0;
    L.ncl = share[l];
    mp.grid_begin[l] = begin;
    begin += uint32_t(share[l] * ns);
  }
  for (int l = n; l <= kMaxFusedLayers; ++l) mp.grid_begin[l] = begin;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(begin);
  cfg.blockDim = dim3(unsigned(kSlicedThreads));
  cfg.dynamicSmemBytes = need;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (flags & VPTQ_FLAG_PDL) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  if (cluster) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = unsigned(ns), attr[nattr].val.clusterDim.y = 1, attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  cfg.attrs = attr, cfg.numAttrs = unsigned(nattr);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, fn, mp);
  if (e != cudaSuccess) {
    set_error("gemv_sliced launch (grid=%u smem=%zu group=%d cluster=%d): %s", begin, need, ns, int(cluster),
              cudaGetErrorString(e));
    return VPTQ_ERR_CUDA;
  }
  return 0;
}

}  // namespace vptq_b200
