// Host side of the decode GEMV: work decomposition (GemvPlan), shared-memory carve-up, launch.
#include <algorithm>
#include <mutex>
#include <unordered_set>

#include "gemv_kernel.cuh"

namespace vptq_b200 {

namespace {

constexpr int kMaxChunkCols = 4096;   // bounds the x' slice in shared memory (16 KB per token)
constexpr int kSegFields = 512;       // index fields per ring stage
constexpr int kSmemReserve = 2048;    // head-room below the opt-in limit

bool supported_vec_len(int v) { return v == 2 || v == 4 || v == 6 || v == 8 || v == 10 || v == 12 || v == 16; }

std::mutex g_attr_mutex;
std::unordered_set<const void*> g_attr_done;

int ensure_smem_attr(const void* fn, int bytes) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  if (g_attr_done.count(fn)) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic smem=%d): %s", bytes, cudaGetErrorString(e));
    return VPTQ_ERR_CUDA;
  }
  g_attr_done.insert(fn);
  return 0;
}

}  // namespace

int gemv_make_plan(const vptq_linear_desc& d, int tokens, const DeviceInfo& dev, GemvPlan* out) {
  GemvPlan pl{};
  const int v = d.vector_len, G = d.num_codebooks, gs = d.group_size;
  const int Ro = (d.out_features + v - 1) / v;
  const int ib = ilog2(d.num_centroids);
  const int rb = d.num_res_centroids > 0 ? ilog2(d.num_res_centroids) : 0;
  const int b = ib + rb;
  const int S = (d.outlier_size > 0 && d.outlier_indices) ? d.outlier_size : 0;
  const int EB = 2 * v;

  pl.nt = (v == 8) ? (tokens >= 4 ? 4 : (tokens >= 2 ? 2 : 1)) : 1;
  pl.seg_fields = kSegFields;
  pl.stage_bytes = uint32_t(align_up(size_t(pl.seg_fields) * b / 8 + 16, 16));

  // ---- codebook placement -------------------------------------------------------------------
  const int budget = dev.smem_optin - kSmemReserve;
  const size_t main_bytes = size_t(d.num_centroids) * EB;
  const size_t res_bytes = rb ? size_t(d.num_res_centroids) * EB : 0;
  pl.res_rep = (rb && v == 8 && res_bytes * 8 <= 32768) ? 8 : 1;
  pl.main_rep = 1;
  pl.main_in_smem = 0;
  if (main_bytes <= 131072) {
    pl.main_in_smem = 1;
    if (v == 8 && main_bytes * 8 <= 32768) pl.main_rep = 8;
  }

  // ---- column chunks and the grid -----------------------------------------------------------
  // Candidates: cpg chunks per group, chunk width a multiple of 128 columns.  Cost model: the SM
  // with the most (row, chunk) units bounds the kernel; more CTAs per chunk than rows is waste;
  // fewer rows per CTA than warps leaves the gather pipeline short of parallelism.
  const int sms = dev.sm_count;
  int best_cpg = 1;
  double best_cost = 1e300;
  for (int cpg = 1; cpg <= 64; cpg *= 2) {
    int cc = int(align_up(size_t((gs + cpg - 1) / cpg), 128));
    if (cc > kMaxChunkCols) continue;
    if (cpg > 1 && cc < 256) break;
    const int real_cpg = (gs + cc - 1) / cc;
    const int nch = G * real_cpg;
    if (nch > sms) break;
    const int cpc = std::max(1, std::min(sms / nch, Ro));
    const int rows_cta = (Ro + cpc - 1) / cpc;
    double cost = double(rows_cta) * cc;                  // fields the busiest SM streams
    cost += 600.0 + 0.35 * cc;                            // per-CTA prologue (x' gather, codebooks)
    cost += 64.0 * rows_cta;                              // per-unit epilogue (reduce, fence, atomic)
    if (rows_cta < 8) cost *= 1.0 + 0.08 * (8 - rows_cta);  // too few warps busy per SM
    if (cost < best_cost) best_cost = cost, best_cpg = real_cpg, pl.chunk_cols = cc;
  }
  if (best_cost == 1e300) {  // very wide single group: fall back to the widest legal chunk
    pl.chunk_cols = kMaxChunkCols;
    best_cpg = (gs + kMaxChunkCols - 1) / kMaxChunkCols;
  }
  pl.cpg = best_cpg;
  pl.nch = G * pl.cpg;
  if (pl.nch > sms) {
    set_error("gemv: %d column chunks exceed %d SMs (num_codebooks=%d)", pl.nch, sms, G);
    return VPTQ_ERR_UNSUPPORTED;
  }
  pl.cpc = std::max(1, std::min(sms / pl.nch, Ro));
  pl.grid = pl.nch * pl.cpc;

  // ---- shared memory carve-up ----------------------------------------------------------------
  auto carve = [&](int warps, int stages, bool main_smem) -> size_t {
    size_t off = 0;
    pl.off_bars = uint32_t(off);
    off += align_up(size_t(1 + warps * stages) * 8, 128);
    pl.off_cbias = uint32_t(off);
    off += align_up(size_t(pl.nt) * (1 + warps) * 4, 128);
    const int n_all = pl.chunk_cols + S;
    pl.sx_stride = int(align_up(size_t(n_all), 32));
    pl.off_pcol = uint32_t(off);
    off += align_up(size_t(n_all) * 2, 128);
    pl.off_wb = uint32_t(off);
    off += align_up(size_t(n_all) * 4, 128);
    pl.off_sx = uint32_t(off);
    off += align_up(size_t(pl.nt) * pl.sx_stride * 4, 128);
    pl.off_res = uint32_t(off);
    off += align_up(res_bytes * pl.res_rep, 128);
    pl.off_main = uint32_t(off);
    if (main_smem) off += align_up(main_bytes * pl.main_rep, 128);
    pl.off_ring = uint32_t(off);
    off += size_t(warps) * stages * pl.stage_bytes;
    return off;
  };
  // prefer 16 warps x 4 stages; shed stages, then warps, then the smem-resident main codebook
  static const int kWarps[] = {16, 8};
  static const int kStages[] = {4, 3, 2};
  bool placed = false;
  for (int pass = 0; pass < 2 && !placed; ++pass) {
    const bool main_smem = pl.main_in_smem && pass == 0;
    for (int w : kWarps) {
      for (int s : kStages) {
        const size_t need = carve(w, s, main_smem);
        if (need <= size_t(budget)) {
          pl.threads = w * 32, pl.stages = s, pl.smem_bytes = uint32_t(need);
          pl.main_in_smem = main_smem;
          if (!main_smem) pl.main_rep = 1;
          placed = true;
          break;
        }
      }
      if (placed) break;
    }
  }
  if (!placed) {
    set_error("gemv: residual codebook of %zu bytes does not fit shared memory", res_bytes);
    return VPTQ_ERR_UNSUPPORTED;
  }

  if (Ro > kMaxIndexRows) {
    set_error("gemv: %d index rows exceed the supported maximum %d", Ro, kMaxIndexRows);
    return VPTQ_ERR_UNSUPPORTED;
  }
  pl.ws_counters_bytes = kCounterRegionBytes;
  pl.ws_partials_bytes = pl.nch > 1 ? align_up(size_t(pl.nch) * pl.nt * Ro * v * 4, 256) : 0;
  *out = pl;
  return 0;
}

int gemv_launch(const vptq_linear_desc& d, const void* x, int64_t x_stride, void* y, int64_t y_stride,
                int tokens, void* workspace, size_t workspace_bytes, uint32_t flags, cudaStream_t stream) {
  const DeviceInfo* dev = device_info();
  if (!dev) return VPTQ_ERR_CUDA;
  if (!supported_vec_len(d.vector_len)) {
    set_error("gemv: vector_len %d not supported (2,4,6,8,10,12,16)", d.vector_len);
    return VPTQ_ERR_UNSUPPORTED;
  }
  GemvPlan pl;
  if (int rc = gemv_make_plan(d, tokens, *dev, &pl)) return rc;
  const size_t need = pl.ws_counters_bytes + pl.ws_partials_bytes;
  if (workspace_bytes < need || (need && !workspace)) {
    set_error("gemv: workspace %zu bytes < required %zu", workspace_bytes, need);
    return VPTQ_ERR_WORKSPACE;
  }

  GemvParams p{};
  p.indices = reinterpret_cast<const uint32_t*>(d.indices);
  p.idx_stride_g = d.index_stride_codebook;
  p.idx_stride_r = d.index_stride_row;
  p.centroids = d.centroids;
  p.cb_stride = d.centroid_stride;
  p.res_centroids = d.res_centroids;
  p.rcb_stride = d.res_centroid_stride;
  p.I = d.in_features, p.O = d.out_features, p.G = d.num_codebooks, p.gs = d.group_size;
  p.Ro = (d.out_features + d.vector_len - 1) / d.vector_len;
  p.K = d.num_centroids, p.ib = ilog2(d.num_centroids);
  p.Kr = d.num_res_centroids > 0 ? d.num_res_centroids : 0;
  p.rb = p.Kr ? ilog2(p.Kr) : 0;
  p.S = (d.outlier_size > 0 && d.outlier_indices) ? d.outlier_size : 0;
  p.vol = p.S ? d.outlier_vector_len : 1;
  p.Kol = p.S ? d.num_outlier_centroids : 0;
  p.Rol = p.S ? (d.out_features + p.vol - 1) / p.vol : 0;
  p.outlier_idx = p.S ? d.outlier_indices : nullptr;
  p.outlier_cb = p.S ? d.outlier_centroids : nullptr;
  p.perm = d.perm;
  p.scale = d.weight_scale;
  p.wbias = d.weight_bias;
  p.bias = d.bias;
  p.x_stride = x_stride, p.y_stride = y_stride;
  p.counters = reinterpret_cast<uint32_t*>(workspace);
  p.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + pl.ws_counters_bytes);
  p.idx_tma_ok = ((reinterpret_cast<uintptr_t>(d.indices) & 15u) == 0 && (d.index_stride_row & 3) == 0 &&
                  (d.index_stride_codebook & 3) == 0)
                     ? 1
                     : 0;
  p.plan = pl;

  const bool res = p.rb > 0;
  const size_t esz = 2;
  for (int t0 = 0; t0 < tokens;) {
    int nt = pl.nt;
    while (nt > tokens - t0) nt >>= 1;  // tail passes: 4 -> 2 -> 1
    GemvKernelFn fn = d.vector_len == 8 ? gemv_kernel_v8(d.dtype, nt, pl.main_in_smem != 0, res)
                                        : gemv_kernel_vx(d.dtype, d.vector_len, pl.main_in_smem != 0, res);
    if (!fn) {
      set_error("gemv: no kernel for dtype=%d v=%d nt=%d", d.dtype, d.vector_len, nt);
      return VPTQ_ERR_UNSUPPORTED;
    }
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(fn), dev->smem_optin)) return rc;
    p.x = reinterpret_cast<const uint8_t*>(x) + size_t(t0) * x_stride * esz;
    p.y = reinterpret_cast<uint8_t*>(y) + size_t(t0) * y_stride * esz;
    // the smem carve-up was sized for pl.nt tokens; a narrower tail pass fits a fortiori, but the
    // partial-sum layout depends on nt, so the kernel is told the pass width through the template.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(pl.grid));
    cfg.blockDim = dim3(unsigned(pl.threads));
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (flags & VPTQ_FLAG_PDL) {
      attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[nattr].val.programmaticStreamSerializationAllowed = 1;
      ++nattr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = unsigned(nattr);
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
    if (e != cudaSuccess) {
      set_error("gemv launch (grid=%d block=%d smem=%u): %s", pl.grid, pl.threads, pl.smem_bytes,
                cudaGetErrorString(e));
      return VPTQ_ERR_CUDA;
    }
    t0 += nt;
  }
  return 0;
}

}  // namespace vptq_b200
