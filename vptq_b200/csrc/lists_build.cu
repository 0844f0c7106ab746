// Host-side (CPU) builder of the slice x tile lists: the C-ABI counterpart of vptq_b200/lists.py for
// hosts that are not python.  Pure data layout, no GPU work; format contract in include/vptq_b200.h
// (vptq_linear_desc::lists_stream / lists_tab), consumer gemv_lists.cu.
#include <cstring>
#include <vector>

#include "kernels.h"

namespace vptq_b200 {
namespace {

constexpr int kSlice = 4096, kTileMax = 4096, kStep = 32;

// field j of a packed row: bits [j*b, (j+1)*b) of its little-endian 32-bit word stream
inline uint32_t field_at(const uint32_t* row, int64_t words, int64_t j, int b) {
  const int64_t bit = j * b, w = bit >> 5;
  uint64_t v = row[w];
  if (w + 1 < words) v |= uint64_t(row[w + 1]) << 32;
  v >>= (bit & 31);
  return b >= 32 ? uint32_t(v) : uint32_t(v & ((uint64_t(1) << b) - 1));
}

}  // namespace
}  // namespace vptq_b200

using namespace vptq_b200;

extern "C" int vptq_b200_lists_build_host(const int32_t* indices_host, int64_t index_stride_row, int32_t out_features,
                                          int32_t in_features, int32_t num_centroids, int32_t num_res_centroids,
                                          const uint16_t* perm_host, void* stream_out, size_t stream_capacity,
                                          uint32_t* tab_out, size_t* steps_out, int32_t* tile_cols_out) {
  const int K = num_centroids, Kr = num_res_centroids > 0 ? num_res_centroids : 0, I = in_features;
  if (!indices_host || !tab_out || !steps_out || !tile_cols_out || out_features < 1 || I < 8 || I > 65535) {
    set_error("lists_build_host: NULL argument or size out of range (in_features %d)", I);
    return VPTQ_ERR_INVALID;
  }
  if (K < 2 * kSlice || K % kSlice || K / kSlice > 16 || (K & (K - 1)) || Kr > 256 || (Kr & (Kr - 1))) {
    set_error("lists_build_host: needs K = NS * 4096 with 2 <= NS <= 16 and Kr <= 256 (K %d, Kr %d)", K, Kr);
    return VPTQ_ERR_UNSUPPORTED;
  }
  const int ib = ilog2(K), rb = Kr ? ilog2(Kr) : 0, b = ib + rb;
  const int NS = K / kSlice, Ro = (out_features + 7) / 8;
  const int NT = (I + kTileMax - 1) / kTileMax, TCW = ((I + NT - 1) / NT + 7) / 8 * 8, Q = NS * NT;
  *tile_cols_out = TCW;
  const int64_t words = (int64_t(I) * b + 31) / 32;
  if (index_stride_row < words) {
    set_error("lists_build_host: index_stride_row %lld < %lld words per row", (long long)index_stride_row, (long long)words);
    return VPTQ_ERR_INVALID;
  }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(indices_host);
  auto feature = [&](int c) { return perm_host ? int(perm_host[c]) : c; };
  if (perm_host)
    for (int c = 0; c < I; ++c)
      if (int(perm_host[c]) >= I) {
        set_error("lists_build_host: perm[%d] = %d is not a feature index (< %d)", c, int(perm_host[c]), I);
        return VPTQ_ERR_INVALID;
      }

  // pass 1: fields per unit -> first step and tail count of every list, combo-major
  std::vector<uint32_t> n(size_t(Q) * Ro, 0);
  for (int r = 0; r < Ro; ++r) {
    const uint32_t* row = base + int64_t(r) * index_stride_row;
    for (int c = 0; c < I; ++c) {
      const uint32_t idx = field_at(row, words, c, b) & uint32_t(K - 1);
      ++n[size_t((feature(c) / TCW) * NS + int(idx >> 12)) * Ro + r];
    }
  }
  uint64_t steps = 0;
  std::vector<uint32_t> first(n.size());
  for (size_t u = 0; u < n.size(); ++u) {
    const uint32_t st = n[u] ? (n[u] + kStep - 1) / kStep : 1u;  // every list has at least one step
    const uint32_t tail = n[u] - kStep * (st - 1);
    first[u] = uint32_t(steps);
    tab_out[u] = uint32_t(steps) | tail << 26;
    steps += st;
  }
  tab_out[n.size()] = uint32_t(steps);
  *steps_out = size_t(steps);
  if (steps >= (1ull << 26)) {
    set_error("lists_build_host: %llu steps exceed the 26-bit step counter", (unsigned long long)steps);
    return VPTQ_ERR_UNSUPPORTED;
  }
  if (!stream_out) return 0;  // sizing call
  if (stream_capacity < steps * 128) {
    set_error("lists_build_host: stream buffer of %zu bytes < %llu needed", stream_capacity, (unsigned long long)(steps * 128));
    return VPTQ_ERR_WORKSPACE;
  }

  // pass 2: zero-fill, then per row bucket the columns by (combo, bank group = index & 7) in column order
  // and deal every list out rank-major, bank-minor: 8 consecutive entries read 8 different bank groups
  uint32_t* out = static_cast<uint32_t*>(stream_out);
  std::memset(out, 0, size_t(steps) * 128);
  std::vector<std::vector<uint32_t>> bucket(size_t(Q) * 8);
  std::vector<uint32_t> fields(I);
  for (int r = 0; r < Ro; ++r) {
    const uint32_t* row = base + int64_t(r) * index_stride_row;
    for (auto& v : bucket) v.clear();
    for (int c = 0; c < I; ++c) {
      const uint32_t f = field_at(row, words, c, b);
      fields[c] = f;
      const uint32_t idx = f & uint32_t(K - 1);
      bucket[(size_t(feature(c) / TCW) * NS + (idx >> 12)) * 8 + (idx & 7)].push_back(uint32_t(c));
    }
    for (int q = 0; q < Q; ++q) {
      uint64_t pos = uint64_t(first[size_t(q) * Ro + r]) * kStep;  // entry index of the list's next slot
      size_t longest = 0;
      for (int k = 0; k < 8; ++k) longest = std::max(longest, bucket[size_t(q) * 8 + k].size());
      for (size_t rank = 0; rank < longest; ++rank)
        for (int k = 0; k < 8; ++k) {
          const std::vector<uint32_t>& bk = bucket[size_t(q) * 8 + k];
          if (rank >= bk.size()) continue;
          const uint32_t c = bk[rank], f = fields[c];
          const uint32_t lcol = uint32_t(feature(int(c)) - (q / NS) * TCW);
          out[pos++] = (f & 4095u) | lcol << 12 | (Kr ? ((f >> ib) & uint32_t(Kr - 1)) << 24 : 0u);
        }
    }
  }
  return 0;
}
