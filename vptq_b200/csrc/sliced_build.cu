// Host-side (CPU) builder of the sliced index lists: the C-ABI counterpart of
// vptq_b200/sliced.py for hosts that are not python.  Pure data layout, no GPU work; format contract
// in include/vptq_b200.h (vptq_linear_desc::sliced_stream / sliced_offsets), consumer gemv_sliced.cu.
#include <cstring>
#include <vector>

#include "kernels.h"

namespace vptq_b200 {
namespace {

constexpr int kSlice = 8192, kStep = 32;

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

extern "C" int vptq_b200_sliced_build_host(const int32_t* indices_host, int64_t index_stride_row, int32_t out_features,
                                           int32_t group_size, int32_t num_centroids, int32_t num_res_centroids,
                                           void* stream_out, size_t stream_capacity, uint32_t* offsets_out,
                                           size_t* steps_out) {
  const int K = num_centroids, Kr = num_res_centroids > 0 ? num_res_centroids : 0, Cq = group_size;
  if (!indices_host || !offsets_out || !steps_out || out_features < 1 || Cq < 1 || Cq >= 65535) {
    set_error("sliced_build_host: NULL argument or size out of range (group_size %d)", Cq);
    return VPTQ_ERR_INVALID;
  }
  if (K < 2 * kSlice || K % kSlice || K / kSlice > 8 || (K & (K - 1)) || Kr > 256 || (Kr & (Kr - 1))) {
    set_error("sliced_build_host: needs K = NS * 8192 with 2 <= NS <= 8 and Kr <= 256 (K %d, Kr %d)", K, Kr);
    return VPTQ_ERR_UNSUPPORTED;
  }
  const int ib = ilog2(K), rb = Kr ? ilog2(Kr) : 0, b = ib + rb;
  const int NS = K / kSlice, Ro = (out_features + 7) / 8;
  const int64_t words = (int64_t(Cq) * b + 31) / 32;
  if (index_stride_row < words) {
    set_error("sliced_build_host: index_stride_row %lld < %lld words per row", (long long)index_stride_row, (long long)words);
    return VPTQ_ERR_INVALID;
  }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(indices_host);
  const size_t rec = Kr ? 160 : 128;

  // pass 1: fields per (slice, row) -> list offsets in steps, slice-major
  std::vector<uint32_t> n(size_t(NS) * Ro, 0);
  for (int r = 0; r < Ro; ++r) {
    const uint32_t* row = base + int64_t(r) * index_stride_row;
    for (int c = 0; c < Cq; ++c) ++n[size_t((field_at(row, words, c, b) & uint32_t(K - 1)) >> 13) * Ro + r];
  }
  uint64_t steps = 0;
  for (size_t i = 0; i < n.size(); ++i) {
    offsets_out[i] = uint32_t(steps);
    steps += (n[i] + kStep - 1) / kStep;
  }
  offsets_out[n.size()] = uint32_t(steps);
  *steps_out = size_t(steps);
  if (!stream_out) return 0;  // sizing call
  if (stream_capacity < steps * rec || steps > 0xffffffffull) {
    set_error("sliced_build_host: stream buffer of %zu bytes < %llu needed", stream_capacity, (unsigned long long)(steps * rec));
    return VPTQ_ERR_WORKSPACE;
  }

  // pass 2: null-fill, then per row bucket the columns by (slice, bank group = index & 7) in column
  // order and deal every list out rank-major, bank-minor: 8 consecutive entries read 8 bank groups
  uint8_t* out = static_cast<uint8_t*>(stream_out);
  const uint32_t null_word = uint32_t(Cq) << 16;
  for (uint64_t t = 0; t < steps; ++t) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out + t * rec);
    for (int i = 0; i < kStep; ++i) w[i] = null_word;
    if (Kr) std::memset(out + t * rec + 128, 0, 32);
  }
  std::vector<std::vector<uint32_t>> bucket(size_t(NS) * 8);  // packed (column << 16 | field bits we need later)
  std::vector<uint32_t> fields(Cq);
  for (int r = 0; r < Ro; ++r) {
    const uint32_t* row = base + int64_t(r) * index_stride_row;
    for (auto& v : bucket) v.clear();
    for (int c = 0; c < Cq; ++c) {
      const uint32_t f = field_at(row, words, c, b);
      fields[c] = f;
      const uint32_t idx = f & uint32_t(K - 1);
      bucket[size_t(idx >> 13) * 8 + (idx & 7)].push_back(uint32_t(c));
    }
    for (int s = 0; s < NS; ++s) {
      uint64_t pos = uint64_t(offsets_out[size_t(s) * Ro + r]) * kStep;  // entry index of the list's next slot
      size_t longest = 0;
      for (int k = 0; k < 8; ++k) longest = std::max(longest, bucket[size_t(s) * 8 + k].size());
      for (size_t rank = 0; rank < longest; ++rank)
        for (int k = 0; k < 8; ++k) {
          const std::vector<uint32_t>& bk = bucket[size_t(s) * 8 + k];
          if (rank >= bk.size()) continue;
          const uint32_t c = bk[rank], f = fields[c];
          const uint64_t t = pos / kStep, i = pos % kStep;
          reinterpret_cast<uint32_t*>(out + t * rec)[i] = (f & 8191u) | (c << 16);
          if (Kr) out[t * rec + 128 + i] = uint8_t((f >> ib) & uint32_t(Kr - 1));
          ++pos;
        }
    }
  }
  return 0;
}
