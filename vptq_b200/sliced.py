"""Load-time re-bucketing of a layer's packed indices for the sliced-codebook decode kernel.

Format contract: include/vptq_b200.h (`vptq_linear_desc::sliced_stream`, `::sliced_offsets`);
consumer: vptq_b200/csrc/gemv_sliced.cu.  A main codebook of K = NS * 8192 entries is cut into NS
slices of 128 KiB; every index row is split into NS lists, list (s, r) holding the fields of row r
whose main index falls into slice s as (index & 8191 | column << 16) plus the residual index.  A sum
does not depend on the order of its terms, so within a list the entries are arranged round-robin
over the eight 16-byte bank groups (index & 7): 8 consecutive entries -- the 8 lanes of a
quarter-warp -- then read 8 different bank groups of shared memory.

Pure tensor code (argsort / scatter), runs on whatever device `indices` lives on; done once per
layer.  This is host-side data layout, not a compute path: the kernel does all arithmetic.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .pack import unpack_index_tensor

SLICE_ENTRIES = 8192
STEP = 32


def eligible(*, vector_len: int, num_centroids: int, num_res_centroids: int, num_codebooks: int,
             outlier_size: int, in_features: int) -> bool:
    """Mirror of gemv_sliced_eligible() (csrc/gemv_sliced.cu) for the shape-only conditions."""
    K = int(num_centroids)
    return (vector_len == 8 and num_codebooks == 1 and outlier_size <= 0 and K >= 2 * SLICE_ENTRIES
            and K % SLICE_ENTRIES == 0 and K // SLICE_ENTRIES <= 8 and num_res_centroids <= 256
            and in_features < 65535)


def build_sliced(indices: torch.Tensor, *, num_centroids: int, num_res_centroids: int, group_size: int,
                 out_features: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """packed int32 [1, >=Ro, W] -> (stream uint8 [T, 160 or 128], offsets int32 [NS*Ro + 1])."""
    K, Kr, Cq = int(num_centroids), int(num_res_centroids), int(group_size)
    Ro = (int(out_features) + 7) // 8
    ib = K.bit_length() - 1
    rb = (Kr.bit_length() - 1) if Kr > 0 else 0
    NS = K // SLICE_ENTRIES
    dev = indices.device
    idx, ridx = unpack_index_tensor(indices[0, :Ro], ib, Cq, rb)          # [Ro, Cq] int64
    sl, low = idx >> 13, idx & (SLICE_ENTRIES - 1)
    key1 = sl * 8 + (low & 7)                                              # (slice, bank group)
    order1 = torch.argsort(key1, dim=1, stable=True)
    k1s = torch.gather(key1, 1, order1)
    cnt = torch.zeros(Ro, NS * 8, dtype=torch.int64, device=dev)
    cnt.scatter_add_(1, key1, torch.ones_like(key1))
    start = cnt.cumsum(1) - cnt
    pos = torch.arange(Cq, device=dev, dtype=torch.int64)[None, :].expand(Ro, Cq)
    rank = pos - torch.gather(start, 1, k1s)                               # rank inside its (slice, bank) bucket
    key2 = ((k1s >> 3) * (Cq + 1) + rank) * 8 + (k1s & 7)                  # slice, then rank, then bank group
    order2 = torch.argsort(key2, dim=1)
    src = torch.gather(order1, 1, order2)                                  # quantised column of each sorted position
    sl_s = torch.gather(k1s >> 3, 1, order2)
    low_s = torch.gather(low, 1, src)
    n_rs = cnt.view(Ro, NS, 8).sum(2)                                      # fields of (row, slice)
    row_start = n_rs.cumsum(1) - n_rs
    within = pos - torch.gather(row_start, 1, sl_s)
    steps_sr = ((n_rs + STEP - 1) // STEP).t().contiguous().view(-1)       # slice-major
    offs = torch.zeros(NS * Ro + 1, dtype=torch.int64, device=dev)
    offs[1:] = steps_sr.cumsum(0)
    T = int(offs[-1].item())
    rows = torch.arange(Ro, device=dev, dtype=torch.int64)[:, None].expand(Ro, Cq)
    dest = (offs[sl_s * Ro + rows] * STEP + within).reshape(-1)
    words = torch.full((T * STEP,), Cq << 16, dtype=torch.int64, device=dev)   # null entries: column Cq, index 0
    words[dest] = (low_s | (src << 16)).reshape(-1)
    words = torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)
    parts = [words.view(T, STEP).view(torch.uint8).view(T, STEP * 4)]
    if Kr > 0:
        rbytes = torch.zeros(T * STEP, dtype=torch.uint8, device=dev)
        rbytes[dest] = torch.gather(ridx, 1, src).reshape(-1).to(torch.uint8)
        parts.append(rbytes.view(T, STEP))
    stream = torch.cat(parts, dim=1).contiguous()
    return stream, offs.to(torch.int32)


def emulate(stream: torch.Tensor, offsets: torch.Tensor, *, num_centroids: int, num_res_centroids: int,
            group_size: int, out_features: int, centroids: torch.Tensor, res_centroids: Optional[torch.Tensor],
            xq: torch.Tensor) -> torch.Tensor:
    """float64 evaluation of sum_c xq[c] * (C[idx] + R[ridx]) straight from the sliced lists
    (test aid: validates the format without a GPU).  Returns [Ro * 8]."""
    K, Kr, Cq = int(num_centroids), int(num_res_centroids), int(group_size)
    Ro = (int(out_features) + 7) // 8
    NS = K // SLICE_ENTRIES
    T = stream.shape[0]
    words = stream[:, :STEP * 4].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF   # [T, 32]
    low, col = words & (SLICE_ENTRIES - 1), words >> 16
    C = centroids.reshape(K, 8).double()
    xpad = torch.cat([xq.double().reshape(-1), torch.zeros(1, dtype=torch.float64)])
    step_list = torch.bucketize(torch.arange(T), offsets[1:].to(torch.int64), right=True)      # list id of each step
    s_of, r_of = step_list // Ro, step_list % Ro
    w = C[(s_of[:, None] * SLICE_ENTRIES + low)]                                               # [T, 32, 8]
    if Kr > 0:
        w = w + res_centroids.reshape(Kr, 8).double()[stream[:, STEP * 4:].to(torch.int64)]
    contrib = (w * xpad[col][:, :, None]).sum(1)                                               # [T, 8]
    y = torch.zeros(Ro, 8, dtype=torch.float64)
    y.index_add_(0, r_of, contrib)
    return y.reshape(-1)
