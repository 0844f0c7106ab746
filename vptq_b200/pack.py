"""The packed-index wire format (host side).

Same contract as the reference's `pack_index` / `unpack_index_tensor`
(vptq/utils/pack.py:26-102 and :105-139): per index row a little-endian bit stream in int32
words; field j occupies bits [j*b, (j+1)*b), b = index_bits + res_bits, and holds
`idx | (res_idx << index_bits)`; rows are padded to a whole word.

Implementation note: the reference explodes every word into 32 bit-planes; here a field is
placed with two shifted adds (fields never overlap, so add == or) and extracted with a
two-word funnel shift, which is also what the CUDA kernels do.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

__all__ = ["pack_index", "unpack_index_tensor"]


def _as_unsigned(t: torch.Tensor, index_dtype: torch.dtype) -> torch.Tensor:
    """indices are stored as uint16 bit patterns viewed as int16 / float16 (vqlinear.py:111-113)."""
    if t.dtype in (torch.int16, torch.float16, torch.bfloat16) and index_dtype == torch.uint16:
        return t.view(torch.uint16).to(torch.int64)
    if t.dtype == torch.uint16:
        return t.to(torch.int64)
    return t.to(torch.int64) & 0xFFFFFFFF


def pack_index(indice: torch.Tensor, index_bits: int, res_indice: Optional[torch.Tensor] = None,
               res_bits: int = 0, index_dtype: torch.dtype = torch.uint16,
               as_dtype: torch.dtype = torch.int32) -> torch.Tensor:
    """[..., n] indices (+ residual indices) -> [..., ceil(n*b/32)] int32 words."""
    total_bits = index_bits + res_bits
    if total_bits > 32 or total_bits <= 0:
        raise ValueError(f"total index bits {total_bits} must be in (0, 32]")
    if as_dtype != torch.int32:
        raise ValueError("as_dtype must be torch.int32")
    merged = _as_unsigned(indice, index_dtype)
    if res_indice is not None and res_bits > 0:
        merged = merged | (_as_unsigned(res_indice, index_dtype) << index_bits)
    n = merged.shape[-1]
    words = (n * total_bits + 31) // 32
    bit = torch.arange(n, device=merged.device, dtype=torch.int64) * total_bits
    w0, sh = bit >> 5, bit & 31
    shifted = merged << sh                                   # < 2^63: merged < 2^32, sh < 32
    out = torch.zeros(*merged.shape[:-1], words + 1, dtype=torch.int64, device=merged.device)
    out.index_add_(-1, w0, shifted & 0xFFFFFFFF)
    out.index_add_(-1, w0 + 1, shifted >> 32)
    out = out[..., :words]
    # two's-complement reinterpretation of the low 32 bits
    return torch.where(out >= (1 << 31), out - (1 << 32), out).to(torch.int32)


def unpack_index_tensor(packed_tensor: torch.Tensor, index_bits: int, num_elements: int, res_bits: int = 0,
                        num_res_elements: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """[..., words] int32 -> (indices, res_indices) int64 [..., num_elements]; res None if res_bits == 0."""
    total_bits = index_bits + res_bits
    w = packed_tensor.to(torch.int64) & 0xFFFFFFFF
    w = torch.nn.functional.pad(w, (0, 1))
    bit = torch.arange(num_elements, device=w.device, dtype=torch.int64) * total_bits
    w0, sh = bit >> 5, bit & 31
    lo = w.index_select(-1, w0)
    hi = w.index_select(-1, w0 + 1)
    field = ((lo >> sh) | (hi << (32 - sh))) & ((1 << total_bits) - 1)
    indices = field & ((1 << index_bits) - 1)
    res = None
    if res_bits > 0:
        res = (field >> index_bits) & ((1 << res_bits) - 1)
    return indices, res
