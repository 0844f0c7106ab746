"""Multi-threaded CPU port of the reference's pure-torch path  --  TEST / BASELINE INFRASTRUCTURE.

`bench.py` times this on the GPU box's host cores as `cpu_baseline` (kind "port") and as the
`--impl reference` arm: /root/reference does not exist on the GPU box, so the reference's own
python cannot travel; this file restates its fallback path (what `vptq.ops.quant_gemm` executes
when the CUDA extension is missing, vptq/ops/quant_gemm.py:247-275) with torch CPU ops, which use
all host threads:

  1. unpack the int32 words into main / residual indices      (vptq/utils/pack.py:105-139)
  2. gather centroid (+ residual) vectors                      (vptq/ops/quant_gemm.py:88-121)
  3. lay them out as W[out, in], un-permute, scale + bias       (:97-103, :151-156)
  4. F.linear(x, W, bias)                                       (:274)

It is checked against the golden fixtures in tests/test_oracle_golden.py::test_torch_port_*.
The unpack here is a two-word funnel shift, which is cheaper than the reference's 32 bit-planes
(pack.py:112-124); the baseline is therefore, if anything, favourable to the CPU.
Never imported by the product package.
"""
from __future__ import annotations

import torch


def unpack(packed: torch.Tensor, index_bits: int, res_bits: int, n: int):
    b = index_bits + res_bits
    w = torch.nn.functional.pad(packed.to(torch.int64) & 0xFFFFFFFF, (0, 1))
    bit = torch.arange(n, dtype=torch.int64) * b
    w0, sh = bit >> 5, bit & 31
    field = ((w.index_select(-1, w0) >> sh) | (w.index_select(-1, w0 + 1) << (32 - sh))) & ((1 << b) - 1)
    idx = field & ((1 << index_bits) - 1)
    ridx = (field >> index_bits) if res_bits else None
    return idx, ridx


def dequant(L: dict) -> torch.Tensor:
    """L: dict of CPU torch tensors / ints with the VQuantLinear names.  Returns W [O, I] fp32."""
    v, G, gs, K, Kr = L["vector_len"], L["num_codebooks"], L["group_size"], L["num_centroids"], L["num_res_centroids"]
    ib = K.bit_length() - 1
    rb = (Kr.bit_length() - 1) if Kr > 0 else 0
    idx, ridx = unpack(L["indices"], ib, rb, gs)                       # [G, Ro, gs]
    C = L["centroids"].float().view(G, K, v)
    Ro = idx.shape[1]
    sel = torch.stack([C[g].index_select(0, idx[g].reshape(-1)) for g in range(G)])  # [G, Ro*gs, v]
    if rb:
        R = L["res_centroids"].float().view(G, Kr, v)
        sel = sel + torch.stack([R[g].index_select(0, ridx[g].reshape(-1)) for g in range(G)])
    W = sel.view(G, Ro, gs, v).permute(1, 3, 0, 2).reshape(Ro * v, G * gs)
    W = W[: L["out_features"]]
    S = L.get("outlier_size", 0)
    if S and L.get("outlier_indices") is not None:
        vol, Kol = L["outlier_vector_len"], L["num_outlier_centroids"]
        oi = L["outlier_indices"].view(torch.uint16).to(torch.int64).view(-1, S)
        Wo = L["outlier_centroids"].float().view(Kol, vol).index_select(0, oi.reshape(-1))
        Wo = Wo.view(-1, S, vol).permute(0, 2, 1).reshape(-1, S)[: L["out_features"]]
        W = torch.cat([Wo, W], dim=1)
    if L.get("perm") is not None:
        inv = torch.argsort(L["perm"].view(torch.uint16).to(torch.int64))
        W = W.index_select(1, inv)
    if L.get("weight_scale") is not None:
        W = W * L["weight_scale"].float() + L["weight_bias"].float()
    return W


def quant_gemm(x: torch.Tensor, L: dict) -> torch.Tensor:
    W = dequant(L)
    b = L.get("bias")
    return torch.nn.functional.linear(x.float(), W, None if b is None else b.float())


def synthetic_layer(in_features: int, out_features: int, vector_len: int = 8, num_centroids: int = 65536,
                    num_res_centroids: int = 256, seed: int = 0, dtype=torch.float16) -> dict:
    """Random layer with the bench's distribution, on the CPU (for the baseline timing)."""
    g = torch.Generator().manual_seed(seed)
    ib = num_centroids.bit_length() - 1
    rb = (num_res_centroids.bit_length() - 1) if num_res_centroids > 0 else 0
    Ro = (out_features + vector_len - 1) // vector_len
    wd = (in_features * (ib + rb) + 31) // 32
    L = dict(in_features=in_features, out_features=out_features, vector_len=vector_len,
             num_centroids=num_centroids, num_res_centroids=num_res_centroids, num_codebooks=1,
             group_size=in_features,
             indices=torch.randint(-2 ** 31, 2 ** 31, (1, Ro, wd), generator=g, dtype=torch.int64).to(torch.int32),
             centroids=(torch.randn(1, num_centroids * vector_len, generator=g) / in_features ** 0.5).to(dtype),
             perm=torch.randperm(in_features, generator=g).to(torch.uint16).view(torch.int16),
             weight_scale=(1 + 0.1 * torch.randn(in_features, generator=g)).to(dtype),
             weight_bias=(0.01 * torch.randn(in_features, generator=g)).to(dtype))
    if rb:
        L["res_centroids"] = (0.25 * torch.randn(1, num_res_centroids * vector_len, generator=g) / in_features ** 0.5).to(dtype)
    return L
