"""Helpers for the GPU parity tests: oracle Layer (numpy) -> tensors / VQuantLinear on cuda:0."""
import numpy as np
import torch

import vptq_oracle as vo


def tdtype(L):
    return torch.float16 if L.dtype == "fp16" else torch.bfloat16


def to_t(a, L, kind="float", device="cuda"):
    if a is None:
        return None
    if kind == "float":
        if L.dtype == "fp16":
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float16)).to(device)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16)).view(torch.bfloat16).to(device)
    if kind == "u16":
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16)).view(torch.int16).to(device)
    if kind == "i32":
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)
    raise ValueError(kind)


def from_t(t):
    """16-bit tensor -> float32 numpy"""
    return t.detach().float().cpu().numpy()


def make_module(L: vo.Layer, device="cuda"):
    """VQuantLinear holding exactly the oracle layer's tensors (packed, as HF loads them)."""
    from vptq_b200 import VQuantLinear
    m = VQuantLinear(
        L.in_features, L.out_features, vector_lens=[L.outlier_vector_len, L.vector_len],
        num_centroids=[L.num_outlier_centroids, L.num_centroids], num_res_centroids=[-1, L.num_res_centroids],
        group_num=L.num_codebooks, group_size=L.group_size, outlier_size=L.outlier_size, indices_as_float=False,
        enable_norm=L.weight_scale is not None, enable_perm=L.perm is not None, is_indice_packed=True,
        bias=L.bias is not None, device=device, dtype=tdtype(L), enable_proxy_error=False)
    with torch.no_grad():
        m.indices.data = to_t(L.indices, L, "i32", device)
        m.centroids.weight.data = to_t(L.centroids, L, device=device).reshape(L.num_codebooks, -1)
        if L.res_bits:
            m.res_centroids.weight.data = to_t(L.res_centroids, L, device=device).reshape(L.num_codebooks, -1)
        if L.enable_outlier:
            m.outlier_centroids.weight.data = to_t(L.outlier_centroids, L, device=device).reshape(1, -1)
            m.outlier_indices.data = to_t(L.outlier_indices, L, "u16", device)
        if L.perm is not None:
            m.perm.data = to_t(L.perm, L, "u16", device)
        if L.weight_scale is not None:
            m.weight_scale.data = to_t(L.weight_scale, L, device=device)
            m.weight_bias.data = to_t(L.weight_bias, L, device=device)
        if L.bias is not None:
            m.bias.data = to_t(L.bias, L, device=device)
    m.eval()
    return m


def x_to_t(x, L, device="cuda"):
    return to_t(x, L, device=device)
