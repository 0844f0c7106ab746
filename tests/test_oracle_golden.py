"""Pins the CPU oracle (oracle/vptq_oracle.py) to the reference's own python.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which runs the
reference's pack_index / unpack_index_tensor / dequant / quant_gemm (vptq/utils/pack.py:26-139,
vptq/ops/quant_gemm.py:43-275).  Integer work is compared bit-exactly, fp32 work to 1e-6.
"""
import numpy as np
import pytest

import vptq_oracle as vo
from _util import golden_names, load_golden

NAMES = golden_names()


def test_golden_present():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", NAMES)
def test_pack_bit_exact(name):
    L, x, ref = load_golden(name)
    packed = vo.pack_index(L.meta["idx"], L.index_bits, L.meta["ridx"], L.res_bits)
    assert packed.dtype == np.int32
    assert np.array_equal(packed, ref["packed_ref"])
    assert np.array_equal(L.indices, ref["packed_ref"])


@pytest.mark.parametrize("name", NAMES)
def test_unpack_bit_exact(name):
    L, x, ref = load_golden(name)
    idx, ridx = vo.unpack_index(ref["packed_ref"], L.index_bits, L.group_size, L.res_bits)
    assert np.array_equal(idx.astype(np.uint16), ref["u_idx"])
    assert np.array_equal(idx, L.meta["idx"])
    if L.res_bits:
        assert np.array_equal(ridx.astype(np.uint16), ref["u_ridx"])
        assert np.array_equal(ridx, L.meta["ridx"])
    else:
        assert ridx is None


@pytest.mark.parametrize("name", NAMES)
def test_dequant_matches_reference_fp32(name):
    L, x, ref = load_golden(name)
    W = vo.dequant(L)
    assert W.shape == (L.out_features, L.in_features)
    # same fp32 operations in the same order as the reference's torch code -> identical
    np.testing.assert_allclose(W, ref["W_ref"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", NAMES)
def test_dequant_close_to_reference_16bit(name):
    """The reference fallback in its native 16-bit arithmetic differs from fp32 math only by
    its intermediate roundings (C+R, *scale, +bias: three 16-bit roundings)."""
    L, x, ref = load_golden(name)
    W = vo.dequant(L)
    ulp = 2.0 ** -10 if L.dtype == "fp16" else 2.0 ** -7
    err = np.abs(W - ref["W_ref16"])
    bound = 2.0 * ulp * np.maximum(np.abs(W), np.abs(ref["W_ref16"]).max() * 0.25)
    assert (err <= bound + 1e-6).all()


@pytest.mark.parametrize("name", NAMES)
def test_quant_gemm_matches_reference_fp32(name):
    L, x, ref = load_golden(name)
    y = vo.quant_gemm(x, L)
    assert y.shape == ref["y_ref"].shape
    scale = np.abs(ref["y_ref"]).max()
    # the reference accumulates in fp32 (torch CPU matmul), the oracle in fp64
    assert np.abs(y - ref["y_ref"]).max() <= 2e-5 * scale


def test_pack_roundtrip_all_widths():
    rng = np.random.default_rng(0)
    for ib in (2, 4, 7, 8, 10, 12, 13, 15, 16):
        for rb in (0, 3, 8, 12, 16):
            gs = 37
            idx = rng.integers(0, 1 << ib, size=(2, 5, gs))
            ridx = rng.integers(0, 1 << rb, size=(2, 5, gs)) if rb else None
            p = vo.pack_index(idx, ib, ridx, rb)
            assert p.shape == (2, 5, (gs * (ib + rb) + 31) // 32)
            i2, r2 = vo.unpack_index(p, ib, gs, rb)
            assert np.array_equal(i2, idx)
            if rb:
                assert np.array_equal(r2, ridx)


def test_gemv_v2_layout_matches_reference_test_loop():
    """quant_gemv_v2 layout = the python loop of reference tests/test_quant_gemv.py:86-105."""
    rng = np.random.default_rng(5)
    I, O, v, K, Kr = 64, 48, 8, 32, 16
    n = I * O // v
    idx = (np.arange(n) % K).astype(np.uint16)          # the reference test's arange-repeat pattern
    ridx = (np.arange(n) % Kr).astype(np.uint8)
    C = vo.f32_to_bf16_bits(rng.normal(0.02, 0.5, (1, K, v)))
    R = vo.f32_to_bf16_bits(rng.normal(0.02, 0.5, (1, Kr, v)))
    sw = vo.f32_to_bf16_bits(rng.normal(0.02, 0.5, (I, 1)))
    sb = vo.f32_to_bf16_bits(rng.normal(0.02, 0.5, (I, 1)))
    x = vo.f32_to_bf16_bits(rng.normal(0.02, 0.5, (1, 1, I)))
    Cf, Rf = vo.bf16_bits_to_f32(C)[0], vo.bf16_bits_to_f32(R)[0]
    W = np.zeros((I, O))
    for i in range(n):
        row, col = i % I, i // I * v
        W[row, col:col + v] = Cf[idx[i]] + Rf[ridx[i]]
    W = vo.bf16_bits_to_f32(sw).astype(np.float64) * W + vo.bf16_bits_to_f32(sb)
    want = vo.bf16_bits_to_f32(x).reshape(1, I).astype(np.float64) @ W
    got = vo.quant_gemv_v2(x, None, idx, C, ridx, R, sw, sb, v, O, dtype="bf16")
    np.testing.assert_allclose(got.reshape(1, O), want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", NAMES)
def test_torch_port_matches_reference(name):
    """oracle/torch_port.py (the multi-threaded CPU baseline bench.py times) == reference python."""
    import torch
    import torch_port as tp
    L, x, ref = load_golden(name)
    f = (lambda a: None if a is None else (torch.from_numpy(np.asarray(a, dtype=np.float16).copy()) if L.dtype == "fp16"
         else torch.from_numpy(np.asarray(a, dtype=np.uint16).copy()).view(torch.bfloat16)))
    u = lambda a: None if a is None else torch.from_numpy(np.asarray(a, dtype=np.uint16).copy()).view(torch.int16)
    d = dict(in_features=L.in_features, out_features=L.out_features, vector_len=L.vector_len,
             num_centroids=L.num_centroids, num_res_centroids=L.num_res_centroids, num_codebooks=L.num_codebooks,
             group_size=L.group_size, outlier_size=L.outlier_size if L.enable_outlier else 0,
             outlier_vector_len=L.outlier_vector_len, num_outlier_centroids=L.num_outlier_centroids,
             indices=torch.from_numpy(L.indices.copy()), centroids=f(L.centroids), res_centroids=f(L.res_centroids),
             outlier_indices=u(L.outlier_indices), outlier_centroids=f(L.outlier_centroids), perm=u(L.perm),
             weight_scale=f(L.weight_scale), weight_bias=f(L.weight_bias), bias=f(L.bias))
    W = tp.dequant(d).numpy()
    np.testing.assert_allclose(W, ref["W_ref"], rtol=1e-6, atol=1e-7)
    y = tp.quant_gemm(f(x), d).numpy()
    assert np.abs(y - ref["y_ref"]).max() <= 2e-5 * np.abs(ref["y_ref"]).max()
