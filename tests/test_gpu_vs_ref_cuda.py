"""GPU-side anchor: our kernels vs the UNMODIFIED reference CUDA kernels (oracle/_ref/libvptq.so,
built by oracle/build_ref.sh from /root/reference/csrc for sm_100a) on identical tensors.

north_star: "Outputs match the reference kernels on identical (indices, centroids,
residual_centroids, perm, outliers, x) within 1e-3 relative fp16".  The reference GEMV accumulates
four columns per thread in fp16 (csrc/kernels/quant_gemv.cuh:34,140-141), so its own distance to
exact arithmetic is a few 1e-4; both distances are asserted.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

import vptq_oracle as vo
from _util import parity_error

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libvptq.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libvptq.so not built (needs /root/reference; see oracle/build_ref.sh)")
    spec = importlib.util.spec_from_file_location("libvptq", REF_SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


CASES = {
    "llama3_k65536_r256": dict(in_features=4096, out_features=1024, vector_len=8, num_centroids=65536, num_res_centroids=256),
    "k65536_r0": dict(in_features=2048, out_features=1024, vector_len=8, num_centroids=65536),
    "cfg1_k256": dict(in_features=4096, out_features=4096, vector_len=8, num_centroids=256),
    "k4096_r4096_v12": dict(in_features=1536, out_features=768, vector_len=12, num_centroids=4096, num_res_centroids=4096),
    "outliers": dict(in_features=2048 + 128, out_features=1024, vector_len=8, num_centroids=4096, num_res_centroids=256,
                     outlier_size=128, outlier_vector_len=4, num_outlier_centroids=4096, bias=True),
    "bf16": dict(in_features=2048, out_features=1024, vector_len=8, num_centroids=65536, num_res_centroids=256, dtype="bf16"),
}


def ref_tensors(L, m):
    G, v = L.num_codebooks, L.vector_len
    cent = m.centroids.weight.view(G, L.num_centroids, v)
    rcent = m.res_centroids.weight.view(G, L.num_res_centroids, v) if L.res_bits else None
    ocent = m.outlier_centroids.weight.view(1, L.num_outlier_centroids, L.outlier_vector_len) if L.enable_outlier else None
    return cent, rcent, ocent


@pytest.mark.parametrize("name", sorted(CASES))
def test_gemv_matches_reference_cuda(ref, name):
    from _gpu import from_t, make_module, x_to_t
    L = vo.make_layer(seed=2024, **CASES[name])
    m = make_module(L)
    cent, rcent, ocent = ref_tensors(L, m)
    tol = 1e-3 if L.dtype == "fp16" else 8e-3
    for tokens in (1, 2):
        x_np = vo.make_x(tokens, L.in_features, L.dtype, seed=tokens)
        x = x_to_t(x_np, L)
        y_ours = from_t(m(x))
        y_ref = from_t(ref.quant_gemv(x, m.indices, cent, None, rcent, m.outlier_indices, ocent, m.perm,
                                      m.weight_scale, m.weight_bias, m.bias, L.in_features, L.out_features))
        torch.cuda.synchronize()
        y_star = vo.quant_gemm(x_np, L)
        e_ours = parity_error(y_ours, y_star)
        assert e_ours <= tol
        if not np.isfinite(y_ref).all():
            # seen on B200 for the outlier configuration: the reference kernel returned NaN in one run and
            # finite values in others on the same inputs; nothing to compare with then -- our distance to
            # exact arithmetic is asserted above
            print(f"{name} tokens={tokens}: ours-vs-exact {e_ours:.2e}  reference kernel output is not finite, skipped")
            continue
        e_ref, e_mut = parity_error(y_ref, y_star), parity_error(y_ours, y_ref)
        print(f"{name} tokens={tokens}: ours-vs-exact {e_ours:.2e}  ref-vs-exact {e_ref:.2e}  ours-vs-ref {e_mut:.2e}")
        assert e_mut <= max(tol, 2 * e_ref), (e_mut, e_ref)


@pytest.mark.parametrize("name", ["llama3_k65536_r256", "outliers", "bf16"])
def test_dequant_matches_reference_cuda(ref, name):
    from _gpu import from_t, make_module
    L = vo.make_layer(seed=2025, **CASES[name])
    m = make_module(L)
    cent, rcent, ocent = ref_tensors(L, m)
    inv = torch.argsort(m.perm.view(torch.uint16).to(torch.int64)).to(torch.uint16).view(torch.int16)
    W_ref = from_t(ref.dequant(m.indices, cent, None, rcent, m.outlier_indices, ocent, inv, m.weight_scale,
                               m.weight_bias, L.vector_len, L.in_features, L.out_features))
    W = from_t(m.dequant())
    torch.cuda.synchronize()
    assert W.shape == W_ref.shape
    # the reference rounds C+R to 16 bit and then fma-rounds again (csrc/kernels/dequant.cuh:87,98);
    # ours evaluates in fp32 and rounds once.  |diff| <= ulp * (|W| + 0.5*|C+R|*|scale|), and
    # |C+R|*|scale| <= |W| + |wbias|.
    ulp = 2.0 ** -10 if L.dtype == "fp16" else 2.0 ** -7
    wb = np.abs(vo.to_f32(L.weight_bias, L.dtype)).max()
    bound = 2 * ulp * (np.abs(W_ref) + wb) + 1e-7
    bad = np.abs(W - W_ref) > bound
    assert not bad.any(), f"{int(bad.sum())} of {bad.size} elements beyond the double-rounding bound"
