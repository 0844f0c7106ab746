"""Generate tests/golden/*.npz by running the REFERENCE's own python on seeded inputs.

Test infrastructure; runs only in the authoring container (needs /root/reference).  For each
case it builds a seeded synthetic layer (oracle/vptq_oracle.make_layer), then calls, through
oracle/ref_shim.py, the reference's

  * pack_index            (vptq/utils/pack.py:26-102)   on the raw indices,
  * unpack_index_tensor   (vptq/utils/pack.py:105-139)  on the packed words,
  * dequant               (vptq/ops/quant_gemm.py:43-158) on fp32-upcast tensors  -> W_ref
  * quant_gemm            (vptq/ops/quant_gemm.py:161-275, torch fallback) fp32   -> y_ref
  * dequant in the native 16-bit dtype (what the fallback would return to F.linear) -> W_ref16

and stores inputs + outputs.  The committed fixtures are what pins oracle/vptq_oracle.py and
oracle/torch_port.py (tests/test_oracle_golden.py), and what the GPU parity tests compare the
CUDA path against (tests/test_gpu_parity.py).

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import vptq_oracle as vo  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> make_layer kwargs (+ tokens).  Small on purpose: fixtures are committed.
CASES = {
    # plain: v=8, K=256, no residual (BASELINE cfg1 parameters at a small size)
    "v8_k256": dict(in_features=256, out_features=128, vector_len=8, num_centroids=256),
    # residual codebook, b = 12 + 8 = 20 (word-straddling fields)
    "v8_k4096_r256": dict(in_features=256, out_features=96, vector_len=8, num_centroids=4096,
                          num_res_centroids=256, bias=True),
    # b = 13 + 4 = 17: odd total width
    "v8_k8192_r16": dict(in_features=192, out_features=64, vector_len=8, num_centroids=8192,
                         num_res_centroids=16),
    # out_features not a multiple of v -> padding rows dropped
    "v6_k1024_pad": dict(in_features=128, out_features=100, vector_len=6, num_centroids=1024,
                         num_res_centroids=-1),
    "v12_k4096_r4096": dict(in_features=128, out_features=96, vector_len=12, num_centroids=4096,
                            num_res_centroids=4096),
    "v16_k256_r256": dict(in_features=128, out_features=64, vector_len=16, num_centroids=256,
                          num_res_centroids=256),
    "v4_k256_nonorm_noperm": dict(in_features=128, out_features=64, vector_len=4, num_centroids=256,
                                  enable_perm=False, enable_norm=False),
    # outlier block (vol=4) + perm + norm
    "v8_k256_outlier": dict(in_features=272, out_features=64, vector_len=8, num_centroids=256,
                            num_res_centroids=16, outlier_size=16, outlier_vector_len=4,
                            num_outlier_centroids=64, bias=True),
    # two codebook groups
    "v8_k256_g2": dict(in_features=256, out_features=64, vector_len=8, num_centroids=256,
                       num_res_centroids=16, num_codebooks=2),
    # bf16 tensors
    "v8_k4096_r256_bf16": dict(in_features=256, out_features=64, vector_len=8, num_centroids=4096,
                               num_res_centroids=256, dtype="bf16"),
    # the reference unit test's value distribution N(0.02, 0.5) (tests/test_quant_gemv.py:129-130)
    "v8_k8192_r256_reftestdist": dict(in_features=256, out_features=64, vector_len=8,
                                      num_centroids=8192, num_res_centroids=256, llm_like=False,
                                      dtype="bf16"),
    # 16-bit main index + 8-bit residual = the b=24 layout of the Llama-3 checkpoints, tiny layer
    "v8_k65536_r256_small": dict(in_features=128, out_features=32, vector_len=8,
                                 num_centroids=65536, num_res_centroids=256),
}
TOKENS = 3


def _t(a, dtype, kind="float"):
    """numpy array from the oracle's Layer -> torch tensor in the reference's storage dtype."""
    if a is None:
        return None
    if kind == "float":
        if dtype == "fp16":
            return torch.from_numpy(np.asarray(a, dtype=np.float16).copy())
        return torch.from_numpy(np.asarray(a, dtype=np.uint16).copy()).view(torch.bfloat16)
    if kind == "u16_as_i16":
        return torch.from_numpy(np.asarray(a, dtype=np.uint16).copy()).view(torch.int16)
    raise ValueError(kind)


def run_reference(L: vo.Layer, x_np, pack, qg):
    tdt = torch.float16 if L.dtype == "fp16" else torch.bfloat16
    idx, ridx = L.meta["idx"], L.meta["ridx"]
    # 1) reference pack_index on uint16 index tensors (as vptq/utils/pack.py:212-232 calls it)
    tidx = torch.from_numpy(idx.astype(np.uint16))
    tridx = torch.from_numpy(ridx.astype(np.uint16)) if ridx is not None else None
    packed_ref = pack.pack_index(tidx, L.index_bits, tridx, L.res_bits, index_dtype=torch.uint16)
    # 2) reference unpack on the packed words
    u_idx, u_ridx = pack.unpack_index_tensor(packed_ref, L.index_bits, L.group_size,
                                             L.res_bits, L.group_size if L.res_bits else 0)
    cent = _t(L.centroids, L.dtype)
    rcent = _t(L.res_centroids, L.dtype)
    ocent = _t(L.outlier_centroids, L.dtype)
    oidx = _t(L.outlier_indices, L.dtype, "u16_as_i16")
    perm = _t(L.perm, L.dtype, "u16_as_i16")
    ws, wb = _t(L.weight_scale, L.dtype), _t(L.weight_bias, L.dtype)
    bias = _t(L.bias, L.dtype)
    x = _t(x_np, L.dtype)

    def call(up):
        f = (lambda t: None if t is None else t.to(up))
        return qg.quant_gemm(
            f(x), f(bias), packed_ref, f(cent), oidx if L.enable_outlier else None,
            f(ocent) if L.enable_outlier else None, None, f(rcent), perm, f(ws), f(wb),
            L.vector_len, L.outlier_vector_len, L.num_codebooks, L.num_centroids,
            L.num_outlier_centroids, L.num_res_centroids, True, L.group_size, L.outlier_size,
            L.in_features, L.out_features, L.padding, L.outlier_padding)

    def call_dequant(up):
        f = (lambda t: None if t is None else t.to(up))
        G, K, Kr, v = L.num_codebooks, L.num_centroids, L.num_res_centroids, L.vector_len
        return qg.dequant(
            indices=packed_ref, centroids=f(cent).view(G, K, v),
            outlier_indices=oidx if L.enable_outlier else None,
            outlier_centroids=f(ocent) if L.enable_outlier else None,
            res_indices=None, res_centroids=f(rcent).view(G, Kr, v) if Kr > 0 else None,
            perm=perm, weight_scale=f(ws), weight_bias=f(wb), is_indice_packed=True,
            enable_outlier=L.enable_outlier, enable_residual=Kr > 0, enable_perm=perm is not None,
            enable_norm=ws is not None, num_centroids=K, num_outlier_centroids=L.num_outlier_centroids,
            num_res_centroids=Kr, padding=L.padding, outlier_padding=L.outlier_padding,
            num_codebooks=G, group_size=L.group_size, outlier_size=L.outlier_size, vector_len=v,
            outlier_vector_len=L.outlier_vector_len)

    with torch.no_grad():
        y32 = call(torch.float32)
        W32 = call_dequant(torch.float32)
        W16 = call_dequant(tdt)        # the reference fallback's own 16-bit arithmetic
    return dict(packed_ref=packed_ref.numpy(), u_idx=u_idx.numpy(),
                u_ridx=None if u_ridx is None else u_ridx.numpy(),
                y_ref=y32.numpy(), W_ref=W32.numpy(), W_ref16=W16.float().numpy())


def save_case(name, L: vo.Layer, x, ref):
    d = dict(
        dtype=L.dtype, in_features=L.in_features, out_features=L.out_features,
        vector_len=L.vector_len, num_centroids=L.num_centroids, num_res_centroids=L.num_res_centroids,
        num_codebooks=L.num_codebooks, group_size=L.group_size, outlier_size=L.outlier_size,
        outlier_vector_len=L.outlier_vector_len, num_outlier_centroids=L.num_outlier_centroids,
        idx=L.meta["idx"].astype(np.uint16), x=x, indices=L.indices, centroids=L.centroids,
        y_ref=ref["y_ref"].astype(np.float32), W_ref=ref["W_ref"].astype(np.float32),
        W_ref16=ref["W_ref16"].astype(np.float32),
        packed_ref=ref["packed_ref"], u_idx=ref["u_idx"].astype(np.uint16))
    if L.meta["ridx"] is not None:
        d["ridx"] = L.meta["ridx"].astype(np.uint16)
        d["u_ridx"] = ref["u_ridx"].astype(np.uint16)
    for k in ("res_centroids", "outlier_indices", "outlier_centroids", "perm", "weight_scale",
              "weight_bias", "bias"):
        if getattr(L, k) is not None:
            d[k] = getattr(L, k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def main():
    pack, qg = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    for i, (name, kw) in enumerate(CASES.items()):
        L = vo.make_layer(seed=1234 + i, **kw)
        x = vo.make_x(TOKENS, L.in_features, L.dtype, seed=77 + i)
        ref = run_reference(L, x, pack, qg)
        # the oracle's own packer must reproduce the reference packer bit for bit before we save
        assert np.array_equal(ref["packed_ref"], L.indices), name
        save_case(name, L, x, ref)
        sz = os.path.getsize(os.path.join(OUT, name + ".npz"))
        print(f"{name:32s} b={L.index_bits + L.res_bits:2d} W{ref['W_ref'].shape} -> {sz/1024:.0f} KiB")


if __name__ == "__main__":
    main()
