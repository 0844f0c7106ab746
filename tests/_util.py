"""Shared helpers for the test-suite (golden loading, parity metric)."""
import glob
import os

import numpy as np

import vptq_oracle as vo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """-> (Layer, x, dict of reference outputs)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = lambda k: z[k] if k in z.files else None
    L = vo.Layer(
        dtype=str(z["dtype"]), in_features=int(z["in_features"]), out_features=int(z["out_features"]),
        vector_len=int(z["vector_len"]), num_centroids=int(z["num_centroids"]),
        num_res_centroids=int(z["num_res_centroids"]), num_codebooks=int(z["num_codebooks"]),
        group_size=int(z["group_size"]), outlier_size=int(z["outlier_size"]),
        outlier_vector_len=int(z["outlier_vector_len"]), num_outlier_centroids=int(z["num_outlier_centroids"]),
        indices=z["indices"], centroids=z["centroids"], res_centroids=g("res_centroids"),
        outlier_indices=g("outlier_indices"), outlier_centroids=g("outlier_centroids"), perm=g("perm"),
        weight_scale=g("weight_scale"), weight_bias=g("weight_bias"), bias=g("bias"))
    L.meta = dict(idx=z["idx"].astype(np.int64), ridx=None if g("ridx") is None else z["ridx"].astype(np.int64))
    ref = {k: z[k] for k in ("y_ref", "W_ref", "W_ref16", "packed_ref", "u_idx")}
    ref["u_ridx"] = g("u_ridx")
    return L, z["x"], ref


def parity_error(y, y_star):
    """SURVEY.md 8(c) metric: max|y - y*| / max|y*|."""
    y = np.asarray(y, dtype=np.float64)
    y_star = np.asarray(y_star, dtype=np.float64)
    return float(np.max(np.abs(y - y_star)) / max(np.max(np.abs(y_star)), 1e-30))


# tolerance north_star states: 1e-3 relative for fp16.  bf16 outputs carry 8 mantissa bits, so one
# output rounding alone is 2^-9 = 1.95e-3 relative; the bf16 bar is 1 bf16 ulp of max|y*| = 4e-3.
TOL = {"fp16": 1e-3, "bf16": 4e-3}
