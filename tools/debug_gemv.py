"""Ad-hoc GPU debugging aid: per-row error map of the fused GEMV against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import vptq_oracle as vo
from _gpu import from_t, make_module, x_to_t

CASES = {
    "v4_k4096": dict(in_features=1024, out_features=512, vector_len=4, num_centroids=4096),
    "noperm": dict(in_features=1024, out_features=1024, vector_len=8, num_centroids=256, enable_perm=False, enable_norm=False, bias=True),
    "k65536_r256": dict(in_features=2048, out_features=1024, vector_len=8, num_centroids=65536, num_res_centroids=256),
    "cfg1": dict(in_features=4096, out_features=4096, vector_len=8, num_centroids=256),
}
for name, kw in CASES.items():
    L = vo.make_layer(seed=4321, **kw)
    m = make_module(L)
    x_np = vo.make_x(1, L.in_features, L.dtype, seed=8)
    x = x_to_t(x_np, L)
    ys = [from_t(m(x)) for _ in range(3)]
    torch.cuda.synchronize()
    y_star = vo.quant_gemm(x_np, L)
    e = np.abs(ys[0] - y_star)[0]
    scale = np.abs(y_star).max()
    v = L.vector_len
    rows_bad = np.unique(np.nonzero(e > 1e-3 * scale)[0] // v)
    print(f"== {name}: max rel err {e.max()/scale:.3e}; deterministic={all(np.array_equal(ys[0], y) for y in ys)}; "
          f"bad outputs {int((e > 1e-3*scale).sum())}/{e.size}; bad rows (first 40) {rows_bad[:40].tolist()} n={len(rows_bad)}")
    bad = np.nonzero(e > 1e-3 * scale)[0][:8]
    for o in bad:
        print(f"   o={o} row={o//v} got {ys[0][0,o]:.5f} want {y_star[0,o]:.5f} diff {ys[0][0,o]-y_star[0,o]:+.5f}")
