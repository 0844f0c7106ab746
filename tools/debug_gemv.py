"""Ad-hoc GPU debugging aid: bench-shaped layers against the oracle, singly and chained."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import vptq_oracle as vo
from _gpu import from_t, make_module, x_to_t
from vptq_b200 import native

def report(name, y, y_star, v):
    e = np.abs(y - y_star)[0]; scale = np.abs(y_star).max()
    bad = np.nonzero(~np.isfinite(y[0]) | (e > 1e-3 * scale))[0]
    rows = np.unique(bad // v)
    print(f"== {name}: max rel err {np.nanmax(e)/scale:.3e} finite={np.isfinite(y).all()} bad {len(bad)}/{e.size} rows(first 20) {rows[:20].tolist()} n={len(rows)}", flush=True)

shapes = {"gate_14336x4096": (4096, 14336), "down_4096x14336": (14336, 4096), "q": (4096, 4096), "kv": (4096, 1024)}
mods = {}
for name, (i, o) in shapes.items():
    L = vo.make_layer(in_features=i, out_features=o, vector_len=8, num_centroids=65536, num_res_centroids=256, seed=11)
    m = make_module(L); mods[name] = (L, m)
    x_np = vo.make_x(1, i, L.dtype, seed=8)
    y = from_t(m(x_to_t(x_np, L))); torch.cuda.synchronize()
    report(name, y, vo.quant_gemm(x_np, L), 8)

# chain q -> gate -> down, back-to-back on one stream with and without PDL, eager and in a graph
Lq, mq = mods["q"]; Lg, mg = mods["gate_14336x4096"]; Ld, md = mods["down_4096x14336"]
x_np = vo.make_x(1, 4096, "fp16", seed=3)
x = x_to_t(x_np, Lq)
def oracle_chain():
    a = vo.quant_gemm(x_np, Lq).astype(np.float16)
    b = vo.quant_gemm(a, Lg).astype(np.float16)
    return vo.quant_gemm(b, Ld)
want = oracle_chain()
bufs = [torch.empty(1, 4096, device="cuda", dtype=torch.float16), torch.empty(1, 14336, device="cuda", dtype=torch.float16),
        torch.empty(1, 4096, device="cuda", dtype=torch.float16)]
for flags in (0, native.FLAG_PDL):
    def chain():
        native.quant_gemv(mq._desc_cache[0], x, bufs[0], flags=flags)
        native.quant_gemv(mg._desc_cache[0], bufs[0], bufs[1], flags=flags)
        native.quant_gemv(md._desc_cache[0], bufs[1], bufs[2], flags=flags)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for b in bufs: b.fill_(float("nan"))
        chain(); s.synchronize()
        report(f"chain eager flags={flags}", from_t(bufs[2]), want, 8)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            chain()
        for b in bufs: b.fill_(float("nan"))
        for _ in range(5): g.replay()
        s.synchronize()
        report(f"chain graph x5 flags={flags}", from_t(bufs[2]), want, 8)
