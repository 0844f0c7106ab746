"""Single-GPU check of the tensor-parallel shard shapes: for world in (2, 4, 8), rank 0's shard of every Llama-3-8B
linear (and the fused q+k+v / gate+up launches) through the list kernel vs the generic kernel on the same tensors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vptq_b200 import native

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m, q = dict(bench.LLAMA3_8B), bench.QUANT
dt = torch.float16
bad = 0
for world in (1, 2, 4, 8):
    layer = {}
    for name, i, o in bench.model_linears(m):
        o_loc = o // world
        t = bench.layer_tensors(m, q, 0, name, i, o, dev, dt, rows=(0, o_loc // 8))
        layer[name] = (t, bench.make_layer_desc(q, t, i, o_loc, dt, lists=True), bench.make_layer_desc(q, t, i, o_loc, dt, lists=False), i, o_loc)
    for group in (("q",), ("k",), ("v",), ("o",), ("gate",), ("up",), ("down",), ("q", "k", "v"), ("gate", "up")):
        i = layer[group[0]][3]
        x = torch.randn(1, i, device=dev).to(dt)
        ys_l = [torch.full((1, layer[n][4]), float("nan"), device=dev, dtype=dt) for n in group]
        ys_g = [torch.full((1, layer[n][4]), float("nan"), device=dev, dtype=dt) for n in group]
        fl = native.FusedGemv([layer[n][1] for n in group], ys_l)
        fg = native.FusedGemv([layer[n][2] for n in group], ys_g)
        for rep in range(3):
            fl(x, native.FLAG_PDL); fg(x)
        torch.cuda.synchronize()
        for n, a, b in zip(group, ys_l, ys_g):
            err = float((a.float() - b.float()).abs().max() / b.float().abs().max())
            flag = "" if err < 3e-3 else "   <-- MISMATCH"
            bad += err >= 3e-3 or not torch.isfinite(a).all()
            print(f"world {world} {'+'.join(group):10s} {n:5s} rows {layer[n][4] // 8:5d}  sep={fl.separate}  rel err {err:.2e}{flag}")
print("BAD" if bad else "ALL OK")
