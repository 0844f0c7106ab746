"""One prefill configuration (BASELINE configs[2]: seq 2048 x batch 4 = 8192 tokens, 4096x4096 layer) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import vptq_oracle as vo
from _gpu import make_module
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = vo.make_layer(in_features=4096, out_features=4096, vector_len=8, num_centroids=65536, num_res_centroids=256, seed=1)
m = make_module(L)
x = torch.randn(T, 4096, device="cuda").half()
for _ in range(4):
    y = m(x)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
