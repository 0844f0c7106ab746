"""Exercise bench.py's tensor-parallel step on ONE GPU (collectives stubbed) to catch host-side errors cheaply."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from vptq_b200 import native
dist.all_reduce = lambda t, **kw: None
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
for world in (2, 4, 8):
    for rank in (0, world - 1):
        m = dict(bench.LLAMA3_8B, layers=1)
        stack = bench.build_stack(m, bench.QUANT, dev, rank, world, torch.float16)
        x_in, step, launches = bench.make_step(m, stack, dev, torch.float16, rank, world, 0)
        x_in.copy_(torch.randn(1, 4096).half())
        h = step(); torch.cuda.synchronize()
        print(f"world {world} rank {rank}: ok, launches {launches[0]}, finite {bool(torch.isfinite(h.float()).all())}", flush=True)
