"""Developer aid: per-warp end times of the list kernel's main loop (first CTA of every launch of two decoder layers,
inside a CUDA graph).  Needs a library built with VPTQ_B200_PROF_WARPS=1 python -m vptq_b200.build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VPTQ_B200_PROF_SLOTS"] = "8"
import torch, bench
from vptq_b200 import native
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
m = dict(bench.LLAMA3_8B, layers=2)
stack = bench.build_stack(m, bench.QUANT, dev, 0, 1, torch.float16)
x_in, step, launches = bench.make_step(m, stack, dev, torch.float16, 0, 1, native.FLAG_PDL)
prof = torch.zeros(32 * 8, dtype=torch.int64, device=dev)
s = torch.cuda.Stream(dev); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    x_in.copy_(torch.randn(1, 4096).half()); step(); step(); s.synchronize()
    native.lib().vptq_b200_debug_phase_stamps(prof.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s): step()
    native.lib().vptq_b200_debug_phase_stamps(None)
    for _ in range(3): g.replay()
    s.synchronize()
st = prof.cpu().view(8, 32).tolist()
for i in sorted(range(8), key=lambda i: st[i][0]):
    r = st[i]; t0 = r[5]   # slice_wait = main loop start
    print("launch", i, "main start->warp ends (us):", " ".join(f"{(r[16+w]-t0)/1e3:.2f}" for w in range(16)), "| sync", f"{(r[7]-t0)/1e3:.2f}")
