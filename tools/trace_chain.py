"""Timeline of chained list-kernel launches (CUDA graph + PDL), from %globaltimer stamps of the first and the last CTA
of every launch:  VPTQ_B200_PROF_SLOTS=<launches> python tools/trace_chain.py [layers]
Prints, per launch, the phases relative to the start of the first launch (us)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
os.environ.setdefault("VPTQ_B200_PROF_SLOTS", str(4 * layers))
import torch
import bench
from vptq_b200 import native

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = dict(bench.LLAMA3_8B, layers=layers)
stack = bench.build_stack(m, bench.QUANT, dev, 0, 1, torch.float16)
x_in, step, launches = bench.make_step(m, stack, dev, torch.float16, 0, 1, native.FLAG_PDL)
nslots = int(os.environ["VPTQ_B200_PROF_SLOTS"])
prof = torch.zeros(32 * nslots, dtype=torch.int64, device=dev)
s = torch.cuda.Stream(dev)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    x_in.copy_(torch.randn(1, m["hidden"]).half())
    step(); step()
    s.synchronize()
    native.lib().vptq_b200_debug_phase_stamps(prof.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    native.lib().vptq_b200_debug_phase_stamps(None)
    for _ in range(3):
        g.replay()
    s.synchronize()
st = prof.cpu().view(nslots, 32).tolist()
NAMES = ["start", "loads+sync", "ring,res+pdl_wait", "x issued", "x'+sync", "slice_wait", "main(w0)", "sync", "arrive", "end"]
t0 = min(r[0] for r in st if r[0])
order = sorted(range(nslots), key=lambda i: st[i][0])
prev_end = None
for i in order:
    r = st[i]
    a = [(r[k] - t0) / 1e3 if r[k] else None for k in range(10)]
    b = [(r[16 + k] - t0) / 1e3 if r[16 + k] else None for k in range(10)]
    end = max(x for x in (a[9], b[9]) if x is not None)
    gap = "" if prev_end is None else f" gap_from_prev_end={min(a[0], b[0]) - prev_end:+.2f}"
    print(f"launch slot {i}:{gap}")
    for who, v in (("cta0", a), ("ctaN", b)):
        print("   " + who + ": " + "  ".join(f"{NAMES[k]}={v[k]:.2f}" for k in range(10) if v[k] is not None))
    prev_end = end
