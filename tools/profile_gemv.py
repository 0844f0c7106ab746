"""Launch the fused GEMV once per Llama-3-8B layer shape (eagerly, distinct weights) for ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_gemv.py
    ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -o gpurun_out/prof python tools/profile_gemv.py
Also prints CUDA-event timings per shape (L2-cold: a 512 MB buffer is written between launches).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vptq_b200 import native

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = dict(bench.LLAMA3_8B, layers=1)
cfgs = {"b24_k65536_r256": bench.QUANT,
        "b16_k65536": dict(vector_len=8, num_centroids=65536, num_res_centroids=-1),
        "b21_k8192_r256": dict(vector_len=8, num_centroids=8192, num_res_centroids=256),
        "b12_k4096": dict(vector_len=8, num_centroids=4096, num_res_centroids=-1),
        "b8_k256": dict(vector_len=8, num_centroids=256, num_res_centroids=-1)}
PHASES = "--phases" in sys.argv
ONCE = "--once" in sys.argv      # one launch per shape (for ncu --set full)
only = [a for a in sys.argv[1:] if not a.startswith("--")] or list(cfgs)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
NAMES = ["start", "bar_init", "issued+staged", "phaseA", "cb_wait+replicate", "pdl_wait", "phaseB+sync", "cluster_wait",
         "main(warp0)", "leader_wait", "end"]
if os.environ.get("VPTQ_B200_LISTS", native.LISTS_DEFAULT) != "0":   # stamps of csrc/gemv_lists.cu
    NAMES = ["start", "loads+sync", "ring,res+pdl_wait", "x issued", "x'+sync", "slice_wait", "main(warp0)", "sync", "arrive",
             "end", "-"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {}
for cname in only:
    q = cfgs[cname]
    K, Kr, v = q["num_centroids"], q["num_res_centroids"], q["vector_len"]
    ib, rb = K.bit_length() - 1, (Kr.bit_length() - 1 if Kr > 0 else 0)
    g = torch.Generator(device=dev).manual_seed(1)
    for name, i, o in (("q_4096x4096", 4096, 4096), ("kv_1024x4096", 4096, 1024), ("gate_14336x4096", 4096, 14336),
                       ("down_4096x14336", 14336, 4096)):
        ro, wd = o // v, (i * (ib + rb) + 31) // 32
        t = dict(indices=torch.randint(-2 ** 31, 2 ** 31 - 1, (1, ro, wd), device=dev, dtype=torch.int32, generator=g),
                 centroids=(torch.randn(1, K * v, device=dev, generator=g) / i ** 0.5).half(),
                 res=(torch.randn(1, max(Kr, 1) * v, device=dev, generator=g) / i ** 0.5).half() if Kr > 0 else None,
                 perm=torch.randperm(i, device=dev, generator=g).to(torch.int16),
                 ws=(1 + 0.1 * torch.randn(i, device=dev, generator=g)).half(),
                 wb=(0.01 * torch.randn(i, device=dev, generator=g)).half())
        desc = native.make_desc(dtype=torch.float16, in_features=i, out_features=o, vector_len=v, num_centroids=K,
                                num_res_centroids=Kr, num_codebooks=1, group_size=i, outlier_size=0,
                                outlier_vector_len=-1, num_outlier_centroids=-1, indices=t["indices"],
                                centroids=t["centroids"], res_centroids=t["res"], outlier_indices=None,
                                outlier_centroids=None, perm=t["perm"], weight_scale=t["ws"], weight_bias=t["wb"], bias=None)
        x = torch.randn(1, i, device=dev).half()
        y = torch.empty(1, o, device=dev, dtype=torch.float16)
        native.quant_gemv(desc, x, y)
        torch.cuda.synchronize()
        if ONCE:
            continue
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); native.quant_gemv(desc, x, y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        # warm: back-to-back launches (weights of this one layer stay in L2 when they fit)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            native.quant_gemv(desc, x, y)
        e1.record(); torch.cuda.synchronize()
        if PHASES:
            flush.fill_(1)
            native.lib().vptq_b200_debug_phase_stamps(prof.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); native.quant_gemv(desc, x, y); e1.record(); torch.cuda.synchronize()
            native.lib().vptq_b200_debug_phase_stamps(None)
            st = prof.cpu().tolist()
            for base, who in ((0, "cta0"), (16, "ctaN")):
                t0 = st[base]
                print(f"   {who}: " + "  ".join(f"{NAMES[k]}={(st[base + k] - t0) / 1e3:.2f}" for k in range(1, 11) if st[base + k]),
                      f"| event {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
            prof.zero_()
        abytes = ro * wd * 4 + i * 2 + o * 2
        us = sorted(ts)[len(ts) // 2]
        out[f"{cname}/{name}"] = dict(cold_us=round(us, 2), warm_us=round(e0.elapsed_time(e1) * 1e3 / 20, 2),
                                      alg_MB=round(abytes / 1e6, 3), cold_GBps=round(abytes / us / 1e3, 1),
                                      Gfields_per_s=round(ro * i / us / 1e3, 1))
        print(f"{cname}/{name}: {out[f'{cname}/{name}']}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"gemv_shapes{os.environ.get('PROFILE_TAG', '')}.json"), "w"), indent=1)
