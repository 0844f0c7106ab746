"""Parity + timing of the sliced decode kernel's reduction variants (run on a B200).

    python tools/check_sliced_variants.py            # spawns one child per variant
    VPTQ_B200_GEMV_TUNE=sliced=2 python tools/check_sliced_variants.py --child

sliced=1: thread-block clusters + st.async (default, validated in tests/test_gpu_sliced.py)
sliced=2: independent CTAs + last-arriver reduction through global memory (experimental; written at the
          end of round 1 without GPU time left to run it -- this script is its first test)
The tuning knob is read once per process, hence the children.
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import numpy as np, torch
    import vptq_oracle as vo
    from _gpu import from_t, tdtype, to_t, x_to_t
    from _util import TOL, parity_error
    from vptq_b200 import native
    out = {"tune": os.environ.get("VPTQ_B200_GEMV_TUNE", "")}

    def desc(L):
        t = dict(indices=to_t(L.indices, L, "i32"), centroids=to_t(L.centroids, L),
                 res_centroids=to_t(L.res_centroids, L) if L.res_bits else None, perm=to_t(L.perm, L, "u16"),
                 weight_scale=to_t(L.weight_scale, L), weight_bias=to_t(L.weight_bias, L), bias=to_t(L.bias, L))
        d = native.make_desc(dtype=tdtype(L), in_features=L.in_features, out_features=L.out_features, vector_len=8,
                             num_centroids=L.num_centroids, num_res_centroids=L.num_res_centroids, num_codebooks=1,
                             group_size=L.group_size, outlier_size=0, outlier_vector_len=-1, num_outlier_centroids=-1,
                             outlier_indices=None, outlier_centroids=None, sliced=True, **t)
        d._t = t
        return d

    cases = {"k65536_r256": dict(in_features=2048, out_features=1024, num_centroids=65536, num_res_centroids=256),
             "k32768_plain": dict(in_features=1024, out_features=256, num_centroids=32768, enable_perm=False, enable_norm=False),
             "k65536_bf16_bias": dict(in_features=1024, out_features=2040, num_centroids=65536, num_res_centroids=16,
                                      dtype="bf16", bias=True)}
    for name, kw in cases.items():
        L = vo.make_layer(vector_len=8, seed=41, **kw)
        x_np = vo.make_x(1, L.in_features, L.dtype, seed=7)
        x = x_to_t(x_np, L)
        d = desc(L)
        ys = []
        for _ in range(3):                                   # repeated: the counters must return to zero
            y = torch.full((1, L.out_features), float("nan"), dtype=x.dtype, device=x.device)
            native.quant_gemv(d, x, y)
            torch.cuda.synchronize()
            ys.append(y)
        err = parity_error(from_t(ys[0]), vo.quant_gemm(x_np, L))
        out[name] = dict(err=err, ok=bool(err <= TOL[L.dtype] and all(torch.equal(ys[0], y) for y in ys[1:])))
    # fused q+k+v launch + timing of the Llama shapes
    shapes = [(2048, 2048), (2048, 512), (2048, 520)]
    Ls = [vo.make_layer(in_features=i, out_features=o, vector_len=8, num_centroids=65536, num_res_centroids=256,
                        seed=50 + k) for k, (i, o) in enumerate(shapes)]
    ds = [desc(L) for L in Ls]
    x_np = vo.make_x(1, 2048, "fp16", seed=9)
    x = x_to_t(x_np, Ls[0])
    ys = [torch.full((1, L.out_features), float("nan"), dtype=x.dtype, device=x.device) for L in Ls]
    fused = native.FusedGemv(ds, ys)
    fused(x); fused(x)
    torch.cuda.synchronize()
    out["fused"] = dict(separate=fused.separate,
                        ok=all(parity_error(from_t(y), vo.quant_gemm(x_np, L)) <= TOL[L.dtype] for L, y in zip(Ls, ys)))
    g = torch.Generator(device="cuda").manual_seed(1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name, i, o in (("q_4096x4096", 4096, 4096), ("gate_14336x4096", 4096, 14336), ("down_4096x14336", 14336, 4096)):
        L = vo.make_layer(in_features=i, out_features=o, vector_len=8, num_centroids=65536, num_res_centroids=256, seed=3)
        d = desc(L)
        x = x_to_t(vo.make_x(1, i, "fp16"), L)
        y = torch.empty(1, o, dtype=x.dtype, device=x.device)
        native.quant_gemv(d, x, y)
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); native.quant_gemv(d, x, y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out[name + "_cold_us"] = round(sorted(ts)[2], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for tune in ("sliced=1", "sliced=2"):
            env = dict(os.environ, VPTQ_B200_GEMV_TUNE=tune)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True,
                               text=True, timeout=600)
            print(tune, "rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1], r.stderr[-800:] if r.returncode else "")
