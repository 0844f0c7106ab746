"""TEST INFRASTRUCTURE.  Writes tests/golden/state_dict_manifest.json: for a set of constructor configurations, the
state_dict of the REFERENCE module (vptq/layers/vqlinear.py:17-240, imported from /root/reference through the stubs
of oracle/ref_shim.py) as {key: [shape, dtype]}.  tests/test_host_logic.py holds vptq_b200.VQuantLinear to it and
round-trips our state_dict through safetensors.  Runs only where /root/reference exists (authoring container)."""
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

CONFIGS = {
    "llama3_packed": dict(in_features=4096, out_features=1024, vector_lens=[-1, 8], num_centroids=[-1, 65536],
                          num_res_centroids=[-1, 256], group_num=1, group_size=4096, outlier_size=0, indices_as_float=False,
                          enable_norm=True, enable_perm=True, is_indice_packed=True, bias=False),
    "outliers_bias": dict(in_features=2048 + 128, out_features=1000, vector_lens=[4, 8], num_centroids=[4096, 4096],
                          num_res_centroids=[-1, 256], group_num=1, group_size=2048, outlier_size=128, indices_as_float=False,
                          enable_norm=True, enable_perm=True, is_indice_packed=True, bias=True),
    "groups_noperm": dict(in_features=1024, out_features=512, vector_lens=[-1, 6], num_centroids=[-1, 1024],
                          num_res_centroids=[-1, -1], group_num=4, group_size=256, outlier_size=0, indices_as_float=False,
                          enable_norm=False, enable_perm=False, is_indice_packed=True, bias=False),
    "unpacked": dict(in_features=512, out_features=256, vector_lens=[-1, 8], num_centroids=[-1, 256],
                     num_res_centroids=[-1, -1], group_num=1, group_size=512, outlier_size=0, indices_as_float=False,
                     enable_norm=True, enable_perm=True, is_indice_packed=False, bias=False),
}


def reference_module_class():
    pack, qg = ref_shim.load()
    saved = {k: sys.modules.get(k) for k in ("vptq", "vptq.ops")}
    pkg = types.ModuleType("vptq")
    pkg.ops = qg
    sys.modules["vptq"], sys.modules["vptq.ops"] = pkg, qg
    try:
        spec = importlib.util.spec_from_file_location("_ref_vqlinear", os.path.join(ref_shim.REF, "vptq/layers/vqlinear.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.VQuantLinear


def main():
    VQ = reference_module_class()
    out = {}
    for name, kw in CONFIGS.items():
        m = VQ(**kw, dtype=torch.float16, device="cpu", enable_proxy_error=False)
        out[name] = {"kwargs": kw, "state": {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}}
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "state_dict_manifest.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, {k: len(v["state"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
