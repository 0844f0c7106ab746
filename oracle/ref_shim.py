"""Import the REFERENCE's own pure-torch hot path from /root/reference (test infrastructure).

This is how the oracle is pinned: `make_golden.py` calls the reference's `pack_index`,
`unpack_index_tensor`, `dequant` and `quant_gemm` (vptq/utils/pack.py:26-139,
vptq/ops/quant_gemm.py:43-275) through this shim and commits their inputs/outputs as fixtures
under tests/golden/.  The reference tree exists only in the authoring container, so nothing
that runs on the GPU box imports this module.  It is never imported by the product path.

The reference package cannot be imported as-is (vptq/__init__.py needs installed dist metadata,
vptq/layers/model_base.py needs `accelerate`, vptq/utils/pack.py needs `accelerate` and
`sentence_transformers`), so the three missing third-party modules are stubbed and the two
files on the hot path are loaded by path under the private alias `_refvptq`-free names below.
"""
import importlib.util
import io
import contextlib
import os
import sys
import types

REF = os.environ.get("VPTQ_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "vptq", "ops", "quant_gemm.py"))


def load():
    """Returns (pack_module, quant_gemm_module) of the reference, loaded from REF."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "vptq" or k.startswith("vptq.")}
    for k in saved:
        del sys.modules[k]

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    made = []
    for name in ("accelerate", "sentence_transformers"):
        if name not in sys.modules:
            stub(name)
            made.append(name)
    if "sentence_transformers.SentenceTransformer" not in sys.modules:
        stub("sentence_transformers.SentenceTransformer",
             SentenceTransformer=type("SentenceTransformer", (), {}))
        made.append("sentence_transformers.SentenceTransformer")
    for pkg in ("vptq", "vptq.utils", "vptq.ops"):
        stub(pkg).__path__ = [os.path.join(REF, *pkg.split("."))]

    def load_file(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(m)
        return m

    try:
        pack = load_file("vptq.utils.pack", os.path.join(REF, "vptq/utils/pack.py"))
        qg = load_file("vptq.ops.quant_gemm", os.path.join(REF, "vptq/ops/quant_gemm.py"))
    finally:
        # leave no trace of the aliasing so the product package `vptq` can be imported afterwards
        for k in [k for k in sys.modules if k == "vptq" or k.startswith("vptq.")]:
            del sys.modules[k]
        for k in made:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return pack, qg
