"""ctypes binding of libvptq_b200.so (the C ABI declared in include/vptq_b200.h).

This is the only bridge between the Python surface and the CUDA kernels.  There is no other
implementation behind it: if the shared library is missing or a call fails, a RuntimeError is
raised -- no CPU path, no torch fallback (the reference silently falls back to a torch
implementation when its extension is absent, vptq/ops/quant_gemm.py:20-40; this package does not).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

from . import lists as _lists

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPTQ_B200_LIB") or os.path.join(_HERE, "libvptq_b200.so")   # (env: developer builds)

VPTQ_FP16, VPTQ_BF16 = 0, 1
OP_GEMV, OP_DEQUANT, OP_GEMM, OP_GEMV_V2 = 0, 1, 2, 3
FLAG_PDL = 1
TP_PLAIN, TP_TAGGED = 0, 1
ABI_VERSION = 6
LISTS_DEFAULT = "1"   # VPTQ_B200_LISTS when unset

EXPORTS = (
    "vptq_b200_abi_version", "vptq_b200_last_error", "vptq_b200_workspace_bytes", "vptq_b200_quant_gemv",
    "vptq_b200_dequant", "vptq_b200_quant_gemm", "vptq_b200_quant_gemv_v2", "vptq_b200_linear_host",
    "vptq_b200_debug_phase_stamps", "vptq_b200_quant_gemv_multi", "vptq_b200_quant_gemv_multi_tp",
    "vptq_b200_lists_build_host", "vptq_b200_quant_gemv_multi_ws", "vptq_b200_tp_untag", "vptq_b200_lists_deal_host",
)

MAX_FUSED, MAX_RANKS = 4, 8


class TpExchange(ctypes.Structure):
    """struct vptq_tp_exchange (include/vptq_b200.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("world", ctypes.c_int32), ("rank", ctypes.c_int32),
        ("slot", ctypes.c_int32), ("wait_slot", ctypes.c_int32),
        ("peer_y", (ctypes.c_void_p * MAX_RANKS) * MAX_FUSED), ("peer_flags", ctypes.c_void_p * MAX_RANKS),
        ("epoch", ctypes.c_void_p), ("done", ctypes.c_void_p), ("error", ctypes.c_void_p),
        ("format", ctypes.c_int32), ("num_slots", ctypes.c_int32),
    ]


class LinearDesc(ctypes.Structure):
    """struct vptq_linear_desc (include/vptq_b200.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("dtype", ctypes.c_int32),
        ("in_features", ctypes.c_int32), ("out_features", ctypes.c_int32),
        ("vector_len", ctypes.c_int32), ("num_centroids", ctypes.c_int32),
        ("num_res_centroids", ctypes.c_int32), ("num_codebooks", ctypes.c_int32),
        ("group_size", ctypes.c_int32), ("outlier_size", ctypes.c_int32),
        ("outlier_vector_len", ctypes.c_int32), ("num_outlier_centroids", ctypes.c_int32),
        ("indices", ctypes.c_void_p), ("index_stride_codebook", ctypes.c_int64),
        ("index_stride_row", ctypes.c_int64),
        ("centroids", ctypes.c_void_p), ("centroid_stride", ctypes.c_int64),
        ("res_centroids", ctypes.c_void_p), ("res_centroid_stride", ctypes.c_int64),
        ("outlier_indices", ctypes.c_void_p), ("outlier_centroids", ctypes.c_void_p),
        ("perm", ctypes.c_void_p), ("weight_scale", ctypes.c_void_p), ("weight_bias", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("weight_scale_q", ctypes.c_void_p), ("weight_bias_q", ctypes.c_void_p),
        ("lists_stream", ctypes.c_void_p), ("lists_tab", ctypes.c_void_p),
        ("lists_tile_cols", ctypes.c_int32), ("lists_reserved", ctypes.c_int32),
    ]


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m vptq_b200.build` "
                "(vptq_b200 has no CPU or torch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.vptq_b200_abi_version.restype = ctypes.c_int
        L.vptq_b200_last_error.restype = ctypes.c_char_p
        L.vptq_b200_workspace_bytes.restype = ctypes.c_size_t
        L.vptq_b200_workspace_bytes.argtypes = [ctypes.POINTER(LinearDesc), ctypes.c_int32, ctypes.c_int32]
        vp, i64, i32, u32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_size_t
        dp = ctypes.POINTER(LinearDesc)
        L.vptq_b200_quant_gemv.argtypes = [dp, vp, i64, vp, i64, i32, vp, sz, u32, vp]
        L.vptq_b200_quant_gemm.argtypes = [dp, vp, i64, vp, i64, i32, vp, sz, u32, vp]
        L.vptq_b200_dequant.argtypes = [dp, vp, vp, sz, vp]
        L.vptq_b200_quant_gemv_v2.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp,
                                              vp, vp, vp, sz, u32, vp]
        L.vptq_b200_linear_host.argtypes = [dp, vp, vp, i32, vp, vp, vp, sz, u32, vp]
        L.vptq_b200_quant_gemv_multi.argtypes = [i32, ctypes.POINTER(dp), vp, i64, ctypes.POINTER(vp),
                                                 ctypes.POINTER(i64), i32, u32, vp]
        L.vptq_b200_quant_gemv_multi.restype = ctypes.c_int
        L.vptq_b200_quant_gemv_multi_ws.argtypes = [i32, ctypes.POINTER(dp), vp, i64, ctypes.POINTER(vp),
                                                    ctypes.POINTER(i64), i32, vp, sz, u32, vp]
        L.vptq_b200_quant_gemv_multi_ws.restype = ctypes.c_int
        L.vptq_b200_quant_gemv_multi_tp.argtypes = [i32, ctypes.POINTER(dp), vp, i64, ctypes.POINTER(vp),
                                                    ctypes.POINTER(i64), i32, ctypes.POINTER(TpExchange), vp, sz, u32, vp]
        L.vptq_b200_quant_gemv_multi_tp.restype = ctypes.c_int
        L.vptq_b200_lists_build_host.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp, sz, vp, ctypes.POINTER(sz),
                                                 ctypes.POINTER(i32)]
        L.vptq_b200_lists_build_host.restype = ctypes.c_int
        L.vptq_b200_lists_deal_host.argtypes = [vp, vp, i64, i32]
        L.vptq_b200_lists_deal_host.restype = ctypes.c_int
        L.vptq_b200_tp_untag.argtypes = [vp, vp, i32, ctypes.POINTER(TpExchange), vp]
        L.vptq_b200_tp_untag.restype = ctypes.c_int
        L.vptq_b200_debug_phase_stamps.argtypes = [vp]
        L.vptq_b200_debug_phase_stamps.restype = None
        for f in ("vptq_b200_quant_gemv", "vptq_b200_quant_gemm", "vptq_b200_dequant", "vptq_b200_quant_gemv_v2",
                  "vptq_b200_linear_host"):
            getattr(L, f).restype = ctypes.c_int
        if L.vptq_b200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libvptq_b200.so ABI {L.vptq_b200_abi_version()} != expected {ABI_VERSION}")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().vptq_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return VPTQ_FP16
    if dt == torch.bfloat16:
        return VPTQ_BF16
    raise RuntimeError(f"vptq_b200 supports float16 and bfloat16 tensors only, got {dt}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def require_cuda(name: str, t: Optional[torch.Tensor], contiguous: bool = True) -> None:
    """The reference's CHECK_INPUT (csrc/util/common.h:11-19): CUDA + contiguous."""
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (vptq_b200 has no CPU path)")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def make_desc(*, dtype: torch.dtype, in_features: int, out_features: int, vector_len: int, num_centroids: int,
              num_res_centroids: int, num_codebooks: int, group_size: int, outlier_size: int,
              outlier_vector_len: int, num_outlier_centroids: int, indices: torch.Tensor,
              centroids: torch.Tensor, res_centroids: Optional[torch.Tensor],
              outlier_indices: Optional[torch.Tensor], outlier_centroids: Optional[torch.Tensor],
              perm: Optional[torch.Tensor], weight_scale: Optional[torch.Tensor],
              weight_bias: Optional[torch.Tensor], bias: Optional[torch.Tensor],
              derive: bool = True, lists: Optional[bool] = None, drop_packed: bool = False) -> LinearDesc:
    """Describe one layer's tensors for the C ABI.  Tensors are borrowed: keep them alive.

    With `derive` (default) the load-time derivatives the ABI accepts are built here, once:
    weight_scale / weight_bias in quantised column order (`t[perm]`).  They hang off the returned
    descriptor (`desc._keep`) so they live as long as it does.

    `lists`: also build the slice x tile index lists (vptq_b200.lists) that let single-token calls of
    large-codebook layers gather from shared memory; costs 4 bytes per index (+ ~6 % padding) on top
    of the packed words (3 bytes per index for the 65536+256 configuration).  None = the
    VPTQ_B200_LISTS environment variable (default on); ignored for layers the list kernel does not cover.

    `drop_packed`: decode-only descriptor -- once the lists are built the descriptor forgets the packed words
    (`indices` = NULL), so the caller may free them: 4.2 instead of 7.2 bytes per index.  Multi-token calls,
    dequant and prefill then return VPTQ_ERR_UNSUPPORTED.  Refused for layers without lists.
    """
    if indices.dtype != torch.int32:
        raise RuntimeError("`indices` must be packed int32 words (is_indice_packed=True); "
                           "see vptq_b200.pack.pack_index")
    for n, t in (("indices", indices), ("centroids", centroids), ("res_centroids", res_centroids),
                 ("outlier_indices", outlier_indices), ("outlier_centroids", outlier_centroids), ("perm", perm),
                 ("weight_scale", weight_scale), ("weight_bias", weight_bias), ("bias", bias)):
        require_cuda(n, t)
    if indices.dim() != 3:
        raise RuntimeError(f"indices must be [num_codebooks, num_indices, packed_groupsize], got {tuple(indices.shape)}")
    use_outlier = outlier_indices is not None and outlier_centroids is not None and outlier_size > 0
    d = LinearDesc()
    d.struct_size = ctypes.sizeof(LinearDesc)
    d.dtype = dtype_code(dtype)
    d.in_features, d.out_features = int(in_features), int(out_features)
    d.vector_len, d.num_centroids = int(vector_len), int(num_centroids)
    d.num_res_centroids = int(num_res_centroids) if res_centroids is not None else -1
    d.num_codebooks, d.group_size = int(num_codebooks), int(group_size)
    d.outlier_size = int(outlier_size) if use_outlier else 0
    d.outlier_vector_len = int(outlier_vector_len)
    d.num_outlier_centroids = int(num_outlier_centroids)
    d.indices = _ptr(indices)
    d.index_stride_codebook, d.index_stride_row = indices.stride(0), indices.stride(1)
    d.centroids = _ptr(centroids)
    d.centroid_stride = int(num_centroids) * int(vector_len)
    d.res_centroids = _ptr(res_centroids)
    d.res_centroid_stride = int(num_res_centroids) * int(vector_len) if res_centroids is not None else 0
    d.outlier_indices = _ptr(outlier_indices) if use_outlier else None
    d.outlier_centroids = _ptr(outlier_centroids) if use_outlier else None
    d.perm = _ptr(perm)
    d.weight_scale, d.weight_bias = _ptr(weight_scale), _ptr(weight_bias)
    d.bias = _ptr(bias)
    d._keep = ()
    if derive and perm is not None and weight_scale is not None and weight_bias is not None:
        pidx = perm.view(torch.uint16).to(torch.int64) if perm.dtype in (torch.int16, torch.uint16) else perm.long()
        ws_q, wb_q = weight_scale[pidx].contiguous(), weight_bias[pidx].contiguous()
        d.weight_scale_q, d.weight_bias_q = ws_q.data_ptr(), wb_q.data_ptr()
        d._keep = (ws_q, wb_q)
    if lists is None:
        lists = os.environ.get("VPTQ_B200_LISTS", LISTS_DEFAULT) != "0"
    if lists and derive and _lists.eligible(
            vector_len=d.vector_len, num_centroids=d.num_centroids, num_res_centroids=d.num_res_centroids,
            num_codebooks=d.num_codebooks, outlier_size=d.outlier_size, in_features=d.in_features) and (
            weight_scale is None or (weight_scale.data_ptr() % 16 == 0 and weight_bias.data_ptr() % 16 == 0)):
        stream, tab, tcw = _lists.build_lists(indices, num_centroids=d.num_centroids,
                                              num_res_centroids=d.num_res_centroids, in_features=d.in_features,
                                              out_features=d.out_features, perm=perm)
        d.lists_stream, d.lists_tab, d.lists_tile_cols = stream.data_ptr(), tab.data_ptr(), int(tcw)
        d._keep = d._keep + (stream, tab)
    if drop_packed:
        if not d.lists_stream:
            raise RuntimeError("drop_packed: this layer has no index lists (not eligible, or lists=False)")
        d.indices = None
    return d


# one zero-initialised workspace per (device, stream): kernels leave it zeroed (include/vptq_b200.h)
_workspaces: dict = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def workspace_bytes(desc: LinearDesc, tokens: int, op: int) -> int:
    return int(lib().vptq_b200_workspace_bytes(ctypes.byref(desc), int(tokens), int(op)))


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs microseconds on
    a path whose kernel takes ten)."""

    def __init__(self, dev: torch.device):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _ws_bytes_cached(desc: LinearDesc, tokens: int, op: int) -> int:
    cache = desc.__dict__.setdefault("_ws_cache", {})
    key = (min(tokens, 4) if op == OP_GEMV else tokens, op)     # the GEMV plans passes of at most 4 tokens
    v = cache.get(key)
    if v is None:
        v = cache[key] = workspace_bytes(desc, tokens, op)
    return v


def quant_gemv(desc: LinearDesc, x2d: torch.Tensor, y2d: torch.Tensor, flags: int = 0) -> None:
    dev = x2d.device
    tokens = x2d.shape[0]
    with _on_device(dev):
        ws = workspace(dev, _ws_bytes_cached(desc, tokens, OP_GEMV))
        rc = lib().vptq_b200_quant_gemv(ctypes.byref(desc), x2d.data_ptr(), x2d.stride(0), y2d.data_ptr(),
                                        y2d.stride(0), tokens, ws.data_ptr(), ws.numel(), flags, _stream(dev))
    check(rc, "vptq_b200_quant_gemv")


class FusedGemv:
    """Prepared argument arrays for vptq_b200_quant_gemv_multi (layers sharing one input)."""

    def __init__(self, descs, ys):
        n = len(descs)
        if not 1 <= n <= MAX_FUSED or len(ys) != n:
            raise RuntimeError(f"FusedGemv takes 1..{MAX_FUSED} layers with one output each")
        if any(d.in_features != descs[0].in_features or d.dtype != descs[0].dtype for d in descs):
            raise RuntimeError("FusedGemv: the layers must read the same x (same in_features and dtype)")
        self.n, self.descs, self.ys = n, list(descs), list(ys)
        self.desc_arr = (ctypes.POINTER(LinearDesc) * n)(*[ctypes.pointer(d) for d in descs])
        self.y_arr = (ctypes.c_void_p * n)(*[y.data_ptr() for y in ys])
        self.stride_arr = (ctypes.c_int64 * n)(*[y.stride(0) for y in ys])
        self.separate = False
        self.ws_bytes = None

    def __call__(self, x2d: torch.Tensor, flags: int = 0) -> None:
        dev = x2d.device
        if not self.separate:
            with torch.cuda.device(dev):
                if self.ws_bytes is None:
                    self.ws_bytes = sum(workspace_bytes(d, x2d.shape[0], OP_GEMV) for d in self.descs)
                ws = workspace(dev, self.ws_bytes)   # the list kernel reduces its partial sums through it
                rc = lib().vptq_b200_quant_gemv_multi_ws(self.n, self.desc_arr, x2d.data_ptr(), x2d.stride(0),
                                                         self.y_arr, self.stride_arr, x2d.shape[0], ws.data_ptr(),
                                                         ws.numel(), flags, _stream(dev))
            if rc != -2:                      # VPTQ_ERR_UNSUPPORTED: these layers cannot share one launch
                check(rc, "vptq_b200_quant_gemv_multi")
                return
            self.separate = True              # same kernels, one launch per layer (identical results)
        for d, y in zip(self.descs, self.ys):
            quant_gemv(d, x2d, y, flags)


class FusedGemvTP:
    """vptq_b200_quant_gemv_multi_tp: fused layers + the tensor-parallel exchange inside the kernel."""

    def __init__(self, descs, ys, exchange: TpExchange):
        n = len(descs)
        self.n, self.descs, self.ys, self.ex = n, list(descs), list(ys), exchange
        self.desc_arr = (ctypes.POINTER(LinearDesc) * n)(*[ctypes.pointer(d) for d in descs])
        self.y_arr = (ctypes.c_void_p * n)(*[y.data_ptr() for y in ys])
        self.stride_arr = (ctypes.c_int64 * n)(*[y.stride(0) for y in ys])
        self.ws_bytes = None

    def __call__(self, x2d: torch.Tensor, flags: int = 0) -> None:
        dev = x2d.device
        with torch.cuda.device(dev):
            if self.ws_bytes is None:
                self.ws_bytes = sum(workspace_bytes(d, x2d.shape[0], OP_GEMV) for d in self.descs)
            ws = workspace(dev, self.ws_bytes)
            rc = lib().vptq_b200_quant_gemv_multi_tp(self.n, self.desc_arr, x2d.data_ptr(), x2d.stride(0), self.y_arr,
                                                     self.stride_arr, x2d.shape[0], ctypes.byref(self.ex), ws.data_ptr(),
                                                     ws.numel(), flags, _stream(dev))
        check(rc, "vptq_b200_quant_gemv_multi_tp")


def tp_untag(tagged: torch.Tensor, y: torch.Tensor, exchange: TpExchange) -> None:
    """vptq_b200_tp_untag: the full-width output of the (VPTQ_TP_TAGGED) launch `exchange` -> plain values in y."""
    dev = y.device
    with torch.cuda.device(dev):
        rc = lib().vptq_b200_tp_untag(tagged.data_ptr(), y.data_ptr(), y.numel(), ctypes.byref(exchange), _stream(dev))
    check(rc, "vptq_b200_tp_untag")


def quant_gemm(desc: LinearDesc, x2d: torch.Tensor, y2d: torch.Tensor, flags: int = 0) -> None:
    dev = x2d.device
    tokens = x2d.shape[0]
    with _on_device(dev):
        ws = workspace(dev, _ws_bytes_cached(desc, tokens, OP_GEMM))
        rc = lib().vptq_b200_quant_gemm(ctypes.byref(desc), x2d.data_ptr(), x2d.stride(0), y2d.data_ptr(),
                                        y2d.stride(0), tokens, ws.data_ptr(), ws.numel(), flags, _stream(dev))
    check(rc, "vptq_b200_quant_gemm")


def dequant(desc: LinearDesc, w_out: torch.Tensor) -> None:
    dev = w_out.device
    with torch.cuda.device(dev):
        ws = workspace(dev, workspace_bytes(desc, 1, OP_DEQUANT))
        rc = lib().vptq_b200_dequant(ctypes.byref(desc), w_out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev))
    check(rc, "vptq_b200_dequant")
