"""world_size-2 gloo tests (CPU) of the tensor-parallel host logic: shard planner + the one
collective per layer.  The per-rank compute is the numpy ORACLE on the rank's shard (no GPU here);
the combined result must equal the unsharded oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vptq_oracle as vo
from _util import load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vptq_b200 import tp
        L, x, ref = load_golden(name)
        t = {"indices": torch.from_numpy(L.indices.copy()),
             "outlier_indices": None if L.outlier_indices is None else torch.from_numpy(L.outlier_indices.astype(np.int16)),
             "bias": None if L.bias is None else torch.from_numpy(np.asarray(L.bias).copy())}
        s = tp.shard_tensors(t, out_features=L.out_features, vector_len=L.vector_len,
                             outlier_vector_len=L.outlier_vector_len if L.enable_outlier else 1, rank=rank, world=world)
        r0, r1, o0, o1 = tp.shard_bounds(L.out_features, L.vector_len, L.outlier_vector_len if L.enable_outlier else 1, rank, world)
        Ls = vo.Layer(**{**L.__dict__})
        Ls.out_features = o1 - o0
        Ls.indices = s["indices"].numpy()
        Ls.outlier_indices = None if s["outlier_indices"] is None else s["outlier_indices"].numpy().astype(np.uint16)
        Ls.bias = None if s["bias"] is None else s["bias"].numpy()
        y_loc = torch.from_numpy(vo.quant_gemm(x, Ls))                # this rank's slice, computed by the oracle
        y = tp.combine(y_loc, L.out_features, rank, world, mode=mode)
        want = vo.quant_gemm(x, L)
        q.put((rank, float(np.abs(y.numpy() - want).max()), list(y.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["all_reduce", "all_gather"])
@pytest.mark.parametrize("name", ["v8_k256", "v8_k256_outlier", "v8_k4096_r256"])
def test_tp_world2_gloo(name, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape in res:
        assert err <= 1e-6, (rank, err)
        assert shape[-1] in (128, 64, 96)


def test_shard_bounds_errors():
    from vptq_b200 import tp
    assert tp.shard_bounds(4096, 8, 1, 3, 8) == (192, 256, 1536, 2048)
    with pytest.raises(ValueError):
        tp.shard_bounds(100, 6, 1, 0, 2)      # padding rows
    with pytest.raises(ValueError):
        tp.shard_bounds(4096, 8, 1, 0, 3)     # rows do not divide
    with pytest.raises(ValueError):
        tp.shard_bounds(36, 6, 4, 1, 2)       # boundary at output 18 splits an outlier vector of 4
