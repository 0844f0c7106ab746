"""Tensor-parallel VQuantLinear on real GPUs over NCCL (needs >= 2 GPUs; skipped otherwise)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
            sys.path.insert(0, p)
        import vptq_oracle as vo
        from _gpu import make_module, x_to_t
        from vptq_b200 import tp
        errs = []
        for kw in (dict(in_features=1024, out_features=1024, vector_len=8, num_centroids=65536, num_res_centroids=256),
                   dict(in_features=1024 + 128, out_features=512, vector_len=8, num_centroids=4096, num_res_centroids=256,
                        outlier_size=128, outlier_vector_len=4, num_outlier_centroids=4096, bias=True)):
            L = vo.make_layer(seed=5, **kw)
            full = make_module(L, f"cuda:{rank}")
            for mode in ("all_reduce", "all_gather"):
                m = tp.shard_module(full, rank, world, mode=mode)
                for tokens in (1, 2, 40):
                    x_np = vo.make_x(tokens, L.in_features, L.dtype, seed=tokens)
                    y = m(x_to_t(x_np, L, f"cuda:{rank}"))
                    torch.cuda.synchronize()
                    y_star = vo.quant_gemm(x_np, L)
                    errs.append(float(np.abs(y.float().cpu().numpy() - y_star).max() / np.abs(y_star).max()))
        q.put((rank, max(errs)))
    finally:
        dist.destroy_process_group()


def test_tp_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err <= 1e-3, (rank, err)
