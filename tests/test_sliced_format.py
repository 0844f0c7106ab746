"""Sliced index lists (vptq_b200/sliced.py): the format the sliced-codebook decode kernel reads.

CPU only: the builder is checked against the oracle through a float64 evaluation of the lists, and the
kernel's work partition arithmetic (csrc/gemv_sliced.cu) is mirrored in integers."""
import numpy as np
import pytest
import torch

from oracle import vptq_oracle as vo
from vptq_b200 import sliced

CASES = [
    dict(I=1024, O=264, K=65536, Kr=256),   # the Llama-3 2-bit configuration, small
    dict(I=512, O=64, K=16384, Kr=16),      # two slices, small residual codebook
    dict(I=768, O=44, K=32768, Kr=-1),      # four slices, no residual, ragged last row
]


def _build(c, dtype="fp16"):
    L = vo.make_layer(c["I"], c["O"], vector_len=8, num_centroids=c["K"], num_res_centroids=c["Kr"], dtype=dtype,
                      seed=11)
    ind = torch.from_numpy(np.ascontiguousarray(L.indices))
    stream, offs = sliced.build_sliced(ind, num_centroids=c["K"], num_res_centroids=c["Kr"], group_size=c["I"],
                                       out_features=c["O"])
    return L, stream, offs


@pytest.mark.parametrize("c", CASES)
def test_lists_reproduce_the_oracle(c):
    L, stream, offs = _build(c)
    x = vo.make_x(1, c["I"], "fp16")
    xf = vo.to_f32(x, "fp16").astype(np.float64).reshape(-1)
    perm = np.asarray(L.perm).astype(np.uint16).astype(np.int64)
    sc = vo.to_f32(L.weight_scale, "fp16").astype(np.float64)
    wb = vo.to_f32(L.weight_bias, "fp16").astype(np.float64)
    xq = xf[perm] * sc[perm]
    C = torch.from_numpy(vo.to_f32(L.centroids, "fp16"))
    R = None if c["Kr"] <= 0 else torch.from_numpy(vo.to_f32(L.res_centroids, "fp16"))
    y = sliced.emulate(stream, offs, num_centroids=c["K"], num_res_centroids=c["Kr"], group_size=c["I"],
                       out_features=c["O"], centroids=C, res_centroids=R, xq=torch.from_numpy(xq)).numpy()
    y = y[:c["O"]] + float((xf * wb).sum())
    if L.bias is not None:
        y = y + vo.to_f32(L.bias, "fp16").astype(np.float64)
    y_star = vo.quant_gemm(x, L).astype(np.float64).reshape(-1)
    assert np.max(np.abs(y - y_star)) <= 1e-5 * max(1.0, np.max(np.abs(y_star)))


@pytest.mark.parametrize("c", CASES)
def test_structure(c):
    L, stream, offs = _build(c)
    I, K = c["I"], c["K"]
    Ro, NS = (c["O"] + 7) // 8, K // 8192
    assert stream.dtype == torch.uint8 and stream.shape[1] == (160 if c["Kr"] > 0 else 128)
    assert offs.shape[0] == NS * Ro + 1 and int(offs[0]) == 0 and int(offs[-1]) == stream.shape[0]
    assert bool((offs[1:] >= offs[:-1]).all())
    words = stream[:, :128].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    col, low = (words >> 16).numpy(), (words & 8191).numpy()
    null = col == I
    assert (low[null] == 0).all() and (col <= I).all()
    # every field exactly once: I real entries per row over its NS lists; padding only at list tails
    o = offs.numpy()
    for r in range(Ro):
        seen = []
        for s in range(NS):
            a, b = o[s * Ro + r], o[s * Ro + r + 1]
            cl = col[a:b].reshape(-1)
            n = int((cl != I).sum())
            assert (cl[:n] != I).all() and (b - a) == (n + 31) // 32
            seen.append(cl[:n])
        assert sorted(np.concatenate(seen).tolist()) == list(range(I))
    # bank-group ordering: most aligned groups of 8 real entries touch 8 different 16-byte bank groups
    g_low, g_null = low.reshape(-1, 8), null.reshape(-1, 8)
    full = ~g_null.any(1)
    distinct = np.array([len(set((row & 7).tolist())) == 8 for row in g_low[full]])
    assert distinct.mean() > 0.6


def _partition(offs_cta, warps=16):
    """Mirror of the kernel's run / piece arithmetic: returns per row the list of (warp, steps)."""
    T0, TT = offs_cta[0], offs_cta[-1] - offs_cta[0]
    nrows = len(offs_cta) - 1
    pieces = {}
    for w in range(warps):
        t, t_end = T0 + TT * w // warps, T0 + TT * (w + 1) // warps
        if t == t_end:
            continue
        lo, hi = 0, nrows - 1
        while lo < hi:
            mid = (lo + hi + 1) >> 1
            if offs_cta[mid] <= t:
                lo = mid
            else:
                hi = mid - 1
        row, acc = lo, 0
        row_end = offs_cta[row + 1]
        while t < t_end:
            while t == row_end:
                pieces[row + w] = (row, w, acc)
                row += 1
                row_end = offs_cta[row + 1]
                acc = 0
            acc += 1
            t += 1
        pieces[row + w] = (row, w, acc)
    return pieces


@pytest.mark.parametrize("seed", range(6))
def test_run_partition_covers_every_step_once(seed):
    rng = np.random.default_rng(seed)
    nrows = int(rng.integers(1, 60))
    lens = rng.integers(0, 40, size=nrows)
    if seed % 3 == 0:
        lens[rng.integers(0, nrows, size=nrows // 2)] = 0          # empty rows
    offs = np.concatenate([[int(rng.integers(0, 1000))], lens]).cumsum()
    T0, TT = int(offs[0]), int(offs[-1] - offs[0])
    pieces = _partition([int(v) for v in offs])
    _check_pieces(offs, pieces)


def _check_pieces(offs, pieces):
    nrows, T0, TT = len(offs) - 1, int(offs[0]), int(offs[-1] - offs[0])
    for row in range(nrows):
        a, b = int(offs[row]) - T0, int(offs[row + 1]) - T0
        total = 0
        if b > a:                                                  # the epilogue's fw / lw formulas
            fw, lw = ((a + 1) * 16 - 1) // TT, (b * 16 - 1) // TT
            for w in range(fw, lw + 1):
                if row + w not in pieces:                          # a warp with an empty run: the kernel
                    assert TT * w // 16 == TT * (w + 1) // 16      # zero-fills the piece table for these
                    continue
                r, ww, n = pieces[row + w]
                assert (r, ww) == (row, w)
                total += n
        assert total == b - a


@pytest.mark.parametrize("lens", [[9], [1], [0, 3, 0], [5, 0, 0, 7], [2, 2, 2], [15], [16], [17, 1]])
def test_run_partition_with_fewer_steps_than_warps(lens):
    offs = np.concatenate([[7], lens]).cumsum()
    _check_pieces(offs, _partition([int(v) for v in offs]))


@pytest.mark.parametrize("c", CASES)
def test_c_abi_host_builder_is_byte_identical(c):
    """vptq_b200_sliced_build_host (plain CPU code in the shared library) == the tensor builder."""
    import ctypes
    from vptq_b200 import native
    L, stream, offs = _build(c)
    lib = native.lib()
    ind = np.ascontiguousarray(L.indices[0])
    Ro, NS = (c["O"] + 7) // 8, c["K"] // 8192
    offs_c = np.zeros(NS * Ro + 1, dtype=np.uint32)
    steps = ctypes.c_size_t(0)
    args = (ind.ctypes.data, ind.shape[1], c["O"], c["I"], c["K"], c["Kr"])
    assert lib.vptq_b200_sliced_build_host(*args, None, 0, offs_c.ctypes.data, ctypes.byref(steps)) == 0
    assert steps.value == stream.shape[0]
    out = np.zeros(steps.value * stream.shape[1], dtype=np.uint8)
    assert lib.vptq_b200_sliced_build_host(*args, out.ctypes.data, out.size, offs_c.ctypes.data, ctypes.byref(steps)) == 0
    assert np.array_equal(offs_c.astype(np.int64), offs.numpy().astype(np.int64))
    assert np.array_equal(out, stream.numpy().reshape(-1))
    # too small a buffer is refused, not overrun
    assert lib.vptq_b200_sliced_build_host(*args, out.ctypes.data, 16, offs_c.ctypes.data, ctypes.byref(steps)) == -3


def test_builders_agree_on_random_shapes():
    """Fuzz: tensor builder == C host builder, and the lists still evaluate to the oracle's sums."""
    import ctypes
    from vptq_b200 import native
    lib = native.lib()
    rng = np.random.default_rng(2024)
    for trial in range(12):
        K = int(rng.choice([16384, 32768, 65536]))
        Kr = int(rng.choice([-1, 2, 64, 256]))
        I = int(rng.integers(33, 700))
        O = int(rng.integers(1, 90))
        L = vo.make_layer(I, O, vector_len=8, num_centroids=K, num_res_centroids=Kr, dtype="fp16", seed=100 + trial)
        ind_t = torch.from_numpy(np.ascontiguousarray(L.indices))
        stream, offs = sliced.build_sliced(ind_t, num_centroids=K, num_res_centroids=Kr, group_size=I, out_features=O)
        ind = np.ascontiguousarray(L.indices[0])
        Ro, NS = (O + 7) // 8, K // 8192
        offs_c = np.zeros(NS * Ro + 1, dtype=np.uint32)
        steps = ctypes.c_size_t(0)
        out = np.zeros(max(stream.numel(), 1), dtype=np.uint8)
        rc = lib.vptq_b200_sliced_build_host(ind.ctypes.data, ind.shape[1], O, I, K, Kr, out.ctypes.data, out.size,
                                             offs_c.ctypes.data, ctypes.byref(steps))
        assert rc == 0, native.last_error()
        assert steps.value == stream.shape[0]
        assert np.array_equal(offs_c.astype(np.int64), offs.numpy().astype(np.int64)), (K, Kr, I, O)
        assert np.array_equal(out[:stream.numel()], stream.numpy().reshape(-1)), (K, Kr, I, O)
        # value check through the float64 evaluator
        x = vo.make_x(1, I, "fp16", seed=trial)
        xf = vo.to_f32(x, "fp16").astype(np.float64).reshape(-1)
        perm = np.asarray(L.perm).astype(np.uint16).astype(np.int64)
        sc, wb = vo.to_f32(L.weight_scale, "fp16").astype(np.float64), vo.to_f32(L.weight_bias, "fp16").astype(np.float64)
        y = sliced.emulate(stream, offs, num_centroids=K, num_res_centroids=Kr, group_size=I, out_features=O,
                           centroids=torch.from_numpy(vo.to_f32(L.centroids, "fp16")),
                           res_centroids=None if Kr <= 0 else torch.from_numpy(vo.to_f32(L.res_centroids, "fp16")),
                           xq=torch.from_numpy(xf[perm] * sc[perm])).numpy()[:O] + float((xf * wb).sum())
        y_star = vo.quant_gemm(x, L).astype(np.float64).reshape(-1)
        assert np.max(np.abs(y - y_star)) <= 1e-5 * max(1.0, np.max(np.abs(y_star))), (K, Kr, I, O)


def test_host_builder_rejects_what_the_kernel_does_not_cover():
    import ctypes
    from vptq_b200 import native
    lib = native.lib()
    ind = np.zeros((4, 64), dtype=np.int32)
    offs = np.zeros(64, dtype=np.uint32)
    steps = ctypes.c_size_t(0)
    for K, Kr in ((8192, 256), (4096, -1), (65536, 512), (24576, 16)):
        rc = lib.vptq_b200_sliced_build_host(ind.ctypes.data, 64, 32, 64, K, Kr, None, 0, offs.ctypes.data, ctypes.byref(steps))
        assert rc in (-1, -2), (K, Kr)
    assert lib.vptq_b200_sliced_build_host(ind.ctypes.data, 1, 32, 64, 65536, 256, None, 0, offs.ctypes.data,
                                           ctypes.byref(steps)) == -1      # stride shorter than a packed row
