"""Slice x tile lists (vptq_b200/lists.py): the format the list-based decode kernel reads.

CPU only: the builder is checked against the oracle through a float64 evaluation of the lists, against
the C host builder (byte-identical), and the kernel's work partition (csrc/gemv_lists.cu: CTA unit ranges,
warp runs, stage sequence, piece merge, row-block arrival counts) is mirrored in integers."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import vptq_oracle as vo
from vptq_b200 import lists

CASES = [
    dict(I=1024, O=264, K=65536, Kr=256),   # the Llama-3 2-bit configuration, small
    dict(I=512, O=64, K=16384, Kr=16),      # four slices, small residual codebook
    dict(I=768, O=44, K=32768, Kr=-1),      # eight slices, no residual, ragged last row
    dict(I=9000, O=24, K=8192, Kr=256),     # three column tiles, two slices
    dict(I=1004, O=100, K=65536, Kr=256),   # in_features not a multiple of 8
]


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _layer(c, dtype="fp16", perm=True):
    return vo.make_layer(c["I"], c["O"], vector_len=8, num_centroids=c["K"], num_res_centroids=c["Kr"], dtype=dtype,
                         seed=11, enable_perm=perm)


def _build(c, dtype="fp16", perm=True):
    L = _layer(c, dtype, perm)
    ind = torch.from_numpy(np.ascontiguousarray(L.indices))
    pt = None if L.perm is None else torch.from_numpy(np.asarray(L.perm).astype(np.uint16).astype(np.int64))
    stream, tab, tcw = lists.build_lists(ind, num_centroids=c["K"], num_res_centroids=c["Kr"], in_features=c["I"],
                                         out_features=c["O"], perm=pt)
    return L, stream, tab, tcw


def _eval(L, c, stream, tab, x):
    xf = vo.to_f32(x, "fp16").astype(np.float64).reshape(-1)
    sc = vo.to_f32(L.weight_scale, "fp16").astype(np.float64)
    wb = vo.to_f32(L.weight_bias, "fp16").astype(np.float64)
    C = torch.from_numpy(vo.to_f32(L.centroids, "fp16"))
    R = None if c["Kr"] <= 0 else torch.from_numpy(vo.to_f32(L.res_centroids, "fp16"))
    y = lists.emulate(stream, tab, num_centroids=c["K"], num_res_centroids=c["Kr"], in_features=c["I"],
                      out_features=c["O"], centroids=C, res_centroids=R, xs=torch.from_numpy(xf * sc)).numpy()
    return y[:c["O"]] + float((xf * wb).sum())


@pytest.mark.parametrize("c", CASES)
@pytest.mark.parametrize("perm", [True, False])
def test_lists_reproduce_the_oracle(c, perm):
    L, stream, tab, _ = _build(c, perm=perm)
    x = vo.make_x(1, c["I"], "fp16")
    y = _eval(L, c, stream, tab, x)
    y_star = vo.quant_gemm(x, L).astype(np.float64).reshape(-1)
    assert np.max(np.abs(y - y_star)) <= 1e-5 * max(1.0, np.max(np.abs(y_star)))


@pytest.mark.parametrize("c", CASES)
def test_structure(c):
    L, stream, tab, tcw = _build(c)
    I, K = c["I"], c["K"]
    NS, NT, TCW = lists.geometry(I, K)
    assert tcw == TCW and TCW % 8 == 0 and TCW <= 4096 and NT * TCW >= I and (NT - 1) * TCW < I
    Ro, Q = (c["O"] + 7) // 8, NS * NT
    assert stream.dtype == torch.int32 and stream.shape[1] == 32
    tb = tab.numpy().astype(np.int64) & 0xFFFFFFFF
    first, tail = tb & lists.STEP_MASK, tb >> 26
    assert tb.shape[0] == Q * Ro + 1 and first[0] == 0 and first[-1] == stream.shape[0]
    assert (np.diff(first) >= 1).all()                                  # every unit has at least one step
    w = stream.numpy().astype(np.int64) & 0xFFFFFFFF
    perm = np.asarray(L.perm).astype(np.uint16).astype(np.int64)
    idx = L.meta["idx"][0]
    ridx = None if L.meta["ridx"] is None else L.meta["ridx"][0]
    inv = np.argsort(perm)
    for r in range(Ro):
        feats = []
        for q in range(Q):
            u = q * Ro + r
            a, b, tl = first[u], first[u + 1], tail[u]
            n = (b - a - 1) * 32 + tl
            ent = w[a:b].reshape(-1)
            assert (ent[n:] == 0).all()                                 # padding only at the list's tail, zero words
            t, s = q // NS, q % NS
            f = t * TCW + ((ent[:n] >> 12) & 4095)
            assert (f < min(I, (t + 1) * TCW)).all()
            c_of = inv[f]                                               # quantised column of each entry
            assert (idx[r, c_of] == s * 4096 + (ent[:n] & 4095)).all()
            if ridx is not None:
                assert (ridx[r, c_of] == ent[:n] >> 24).all()
            else:
                assert (ent[:n] >> 24 == 0).all()
            feats.append(f)
        assert sorted(np.concatenate(feats).tolist()) == list(range(I))  # every field exactly once
    # bank-group ordering: most aligned groups of 8 real entries touch 8 different 16-byte bank groups
    unit_of = np.searchsorted(first[1:], np.arange(stream.shape[0]), side="right")
    last = np.arange(stream.shape[0]) + 1 == first[unit_of + 1]
    full_steps = ~last | (tail[unit_of] == 32)
    g = (w[full_steps] & 7).reshape(-1, 8)
    if len(g):
        distinct = np.array([len(set(row.tolist())) == 8 for row in g])
        assert distinct.mean() > 0.5


# ------------------------------------------------------------------------------------------------------
# integer mirror of the kernel's partition arithmetic
# ------------------------------------------------------------------------------------------------------
def _kernel_partition(first, tail, Ro, Q, ncta, warps=16, sps=8, rb=32):
    """Walks every CTA / warp / stage exactly as gemv_lists_kernel does.  Returns (entries counted per unit,
    units written per writer kind, arrivals per row block)."""
    U = Q * Ro
    assert ncta >= Q and ncta <= U
    got = np.zeros(U, dtype=np.int64)          # valid entries accumulated per unit (through pieces or directly)
    writes = np.zeros(U, dtype=np.int64)       # how many times part[u] is written (must be exactly 1)
    arrivals = np.zeros((Ro + rb - 1) // rb, dtype=np.int64)
    for q in range(ncta):
        u0, u1 = U * q // ncta, U * (q + 1) // ncta
        nun = u1 - u0
        cA, cB = u0 // Ro, (u1 - 1) // Ro
        assert cB <= cA + 1
        two = cB != cA
        nA = cB * Ro - u0 if two else nun
        tab_first = first[u0:u1 + 1]
        T0, T1 = int(tab_first[0]), int(tab_first[nun])
        TT = T1 - T0
        TB = int(tab_first[nA]) if two else T1
        pieces = []                                                    # (slot index, unit, count)
        for w in range(warps):
            tb_, te_ = T0 + TT * w // warps, T0 + TT * (w + 1) // warps
            e1, b2 = min(te_, TB), max(tb_, TB)
            n1, n2 = (max(e1 - tb_, 0) + sps - 1) // sps, (max(te_ - b2, 0) + sps - 1) // sps
            if n1 + n2 == 0:
                continue
            lo, hi = 0, nun - 1
            while lo < hi:
                mid = (lo + hi + 1) >> 1
                if tab_first[mid] <= tb_:
                    lo = mid
                else:
                    hi = mid - 1
            u = uF = lo
            u_end, tl = int(tab_first[u + 1]), int(tail[u0 + u])
            acc = 0

            def flush(complete):
                nonlocal acc
                boundary = u == uF or (not complete) or u_end >= te_
                if not boundary:
                    got[u0 + u] += acc
                    writes[u0 + u] += 1
                else:
                    pieces.append((w * 2 + (0 if u == uF else 1), u, acc))
                acc = 0

            for qi in range(n1 + n2):
                if qi < n1:
                    t = tb_ + qi * sps
                    cnt = min(sps, e1 - t)
                    assert t + cnt <= TB
                else:
                    t = b2 + (qi - n1) * sps
                    cnt = min(sps, te_ - t)
                    assert t >= TB
                segB = qi >= n1
                for j in range(cnt):
                    assert (u >= nA) == segB                           # a stage never straddles the segments
                    last = t + 1 == u_end
                    acc += tl if last else 32
                    t += 1
                    if last:
                        flush(True)
                        u += 1
                        if u < nun:
                            u_end, tl = int(tab_first[u + 1]), int(tail[u0 + u])
            if u < nun and tab_first[u] < te_:
                flush(False)
        slots = dict((k, (un, n)) for k, un, n in pieces)
        assert len(slots) == len(pieces)                               # one piece per slot
        order = sorted(slots)
        i = 0
        while i < len(order):                                          # the merge: leaders sum their followers
            un, n = slots[order[i]]
            j = i + 1
            while j < len(order) and slots[order[j]][0] == un:
                n += slots[order[j]][1]
                j += 1
            got[u0 + un] += n
            writes[u0 + un] += 1
            i = j
        rA0 = u0 - cA * Ro
        rA1 = rA0 + nA
        nB = nun - nA
        for b in range(rA0 // rb, (rA1 - 1) // rb + 1):
            arrivals[b] += min(rA1, (b + 1) * rb) - max(rA0, b * rb)
        if two:
            for b in range((nB - 1) // rb + 1):
                arrivals[b] += min(nB, (b + 1) * rb) - b * rb
    return got, writes, arrivals


@pytest.mark.parametrize("c,ncta", [(CASES[0], 148), (CASES[0], 16), (CASES[0], 37), (CASES[1], 148), (CASES[1], 5),
                                    (CASES[2], 48), (CASES[3], 6), (CASES[3], 18), (CASES[4], 148), (CASES[4], 208)])
def test_kernel_partition_covers_every_entry_once(c, ncta):
    L, stream, tab, tcw = _build(c)
    NS, NT, _ = lists.geometry(c["I"], c["K"])
    Ro, Q = (c["O"] + 7) // 8, NS * NT
    ncta = max(Q, min(ncta, Q * Ro))
    tb = tab.numpy().astype(np.int64) & 0xFFFFFFFF
    first, tail = tb & lists.STEP_MASK, tb >> 26
    got, writes, arrivals = _kernel_partition(first, tail, Ro, Q, ncta)
    n_valid = (np.diff(first) - 1) * 32 + tail[:-1]
    assert (got == n_valid).all() and (writes == 1).all()
    assert got.reshape(Q, Ro).sum(0).tolist() == [c["I"]] * Ro
    rows_b = np.minimum(32, Ro - 32 * np.arange(len(arrivals)))
    assert (arrivals == Q * rows_b).all()


@pytest.mark.parametrize("seed", range(8))
def test_kernel_partition_random_list_lengths(seed):
    """Lists of arbitrary (also empty -> one padded step) lengths, few steps per warp, odd CTA counts."""
    rng = np.random.default_rng(seed)
    Ro, Q = int(rng.integers(1, 70)), int(rng.choice([2, 4, 6, 16]))
    n = rng.integers(0, 150, size=Q * Ro)
    if seed % 2:
        n[rng.integers(0, Q * Ro, size=Q * Ro // 3)] = 0
    steps = np.maximum((n + 31) // 32, 1)
    first = np.concatenate([[0], steps.cumsum()])
    tail = np.concatenate([n - 32 * (steps - 1), [0]])
    ncta = int(rng.integers(Q, max(Q + 1, min(Q * Ro, 300) + 1)))
    got, writes, arrivals = _kernel_partition(first, tail, Ro, Q, ncta)
    assert (got == n).all() and (writes == 1).all()
    rows_b = np.minimum(32, Ro - 32 * np.arange(len(arrivals)))
    assert (arrivals == Q * rows_b).all()


# ------------------------------------------------------------------------------------------------------
# C host builder
# ------------------------------------------------------------------------------------------------------
def _c_build(L, c, perm=True, deal=True):
    from vptq_b200 import native
    lib = native.lib()
    ind = np.ascontiguousarray(L.indices[0])
    NS, NT, TCW = lists.geometry(c["I"], c["K"])
    Ro = (c["O"] + 7) // 8
    tab_c = np.zeros(NS * NT * Ro + 1, dtype=np.uint32)
    steps, tcw = ctypes.c_size_t(0), ctypes.c_int32(0)
    pp = np.ascontiguousarray(np.asarray(L.perm).astype(np.uint16)) if (perm and L.perm is not None) else None
    args = (ind.ctypes.data, ind.shape[1], c["O"], c["I"], c["K"], c["Kr"], None if pp is None else pp.ctypes.data)
    rc = lib.vptq_b200_lists_build_host(*args, None, 0, tab_c.ctypes.data, ctypes.byref(steps), ctypes.byref(tcw))
    assert rc == 0, native.last_error()
    out = np.zeros(steps.value * 32, dtype=np.uint32)
    rc = lib.vptq_b200_lists_build_host(*args, out.ctypes.data, out.nbytes, tab_c.ctypes.data, ctypes.byref(steps),
                                        ctypes.byref(tcw))
    assert rc == 0, native.last_error()
    if deal:   # the second load-time pass lists.build_lists applies by default
        assert lib.vptq_b200_lists_deal_host(out.ctypes.data, tab_c.ctypes.data, len(tab_c) - 1, 0) == 0, native.last_error()
    return out, tab_c, steps.value, tcw.value, (args, ind, pp)   # (ind / pp keep the host buffers alive)


@pytest.mark.parametrize("c", CASES)
def test_c_abi_host_builder_is_byte_identical(c):
    """vptq_b200_lists_build_host (plain CPU code in the shared library) == the tensor builder."""
    from vptq_b200 import native
    L, stream, tab, tcw = _build(c)
    out, tab_c, steps, tcw_c, (args, *_keep) = _c_build(L, c)
    assert steps == stream.shape[0] and tcw_c == tcw
    assert np.array_equal(tab_c.astype(np.int64), tab.numpy().astype(np.int64) & 0xFFFFFFFF)
    assert np.array_equal(out, stream.numpy().reshape(-1).view(np.uint32))
    # too small a buffer is refused, not overrun
    st, tc = ctypes.c_size_t(0), ctypes.c_int32(0)
    assert native.lib().vptq_b200_lists_build_host(*args, out.ctypes.data, 16, tab_c.ctypes.data, ctypes.byref(st),
                                                   ctypes.byref(tc)) == -3


def test_builders_agree_on_random_shapes():
    """Fuzz: tensor builder == C host builder, and the lists still evaluate to the oracle's sums."""
    rng = np.random.default_rng(2024)
    for trial in range(12):
        c = dict(K=int(rng.choice([8192, 16384, 32768, 65536])), Kr=int(rng.choice([-1, 2, 64, 256])),
                 I=int(rng.integers(33, 700)), O=int(rng.integers(1, 90)))
        perm = bool(trial % 3)
        L = vo.make_layer(c["I"], c["O"], vector_len=8, num_centroids=c["K"], num_res_centroids=c["Kr"], dtype="fp16",
                          seed=100 + trial, enable_perm=perm)
        pt = None if L.perm is None else torch.from_numpy(np.asarray(L.perm).astype(np.uint16).astype(np.int64))
        stream, tab, tcw = lists.build_lists(torch.from_numpy(np.ascontiguousarray(L.indices)), num_centroids=c["K"],
                                             num_res_centroids=c["Kr"], in_features=c["I"], out_features=c["O"], perm=pt)
        out, tab_c, steps, tcw_c, _ = _c_build(L, c, perm)
        assert steps == stream.shape[0] and tcw_c == tcw, c
        assert np.array_equal(tab_c.astype(np.int64), tab.numpy().astype(np.int64) & 0xFFFFFFFF), c
        assert np.array_equal(out, stream.numpy().reshape(-1).view(np.uint32)), c
        x = vo.make_x(1, c["I"], "fp16", seed=trial)
        y = _eval(L, c, stream, tab, x)
        y_star = vo.quant_gemm(x, L).astype(np.float64).reshape(-1)
        assert np.max(np.abs(y - y_star)) <= 1e-5 * max(1.0, np.max(np.abs(y_star))), c


def _list_bounds(tab):
    t = tab.astype(np.int64) & 0xFFFFFFFF
    first, end, tail = t[:-1] & lists.STEP_MASK, t[1:] & lists.STEP_MASK, t[:-1] >> 26
    return first * 32, (end - 1) * 32 + tail, end * 32      # entry range [a, b) of every list, c = end of its padding


@pytest.mark.parametrize("c", CASES)
def test_dealing_only_reorders_inside_lists(c):
    """vptq_b200_lists_deal_host: every list keeps its multiset of entries, padding stays zero, the result does not
    depend on the number of threads, and running it again on its own output changes nothing a sum could see."""
    from vptq_b200 import native
    L = _layer(c)
    pt = None if L.perm is None else torch.from_numpy(np.asarray(L.perm).astype(np.uint16).astype(np.int64))
    kw = dict(num_centroids=c["K"], num_res_centroids=c["Kr"], in_features=c["I"], out_features=c["O"], perm=pt)
    ind = torch.from_numpy(np.ascontiguousarray(L.indices))
    raw, tab, _ = lists.build_lists(ind, deal=False, **kw)
    dealt, tab2, _ = lists.build_lists(ind, deal=True, **kw)
    assert torch.equal(tab, tab2)
    r, d = raw.numpy().reshape(-1).view(np.uint32), dealt.numpy().reshape(-1).view(np.uint32)
    a, b, e = _list_bounds(tab.numpy())
    for u in range(len(a)):
        assert np.array_equal(np.sort(r[a[u]:b[u]]), np.sort(d[a[u]:b[u]])), u
        assert not d[b[u]:e[u]].any(), u
    one = raw.clone()
    th = tab.contiguous()
    assert native.lib().vptq_b200_lists_deal_host(one.data_ptr(), th.data_ptr(), th.numel() - 1, 1) == 0
    assert torch.equal(one, dealt)


@pytest.mark.parametrize("kind", ["constant", "two_values", "one_slice", "long_lists"])
def test_dealing_survives_degenerate_index_patterns(kind):
    """All entries in one class, all in one slice, lists of 64 steps: still a permutation inside every list."""
    from vptq_b200 import pack
    g = torch.Generator().manual_seed(5)
    K, Kr, I, O = 65536, 256, 1024, 40
    Ro = O // 8
    if kind == "constant":
        idx = torch.full((1, Ro, I), 4099, dtype=torch.int64)
    elif kind == "two_values":
        idx = torch.where(torch.rand(1, Ro, I, generator=g) < 0.9, 8, 4097)
    elif kind == "one_slice":    # 1024 entries in one list, the other 15 lists of the row empty
        idx = torch.randint(0, 4096, (1, Ro, I), generator=g)
    else:
        K, I = 8192, 4096
        idx = torch.randint(0, K, (1, Ro, I), generator=g)
    ridx = torch.randint(0, Kr, idx.shape, generator=g)
    packed = pack.pack_index(idx, K.bit_length() - 1, ridx, 8)
    kw = dict(num_centroids=K, num_res_centroids=Kr, in_features=I, out_features=O, perm=None)
    raw, tab, _ = lists.build_lists(packed, deal=False, **kw)
    dealt, _, _ = lists.build_lists(packed, deal=True, **kw)
    r, d = raw.numpy().reshape(-1).view(np.uint32), dealt.numpy().reshape(-1).view(np.uint32)
    a, b, e = _list_bounds(tab.numpy())
    assert int((b - a).sum()) == Ro * I
    for u in range(len(a)):
        assert np.array_equal(np.sort(r[a[u]:b[u]]), np.sort(d[a[u]:b[u]])), (kind, u)
        assert not d[b[u]:e[u]].any(), (kind, u)


def test_dealing_lowers_the_modelled_bank_conflicts():
    """On a K = 65536 layer with random indices the re-ordering must not make either gather worse, and must cut the
    x' gather's conflicts (the model of tools/deal_stats.py: distinct addresses per bank group / bank)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("deal_stats", os.path.join(ROOT, "tools", "deal_stats.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    c = dict(K=65536, Kr=256, I=4096, O=128)
    L = _layer(c)
    pt = torch.from_numpy(np.asarray(L.perm).astype(np.uint16).astype(np.int64))
    kw = dict(num_centroids=c["K"], num_res_centroids=c["Kr"], in_features=c["I"], out_features=c["O"], perm=pt)
    ind = torch.from_numpy(np.ascontiguousarray(L.indices))
    raw, tab, _ = lists.build_lists(ind, deal=False, **kw)
    dealt, _, _ = lists.build_lists(ind, deal=True, **kw)
    tb = tab.numpy().astype(np.int64) & 0xFFFFFFFF
    c0, x0 = ds.wavefronts(raw.numpy().view(np.uint32), tb)
    c1, x1 = ds.wavefronts(dealt.numpy().view(np.uint32), tb)
    assert c1 <= c0 + 1e-9 and x1 < x0 - 0.5, (c0, x0, c1, x1)


def test_host_builder_rejects_what_the_kernel_does_not_cover():
    from vptq_b200 import native
    lib = native.lib()
    ind = np.zeros((4, 64), dtype=np.int32)
    tab = np.zeros(1024, dtype=np.uint32)
    steps, tcw = ctypes.c_size_t(0), ctypes.c_int32(0)
    for K, Kr in ((4096, 256), (2048, -1), (65536, 512), (24576, 16)):
        rc = lib.vptq_b200_lists_build_host(ind.ctypes.data, 64, 32, 64, K, Kr, None, None, 0, tab.ctypes.data,
                                            ctypes.byref(steps), ctypes.byref(tcw))
        assert rc in (-1, -2), (K, Kr)
    assert lib.vptq_b200_lists_build_host(ind.ctypes.data, 1, 32, 64, 65536, 256, None, None, 0, tab.ctypes.data,
                                          ctypes.byref(steps), ctypes.byref(tcw)) == -1   # stride shorter than a packed row
