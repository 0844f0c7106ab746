"""Modelled shared-memory wavefronts per 32-entry step of a layer's lists, before and after the bank-aware
re-ordering (vptq_b200_lists_deal_host).  CPU only.

  codebook gather (LDS.128): per quarter-warp, max number of DISTINCT 16-byte addresses in one bank group (index & 7);
  x' gather (LDS.U16)      : per warp, max number of DISTINCT 4-byte words in one bank ((column >> 1) & 31).

usage: python tools/deal_stats.py [--O 4096] [--I 4096] [--K 65536] [--Kr 256] [--rows 64]
"""
import argparse
import time

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptq_b200 import lists, pack  # noqa: E402


def wavefronts(stream: np.ndarray, tab: np.ndarray):
    first = tab[:-1] & lists.STEP_MASK
    tail = tab[:-1] >> 26
    end = tab[1:] & lists.STEP_MASK
    T = stream.shape[0]
    valid = np.ones((T, 32), dtype=bool)
    last = end - 1
    lane = np.arange(32)[None, :]
    valid[last] = lane < tail[:, None]
    e = stream.astype(np.uint32)
    cls = (e & 7).astype(np.int64)
    addr = (e & 4095).astype(np.int64)
    word = ((e >> 13) & 2047).astype(np.int64)   # 4-byte word of x'
    bank = word & 31
    code = np.zeros(T, dtype=np.int64)
    for q in range(4):
        sl = slice(8 * q, 8 * q + 8)
        worst = np.zeros(T, dtype=np.int64)
        for g in range(8):
            m = valid[:, sl] & (cls[:, sl] == g)
            a = np.where(m, addr[:, sl], -1 - np.arange(8)[None, :])
            a = np.sort(a, axis=1)
            distinct = (m.any(1)).astype(np.int64) + ((a[:, 1:] != a[:, :-1]) & (a[:, 1:] >= 0) & (a[:, :-1] >= 0)).sum(1)
            worst = np.maximum(worst, distinct)
        code += np.maximum(worst, 1)
    xw = np.zeros(T, dtype=np.int64)
    for b in range(32):
        m = valid & (bank == b)
        a = np.where(m, word, -1 - lane)
        a = np.sort(a, axis=1)
        distinct = (m.any(1)).astype(np.int64) + ((a[:, 1:] != a[:, :-1]) & (a[:, 1:] >= 0) & (a[:, :-1] >= 0)).sum(1)
        xw = np.maximum(xw, distinct)
    xw = np.maximum(xw, 1)
    return code.mean(), xw.mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--O", type=int, default=4096)
    ap.add_argument("--I", type=int, default=4096)
    ap.add_argument("--K", type=int, default=65536)
    ap.add_argument("--Kr", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(a.seed)
    Ro = a.O // 8
    idx = torch.randint(0, a.K, (1, Ro, a.I), generator=g)
    ridx = torch.randint(0, a.Kr, (1, Ro, a.I), generator=g) if a.Kr else None
    ib, rb = a.K.bit_length() - 1, (a.Kr.bit_length() - 1 if a.Kr else 0)
    packed = pack.pack_index(idx, ib, ridx, rb)
    perm = torch.randperm(a.I, generator=g).to(torch.int16)
    kw = dict(num_centroids=a.K, num_res_centroids=a.Kr, in_features=a.I, out_features=a.O, perm=perm)
    s0, tab, _ = lists.build_lists(packed, deal=False, **kw)
    t0 = time.time()
    s1 = lists.deal_lists(s0.clone(), tab)
    dt = time.time() - t0
    tb = tab.numpy().astype(np.int64) & 0xffffffff
    c0, x0 = wavefronts(s0.numpy().view(np.uint32), tb)
    c1, x1 = wavefronts(s1.numpy().view(np.uint32), tb)
    print(f"steps {s0.shape[0]}  deal {dt:.2f} s")
    print(f"round-robin : codebook {c0:.2f} + x' {x0:.2f} = {c0 + x0:.2f} wavefronts / step")
    print(f"matched     : codebook {c1:.2f} + x' {x1:.2f} = {c1 + x1:.2f} wavefronts / step")


if __name__ == "__main__":
    main()
