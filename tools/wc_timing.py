#!/usr/bin/env python
"""Timing of the W&C path (config 4 shape) through the C ABI: partition repack, per-group count
passes (K2 x 27) and K4 + fold.  usage: wc_timing.py [sites] [repeat]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ferromic_b200 import _lib  # noqa: E402

# 26 subpopulation sizes (diploid samples), 1000G-like: 61..113, sum 2504
POP_SIZES = [96, 61, 86, 93, 99, 103, 105, 94, 99, 99, 91, 103, 113, 107, 102, 104, 99, 99, 85, 64, 85, 96, 104,
             102, 107, 108]
assert sum(POP_SIZES) == 2504 and len(POP_SIZES) == 26


def membership(n_samples=2504):
    left = np.full(n_samples, 0xFFFF, dtype=np.uint16)
    s = 0
    for p, k in enumerate(POP_SIZES):
        left[s:s + k] = p
        s += k
    return left, left.copy()


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    S = 2504
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    _lib.check(L.fm_set_device(0))
    pos = bench.make_positions(V, 10_002_504)
    d_data, d_bitmap = bench.gen_device(V, S, 10_002_504, dev)
    torch.cuda.synchronize()
    m = C.c_void_p()
    _lib.check(L.fm_matrix_create_device(d_data.data_ptr(), d_bitmap.data_ptr(), V, S, 2, 1, pos.ctypes.data,
                                         C.byref(m)))
    left, right = membership(S)
    G, NP = 26, 325
    tim = _lib.Timings()
    ph = None
    for attempt in ("cold (first launch of the process)", "warm"):
        if ph is not None:
            L.fm_partition_release(ph)
        L.fm_timings_reset()
        t0 = time.perf_counter()
        ph = C.c_void_p()
        _lib.check(L.fm_partition_create(m, left.ctypes.data, right.ctypes.data, S, G, C.byref(ph)))
        t_part = time.perf_counter() - t0
        L.fm_timings_get(C.byref(tim))
        print(f"partition_create {attempt}: {t_part * 1e3:.1f} ms wall, count kernel (K1, all {G + 1} groups) "
              f"{tim.repack_ms:.2f} ms = {V * S * 2 * 1.125 / (tim.repack_ms * 1e-3) / 1e12:.2f} TB/s of u8 + bitmap")
    w = np.array([int(pos[0]), int(pos[-1])], dtype=np.int64)
    nv = np.zeros(1, dtype=np.uint64); os_ = np.zeros(1, dtype=np.uint64)
    oa = np.zeros(1); ob = np.zeros(1)
    pa = np.zeros(NP); pb = np.zeros(NP); pn = np.zeros(NP, dtype=np.uint64)
    for i in range(rep + 1):
        L.fm_timings_reset()
        t0 = time.perf_counter()
        _lib.check(L.fm_wc_window_sums(ph, w.ctypes.data, 1, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                       os_.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
        dt = time.perf_counter() - t0
        L.fm_timings_get(C.byref(tim))
        tag = "first (27 count passes + K4)" if i == 0 else "cached counts (K4 + fold)"
        print(f"wc_window_sums {tag}: {dt * 1e3:.2f} ms wall; stats_ms total {tim.stats_ms:.3f}; "
              f"genotypes/s {V * S * 2 / dt:.3e}; launches {tim.kernel_launches}")
    print("overall", oa[0], ob[0], int(os_[0]), "fst", oa[0] / (oa[0] + ob[0]))
    # many windows (100 kb)
    edges = np.arange(int(pos[0]), int(pos[-1]) + 1, 100_000, dtype=np.int64)
    ww = np.stack([edges, edges + 99_999], axis=1).reshape(-1).copy()
    n = len(edges)
    nv = np.zeros(n, dtype=np.uint64); os_ = np.zeros(n, dtype=np.uint64)
    oa = np.zeros(n); ob = np.zeros(n)
    pa = np.zeros(n * NP); pb = np.zeros(n * NP); pn = np.zeros(n * NP, dtype=np.uint64)
    t0 = time.perf_counter()
    _lib.check(L.fm_wc_window_sums(ph, ww.ctypes.data, n, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                   os_.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
    print(f"{n} windows of 100 kb: {(time.perf_counter() - t0) * 1e3:.2f} ms wall")
    L.fm_partition_release(ph)
    L.fm_matrix_release(m)


if __name__ == "__main__":
    main()
