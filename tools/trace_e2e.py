"""Wall time of every C-ABI call of one packed e2e step (bench.py's e2e), to see where the host overhead is."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ferromic_b200 import _lib  # noqa: E402


def main():
    import torch
    V, S = 1_000_000, 2504
    L = _lib.lib()
    _lib.check(L.fm_set_device(0))
    pos = bench.make_positions(V, 1)
    mask = bench.make_mask(pos, 1)
    g0, g1 = bench.make_groups(S, 1)
    garrs = [bench.group_arrays(h) for h in (g0, g1)]
    rw = (S * 2 + 31) // 32
    h_ab = torch.randint(0, 2 ** 31 - 1, (V * rw,), dtype=torch.int32).pin_memory()
    h_cb = torch.full((V * rw,), -1, dtype=torch.int32).pin_memory()
    # sparse missing list: 50 missing cells per row (1 %), ascending columns
    rng = np.random.default_rng(7)
    cols = np.sort(rng.integers(0, S * 2, size=(V, 50), dtype=np.int64).astype(np.uint16), axis=1)
    h_cols = torch.from_numpy(cols.reshape(-1).view(np.int16).copy()).pin_memory()
    h_start = torch.arange(0, 50 * (V + 1), 50, dtype=torch.int64).pin_memory()
    out_pos = torch.empty(V, dtype=torch.int64, pin_memory=True).numpy()
    out_pi = torch.empty((2, V), dtype=torch.float64, pin_memory=True).numpy()
    out_th = torch.empty((2, V), dtype=torch.float64, pin_memory=True).numpy()
    raw_n = (C.c_size_t * 2)(len(g0), len(g1))
    rows = []
    for it in range(5):
        T = {}

        def call(name, fn, *a):
            t = time.perf_counter()
            _lib.check(fn(*a))
            T[name] = T.get(name, 0.0) + (time.perf_counter() - t) * 1e3

        ih = C.c_void_p()
        call("begin", L.fm_ingest_begin, V, S, 2, 1, 1, pos.ctypes.data, 0, C.byref(ih))
        for idx, side in garrs:
            call("add_group", L.fm_ingest_add_group, ih, idx.ctypes.data, side.ctypes.data, len(idx), None)
        if it % 2 == 0:
            call("rows_packed", L.fm_ingest_rows_packed, ih, h_ab.data_ptr(), h_cb.data_ptr(), 0, V)
        else:
            call("rows_packed_sparse", L.fm_ingest_rows_packed_sparse, ih, h_ab.data_ptr(), h_start.data_ptr(),
                 h_cols.data_ptr(), 2, 0, V)
        mh = C.c_void_p()
        gh = (C.c_void_p * 2)()
        call("finish", L.fm_ingest_finish, ih, C.byref(mh), gh, None)
        n = C.c_size_t()
        call("per_site_multi", L.fm_per_site_diversity_multi, gh, raw_n, 2, int(pos[0]), int(pos[-1]), mask.ctypes.data,
             mask.size // 2, None, 0, out_pos.ctypes.data, out_pi.ctypes.data, out_th.ctypes.data, V, C.byref(n))
        call("per_site_multi_nopos", L.fm_per_site_diversity_multi, gh, raw_n, 2, int(pos[0]), int(pos[-1]),
             mask.ctypes.data, mask.size // 2, None, 0, None, out_pi.ctypes.data, out_th.ctypes.data, V, C.byref(n))
        for g in gh:
            call("release", L.fm_group_release, C.c_void_p(g))
        call("release", L.fm_matrix_release, mh)
        rows.append(T)
    print(json.dumps(rows[1:]))


if __name__ == "__main__":
    main()
