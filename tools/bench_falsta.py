"""Times the device FALSTA track renderer (fm_falsta_tracks) on a 5 Mb region with 150k records, six FST
tracks (the shape append_fst_falsta writes for W&C), next to the reference's algorithm restated in Python
on a bounded sample.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from ferromic_b200 import falsta
    from oracle import falsta as ofa
    rng = np.random.default_rng(5)
    region_len, n, T = 5_000_000, 150_000, 6
    pos = np.sort(rng.choice(np.arange(1, region_len + 1), size=n, replace=False)).astype(np.int64)
    vals = rng.random((T, n))
    vals[rng.random((T, n)) < 0.1] = np.nan
    vals[rng.random((T, n)) < 0.1] = 0.0
    best = None
    for _ in range(4):
        t0 = time.perf_counter()
        lines = falsta.track_lines(pos, vals, 1, region_len, falsta.FST)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    nbytes = sum(len(l) for l in lines)
    # the C-ABI call alone (host arrays in, text out into a caller buffer), without the Python line splitting
    import ctypes as C
    from ferromic_b200 import _lib
    L = _lib.lib()
    cap = nbytes + 64
    buf = np.empty(cap, dtype=np.uint8)
    lens = (C.c_size_t * T)()
    total = C.c_size_t(0)
    v = np.ascontiguousarray(vals)
    call = None
    for _ in range(4):
        t0 = time.perf_counter()
        _lib.check(L.fm_falsta_tracks(pos.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), n, T, 1, region_len,
                                      falsta.FST, buf.ctypes.data_as(C.c_void_p), cap, lens, C.byref(total)))
        dt = time.perf_counter() - t0
        call = dt if call is None or dt < call else call
    # CPU: the reference's algorithm (Vec<String> of region length per track, one pass over the records per track)
    small = 200_000
    recs = [(int(p), *vals[:, i]) for i, p in enumerate(pos[pos <= small])]
    t0 = time.perf_counter()
    ref = ofa.fst_falsta_text("1", 1, small, recs, [])
    cpu_dt = time.perf_counter() - t0
    got = falsta.fst_falsta_text("1", 1, small, recs, []).decode()
    assert got == ref
    print(json.dumps({"what": "fm_falsta_tracks, 6 FST tracks over a 5 Mb region, 150k records", "text_MB": nbytes / 1e6,
                      "wall_ms": best * 1e3, "cabi_call_ms": call * 1e3, "text_GBps_cabi": nbytes / call / 1e9,
                      "positions_x_tracks_per_s": region_len * T / best,
                      "cpu_port": {"kind": "port (pure Python)", "region": small, "seconds": cpu_dt,
                                   "positions_x_tracks_per_s": small * T / cpu_dt}}))


if __name__ == "__main__":
    main()
