"""Host packer throughput (fm_pack_rows) against the number of threads, pinned vs pageable source, next to a
plain memcpy of the same bytes; and the wall time of each C-ABI call of one packed e2e step.
usage: python tools/bench_pack.py [sites]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ferromic_b200 import _lib  # noqa: E402


def main():
    import torch
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    S = 2504
    stride = S * 2
    rw = (stride + 31) // 32
    L = _lib.lib()
    rng = np.random.default_rng(1)
    cells = (rng.random((V, stride), dtype=np.float32) < 0.3).astype(np.uint8)
    words = (V * stride + 63) // 64
    bitmap = rng.integers(0, 2 ** 63, size=words, dtype=np.uint64) & rng.integers(0, 2 ** 63, size=words, dtype=np.uint64)
    out = {"sites": V, "cores": os.cpu_count(), "pack": []}
    pin = torch.empty(V * stride, dtype=torch.uint8, pin_memory=True)
    pin.numpy()[:] = cells.reshape(-1)
    ab = torch.empty(V * rw, dtype=torch.int32, pin_memory=True)
    cb = torch.empty(V * rw, dtype=torch.int32, pin_memory=True)
    for name, src in (("pageable", cells.ctypes.data), ("pinned", pin.data_ptr())):
        for mode in (1, 0):
            for th in (1, 2, 4, 8, 16, 32):
                if th > 2 * (os.cpu_count() or 1):
                    continue
                best = 1e9
                for _ in range(3):
                    t = time.perf_counter()
                    _lib.check(L.fm_pack_rows(src, bitmap.ctypes.data, mode, 0, V, V, stride, ab.data_ptr(),
                                              cb.data_ptr() if mode else None, th))
                    best = min(best, time.perf_counter() - t)
                out["pack"].append({"src": name, "mode": mode, "threads": th, "ms": best * 1e3,
                                    "u8_GBps": V * stride / best / 1e9})
    dst = np.empty_like(cells)
    t = time.perf_counter()
    np.copyto(dst, cells)
    out["memcpy_1thread_GBps_read"] = V * stride / (time.perf_counter() - t) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
