"""Concurrent host-to-device bandwidth of N ranks (torchrun), from (a) cudaHostAlloc memory (torch pin_memory) and
(b) an anonymous mapping advised to transparent huge pages and registered with cudaHostRegister.  Answers whether the
multi-GPU e2e ceiling of the bench boxes (KVM guests, GPUs behind an IOMMU) depends on the page size of the pinned
source.  usage: torchrun --nproc-per-node N tools/h2d_probe.py"""
import ctypes
import json
import mmap
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 704 << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def bw(src_ptr, reps=6):
        best = 0.0
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = rt.cudaMemcpyAsync(d.data_ptr(), src_ptr, n, 1, None)
            assert rc == 0, rc
            torch.cuda.synchronize()
            best = max(best, n / (time.perf_counter() - t0) / 1e9)
        t = torch.tensor([best], device=dev)
        if world > 1:
            dist.all_reduce(t)  # aggregate of the per-rank best (an upper bound of the concurrent rate)
        return float(t.item())

    a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    a.fill_(1)
    r_alloc = bw(a.data_ptr())

    def bw_chunked(src_ptr, chunk, reps=6):  # the ingest's pattern: 32 MB pieces queued back to back on one stream
        st = torch.cuda.Stream()
        rates = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for o in range(0, n, chunk):
                m = min(chunk, n - o)
                rc = rt.cudaMemcpyAsync(d.data_ptr() + o, src_ptr + o, m, 1, ctypes.c_void_p(st.cuda_stream))
                assert rc == 0, rc
            st.synchronize()
            rates.append(n / (time.perf_counter() - t0) / 1e9)
        t = torch.tensor([max(rates), sorted(rates)[len(rates) // 2]], device=dev)
        if world > 1:
            dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    r_chunk = bw_chunked(a.data_ptr(), 32 << 20)
    del a
    # (b) THP-advised anonymous memory, registered
    two_mb = 2 << 20
    mm = mmap.mmap(-1, n + two_mb, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    thp = "n/a"
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
        thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    except Exception as e:  # noqa: BLE001
        thp = f"madvise failed: {e}"
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    start = (base + two_mb - 1) & ~(two_mb - 1)
    ctypes.memset(start, 1, n)  # touch
    huge = None
    try:
        for line in open("/proc/self/smaps"):
            if line.startswith("AnonHugePages") and int(line.split()[1]) > 0:
                huge = (huge or 0) + int(line.split()[1])
    except Exception:  # noqa: BLE001
        pass
    rc = int(torch.cuda.cudart().cudaHostRegister(start, n, 0))
    r_reg = bw(start) if rc == 0 else None
    if rc == 0:
        torch.cuda.cudart().cudaHostUnregister(start)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_rank": n, "aggregate_GBps_cudaHostAlloc": r_alloc,
                          "aggregate_GBps_thp_registered": r_reg,
                          "aggregate_GBps_32MB_chunks_best_median": r_chunk, "thp_enabled": thp, "anon_huge_kb": huge,
                          "register_rc": rc}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
