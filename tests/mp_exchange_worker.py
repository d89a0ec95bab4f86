"""Worker of tests/test_gpu_multiprocess.py: one process per rank (torchrun), gloo for the host-side plumbing, the
library's cudaIpc mailbox for the data path.  Rank r works on GPU r % device_count, so the cross-process path is
exercised on a single-GPU box too (two processes mapping each other's mailbox on the same device)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from ferromic_b200 import _lib, sharded

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    dev = rank % torch.cuda.device_count()
    L = _lib.lib()
    _lib.check(L.fm_set_device(dev))

    def exchange(h):
        out = [None] * world
        dist.all_gather_object(out, h)
        return out

    comm = sharded.PeerComm(rank, world, exchange_handles=exchange)
    # 1. raw words: gathered exact, merged = rank-ordered sums, identical bits on every rank
    for step, n in enumerate([5, 2048, 1, 77]):
        f = np.random.default_rng(1000 * step + rank).normal(size=n)
        gathered, merged = comm.allgather_words(f, n_double=n // 2)
        allf = [None] * world
        dist.all_gather_object(allf, f)
        for r in range(world):
            assert np.array_equal(gathered[r], allf[r].view(np.uint64)), "gathered words differ"
        tot = np.zeros(n // 2)
        for r in range(world):
            tot = tot + allf[r][: n // 2]
        assert np.array_equal(merged[: n // 2].view(np.float64), tot), "merged doubles are not the rank-ordered sum"
        allm = [None] * world
        dist.all_gather_object(allm, merged)
        assert all(np.array_equal(allm[0], m) for m in allm), "ranks hold different merged bits"
    # 2. the sharded Hudson call: merged totals == rank-ordered sum of the ranks' local totals, same outcome everywhere
    V_total, S = 8192 * world + 8192, 120
    per = ((V_total + world - 1) // world + 8191) // 8192 * 8192
    lo, hi = min(V_total, rank * per), min(V_total, (rank + 1) * per)
    V = hi - lo
    pop = (np.arange(S) >= S // 2).astype(np.uint16)
    d_data = torch.empty(max(V * S * 2, 16), dtype=torch.uint8, device=f"cuda:{dev}")
    d_miss = torch.empty((V * S * 2 + 63) // 64 + 2, dtype=torch.int64, device=f"cuda:{dev}")
    _lib.check(L.fm_synth_fill(d_data.data_ptr(), d_miss.data_ptr(), V, S, 2, lo, 4242, pop.ctypes.data, 0.05, 0.02))
    pos = np.arange(lo, hi, dtype=np.int64) * 3
    m = C.c_void_p()
    _lib.check(L.fm_matrix_create_device(d_data.data_ptr(), d_miss.data_ptr(), V, S, 2, 1, pos.ctypes.data, C.byref(m)))
    groups = []
    for members in (range(S // 2), range(S // 2, S)):
        idx = np.repeat(np.asarray(list(members), dtype=np.uint64), 2)
        side = np.tile(np.array([0, 1], dtype=np.uint8), len(idx) // 2)
        g = C.c_void_p()
        _lib.check(L.fm_group_create(m, idx.ctypes.data, side.ctypes.data, len(idx), C.byref(g)))
        groups.append(g)
    Lr = V_total * 3
    out, sums = _lib.HudsonOutcome(), _lib.HudsonSums()
    lout, lsums = _lib.HudsonOutcome(), _lib.HudsonSums()
    for _ in range(3):
        _lib.check(L.fm_hudson_pair_sharded(groups[0], groups[1], Lr, S, S, comm.handle, C.byref(out), C.byref(sums)))
    _lib.check(L.fm_hudson_pair_sharded(groups[0], groups[1], Lr, S, S, None, C.byref(lout), C.byref(lsums)))
    vec = np.array([lsums.num, lsums.den, lsums.dxy, lsums.pi1, lsums.pi2, float(lsums.dxy_uncallable), float(lsums.unc1),
                    float(lsums.unc2)])
    allv = [None] * world
    dist.all_gather_object(allv, vec)
    tot = np.zeros(8)
    for r in range(world):
        tot = tot + allv[r]
    got = np.array([sums.num, sums.den, sums.dxy, sums.pi1, sums.pi2, float(sums.dxy_uncallable), float(sums.unc1),
                    float(sums.unc2)])
    assert np.array_equal(got, tot), f"rank {rank}: exchanged Hudson totals differ from the rank-ordered sum"
    res = (out.fst, out.d_xy, out.pi_pop1, out.pi_pop2, out.some)
    allr = [None] * world
    dist.all_gather_object(allr, res)
    assert all(r == allr[0] for r in allr), "ranks computed different outcomes"
    assert out.some == 31 and 0.0 <= out.d_xy <= 1.0
    for g in groups:
        L.fm_group_release(g)
    L.fm_matrix_release(m)
    dist.barrier()
    comm.close()  # collective: closing handshake with every peer
    dist.barrier()
    if rank == 0:
        print("MP_EXCHANGE_OK", world, res)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
