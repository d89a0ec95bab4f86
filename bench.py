#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-site estimator path (BASELINE.json metric:
site x haplotype genotypes / second, and % of HBM peak).

Workload (BASELINE.json configs[1]): per-site pi/theta tracks (per_site_diversity) for 1M
biallelic sites x 2,504 diploid samples (5,008 haplotypes) split into two inversion-orientation
haplotype groups, mask BED applied, missing-data bitmap present.  One "step" = one pass of the
hot path over the whole shard: for both groups, stream the group's bitplanes once, derive
alt/called counts, evaluate per-site pi and theta (NaN for masked / uncallable sites) and the
region partials (S, sum pi, uncallable sites).

  value : device-resident throughput (planes already in HBM), CUDA events on the launching stream
  e2e   : same metric through the C-ABI with HOST (pinned) buffers: H2D of the u8 matrix +
          bitmap, on-device repack into bitplanes, the stats pass, D2H of the tracks
  roofline / cpu_baseline : see DESIGN.md "Measurement"

python bench.py --gpus N --steps K --warmup W          (our arm; torchrun for N > 1)
python bench.py --impl reference ...                   (CPU arm: oracle port of the Rust path)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "site_x_haplotype_genotypes_per_sec"
UNIT = "genotypes/s"
N_SITES = 1_000_000
N_SAMPLES = 2_504
MISSING_RATE = 0.01
N_MASK = 2_000
CPU_SAMPLE_SITES = 100_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- synthetic data
def make_positions(n, seed):
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.integers(1, 50, size=n, dtype=np.int64))


def make_mask(positions, seed, n_mask=N_MASK):
    rng = np.random.default_rng(seed + 1)
    lo, hi = int(positions[0]), int(positions[-1])
    starts = rng.integers(lo, max(hi, lo + 1), size=n_mask, dtype=np.int64)
    lens = rng.integers(1_000, 5_000, size=n_mask, dtype=np.int64)
    return np.ascontiguousarray(np.stack([starts, starts + lens], axis=1).reshape(-1))


def make_groups(n_samples, seed):
    """Inversion orientation per haplotype ~ Bernoulli(0.3) -> haplotype groups 0 / 1."""
    rng = np.random.default_rng(seed + 2)
    orient = rng.random((n_samples, 2)) < 0.3
    g0 = [(s, side) for s in range(n_samples) for side in (0, 1) if not orient[s, side]]
    g1 = [(s, side) for s in range(n_samples) for side in (0, 1) if orient[s, side]]
    return g0, g1


def gen_device(n_sites, n_samples, seed, device, missing_rate=MISSING_RATE, inband=None):
    """u8 matrix [V, S, 2] + packed missing bitmap generated on the GPU (host never has to
    synthesise 5 GB with numpy).  Beta(0.8, 0.8) site frequencies, Bernoulli alleles."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    stride = n_samples * 2
    total = n_sites * stride
    data = torch.empty(total, dtype=torch.uint8, device=device)
    words = (total + 63) // 64
    bitmap = torch.zeros(words, dtype=torch.int64, device=device)
    weights = (torch.ones(64, dtype=torch.int64, device=device) << torch.arange(64, device=device))
    beta = torch.distributions.Beta(torch.tensor(0.8), torch.tensor(0.8))
    chunk = max(64, min(64 * 512, ((1 << 28) // max(n_samples * 2, 1)) // 64 * 64))  # bounds the temporaries
    for s0 in range(0, n_sites, chunk):
        s1 = min(n_sites, s0 + chunk)
        n = s1 - s0
        torch.manual_seed(seed * 1_000_003 + s0)
        f = beta.sample((n,)).to(device).clamp_(0.001, 0.999)
        a = (torch.rand((n, stride), generator=gen, device=device) < f[:, None])
        miss = torch.rand((n, stride), generator=gen, device=device) < missing_rate
        a &= ~miss
        data[s0 * stride:s1 * stride] = a.reshape(-1).to(torch.uint8)
        if inband is not None:  # the same cohort as an int8 array: missing cells are negative (0xFF)
            a8 = a.to(torch.uint8)
            a8[miss] = 255
            inband[s0 * stride:s1 * stride] = a8.reshape(-1)
        flat = miss.reshape(-1)
        e0 = s0 * stride  # multiple of 64 because chunk is
        pad = (-flat.numel()) % 64
        if pad:
            flat = torch.cat([flat, torch.zeros(pad, dtype=torch.bool, device=device)])
        w = (flat.reshape(-1, 64).to(torch.int64) * weights).sum(dim=1)
        bitmap[e0 // 64:e0 // 64 + w.numel()] = w
    return data, bitmap


def gen_host_sample(n_sites, n_samples, seed):
    """CPU-only generator for the reference arm / cpu baseline sample (numpy)."""
    rng = np.random.default_rng(seed)
    stride = n_samples * 2
    f = np.clip(rng.beta(0.8, 0.8, size=n_sites), 0.001, 0.999).astype(np.float32)
    data = np.empty((n_sites, stride), dtype=np.uint8)
    miss = np.empty((n_sites, stride), dtype=bool)
    step = 8192
    for s0 in range(0, n_sites, step):
        s1 = min(n_sites, s0 + step)
        r = rng.random((s1 - s0, stride), dtype=np.float32)
        m = rng.random((s1 - s0, stride), dtype=np.float32) < MISSING_RATE
        data[s0:s1] = (r < f[s0:s1, None]) & ~m
        miss[s0:s1] = m
    return data, miss


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max = index, [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                p = [x.strip() for x in out.stdout.strip().split(",")]
                self.samples.append(float(p[0]))
                self.max = float(p[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   p[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_pass(data2d, miss2d, positions, groups, mask, threads):
    """One pass of the reference's CPU path over a block of sites, per group: the dense summary
    (rayon-parallel in the reference, stats.rs:1415-1460 -> threaded here) plus the serial
    per-site track loop (stats.rs:4693).  Returns seconds."""
    from oracle import pyoracle as orc

    V, stride = data2d.shape
    S = stride // 2
    bits = orc.pack_missing_bits(miss2d.reshape(-1))
    dense = orc.Dense(data2d.reshape(-1), bits, V, S, 2, 1)
    gt = data2d.reshape(V, S, 2).copy()
    gt[np.broadcast_to(miss2d.reshape(V, S, 2).any(axis=2, keepdims=True), gt.shape)] = 0xFF
    vs = orc.Variants(positions, gt)
    region = (int(positions[0]), int(positions[-1]))
    t0 = time.perf_counter()
    for haps in groups:
        orc.build_summary(dense, haps, threads=threads)
        orc.per_site_diversity(vs, haps, region, mask=mask.reshape(-1, 2))
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_sample_sites
    seed = N_SITES + N_SAMPLES
    pos = make_positions(n, seed)
    data, miss = gen_host_sample(n, args.samples, seed)
    groups = make_groups(args.samples, seed)
    mask = make_mask(pos, seed, max(1, N_MASK * n // N_SITES))
    for _ in range(args.warmup):
        cpu_pass(data[: n // 10], miss[: n // 10], pos[: n // 10], groups, mask, threads)
    times = [cpu_pass(data, miss, pos, groups, mask, threads) for _ in range(args.steps)]
    t = sum(times) / len(times)
    geno = n * args.samples * 2
    val = geno / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 -> popcount u32 + f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} of {args.sites} sites x {args.samples * 2} haplotypes per step "
                                   "(oracle port of the Rust path; the Rust crate cannot be built here)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "configs[1]: per-site pi/theta tracks, two inversion-orientation haplotype groups, "
                        "mask BED, missing bitmap",
            "sites_per_gpu": args.sites, "diploid_samples": args.samples, "haplotypes": args.samples * 2,
            "groups": 2, "mask_intervals": N_MASK, "missing_rate": MISSING_RATE,
            "l2_policy": "inputs larger than L2 (bitplanes ~1.3 GB per pass vs 126 MB L2)"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    # libraries (NCCL's version banner) may write to fd 1: keep the real stdout for the one JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    import ferromic_b200 as fm
    from ferromic_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    _lib.check(L.fm_set_device(local))

    V, S = args.sites, args.samples
    seed = N_SITES + N_SAMPLES + rank
    pos = make_positions(V, seed)
    mask = make_mask(pos, seed)
    g0, g1 = make_groups(S, N_SITES + N_SAMPLES)
    t0 = time.perf_counter()
    h_i8 = None
    if not args.skip_e2e and not args.skip_inband:
        d_i8 = torch.empty(V * S * 2, dtype=torch.uint8, device=device)
        d_data, d_bitmap = gen_device(V, S, seed, device, inband=d_i8)
        h_i8 = torch.empty(d_i8.numel(), dtype=torch.uint8, pin_memory=True)
        h_i8.copy_(d_i8)
        torch.cuda.synchronize()
        del d_i8
    else:
        d_data, d_bitmap = gen_device(V, S, seed, device)
    torch.cuda.synchronize()
    log(f"[rank {rank}] generated {V}x{S * 2} u8 matrix on device in {time.perf_counter() - t0:.1f}s")

    def group_arrays(haps):
        return (np.asarray([h[0] for h in haps], dtype=np.uint64), np.asarray([h[1] for h in haps], dtype=np.uint8))

    def make_groups_on(matrix):
        hs = []
        for haps in (g0, g1):
            idx, side = group_arrays(haps)
            h = C.c_void_p()
            _lib.check(L.fm_group_create(matrix, idx.ctypes.data, side.ctypes.data, len(haps), C.byref(h)))
            hs.append(h)
        return hs

    # ---------------- device-resident (value)
    m = C.c_void_p()
    _lib.check(L.fm_matrix_create_device(d_data.data_ptr(), d_bitmap.data_ptr(), V, S, 2, 1, pos.ctypes.data,
                                         C.byref(m)))
    groups = make_groups_on(m)
    garr = (C.c_void_p * 2)(*[g.value for g in groups])
    res = _lib.BenchResult()
    # N > 1: every step ends with the exchange of the groups' region totals (S, sum pi, uncallable
    # sites) -- a fused fold + P2P mailbox kernel over NVLink (csrc/fm_comm.cuh); torch.distributed
    # only carries the 64-byte mailbox handles at start-up and the barriers.
    comm = None
    if world == 1 and os.environ.get("FM_BENCH_SELF_COMM"):  # diagnostic: the exchange kernel alone
        comm = C.c_void_p()
        _lib.check(L.fm_comm_create(0, 1, C.byref(comm)))
    if world > 1:
        comm = C.c_void_p()
        _lib.check(L.fm_comm_create(rank, world, C.byref(comm)))
        hb = (C.c_uint8 * 64)()
        _lib.check(L.fm_comm_export(comm, hb))
        mine = torch.tensor(list(hb), dtype=torch.uint8, device=device)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine)
        handles = np.ascontiguousarray(allh.cpu().numpy())
        _lib.check(L.fm_comm_connect(comm, handles.ctypes.data))
        dist.barrier()
    _lib.check(L.fm_bench_diversity(garr, 2, 1, mask.ctypes.data, mask.size // 2, max(args.warmup, 3), comm,
                                    C.byref(res)))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        wall0 = time.perf_counter()
        _lib.check(L.fm_bench_diversity(garr, 2, 1, mask.ctypes.data, mask.size // 2, args.steps, comm,
                                        C.byref(res)))
        dev_ms = res.step_ms_avg * args.steps
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - wall0) * 1e3
        if world > 1:
            dist.barrier()
        # the timed region is a few milliseconds: keep the same kernels running (untimed) until the
        # sampler has seen the clocks under this load a few times
        t_keep = time.perf_counter()
        while len(clocks.samples) < 8 and time.perf_counter() - t_keep < 4.0:
            _lib.check(L.fm_bench_diversity(garr, 2, 1, mask.ctypes.data, mask.size // 2, 200, None,
                                            C.byref(_lib.BenchResult())))
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    geno_per_rank = V * S * 2
    value = world * geno_per_rank * args.steps / (dev_ms_max * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    launches_per_step = max(1, int(round(res.plane_launches / args.steps)))
    plane_launch_bytes = res.plane_bytes_per_step / launches_per_step
    achieved = plane_launch_bytes / (res.plane_ms_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_plane_pass.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch")
            if tj.get("launches_per_step") and tj["launches_per_step"] != launches_per_step:
                traffic = traffic * tj["launches_per_step"] / launches_per_step  # same bytes, other launch split
        except Exception:
            traffic = None
    kname = ("fm_k_plane_pass_seq (both groups' planes streamed by one persistent launch)" if launches_per_step == 1
             else "fm_k_plane_pass<1>")
    roofline = {"bound": "hbm", "kernel": kname, "launches_per_step": launches_per_step, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "peak_source": peak_src,
                "traffic": traffic, "bytes_per_launch": plane_launch_bytes, "ms_per_launch": res.plane_ms_avg,
                "per_group": None if launches_per_step == 1 else
                [{"haplotypes": len(h), "ms": res.group_ms_avg[i],
                  "GBps": res.group_bytes[i] / (res.group_ms_avg[i] * 1e-3) / 1e9} for i, h in enumerate((g0, g1))]}

    for g in groups:
        L.fm_group_release(g)
    L.fm_matrix_release(m)
    if comm is not None:
        if world > 1:
            dist.barrier()  # nobody may still be writing into a mailbox that is about to be freed
        L.fm_comm_destroy(comm)

    # ---------------- end to end through the C ABI with host (pinned) buffers
    e2e = None
    if not args.skip_e2e:
        h_data = torch.empty(d_data.numel(), dtype=torch.uint8, pin_memory=True)
        h_bitmap = torch.empty(d_bitmap.numel(), dtype=torch.int64, pin_memory=True)
        h_data.copy_(d_data)
        h_bitmap.copy_(d_bitmap)
        torch.cuda.synchronize()
        if args.free_device_copy:
            del d_data, d_bitmap
            torch.cuda.empty_cache()
        out_pos = torch.empty(V, dtype=torch.int64, pin_memory=True).numpy()  # caller-owned result buffers are pinned
        out_pi = torch.empty((2, V), dtype=torch.float64, pin_memory=True).numpy()
        out_th = torch.empty((2, V), dtype=torch.float64, pin_memory=True).numpy()

        phases = {}

        def e2e_step(inband=False):
            # streaming ingest: chunked H2D overlapped with the repack into both groups' bitplanes
            t = [time.perf_counter()]

            def lap(name):
                t.append(time.perf_counter())
                phases[name] = phases.get(name, 0.0) + (t[-1] - t[-2]) * 1e3

            ih = C.c_void_p()
            _lib.check(L.fm_ingest_begin(V, S, 2, 2 if inband else 1, 1, pos.ctypes.data, 0, C.byref(ih)))
            for idx, side in garrs:
                _lib.check(L.fm_ingest_add_group(ih, idx.ctypes.data, side.ctypes.data, len(idx), None))
            lap("begin+declare_groups")
            if inband:  # the caller's int8 array as it is: negative cells are missing, no bitmap
                _lib.check(L.fm_ingest_rows(ih, h_i8.data_ptr(), None, 0, V))
            else:
                _lib.check(L.fm_ingest_rows(ih, h_data.data_ptr(), h_bitmap.data_ptr(), 0, V))
            lap("ingest_rows")
            mh = C.c_void_p()
            gh = (C.c_void_p * 2)()
            _lib.check(L.fm_ingest_finish(ih, C.byref(mh), gh, None))
            gs = [C.c_void_p(gh[0]), C.c_void_p(gh[1])]
            lap("finish")
            n = C.c_size_t()
            for k, (g, haps) in enumerate(zip(gs, (g0, g1))):
                _lib.check(L.fm_per_site_diversity(g, len(haps), int(pos[0]), int(pos[-1]), mask.ctypes.data,
                                                   mask.size // 2, None, 0, out_pos.ctypes.data,
                                                   out_pi[k].ctypes.data, out_th[k].ctypes.data, V, C.byref(n)))
            lap("per_site_diversity_x2")
            for g in gs:
                L.fm_group_release(g)
            L.fm_matrix_release(mh)
            lap("release")
            return n.value

        garrs = [group_arrays(h) for h in (g0, g1)]
        e2e_step()  # warm-up
        phases.clear()
        k = max(1, min(args.steps, args.e2e_steps))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        L.fm_timings_reset()
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        tim = _lib.Timings()
        L.fm_timings_get(C.byref(tim))
        phases_timed = dict(phases)
        inband_info = None
        if h_i8 is not None:
            e2e_step(inband=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(k):
                e2e_step(inband=True)
            torch.cuda.synchronize()
            t_in = (time.perf_counter() - t0) / k
            inband_info = {"value": geno_per_rank / t_in, "ms_per_step": t_in * 1e3,
                           "h2d_bytes_per_step": int(h_i8.numel() + V * 8),
                           "what": "same step from the int8 array Population.from_numpy receives (negative = missing, "
                                   "FM_MISSING_IN_BAND): no bitmap, no host conversion pass; per rank"}
        # the same call with PAGEABLE host memory (what a Rust Vec<u8> or a numpy array is): the library
        # fills pinned bounce buffers with several host threads and overlaps them with the DMA
        pageable_ms = None
        if rank == 0 and world == 1 and not args.skip_pageable:
            p_data = np.array(h_data.numpy(), copy=True)
            p_bitmap = np.array(h_bitmap.numpy(), copy=True)
            pinned_ptrs = (h_data, h_bitmap)

            class _P:  # minimal stand-in exposing data_ptr() like the pinned tensors
                def __init__(self, a):
                    self.a = a

                def data_ptr(self):
                    return self.a.ctypes.data
            h_data, h_bitmap = _P(p_data), _P(p_bitmap)
            e2e_step()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            pageable_ms = (time.perf_counter() - t0) * 1e3
            h_data, h_bitmap = pinned_ptrs
            del p_data, p_bitmap
        te = torch.tensor([t_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * geno_per_rank * k / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_data.numel() + h_bitmap.numel() * 8 + V * 8),
               "d2h_bytes_per_step": int(2 * 2 * V * 8), "steps": k, "ms_per_step": float(te.item()) / k * 1e3,
               "breakdown_ms_per_step": {"h2d": tim.h2d_ms / k, "repack": tim.repack_ms / k,
                                         "stats": tim.stats_ms / k, "reduce": tim.reduce_ms / k,
                                         "d2h": tim.d2h_ms / k},
               "host_phase_ms_per_step": {k_: v_ / k for k_, v_ in phases_timed.items()},
               "pageable_host_ms_per_step": pageable_ms, "inband_int8": inband_info,
               "api": "fm_ingest_begin/add_group/rows/finish (chunked H2D overlapped with repack) + "
                      "fm_per_site_diversity per group; h2d and repack spans overlap",
               "timing": "wall clock around synchronous C-ABI calls, cuda-synchronised on both sides"}

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        n = min(args.cpu_sample_sites, V)
        threads = os.cpu_count() or 1
        sd, sm = gen_host_sample(n, S, N_SITES + N_SAMPLES)
        spos = pos[:n]
        smask = make_mask(spos, N_SITES + N_SAMPLES, max(1, N_MASK * n // N_SITES))
        t = cpu_pass(sd, sm, spos, (g0, g1), smask, threads)
        cpu = {"value": n * S * 2 / t, "unit": UNIT, "cores": threads, "kind": "port", "seconds": t,
               "sample": f"{n} of {V} sites x {S * 2} haplotypes, both groups (oracle port: threaded dense summary "
                         "+ serial per-site track loop, as in the reference)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8 -> 1-bit planes, popcount u32 + f64",
                "data": "synthetic", "config": workload_config(args), "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": int(res.plane_launches + res.other_launches),
                "clocks": clocks.summary(), "wall_ms_per_step": wall_ms / args.steps,
                "exchange_ms_per_step": res.comm_ms_avg if comm is not None else None,
                "collective": None if world == 1 else "per step: fused fold + P2P mailbox exchange of region totals "
                                                      "(fm_k_comm_exchange over NVLink peer memory), inside the timed region",
                "timing": "CUDA events on the launching stream (cudaStreamPerThread), max over ranks"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=N_SITES)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--cpu-sample-sites", type=int, default=CPU_SAMPLE_SITES)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-pageable", action="store_true")
    ap.add_argument("--skip-inband", action="store_true")
    ap.add_argument("--free-device-copy", action="store_true", default=True)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
