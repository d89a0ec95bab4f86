"""Strong-scaling block of bench.py alone (config 3 split over the ranks, one sharded Hudson call per step).
usage: torchrun --nproc-per-node N tools/bench_strong.py   (FM_SHARDED_TRACE=1 prints the stage times)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import bench_configs
    from ferromic_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    _lib.check(L.fm_set_device(local))

    class A:
        steps = 10
    peak = 6553.3
    out = bench_configs.strong_cfg3(L, _lib, A, rank, world, device, dist if world > 1 else None, peak,
                                    float(os.environ.get("FM_BENCH_CONFIG_SCALE", "1.0")))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
