"""The NVLink / cudaIpc mailbox exchange ACROSS PROCESSES: torchrun starts one process per rank;
every rank maps every peer's mailbox through cudaIpc, runs raw exchanges and the sharded Hudson call, and checks that
gathered words are exact, that the merged totals are the rank-ordered sum of all ranks' local totals bit for bit and
that every rank holds identical results (tests/mp_exchange_worker.py).  With fewer GPUs than ranks the ranks share a
device -- the inter-process path is the same."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_mailbox_exchange_across_processes(world):
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_exchange_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert f"MP_EXCHANGE_OK {world}" in r.stdout


def test_device_list_api_and_environment_variable():
    """fm_set_devices / fm_get_devices / FERROMIC_GPU_DEVICES: the allowed-device list is process-wide, a
    thread's default device is its first entry, fm_set_device rejects ordinals outside it."""
    code = r'''
import ctypes as C, os, sys
from ferromic_b200 import _lib
L = _lib.lib()
n = C.c_size_t()
devs = (C.c_int * 64)()
assert L.fm_get_devices(devs, 64, C.byref(n)) == 0
expect = os.environ.get("EXPECT")
if expect is not None:
    assert [devs[i] for i in range(n.value)] == [int(x) for x in expect.split(",") if x], list(devs[:n.value])
cnt = C.c_int()
L.fm_device_count(C.byref(cnt))
one = (C.c_int * 1)(0)
assert L.fm_set_devices(one, 1) == 0
assert L.fm_get_devices(devs, 64, C.byref(n)) == 0 and n.value == 1 and devs[0] == 0
if cnt.value > 1:
    assert L.fm_set_device(1) == _lib.FM_ERR_INVALID_ARG
assert L.fm_set_device(0) == 0
bad = (C.c_int * 1)(99)
assert L.fm_set_devices(bad, 1) == _lib.FM_ERR_INVALID_ARG
assert L.fm_set_devices(None, 0) == 0
assert L.fm_get_devices(devs, 64, C.byref(n)) == 0 and n.value == cnt.value
print("DEVICES_OK")
'''
    for env_val, expect in ((None, None), ("0", "0"), ("0, 0,77", "0")):
        env = dict(os.environ)
        env.pop("FERROMIC_GPU_DEVICES", None)
        env.pop("EXPECT", None)
        if env_val is not None:
            env["FERROMIC_GPU_DEVICES"] = env_val
            env["EXPECT"] = expect
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "DEVICES_OK" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
