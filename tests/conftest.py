import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["packed", "u8"])
def both_ingest_modes(request, monkeypatch):
    """Runs a test once per host->device ingest format of the Python mirror (FERROMIC_GPU_INGEST): the 2-bit
    packed rows (fm_pack_rows + fm_matrix_create_packed, the default) and the reference-layout u8 upload that is
    repacked on the device.  Both must give the oracle's results."""
    monkeypatch.setenv("FERROMIC_GPU_INGEST", request.param)
    return request.param
