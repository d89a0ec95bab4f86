"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
declared in include/ferromic_gpu.h, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ferromic_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ferromic_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in ferromic_gpu.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_host_only_entry_points_need_no_gpu():
    import ferromic_b200 as fm
    assert fm.adjusted_sequence_length(100, 200, None, [(100, 101)]) == 100
    assert fm.adjusted_sequence_length(1, 100, [(11, 20), (40, 60)], [(45, 50)]) == 24
    assert abs(fm.watterson_theta(2, 4, 100) - 12.0 / 11.0 / 100.0) < 1e-15
    with pytest.raises(ValueError):
        fm.watterson_theta(1, 1, 100)
    with pytest.raises(ValueError):
        fm.adjusted_sequence_length(10, 1)
    assert fm.inversion_allele_frequency({"a": (0, 1), "b": (1, 1), "c": (2, 255)}) == 0.75


def test_compute_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import ferromic_b200 as fm
    from ferromic_b200 import _lib
    n = C.c_int(-1)
    assert _lib.lib().fm_device_count(C.byref(n)) == 0 and n.value == 0
    with pytest.raises(fm.FerromicGpuError) as exc:
        fm.segregating_sites([{"position": 1, "genotypes": [[0, 1]]}])
    assert exc.value.code == _lib.FM_ERR_NO_DEVICE
    assert "no CPU fallback" in str(exc.value)


def test_adjacent_stages_fail_loudly_without_device():
    """The VCF parser and the FALSTA renderer are device code too: no CPU path behind them."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import ferromic_b200 as fm
    from ferromic_b200 import _lib, falsta, vcf
    with pytest.raises(fm.FerromicGpuError) as exc:
        vcf.process_lines(b"1\t5\t.\tA\tC\t.\t.\t.\tGT:GQ\t0|1:40\n", "1", [(0, 10)], [9], 30)
    assert exc.value.code == _lib.FM_ERR_NO_DEVICE
    with pytest.raises(fm.FerromicGpuError) as exc:
        falsta.track_lines([2], [0.5], 1, 3, falsta.FST)
    assert exc.value.code == _lib.FM_ERR_NO_DEVICE
    assert falsta.format_value(0.5, falsta.FST) == "0.500000"  # the token routine itself is host code


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ferromic_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("no CPU", ""), f
