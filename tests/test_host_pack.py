"""Host packer of the 2-bit ingest format (fm_pack_rows, csrc/fm_host_pack.cpp) against a numpy restatement of
the format's definition (include/ferromic_gpu.h "packed rows"): allele bit = called & (cell != 0), called bit =
not missing, bit (c & 31) of word (c >> 5), zero past the last cell.  Runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from ferromic_b200 import _lib
from ferromic_b200.api import _pack_bits, pack_rows, pack_rows_sparse


def numpy_pack(cells_u8, missing_bool):
    rows, stride = cells_u8.shape
    rw = (stride + 31) // 32

    def words(bits):
        pad = np.zeros((rows, rw * 32), dtype=np.uint8)
        pad[:, :stride] = bits
        return np.packbits(pad.reshape(rows, rw, 32), axis=2, bitorder="little").view(np.uint32).reshape(rows, rw)

    called = ~missing_bool
    return words((cells_u8 != 0) & called), words(called)


@pytest.mark.parametrize("stride", [1, 7, 31, 32, 33, 64, 100, 257, 5008])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pack_rows_matches_numpy(stride, mode):
    rng = np.random.default_rng(1000 * stride + mode)
    rows = 37
    cells = rng.integers(0, 2, size=(rows, stride), dtype=np.uint8)
    cells[rng.random(cells.shape) < 0.05] = 3  # any non-zero allele index sets the allele bit
    miss = rng.random(cells.shape) < 0.1 if mode else np.zeros(cells.shape, dtype=bool)
    bitmap = None
    src = cells
    if mode == 1:
        bitmap = _pack_bits(miss.reshape(-1).astype(np.uint8))
        src = cells.copy()
        src[miss & (rng.random(cells.shape) < 0.5)] = 1  # the byte under a missing bit is arbitrary
    elif mode == 2:
        src = cells.copy()
        src[miss] = rng.integers(0x80, 0x100, size=int(miss.sum()), dtype=np.uint16).astype(np.uint8)
    want_a, want_c = numpy_pack(cells, miss)
    for generic in (False, True):
        for threads in ((1, 3) if not generic else (0,)):
            got_a, got_c = pack_rows(src, mode, bitmap, threads=threads, generic=generic)
            assert np.array_equal(got_a, want_a)
            if mode == 0:
                assert got_c is None
            else:
                assert np.array_equal(got_c, want_c)


def test_pack_rows_chunk_of_a_larger_matrix():
    """rows points at row first_row; the bitmap is the WHOLE matrix's (bit index = linear cell index)."""
    rng = np.random.default_rng(7)
    V, stride = 50, 77  # row starts fall on every bit offset of the u64 bitmap words
    cells = rng.integers(0, 2, size=(V, stride), dtype=np.uint8)
    miss = rng.random(cells.shape) < 0.2
    bitmap = _pack_bits(miss.reshape(-1).astype(np.uint8))
    want_a, want_c = numpy_pack(cells, miss)
    for r0, r1 in [(0, 50), (0, 13), (13, 41), (41, 50), (49, 50)]:
        for generic in (False, True):
            a, c = pack_rows(cells[r0:r1], 1, bitmap, first_row=r0, n_total_rows=V, generic=generic)
            assert np.array_equal(a, want_a[r0:r1]) and np.array_equal(c, want_c[r0:r1])


def test_pack_rows_many_threads_large():
    rng = np.random.default_rng(11)
    V, stride = 3000, 2048 + 40
    cells = (rng.random((V, stride)) < 0.3).astype(np.uint8)
    i8 = cells.astype(np.int8)
    miss = rng.random(cells.shape) < 0.01
    i8[miss] = -1
    want_a, want_c = numpy_pack(cells, miss)
    a, c = pack_rows(i8, 2, threads=8)
    assert np.array_equal(a, want_a) and np.array_equal(c, want_c)


def test_pack_rows_argument_errors():
    L = _lib.lib()
    cells = np.zeros((2, 40), dtype=np.uint8)
    out = np.zeros((2, 2), dtype=np.uint32)
    n = C.c_size_t()
    assert L.fm_packed_row_words(2504, 2, C.byref(n)) == 0 and n.value == 157
    assert L.fm_pack_rows(cells.ctypes.data, None, 1, 0, 2, 2, 40, out.ctypes.data, out.ctypes.data, 1) == _lib.FM_ERR_INVALID_ARG
    assert b"bitmap" in L.fm_last_error()
    assert L.fm_pack_rows(cells.ctypes.data, None, 2, 0, 2, 2, 40, out.ctypes.data, None, 1) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_pack_rows(cells.ctypes.data, None, 7, 0, 2, 2, 40, out.ctypes.data, None, 1) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_pack_rows(cells.ctypes.data, None, 0, 1, 2, 2, 40, out.ctypes.data, None, 1) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_pack_rows(None, None, 0, 0, 0, 0, 40, None, None, 1) == 0  # nothing to do


@pytest.mark.parametrize("stride", [1, 33, 100, 5008, 70000])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pack_rows_sparse_matches_the_dense_packer(stride, mode):
    """Allele bits identical to fm_pack_rows; the CSR list holds exactly the zero bits of its called plane, ascending."""
    rng = np.random.default_rng(31 * stride + mode)
    rows = 41 if stride < 10000 else 5
    cells = (rng.random((rows, stride)) < 0.4).astype(np.uint8)
    miss = rng.random(cells.shape) < 0.03 if mode else np.zeros(cells.shape, dtype=bool)
    miss[rows // 2] = mode != 0          # a row that is missing altogether
    miss[0] = False                      # and one without any missing cell
    bitmap, src = None, cells
    if mode == 1:
        bitmap = _pack_bits(miss.reshape(-1).astype(np.uint8))
    elif mode == 2:
        src = cells.copy()
        src[miss] = 0xFF
    dense_a, dense_c = pack_rows(src, mode, bitmap)
    for threads in (1, 4):
        ab, start, cols = pack_rows_sparse(src, mode, bitmap, threads=threads)
        assert np.array_equal(ab, dense_a)
        assert cols.dtype == (np.uint16 if stride <= 65536 else np.uint32)
        assert start[0] == 0 and np.all(np.diff(start.astype(np.int64)) >= 0) and int(start[-1]) == len(cols) == int(miss.sum())
        for r in range(rows):
            want = np.nonzero(miss[r])[0]
            assert np.array_equal(cols[int(start[r]):int(start[r + 1])].astype(np.int64), want)


def _decode_gap_code(b):
    """One row of the one-byte gap code (include/ferromic_gpu.h): the ascending columns it lists."""
    out, at = [], -1
    for x in b.tolist():
        if x == 255:
            at += 255
        else:
            at += x + 1
            out.append(at)
    return np.asarray(out, dtype=np.int64)


@pytest.mark.parametrize("stride", [1, 33, 700, 5008, 70000])
@pytest.mark.parametrize("mode", [1, 2])
def test_pack_rows_sparse_gap_code(stride, mode):
    """col_bytes == 1: every row's bytes decode to exactly the missing columns (gaps above 255 take escape bytes)."""
    rng = np.random.default_rng(77 * stride + mode)
    rows = 37 if stride < 10000 else 5
    cells = (rng.random((rows, stride)) < 0.4).astype(np.uint8)
    miss = rng.random(cells.shape) < (0.002 if stride > 600 else 0.05)   # wide rows: mostly gaps > 255
    miss[rows // 2] = True               # a row that is missing altogether (all gaps are 1)
    miss[0] = False                      # one without any missing cell
    miss[1] = False
    miss[1, stride - 1] = True           # a single cell at the far end: only escapes before it
    bitmap, src = None, cells
    if mode == 1:
        bitmap = _pack_bits(miss.reshape(-1).astype(np.uint8))
    else:
        src = cells.copy()
        src[miss] = 0xFF
    dense_a, _ = pack_rows(src, mode, bitmap)
    for threads in (1, 4):
        ab, start, code = pack_rows_sparse(src, mode, bitmap, threads=threads, gap_code=True)
        assert code.dtype == np.uint8 and np.array_equal(ab, dense_a)
        assert start[0] == 0 and int(start[-1]) == len(code)
        for r in range(rows):
            got = _decode_gap_code(code[int(start[r]):int(start[r + 1])])
            assert np.array_equal(got, np.nonzero(miss[r])[0])
        n_escape = int((code == 255).sum())
        assert len(code) == int(miss.sum()) + n_escape


def test_pack_rows_sparse_chunk_and_capacity():
    from ferromic_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    V, stride = 30, 77
    cells = rng.integers(0, 2, size=(V, stride), dtype=np.uint8)
    miss = rng.random(cells.shape) < 0.2
    bitmap = _pack_bits(miss.reshape(-1).astype(np.uint8))
    ab, start, cols = pack_rows_sparse(cells[7:19], 1, bitmap, first_row=7, n_total_rows=V)
    for r in range(12):
        assert np.array_equal(cols[int(start[r]):int(start[r + 1])].astype(np.int64), np.nonzero(miss[7 + r])[0])
    # capacity too small: the needed size and the row starts are still reported
    need = C.c_size_t()
    st2 = np.zeros(V + 1, dtype=np.uint64)
    a2 = np.zeros((V, 3), dtype=np.uint32)
    tiny = np.zeros(4, dtype=np.uint16)
    rc = L.fm_pack_rows_sparse(cells.ctypes.data, bitmap.ctypes.data, 1, 0, V, V, stride, a2.ctypes.data, st2.ctypes.data,
                               tiny.ctypes.data, 4, 2, 2, C.byref(need))
    assert rc == _lib.FM_ERR_INVALID_ARG and need.value == int(miss.sum()) == int(st2[-1])
    assert L.fm_pack_rows_sparse(cells.ctypes.data, bitmap.ctypes.data, 1, 0, V, V, stride, a2.ctypes.data, st2.ctypes.data,
                                 tiny.ctypes.data, 4, 3, 2, C.byref(need)) == _lib.FM_ERR_INVALID_ARG  # bad col_bytes
