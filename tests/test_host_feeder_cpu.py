"""Host side of the genotype ingress (csrc/feeder.cpp) without a GPU: the conversion of host matrices to int8 dosages that the
feeder's worker threads run -- values, integrality / range verdicts, strided sources, every element type."""
import ctypes

import numpy as np
import pytest

from cellregmap_b200 import _lib
from cellregmap_b200._cellregmap import _G_DTYPES


def _narrow(arr, cols=None):
    lib = _lib.load()
    rows, width = arr.shape
    cols = width if cols is None else cols
    out = np.full((rows, cols + 3), 99, dtype=np.int8)
    bad, gmax = ctypes.c_int32(-1), ctypes.c_int32(-1)
    item = arr.dtype.itemsize
    _lib.call("crm_host_narrow", ctypes.c_void_p(arr.ctypes.data), _G_DTYPES[arr.dtype], arr.strides[0] // item, rows, cols,
              ctypes.c_void_p(out.ctypes.data), out.strides[0], ctypes.byref(bad), ctypes.byref(gmax))
    assert lib.crm_host_threads() >= 1
    assert np.all(out[:, cols:] == 99)          # nothing written beyond the block
    return out[:, :cols], bad.value, gmax.value


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int8, np.uint8, np.int16, np.int32, np.int64])
@pytest.mark.parametrize("shape", [(1, 1), (7, 15), (130, 16), (257, 333), (1000, 49)])
def test_narrow_dosages(dtype, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    g = rng.integers(0, 3, shape).astype(dtype)
    out, bad, gmax = _narrow(g)
    assert bad == 0 and gmax == int(g.max())
    np.testing.assert_array_equal(out, g.astype(np.int8))


def test_narrow_signed_range_and_strides():
    rng = np.random.default_rng(3)
    wide = rng.integers(-127, 128, (300, 500)).astype(np.float64)
    block = wide[:, 37:37 + 401]                       # a column block of a wider row-major matrix
    out, bad, gmax = _narrow(block)
    assert bad == 0 and gmax == int(np.abs(block).max())
    np.testing.assert_array_equal(out, block.astype(np.int8))
    neg0 = np.zeros((5, 40))
    neg0[2, 3] = -0.0
    out, bad, gmax = _narrow(neg0)
    assert bad == 0 and gmax == 0 and not out.any()


@pytest.mark.parametrize("value", [0.5, 1e-300, 127.5, 128.0, -128.0, 1e10, -3e200, np.nan, np.inf, -np.inf, 2.0000000000000004])
@pytest.mark.parametrize("position", [(0, 0), (11, 15), (11, 16), (63, 39)])
def test_narrow_flags_anything_but_small_integers(value, position):
    g = np.random.default_rng(0).integers(0, 3, (64, 40)).astype(np.float64)
    g[position] = value
    _, bad, _ = _narrow(g)
    assert bad == 1
    v32 = np.float32(value)
    if not np.isfinite(v32) or abs(v32) > 127 or v32 != np.floor(v32):        # still not a small integer after rounding to float32
        assert _narrow(g.astype(np.float32))[1] == 1


def test_narrow_integer_types_out_of_range():
    for dtype, value in ((np.int16, 128), (np.int16, -128), (np.int32, 70000), (np.uint8, 200), (np.int8, -128), (np.int64, 2 ** 40)):
        g = np.ones((33, 21), dtype=dtype)
        g[32, 20] = value
        _, bad, _ = _narrow(g)
        assert bad == 1, (dtype, value)


def _blocks(p, basis_cols, monkeypatch, cap=2816, env_block=None):
    monkeypatch.setenv("CRM_FEEDER_CAP", str(cap))           # the cap of a pool of 16 threads, whatever this process has
    if env_block is not None:
        monkeypatch.setenv("CRM_FEEDER_BLOCK", str(env_block))
    starts = (ctypes.c_int64 * 4096)()
    nb = ctypes.c_int32(0)
    _lib.call("crm_feeder_blocks", p, basis_cols, starts, 4096, ctypes.byref(nb))
    return np.array(starts[: nb.value + 1], dtype=np.int64)


def _waves(widths, basis_cols, sms=148):
    m_tiles = -(-basis_cols // 128)
    return sum(-(-(m_tiles * -(-int(w) // 256)) // sms) for w in widths)


def test_feeder_block_schedule(monkeypatch):
    """Column blocks of the feeder (abi.cu: feeder_block_starts): they tile [0, p) in order, stay under the cap, and -- when the width of
    the basis operand is known -- are whole 256-SNP tiles whose int8 contractions end on full waves of the persistent grid: never more
    waves than equal blocks, wider blocks first."""
    for p, basis, cap in ((10000, 11964, 2816), (10000, 21462, 2816), (2000, 2500, 2816), (777, 11964, 2816), (100000, 11964, 3072), (5000, 0, 2816),
                          (1250, 11964, 512), (10000, 11964, 1536)):
        starts = _blocks(p, basis, monkeypatch, cap=cap)
        widths = np.diff(starts)
        assert starts[0] == 0 and starts[-1] == p and np.all(widths > 0)
        assert widths.max() <= cap
        if basis > 0 and len(widths) > 1:
            assert np.all(widths[:-1] % 256 == 0)
            cap = int(widths.max())
            nb = -(-p // cap)
            eq = min(p, -(-(-(-p // nb)) // 256) * 256)
            equal = [min(eq, p - s) for s in range(0, p, eq)]
            assert _waves(widths, basis) <= _waves(equal, basis)
            assert np.all(np.diff(widths[:-1]) <= 0)            # wider blocks first (the last one takes the ragged rest)
    # the bench case with 16 feeder threads: {11, 11, 11, 7} tiles = 26 waves where four blocks of 10 tiles cost 28
    starts = _blocks(10000, 11964, monkeypatch)
    assert list(np.diff(starts)) == [2816, 2816, 2816, 1552]
    # explicit block width (tests of the scan use many small blocks)
    starts = _blocks(1000, 11964, monkeypatch, env_block=96)
    assert list(np.diff(starts)) == [96] * 10 + [40]
