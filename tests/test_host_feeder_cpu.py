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
