"""Property-based pinning of the oracle's Grid3D layer against the reference's own code compiled in place
(oracle/_ref): arbitrary float32 inputs including NaN, infinities, subnormals and huge values."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgrid3d_ref.so not built")

f32 = hnp.from_dtype(np.dtype(np.float32), allow_nan=True, allow_infinity=True, allow_subnormal=True)
coords = st.one_of(f32, st.floats(-2, 40, width=32))
COMMON = dict(deadline=None, max_examples=60, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(**COMMON)
@given(dimX=st.integers(1, 33), dimY=st.integers(1, 17), pts=st.lists(st.tuples(coords, coords), min_size=1, max_size=60))
def test_vote_any_float(O, dimX, dimY, pts):
    x = np.array([p[0] for p in pts], np.float32)
    y = np.array([p[1] for p in pts], np.float32)
    with np.errstate(all="ignore"):
        want = ref.vote(np.zeros((1, dimY, dimX), np.float32), 0, x, y)
        got = np.zeros((1, dimY, dimX), np.float32)
        for xi, yi in zip(x, y):
            O.lib().oracle_vote(C.c_float(xi), C.c_float(yi), got[0].ctypes.data_as(C.c_void_p), dimX, dimY)
    assert got.tobytes() == want.tobytes()


vols = hnp.arrays(np.float32, (2, 3, 4), elements=st.one_of(st.floats(0, 1e4, width=32), f32))


@settings(**COMMON)
@given(a=vols, b=vols, op=st.integers(0, 10), n=st.integers(2, 6))
def test_voxel_ops_any_float(O, a, b, op, n):
    eps = {2: 0.1, 7: 0.1, 8: 1e-2}.get(op, 0.0)
    with np.errstate(all="ignore"):
        want = ref.grid_op(op, a.copy(), None if op >= 9 else b, n=n, eps=eps)
        got = O.fuse_op(op, a.copy(), None if op >= 9 else b, n=n, eps=eps)
    assert got.tobytes() == want.tobytes()


@settings(**COMMON)
@given(v=hnp.arrays(np.float32, (7, 3, 5), elements=st.one_of(st.floats(-10, 10, width=32), st.just(np.float32(0)))))
def test_collapse_any_finite(O, v):
    conf_r, idx_r = ref.collapse_max(v)
    conf_o, idx_o = O.collapse_max(v)
    assert conf_o.tobytes() == conf_r.tobytes() and np.array_equal(idx_o.astype(np.uint8), idx_r)
