"""Pins the CPU oracle (oracle/emvs_oracle.cpp) against the REFERENCE's own code.

Two sources, both produced from the reference's cartesian3dgrid.h/.cpp and depth_vector.hpp
compiled in place (oracle/Makefile target `ref`):
  * tests/golden/grid3d_ref.npz — committed fixtures (tests/golden/make_golden.py);
  * oracle/_ref/libgrid3d_ref.so — when present (it is git-ignored and travels with the gpurun
    snapshot), larger randomised comparisons.
Every comparison is bit-exact: same IEEE ops on identical inputs.  What stays UNPINNED is the
Eigen / minkindr arithmetic of the packet and event stages (the reference's mapper cannot be
compiled here) — see DESIGN.md §3.
"""
import os

import numpy as np
import pytest

from oracle import ref

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid3d_ref.npz"))
OPS = {"add": 0, "min": 1, "hm": 2, "gm": 3, "am": 4, "rms": 5, "max": 6, "hm_n": 7, "add_inv": 8,
       "hm_from_suminv": 9, "am_from_sum": 10}
OP_KW = {"hm": dict(eps=0.1), "hm_n": dict(n=3, eps=0.1), "add_inv": dict(eps=1e-2), "hm_from_suminv": dict(n=5),
         "am_from_sum": dict(n=5)}


def oracle_vote(O, vol, k, x, y):
    import ctypes as C
    dimZ, dimY, dimX = vol.shape
    sl = vol[k]
    acc = 0
    for xi, yi in zip(x, y):
        acc += O.lib().oracle_vote(C.c_float(xi), C.c_float(yi), sl.ctypes.data_as(C.c_void_p), dimX, dimY)
    return acc


def test_golden_vote(O):
    dimX, dimY, dimZ = GOLD["vote_dims"]
    vol = np.zeros((dimZ, dimY, dimX), np.float32)
    n_acc = oracle_vote(O, vol, 1, GOLD["vote_x"], GOLD["vote_y"])
    assert vol.tobytes() == GOLD["vote_out"].tobytes()
    assert not vol[0].any() and not vol[2].any()
    # every accepted vote adds weights summing to 1 (up to rounding): plane sum ~ accepted count
    assert abs(float(vol[1].sum(dtype=np.float64)) - n_acc) < 1e-2 and n_acc > 3000


@pytest.mark.parametrize("name", list(OPS))
def test_golden_ops(O, name):
    op = OPS[name]
    a, b = GOLD["op_a"].copy(), GOLD["op_b"]
    kw = OP_KW.get(name, {})
    with np.errstate(all="ignore"):
        got = O.fuse_op(op, a, None if op >= 9 else b, **kw)
    assert got.tobytes() == GOLD["op_" + name].tobytes()


def test_golden_collapse_and_mean_square(O):
    conf, idx = O.collapse_max(GOLD["collapse_in"])
    assert conf.tobytes() == GOLD["collapse_conf"].tobytes()
    assert np.array_equal(idx.astype(np.uint8), GOLD["collapse_idx"]) and idx.max() < 256
    assert idx[0, 0] == 0 and conf[0, 0] == 0 and idx[1, 1] == 5 and idx[2, 2] == 33 and idx[3, 3] == 39
    assert O.mean_square(GOLD["collapse_in"]) == GOLD["mean_square"][0]
    assert O.mean_square(GOLD["op_a"]) == GOLD["mean_square"][1]
    # numpy twin agrees too
    conf2, idx2 = O.np_collapse_max(GOLD["collapse_in"])
    assert np.array_equal(conf2, conf) and np.array_equal(idx2, idx)


def test_golden_depth_tables(O):
    for i, (inv, zmin, zmax, nz) in enumerate(GOLD["depth_cases"]):
        got = O.depth_vector(float(zmin), float(zmax), int(nz), bool(inv))
        assert got.tobytes() == GOLD[f"depth_{i}"].tobytes(), (inv, zmin, zmax, nz)


def test_host_geometry_depth_tables_match_golden():
    """The product's own host-side depth table (C-ABI, no GPU needed) equals the reference's."""
    import ctypes as C
    from dvs_mcemvs_b200 import _capi as capi
    lib = capi.load()
    for i, (inv, zmin, zmax, nz) in enumerate(GOLD["depth_cases"]):
        sh = capi.Shape(0, 0, int(nz), float(zmin), float(zmax), 0.0, int(inv))
        out = np.zeros(int(nz), np.float32)
        if zmin > zmax:   # the C-ABI validates like MapperEMVS::setupDSI (CHECK_GT max > min): error, not swap
            assert lib.emvs_depth_vector(C.byref(sh), capi.ptr(out)) == capi.EMVS_ERR_INVALID
            continue
        capi.check(lib.emvs_depth_vector(C.byref(sh), capi.ptr(out)))
        assert out.tobytes() == GOLD[f"depth_{i}"].tobytes()


# ---- live comparison against the compiled reference (randomised, larger) -------------------------
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgrid3d_ref.so not built")


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_vote_matches_reference(O, seed):
    rng = np.random.default_rng(seed)
    dimX, dimY = int(rng.integers(2, 90)), int(rng.integers(2, 70))
    n = 20000
    x = rng.uniform(-2, dimX + 2, n).astype(np.float32)
    y = rng.uniform(-2, dimY + 2, n).astype(np.float32)
    want = ref.vote(np.zeros((2, dimY, dimX), np.float32), 0, x, y)
    got = np.zeros((2, dimY, dimX), np.float32)
    oracle_vote(O, got, 0, x, y)
    assert got.tobytes() == want.tobytes()


@needs_ref
@pytest.mark.parametrize("name", list(OPS))
def test_live_ops_match_reference(O, name):
    rng = np.random.default_rng(OPS[name] + 50)
    shape = (6, 40, 50)
    a = np.where(rng.random(shape) < 0.3, 0, rng.gamma(1.0, 5.0, shape)).astype(np.float32)
    b = np.where(rng.random(shape) < 0.3, 0, rng.gamma(1.0, 5.0, shape)).astype(np.float32)
    op, kw = OPS[name], OP_KW.get(name, {})
    with np.errstate(all="ignore"):
        want = ref.grid_op(op, a.copy(), None if op >= 9 else b, **kw)
        got = O.fuse_op(op, a.copy(), None if op >= 9 else b, **kw)
    assert got.tobytes() == want.tobytes()


@needs_ref
def test_live_process1_fold_matches_reference(O):
    """process1.cpp:126-191 fold (reset + add, pairwise op, third camera) on the reference's Grid3D."""
    rng = np.random.default_rng(77)
    vols = [np.where(rng.random((8, 30, 31)) < 0.4, 0, rng.gamma(1, 4, (8, 30, 31))).astype(np.float32) for _ in range(3)]
    for method, pair in {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}.items():
        fused = np.zeros_like(vols[0])
        ref.grid_op(0, fused, vols[0])
        ref.grid_op(pair, fused, vols[1], n=2, eps=0.1)
        if method == 1:
            ref.grid_op(1, fused, vols[2])
        elif method == 2:
            ref.grid_op(7, fused, vols[2], n=3, eps=0.1)
        elif method == 6:
            ref.grid_op(6, fused, vols[2])
        assert O.fuse_reference(method, vols).tobytes() == fused.tobytes()
        conf_r, idx_r = ref.collapse_max(fused)
        conf_o, idx_o = O.collapse_max(fused)
        assert conf_o.tobytes() == conf_r.tobytes() and np.array_equal(idx_o.astype(np.uint8), idx_r)
        assert O.mean_square(fused) == ref.mean_square(fused)


@needs_ref
def test_live_depth_tables(O):
    rng = np.random.default_rng(5)
    for _ in range(50):
        zmin = float(np.float32(rng.uniform(0.1, 10)))
        zmax = float(np.float32(zmin + rng.uniform(0.1, 300)))
        nz = int(rng.integers(1, 600))
        for inv in (False, True):
            assert O.depth_vector(zmin, zmax, nz, inv).tobytes() == ref.depth_vector(zmin, zmax, nz, inv).tobytes()
