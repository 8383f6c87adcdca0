"""CUDA path vs the committed golden vectors produced by the REFERENCE's own Grid3D /
depth_vector.hpp (tests/golden/make_golden.py) — runs on the GPU box, where /root/reference
does not exist.  Voxel ops, collapse, depth tables and the checksum are bit-exact; the bilinear
vote sums the same float32 weights in a different order (atomics), so voxels agree to float-sum
tolerance and the accepted-vote count is exact."""
import os

import numpy as np
import pytest

from dvs_mcemvs_b200 import _capi as capi
from dvs_mcemvs_b200 import api

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid3d_ref.npz"))
OPS = {"add": 0, "min": 1, "hm": 2, "gm": 3, "am": 4, "rms": 5, "max": 6, "hm_n": 7, "add_inv": 8,
       "hm_from_suminv": 9, "am_from_sum": 10}
OP_KW = {"hm": dict(eps=0.1), "hm_n": dict(n=3, eps=0.1), "add_inv": dict(eps=1e-2), "hm_from_suminv": dict(n=5),
         "am_from_sum": dict(n=5)}


def test_vote_matches_reference_grid3d(ctx):
    """Feed the golden sub-pixel positions through the engine unchanged: LUT = the positions,
    H = identity, C = 0, one plane at z0 = 1 -> X = (X0*1 + 0)/1 = X0 exactly."""
    x, y = GOLD["vote_x"], GOLD["vote_y"]
    dimX, dimY, _ = [int(v) for v in GOLD["vote_dims"]]
    W, H = 128, 48
    lut = np.full((W * H, 2), np.nan, np.float32)
    lut[:x.shape[0], 0], lut[:x.shape[0], 1] = x, y
    cam = api.CameraModel(W, H, 100., 100., 64., 24., lut=lut)
    m = api.MapperEMVS(ctx, cam, api.ShapeDSI(dimX, dimY, 1, 1.0, 2.0, 0.0))
    assert m.raw_depths_vec_[0] == 1.0
    ev = np.zeros(W * H, capi.EVENT_DTYPE)
    ev["x"], ev["y"] = np.tile(np.arange(W), H), np.repeat(np.arange(H), W)
    pk = np.zeros(W * H // 1024, capi.PACKET_DTYPE)
    pk["H"][:] = np.eye(3, dtype=np.float32).ravel()
    pk["first_event"] = np.arange(pk.shape[0]) * 1024
    m.build(ev, pk)
    want = GOLD["vote_out"][1]
    got = m.dsi_.download()[0]
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # accepted votes: weights of one vote sum to 1 (up to rounding)
    assert int(m.counts()[0]) == int(round(float(want.sum(dtype=np.float64))))
    m.close()


@pytest.mark.parametrize("name", list(OPS))
def test_ops_match_reference_grid3d(ctx, name):
    a, b = GOLD["op_a"], GOLD["op_b"]
    dimZ, dimY, dimX = a.shape
    ga, gb = api.Grid3D(ctx, dimX, dimY, dimZ), api.Grid3D(ctx, dimX, dimY, dimZ)
    ga.upload(a)
    gb.upload(b)
    op, kw = OPS[name], OP_KW.get(name, {})
    ga._op(None if op >= 9 else gb, op, kw.get("n", 0), kw.get("eps", 0.0))
    assert ga.download().tobytes() == GOLD["op_" + name].tobytes()
    ga.close()
    gb.close()


def test_collapse_and_checksum_match_reference_grid3d(ctx):
    c = GOLD["collapse_in"]
    dimZ, dimY, dimX = c.shape
    g = api.Grid3D(ctx, dimX, dimY, dimZ)
    g.upload(c)
    conf, idx = g.collapseMaxZSlice()
    assert idx.dtype == np.uint8
    assert conf.tobytes() == GOLD["collapse_conf"].tobytes() and np.array_equal(idx, GOLD["collapse_idx"])
    assert g.computeMeanSquare() == pytest.approx(GOLD["mean_square"][0], rel=1e-12)
    g.close()


def test_depth_tables_match_reference(ctx):
    cam = api.CameraModel(64, 48, 50., 50., 32., 24.)
    for i, (inv, zmin, zmax, nz) in enumerate(GOLD["depth_cases"]):
        if zmin > zmax:
            continue  # MapperEMVS::setupDSI CHECKs max > min before the depth vector would swap them
        m = api.MapperEMVS(ctx, cam, api.ShapeDSI(8, 8, int(nz), float(zmin), float(zmax), 0.0, bool(inv)))
        assert m.raw_depths_vec_.tobytes() == GOLD[f"depth_{i}"].tobytes()
        m.close()
