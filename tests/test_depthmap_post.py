"""Depth-map post-processing (mapper_emvs_stereo.cpp:393-436 minus inpainting).

CPU: the oracle's restatement of the OpenCV arithmetic against tests/golden/depthmap_ref.npz, which
was produced with cv2.normalize / cv2.adaptiveThreshold and the reference's own huangMedianFilter
compiled in place (tests/golden/make_golden_depthmap.py).  GPU: the CUDA path against both.
All outputs are bytes / indices: bit-exact."""
import os

import numpy as np
import pytest

from oracle import ref

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depthmap_ref.npz"))


def _cases():
    for i in range(int(GOLD["n_cases"])):
        ks, c, mc, ms = GOLD[f"case{i}_params"]
        yield i, dict(ks=int(ks), c=float(c), max_confidence=float(mc), median_size=int(ms))


def _check(out, i, kw):
    assert out["conf8"].tobytes() == GOLD[f"case{i}_conf8"].tobytes()
    assert out["mask"].tobytes() == GOLD[f"case{i}_mask"].tobytes()
    assert out["idx_filtered"].tobytes() == GOLD[f"case{i}_idx_filtered"].tobytes()
    assert np.array_equal(out["depth"], GOLD["depths"][GOLD[f"case{i}_idx_filtered"]])
    want_conf = GOLD["conf"].copy()
    want_conf[0, 0] = kw["max_confidence"]
    assert out["conf"].tobytes() == want_conf.tobytes()


@pytest.mark.parametrize("i,kw", list(_cases()))
def test_oracle_matches_opencv_and_reference_median(O, i, kw):
    _check(O.depth_map_post(GOLD["conf"], GOLD["idx"], GOLD["depths"], **kw), i, kw)


def test_oracle_rejects_unsupported_sizes(O):
    with pytest.raises(ValueError):
        O.depth_map_post(GOLD["conf"], GOLD["idx"], GOLD["depths"], ks=9)      # non-dyadic Gaussian kernel
    with pytest.raises(ValueError):
        O.depth_map_post(GOLD["conf"], GOLD["idx"], GOLD["depths"], median_size=4)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgrid3d_ref.so not built")
def test_oracle_median_matches_reference_live(O):
    rng = np.random.default_rng(12)
    for ms in (3, 5, 9, 15):
        img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
        mask = (rng.random((37, 53)) < 0.35).astype(np.uint8)
        conf = np.zeros((37, 53), np.float32)
        # drive the oracle's median with a given mask: ks=3, c=-300 accepts nothing, so call the reference directly
        want = ref.huang_median(img, mask, ms)
        p = ms // 2
        got = np.zeros_like(img)
        for y in range(37):
            for x in range(53):
                sl = (slice(max(y - p, 0), y + p + 1), slice(max(x - p, 0), x + p + 1))
                v = np.sort(img[sl][mask[sl] > 0])
                got[y, x] = v[(len(v) + 1) // 2 - 1] if len(v) else 0      # lower median, 0 for an empty window
        assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("i,kw", list(_cases()))
def test_cuda_depth_map_matches_golden(ctx, O, i, kw):
    from dvs_mcemvs_b200 import api
    out = api.depth_map_postprocess(ctx, GOLD["conf"], GOLD["idx"], GOLD["depths"], **kw)
    _check(out, i, kw)


@pytest.mark.gpu
def test_cuda_get_depth_map_from_dsi(ctx, O, small_case):
    """MapperEMVS.getDepthMapFromDSI(options) on a GPU-built DSI == oracle collapse + post-processing."""
    from dvs_mcemvs_b200 import api
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    m.build(small_case.events[0], small_case.packets[0])
    opts = api.OptionsDepthMap(adaptive_threshold_kernel_size=5, adaptive_threshold_c=5.0, median_filter_size=5)
    depth, conf, mask, idx_f = m.getDepthMapFromDSI(opts)
    conf_o, idx_o = O.collapse_max(m.dsi_.download())
    want = O.depth_map_post(conf_o, idx_o.astype(np.uint8), small_case.depths, ks=5, c=5.0, max_confidence=0.0, median_size=5)
    assert np.array_equal(mask, want["mask"]) and np.array_equal(idx_f, want["idx_filtered"])
    assert np.array_equal(depth, want["depth"]) and np.array_equal(conf, want["conf"])
    assert mask.any() and mask[:3].sum() == 0 and mask[:, :3].sum() == 0     # border stripped
    m.close()
