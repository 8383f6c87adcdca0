"""The benchmarked configurations at FULL size against the CPU oracle (pytest -m gpu): BASELINE.json configs[1]
(DSEC-like stereo pair, 5 M events per camera, 640x480x256, harmonic fusion — what bench.py times) and configs[2]
(4-camera bar, 10 M events per camera, n-ary geometric mean).  Every plane, every camera, the whole chain
evaluateDSI -> fusion -> Z-argmax -> depth:

  * per-plane accepted-vote counts                bit-exact
  * computeMeanSquare of every DSI                1e-5 relative (the reference's own run-time checksum, process1.cpp:86)
  * fused confidence                              1e-4 relative (+1e-6 absolute), the north-star tolerance
  * arg-max index / depth                         identical except at near-ties of the oracle's own fused column

The oracle needs a few seconds per camera on the GPU box's host cores at these sizes."""
import numpy as np
import pytest

from dvs_mcemvs_b200 import api

from conftest import Case

pytestmark = pytest.mark.gpu


def _check_full(ctx, O, case, lists_of=lambda ev: ev):
    from oracle import parity as P
    trs = [api.LinearTrajectory(t) for t in case.trajs]
    mappers = [api.MapperEMVS(ctx, c, case.shape) for c in case.cams]
    vols = []
    try:
        for i, m in enumerate(mappers):
            assert m.evaluateDSI(lists_of(case.events[i]), trs[i], case.T_rv_w)
            dsi_o, inb_o = case.oracle_dsi(i)
            assert np.array_equal(m.counts(), inb_o), f"camera {i}: per-plane vote counts differ from the oracle"
            ms_g, ms_o = m.dsi_.computeMeanSquare(), O.mean_square(dsi_o)
            assert ms_g == pytest.approx(ms_o, rel=P.MS_RTOL), f"camera {i}: mean square {ms_g} vs {ms_o}"
            vols.append(dsi_o)
        conf, idx, depth = api.fuse_collapse([m.dsi_ for m in mappers], case.method, case.depths)
        fused_o = O.fuse_reference(case.method, vols) if len(vols) <= 3 else O.fuse_nary(case.method, vols)
        del vols
        conf_o, idx_o, depth_o = O.collapse_max(fused_o, case.depths)
        p = P.compare_maps(conf, idx, depth, conf_o, idx_o, depth_o, fused_o)
        p["counts_exact"] = True
        assert p["conf_max_err_over_tol"] <= 1.0, p
        assert p["idx_mismatches_are_near_ties"] and p["idx_agreement"] > 0.999, p
        assert p["depth_exact_where_idx_agrees"], p
        assert P.verdict(p)
        assert conf.max() > 10.0          # a structured scene: rays really intersect
        return p
    finally:
        for m in mappers:
            m.close()


def test_configs1_dsec_stereo_full(ctx, O):
    """bench.py's default workload, exactly: both cameras, 5 M events each, all 256 planes, harmonic mean."""
    case = Case("dsec_stereo", events_per_cam=5_000_000)
    assert case.n_cams == 2 and case.method == 2 and (case.dimX, case.dimY, case.dimZ) == (640, 480, 256)
    assert all(len(pk) == 4882 for pk in case.packets)
    _check_full(ctx, O, case)


def test_configs1_dsec_stereo_full_soa(ctx, O):
    """The same through the structure-of-arrays entry point (emvs_mapper_evaluate_dsi_soa)."""
    case = Case("dsec_stereo", events_per_cam=2_000_000)
    _check_full(ctx, O, case, lists_of=api.EventsSoA.from_events)


def test_configs2_bar4_full(ctx, O):
    """BASELINE.json configs[2] at its stated size: 4 cameras x 10 M events, n-ary geometric mean (an extension
    whose n = 2 case is the reference's sqrt(a*b), tests/test_gpu_parity.py::test_config3_four_camera_geometric_mean)."""
    case = Case("bar4", events_per_cam=10_000_000)
    assert case.n_cams == 4 and case.method == 3
    _check_full(ctx, O, case)
