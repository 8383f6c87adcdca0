"""Alg. 2 (process_2) and its shuffled variant (process_5) through the Python mirror vs the same sequence of
oracle calls (process2.cpp:98-289, process5.cpp:89-250)."""
import numpy as np
import pytest

from dvs_mcemvs_b200 import api

pytestmark = pytest.mark.gpu

PAIR = {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}


def oracle_process_2(O, case, n_sub, stereo, temporal, shuffle):
    shape = (case.dimZ, case.dimY, case.dimX)
    fused, left, right = (np.zeros(shape, np.float32) for _ in range(3))
    sl = [api.subinterval_slices(len(case.events[0]), n_sub),
          api.subinterval_slices(len(case.events[1]), n_sub, n_sub // 2 if shuffle else 0)]
    for k in range(n_sub):
        d = []
        for i in range(2):
            c, e = case.cams[i], case.events[i]
            sub = np.concatenate([e[lo:hi] for lo, hi in sl[i][k]])
            pk = O.packetize(sub, case.trajs[i], case.T_rv_w, np.array([c.fx, c.fy, c.cx, c.cy], np.float32), case.virts[i],
                             case.depths[0])
            d.append(O.build_dsi(sub, pk, c.lut, c.width, case.depths, case.virts[i], case.dimX, case.dimY)[0])
        fs = O.fuse_reference(stereo, d)
        if temporal == 2:
            for acc, v in ((left, d[0]), (right, d[1]), (fused, fs)):
                O.fuse_op(O.OP_ADD_INV, acc, v, eps=1e-2)
        elif temporal == 4:
            for acc, v in ((left, d[0]), (right, d[1]), (fused, fs)):
                O.fuse_op(O.OP_ADD, acc, v)
    fin = {2: O.OP_HM_FROM_SUMINV, 4: O.OP_AM_FROM_SUM}.get(temporal)
    if fin is not None:
        for acc in (left, right, fused):
            O.fuse_op(fin, acc, None, n=n_sub)
    return dict(fused=fused, left=left, right=right, camera_time=O.fuse_reference(stereo, [left, right]))


def test_subinterval_slices_follow_the_reference():
    assert api.subinterval_slices(103, 4) == [[(0, 25)], [(25, 50)], [(50, 75)], [(75, 100)]]          # remainder dropped
    # process_5: start at sub-interval n/2 and wrap (process5.cpp:136-150)
    assert api.subinterval_slices(100, 4, 2) == [[(50, 75)], [(75, 100), (0, 0)], [(0, 25)], [(25, 50)]]
    assert api.subinterval_slices(103, 4, 2) == [[(50, 75)], [(75, 100)], [(100, 103), (0, 22)], [(22, 47)]]


@pytest.mark.parametrize("stereo,temporal,shuffle", [(2, 4, False), (2, 2, False), (4, 4, True), (1, 2, True), (3, 1, False)])
def test_process_2_matches_oracle(ctx, O, small_case, stereo, temporal, shuffle):
    n_sub = 4
    got = api.process_2(ctx, small_case.cams, [api.LinearTrajectory(t) for t in small_case.trajs], small_case.events,
                        small_case.shape, n_sub, small_case.T_rv_w, stereo, temporal, shuffle=shuffle)
    want = oracle_process_2(O, small_case, n_sub, stereo, temporal, shuffle)
    for name in ("left", "right", "fused", "camera_time"):
        g = got[name].download()
        w = want[name]
        if temporal == 2:   # HM over time divides by sums of 1/(0.01 + x): compare where the result is not dominated by eps
            np.testing.assert_allclose(g, w, rtol=2e-4, atol=1e-5, err_msg=name)
        else:
            np.testing.assert_allclose(g, w, rtol=1e-4, atol=1e-5, err_msg=name)
    if temporal not in (2, 4):   # the reference's empty cases: nothing is accumulated over time
        assert not got["fused"].download().any()
    for g in got.values():
        g.close()
    with pytest.raises(ValueError):
        api.process_2(ctx, small_case.cams, [], small_case.events, small_case.shape, 4, small_case.T_rv_w, 9, 4)
