"""CUDA path vs CPU oracle through the C-ABI (pytest -m gpu).

Bars (BASELINE.json north_star / SURVEY.md §8c):
  * integer observables (per-plane accepted-vote counts inb[k], argmax indices away from
    float ties) — bit-exact;
  * float DSI — the GPU sums the same float32 weights in a different order, so voxels agree to
    float-summation tolerance (rtol 1e-5 + atol 1e-5 on values that are sums of <= ~1e3 weights);
  * fused confidence / depth — within 1e-4 relative (the tolerance the north star states).
"""
import numpy as np
import pytest

from dvs_mcemvs_b200 import _capi as capi
from dvs_mcemvs_b200 import api

from conftest import Case

pytestmark = pytest.mark.gpu

DSI_RTOL, DSI_ATOL = 1e-5, 1e-5
MAP_RTOL = 1e-4  # north-star tolerance for depth / confidence maps


def make_mappers(ctx, case):
    return [api.MapperEMVS(ctx, c, case.shape) for c in case.cams]


def assert_argmax_matches(idx_gpu, conf_gpu, fused_oracle, idx_o, conf_o):
    """Indices must be identical except where the oracle's own column has a near-tie."""
    np.testing.assert_allclose(conf_gpu, conf_o, rtol=MAP_RTOL, atol=1e-6)
    diff = idx_gpu.astype(np.int64) != idx_o.astype(np.int64)
    if diff.any():
        ys, xs = np.nonzero(diff)
        a = fused_oracle[idx_gpu[ys, xs].astype(np.int64), ys, xs]
        b = fused_oracle[idx_o[ys, xs].astype(np.int64), ys, xs]
        np.testing.assert_allclose(a, b, rtol=MAP_RTOL, atol=1e-6)  # only float near-ties may flip
        assert diff.mean() < 1e-3


# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def built_small(ctx, small_case):
    mappers = make_mappers(ctx, small_case)
    for m, ev, pk in zip(mappers, small_case.events, small_case.packets):
        m.build(ev, pk)
    oracle = [small_case.oracle_dsi(i) for i in range(small_case.n_cams)]
    yield mappers, oracle
    for m in mappers:
        m.close()


def test_setup_matches_oracle(ctx, small_case, built_small):
    mappers, _ = built_small
    for m, virt in zip(mappers, small_case.virts):
        assert m.raw_depths_vec_.tobytes() == small_case.depths.tobytes()
        assert m.virtual_cam_.tobytes() == virt.tobytes()


def test_build_counts_bit_exact_and_dsi_close(small_case, built_small):
    mappers, oracle = built_small
    for m, (dsi_o, inb_o) in zip(mappers, oracle):
        assert np.array_equal(m.counts(), inb_o)  # integer observable: bit-exact
        dsi = m.dsi_.download()
        np.testing.assert_allclose(dsi, dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
        # plane-sum KAT: bilinear weights sum to 1 -> plane sum == accepted votes
        np.testing.assert_allclose(dsi.reshape(dsi.shape[0], -1).sum(1, dtype=np.float64), inb_o.astype(np.float64),
                                   rtol=1e-5)


def test_evaluate_dsi_whole_path(ctx, small_case, built_small, O):
    """evaluateDSI through the product's own host packet stage equals oracle packet stage + build."""
    _, oracle = built_small
    m = api.MapperEMVS(ctx, small_case.cams[1], small_case.shape)
    assert m.evaluateDSI(small_case.events[1], api.LinearTrajectory(small_case.trajs[1]), small_case.T_rv_w) is True
    assert np.array_equal(m.counts(), oracle[1][1])
    np.testing.assert_allclose(m.dsi_.download(), oracle[1][0], rtol=DSI_RTOL, atol=DSI_ATOL)
    # fewer than one packet -> false, like mapper_emvs_stereo.cpp:71-75
    assert m.evaluateDSI(small_case.events[1][:1000], api.LinearTrajectory(small_case.trajs[1]),
                         small_case.T_rv_w) is False
    m.close()


@pytest.mark.parametrize("percent", [10, 25, 50, 90])
def test_evaluate_dsi_split_upload(ctx, small_case, built_small, percent):
    """evaluateDSI on an idle pipeline votes the head of the list while the tail is uploaded, then accumulates the
    tail: same per-plane counts (bit-exact) and the same DSI up to float order as the one-piece build / the oracle."""
    _, oracle = built_small
    tr = api.LinearTrajectory(small_case.trajs[0])
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    try:
        ctx.sync()
        ctx.set_upload_split(percent, 4096)
        n0 = ctx.launch_count()
        assert m.evaluateDSI(small_case.events[0], tr, small_case.T_rv_w) is True
        n_split = ctx.launch_count() - n0
        assert np.array_equal(m.counts(), oracle[0][1])
        np.testing.assert_allclose(m.dsi_.download(), oracle[0][0], rtol=DSI_RTOL, atol=DSI_ATOL)
        ctx.sync()
        ctx.set_upload_split(0)
        n0 = ctx.launch_count()
        assert m.evaluateDSI(small_case.events[0], tr, small_case.T_rv_w) is True
        n_whole = ctx.launch_count() - n0
        assert n_split == 2 * n_whole      # head build + tail build really ran as two passes
        assert np.array_equal(m.counts(), oracle[0][1])
        np.testing.assert_allclose(m.dsi_.download(), oracle[0][0], rtol=DSI_RTOL, atol=DSI_ATOL)
        # a second evaluateDSI with the split on must RESET, not keep accumulating into the previous result
        ctx.set_upload_split(percent, 4096)
        ctx.sync()
        assert m.evaluateDSI(small_case.events[0], tr, small_case.T_rv_w) is True
        assert np.array_equal(m.counts(), oracle[0][1])
        np.testing.assert_allclose(m.dsi_.download(), oracle[0][0], rtol=DSI_RTOL, atol=DSI_ATOL)
    finally:
        ctx.set_upload_split(15)
        m.close()


def test_prefetch_events(ctx, small_case, built_small):
    """A list announced with prefetch_events is uploaded ahead of the NEXT evaluateDSI; that call consumes it when it
    names the same list and drops it otherwise (emvs_b200.h lifetime rule: an announcement never outlives the next
    host-buffer build, so it cannot be matched to an unrelated later list at the same address)."""
    import time
    _, oracle = built_small
    trs = [api.LinearTrajectory(t) for t in small_case.trajs]
    ms = [api.MapperEMVS(ctx, c, small_case.shape) for c in small_case.cams[:2]]
    pinned = []
    for i in range(2):
        buf = api.pinned_empty(small_case.events[i].shape, small_case.events[i].dtype)
        buf[...] = small_case.events[i]
        pinned.append(buf)

    def check(i):
        assert np.array_equal(ms[i].counts(), oracle[i][1])
        np.testing.assert_allclose(ms[i].dsi_.download(), oracle[i][0], rtol=DSI_RTOL, atol=DSI_ATOL)
    try:
        assert ctx.prefetch_pending() == 0
        ctx.prefetch_events(pinned[0])
        g1 = ctx.prefetch_pending()
        assert g1 > 0
        # white box: once the prefetch copy has landed, scribble over the pixel coordinates on the host (timestamps,
        # which the host packet stage reads, stay).  Only the prefetched device copy still holds the real events.
        time.sleep(0.2)
        saved = pinned[0]["x"].copy()
        pinned[0]["x"] = 0
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        pinned[0]["x"] = saved
        check(0)
        assert ctx.prefetch_pending() == 0                                   # consumed
        # an unrelated list in between DROPS the announcement; the announced list is then uploaded the ordinary way
        ctx.prefetch_events(pinned[0])
        assert ctx.prefetch_pending() > g1                                   # every prefetch gets a new generation
        assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
        assert ctx.prefetch_pending() == 0
        time.sleep(0.2)
        half = pinned[0].copy()
        pinned[0]["x"][::2] = 1                                              # the list changed after the (dropped) prefetch ...
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)       # ... and THIS content is what gets voted
        changed = ms[0].dsi_.download()
        pinned[0][...] = half
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        check(0); check(1)
        assert not np.allclose(changed, oracle[0][0], rtol=1e-3, atol=1e-3)  # the stale device copy was NOT used
        # a second prefetch replaces the first
        ctx.prefetch_events(pinned[0])
        ctx.prefetch_events(pinned[1])
        assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
        check(1)
        # cancel: the arrays are the caller's again, the next call uploads what is there now
        ctx.prefetch_events(pinned[0])
        ctx.prefetch_cancel()
        assert ctx.prefetch_pending() == 0
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        check(0)
        # steady state of a streaming caller: prefetch the next window's first list before collecting this window's maps
        for _ in range(3):
            assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
            assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
            ctx.prefetch_events(pinned[0])
            conf, idx, depth = api.fuse_collapse([m.dsi_ for m in ms], 2, ms[0].raw_depths_vec_)
        check(0); check(1)
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)      # consume the last prefetch
        check(0)
        with pytest.raises(ValueError):
            ctx.prefetch_events(small_case.events[0][::2])
    finally:
        ctx.prefetch_cancel()
        for m in ms:
            m.close()


def test_prefetch_dsi(ctx, small_case, built_small):
    """MapperEMVS.prefetch runs the upload and the host packet stage of the next evaluateDSI ahead of time; that
    call with the same arguments only launches kernels, any other call falls back / drops the announcement."""
    import time
    _, oracle = built_small
    trs = [api.LinearTrajectory(t) for t in small_case.trajs]
    ms = [api.MapperEMVS(ctx, c, small_case.shape) for c in small_case.cams[:2]]
    m0b = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    pinned = []
    for i in range(2):
        buf = api.pinned_empty(small_case.events[i].shape, small_case.events[i].dtype)
        buf[...] = small_case.events[i]
        pinned.append(buf)

    def check(m, i):
        assert np.array_equal(m.counts(), oracle[i][1])
        np.testing.assert_allclose(m.dsi_.download(), oracle[i][0], rtol=DSI_RTOL, atol=DSI_ATOL)
    try:
        # white box: after the prefetch has landed, destroy coordinates AND timestamps on the host; only a call that
        # uses the prefetched events and the prefetched packets can still produce the right DSI.  Builds from
        # device-resident inputs in between leave the announcement alone.
        assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
        ms[0].prefetch(pinned[0], trs[0], small_case.T_rv_w)
        time.sleep(0.2)
        saved = pinned[0].copy()
        pinned[0]["x"] = 0
        pinned[0]["sec"] = 0
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        pinned[0][...] = saved
        check(ms[0], 0); check(ms[1], 1)
        # another mapper than the one announced: falls back to the prefetched events + its own packet stage
        ms[0].prefetch(pinned[0], trs[0], small_case.T_rv_w)
        assert m0b.evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        check(m0b, 0)
        # an unrelated evaluateDSI drops the announcement (and its reserved packet buffer)
        ms[0].prefetch(pinned[0], trs[0], small_case.T_rv_w)
        assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
        assert ctx.prefetch_pending() == 0
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        check(ms[0], 0); check(ms[1], 1)
        # steady state of a streaming caller
        for _ in range(3):
            assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
            assert ms[1].evaluateDSI(pinned[1], trs[1], small_case.T_rv_w)
            ms[0].prefetch(pinned[0], trs[0], small_case.T_rv_w)
            api.fuse_collapse([m.dsi_ for m in ms], 2, ms[0].raw_depths_vec_)
            check(ms[0], 0); check(ms[1], 1)
        assert ms[0].evaluateDSI(pinned[0], trs[0], small_case.T_rv_w)
        check(ms[0], 0)
        # fewer than 1024 events: nothing is prefetched, the later call returns False as always
        ms[0].prefetch(pinned[0][:1000].copy(), trs[0], small_case.T_rv_w)
        assert ms[0].evaluateDSI(pinned[0][:1000], trs[0], small_case.T_rv_w) is False
    finally:
        ctx.prefetch_cancel()
        for m in ms + [m0b]:
            m.close()


def test_prepared_division_is_ieee(ctx):
    """The vote kernel's shared-divisor division (prepared reciprocal + 3 FFMAs per numerator, EMVS_VOTE_FASTDIV)
    must equal __fdiv_rn bit for bit over its whole operand range: device self-test over ~2^30 random and adversarial
    pairs (all-ones / power-of-two mantissas, exact and one-ulp-off multiples of the divisor)."""
    for seed in (1, 2, 3, 4):
        assert ctx.selftest_division(1 << 28, seed) == 0


def test_mean_square(built_small, O):
    mappers, oracle = built_small
    for m, (dsi_o, _) in zip(mappers, oracle):
        assert m.dsi_.computeMeanSquare() == pytest.approx(O.mean_square(m.dsi_.download()), rel=1e-12)
        assert m.dsi_.computeMeanSquare() == pytest.approx(O.mean_square(dsi_o), rel=1e-5)


@pytest.mark.parametrize("slab", [1, 3, 7, 64])
def test_slab_size_does_not_change_counts(ctx, small_case, built_small, slab):
    _, oracle = built_small
    ctx.set_slab(slab)
    try:
        m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
        m.build(small_case.events[0], small_case.packets[0])
        assert np.array_equal(m.counts(), oracle[0][1])
        np.testing.assert_allclose(m.dsi_.download(), oracle[0][0], rtol=DSI_RTOL, atol=DSI_ATOL)
        m.close()
    finally:
        ctx.set_slab(0)


def test_accumulate_equals_unsharded(ctx, small_case, built_small):
    """Packet-aligned sub-interval shards voted with ACCUMULATE add up to the unsharded DSI."""
    _, oracle = built_small
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    pk = small_case.packets[0]
    cut = [0, 11, 30, len(pk)]
    for s in range(3):
        m.build(small_case.events[0], pk[cut[s]:cut[s + 1]], accumulate=(s > 0))
    assert np.array_equal(m.counts(), oracle[0][1])
    np.testing.assert_allclose(m.dsi_.download(), oracle[0][0], rtol=DSI_RTOL, atol=DSI_ATOL)
    m.close()


# ------------------------------------------------------------------------------------------
# fusion ops: every Grid3D op, bit-exact (same IEEE ops on identical inputs)
# ------------------------------------------------------------------------------------------
OPS = [("addTwoGrids", capi.OP_ADD, {}), ("minTwoGrids", capi.OP_MIN, {}), ("harmonicMeanTwoGrids", capi.OP_HM, {"eps": 0.1}),
       ("geometricMeanTwoGrids", capi.OP_GM, {}), ("arithmeticMeanTwoGrids", capi.OP_AM, {}),
       ("rmsTwoGrids", capi.OP_RMS, {}), ("maxTwoGrids", capi.OP_MAX, {}),
       ("harmonicMean3", capi.OP_HM_N, {"n": 3, "eps": 0.1}), ("addInverseOfTwoGrids", capi.OP_ADD_INV, {"eps": 1e-2}),
       ("computeHMfromSumOfInv", capi.OP_HM_FROM_SUMINV, {"n": 5}), ("computeAMfromSum", capi.OP_AM_FROM_SUM, {"n": 5})]


@pytest.mark.parametrize("name,op,kw", OPS, ids=[o[0] for o in OPS])
def test_grid_ops_bit_exact(ctx, O, name, op, kw):
    rng = np.random.default_rng(11)
    shape = (5, 33, 47)  # odd sizes: exercises tails
    a = rng.gamma(1.0, 3.0, shape).astype(np.float32)
    b = rng.gamma(1.0, 3.0, shape).astype(np.float32)
    a[rng.random(shape) < 0.3] = 0.0  # DSIs are sparse: HM(x,0)=0, GM(x,0)=0
    b[rng.random(shape) < 0.3] = 0.0
    ga, gb = api.Grid3D(ctx, 47, 33, 5), api.Grid3D(ctx, 47, 33, 5)
    ga.upload(a)
    gb.upload(b)
    unary = op in (capi.OP_HM_FROM_SUMINV, capi.OP_AM_FROM_SUM)
    ga._op(None if unary else gb, op, kw.get("n", 0), kw.get("eps", 0.0))
    want = O.fuse_op(op, a.copy(), None if unary else b, n=kw.get("n", 0), eps=kw.get("eps", 0.0))
    got = ga.download()
    assert got.tobytes() == want.tobytes(), f"{name}: max abs diff {np.nanmax(np.abs(got - want))}"
    ga.close()
    gb.close()


def test_grid_reset_copy_roundtrip(ctx):
    rng = np.random.default_rng(5)
    a = rng.random((3, 8, 9)).astype(np.float32)
    g, h = api.Grid3D(ctx, 9, 8, 3), api.Grid3D(ctx, 9, 8, 3)
    assert not g.download().any()  # allocate() zero-fills (cartesian3dgrid.cpp:37-45)
    g.upload(a)
    h.copyFrom(g)
    assert np.array_equal(h.download(), a)
    g.resetGrid()
    assert not g.download().any()
    with pytest.raises(api.EmvsError):
        g.addTwoGrids(api.Grid3D(ctx, 9, 8, 4))  # shape mismatch is an error, not UB


# ------------------------------------------------------------------------------------------
# collapse + fused fuse/collapse
# ------------------------------------------------------------------------------------------
def test_collapse_ties_and_zero_columns(ctx, O):
    dsi = np.zeros((6, 4, 5), np.float32)
    dsi[2, 1, 1] = dsi[4, 1, 1] = 3.0      # tie -> first index (std::max_element)
    dsi[5, 0, 0] = 1.0
    dsi[0, 3, 4] = 2.0
    dsi[:, 2, 2] = -1.0
    dsi[3, 2, 2] = -0.5                    # negative column
    g = api.Grid3D(ctx, 5, 4, 6)
    g.upload(dsi)
    depths = np.linspace(1, 2, 6).astype(np.float32)
    conf, idx, depth = g.collapseMaxZSlice(depths)
    conf_o, idx_o, depth_o = O.collapse_max(dsi, depths)
    assert idx.dtype == np.uint8
    assert np.array_equal(idx, idx_o) and np.array_equal(conf, conf_o) and np.array_equal(depth, depth_o)
    assert idx[1, 1] == 2 and idx[0, 0] == 5 and idx[3, 3] == 0 and conf[3, 3] == 0.0 and idx[2, 2] == 3


@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6])
def test_process1_two_cameras(ctx, O, small_case, built_small, method):
    mappers, oracle = built_small
    vols = [o[0] for o in oracle]
    fused_o = O.fuse_reference(method, vols)
    conf_o, idx_o, depth_o = O.collapse_max(fused_o, small_case.depths)
    # (a) fused sweep on the GPU-built DSIs
    fused_g = api.Grid3D(ctx, small_case.dimX, small_case.dimY, small_case.dimZ)
    conf, idx, depth = api.fuse_collapse([m.dsi_ for m in mappers], method, small_case.depths, fused_out=fused_g)
    np.testing.assert_allclose(fused_g.download(), fused_o, rtol=MAP_RTOL, atol=1e-5)
    assert_argmax_matches(idx, conf, fused_o, idx_o, conf_o)
    same = idx == idx_o
    assert np.array_equal(depth[same], depth_o[same])
    # (b) on identical inputs (oracle DSIs uploaded) everything is bit-exact
    ga, gb = (api.Grid3D(ctx, small_case.dimX, small_case.dimY, small_case.dimZ) for _ in range(2))
    ga.upload(vols[0])
    gb.upload(vols[1])
    conf2, idx2, depth2 = api.fuse_collapse([ga, gb], method, small_case.depths, fused_out=fused_g)
    assert fused_g.download().tobytes() == fused_o.tobytes()
    assert np.array_equal(conf2, conf_o) and np.array_equal(idx2, idx_o) and np.array_equal(depth2, depth_o)
    # (c) pairwise Grid3D calls, as process_1 issues them
    fused_g.copyFrom(ga)
    getattr(fused_g, {1: "minTwoGrids", 2: "harmonicMeanTwoGrids", 3: "geometricMeanTwoGrids",
                      4: "arithmeticMeanTwoGrids", 5: "rmsTwoGrids", 6: "maxTwoGrids"}[method])(gb)
    assert fused_g.download().tobytes() == fused_o.tobytes()
    for g in (ga, gb, fused_g):
        g.close()


@pytest.mark.parametrize("method", [1, 2, 6])
def test_three_camera_fusion_matches_reference_fold(ctx, O, method):
    rng = np.random.default_rng(method)
    vols = [np.where(rng.random((7, 10, 12)) < 0.4, 0, rng.gamma(1, 4, (7, 10, 12))).astype(np.float32) for _ in range(3)]
    grids = []
    for v in vols:
        g = api.Grid3D(ctx, 12, 10, 7)
        g.upload(v)
        grids.append(g)
    fused_o = O.fuse_reference(method, vols)
    conf_o, idx_o = O.collapse_max(fused_o)
    out = api.Grid3D(ctx, 12, 10, 7)
    conf, idx = api.fuse_collapse(grids, method, fused_out=out)
    assert out.download().tobytes() == fused_o.tobytes()
    assert np.array_equal(conf, conf_o) and np.array_equal(idx, idx_o)


@pytest.mark.parametrize("method,n", [(3, 4), (4, 4), (5, 4), (3, 3), (2, 4), (4, 5)])
def test_nary_extension(ctx, O, method, n):
    rng = np.random.default_rng(100 + method + n)
    vols = [np.where(rng.random((4, 9, 11)) < 0.3, 0, rng.gamma(1, 4, (4, 9, 11))).astype(np.float32) for _ in range(n)]
    grids = []
    for v in vols:
        g = api.Grid3D(ctx, 11, 9, 4)
        g.upload(v)
        grids.append(g)
    want = O.fuse_nary(method, vols)
    out = api.Grid3D(ctx, 11, 9, 4)
    api.fuse_collapse(grids, method, fused_out=out)
    got = out.download()
    if method == 3 and n not in (2, 4):
        np.testing.assert_allclose(got, want, rtol=1e-6)  # pow() differs by an ulp between libm and CUDA
    else:
        assert got.tobytes() == want.tobytes()
    # n == 2 of the extension is the reference op
    assert O.fuse_nary(method, vols[:2]).tobytes() == O.fuse_reference(method, vols[:2]).tobytes()


def test_temporal_fusion_alg2(ctx, O, small_case):
    """process_2 temporal fusion (process2.cpp:211-242): HM = n / sum 1/(0.01+x); AM = sum/n."""
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    pk = small_case.packets[0]
    n_sub = 4
    per = len(pk) // n_sub
    acc_hm = api.Grid3D(ctx, small_case.dimX, small_case.dimY, small_case.dimZ)
    acc_am = api.Grid3D(ctx, small_case.dimX, small_case.dimY, small_case.dimZ)
    want_hm = np.zeros((small_case.dimZ, small_case.dimY, small_case.dimX), np.float32)
    want_am = want_hm.copy()
    c = small_case.cams[0]
    for k in range(n_sub):
        sub = pk[k * per:(k + 1) * per]
        m.build(small_case.events[0], sub)
        acc_hm.addInverseOfTwoGrids(m.dsi_)
        acc_am.addTwoGrids(m.dsi_)
        dsi_o, _ = O.build_dsi(small_case.events[0], sub, c.lut, c.width, small_case.depths, small_case.virts[0],
                               small_case.dimX, small_case.dimY)
        O.fuse_op(O.OP_ADD_INV, want_hm, dsi_o, eps=1e-2)
        O.fuse_op(O.OP_ADD, want_am, dsi_o)
    acc_hm.computeHMfromSumOfInv(n_sub)
    acc_am.computeAMfromSum(n_sub)
    O.fuse_op(O.OP_HM_FROM_SUMINV, want_hm, None, n=n_sub)
    O.fuse_op(O.OP_AM_FROM_SUM, want_am, None, n=n_sub)
    np.testing.assert_allclose(acc_am.download(), want_am, rtol=MAP_RTOL, atol=1e-6)
    np.testing.assert_allclose(acc_hm.download(), want_hm, rtol=MAP_RTOL, atol=1e-6)
    m.close()


# ------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------
def test_zero_packets_resets(ctx, small_case):
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    m.build(small_case.events[0], small_case.packets[0])
    assert m.dsi_.download().any()
    m.build(small_case.events[0], small_case.packets[0][:0])
    assert not m.dsi_.download().any() and not m.counts().any()
    m.close()


def test_out_of_view_and_degenerate_coordinates(ctx, O):
    """Votes landing outside [0, dim-1) on either axis, NaN/Inf coordinates (d == 0) are rejected."""
    cam = api.CameraModel(64, 48, 50., 50., 32., 24.)
    shape = api.ShapeDSI(0, 0, 5, 1.0, 2.0, 0.0)
    rng = np.random.default_rng(3)
    ev = np.zeros(4096, capi.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 64, 4096)
    ev["y"] = rng.integers(0, 48, 4096)
    pk = np.zeros(4, capi.PACKET_DTYPE)
    pk["first_event"] = np.arange(4) * 1024
    H = np.eye(3, dtype=np.float32)
    pk["H"][0] = H.ravel()
    pk["C"][0] = (0.3, -0.2, 0.1)               # strong parallax: many votes leave the grid
    H2 = H.copy(); H2[0, 2] = -40; H2[1, 2] = 30
    pk["H"][1] = H2.ravel()                     # shifted: negative x, y beyond dimY
    pk["C"][1] = (0.0, 0.0, 1.0)                # C_z == z0 -> d = z_k * (z0 - C_z) == 0 -> Inf / NaN
    pk["H"][2] = (H * np.float32(1e30)).ravel() # huge but finite ratios
    pk["C"][2] = (1e20, 0, 0)
    H3 = H.copy(); H3[2, 2] = 0.0
    pk["H"][3] = H3.ravel()                     # p2 == 0 -> division by zero in the event stage
    pk["C"][3] = (0, 0, 0)
    m = api.MapperEMVS(ctx, cam, shape)
    m.build(ev, pk)
    depths = O.depth_vector(1.0, 2.0, 5)
    virt = O.virtual_camera(50., 32., 24., 64, 0.0)
    with np.errstate(all="ignore"):
        dsi_o, inb_o = O.build_dsi(ev, pk, cam.lut, 64, depths, virt, 64, 48)
    assert np.array_equal(m.counts(), inb_o)
    got = m.dsi_.download()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
    m.close()


@pytest.mark.parametrize("dims", [(31, 17, 3), (2, 2, 1), (1, 9, 2), (65, 33, 9)])
def test_odd_and_tiny_dsi_sizes(ctx, O, dims):
    dimX, dimY, dimZ = dims
    cam = api.CameraModel(40, 30, 30., 30., 20., 15.)
    shape = api.ShapeDSI(dimX, dimY, dimZ, 1.0, 3.0, 0.0)
    rng = np.random.default_rng(dimX * 7 + dimY)
    ev = np.zeros(3000, capi.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 40, 3000)
    ev["y"] = rng.integers(0, 30, 3000)
    pk = np.zeros(2, capi.PACKET_DTYPE)
    pk["first_event"] = (0, 1500)
    for j in range(2):
        pk["H"][j] = np.array([[0.9, 0.01, -1.0 + j], [0.0, 0.95, -2.0], [0, 0, 1]], np.float32).ravel()
        pk["C"][j] = (0.05 * j, 0.02, -0.01)
    m = api.MapperEMVS(ctx, cam, shape)
    m.build(ev, pk)
    depths = O.depth_vector(1.0, 3.0, dimZ)
    virt = O.virtual_camera(30., 20., 15., dimX, 0.0)
    dsi_o, inb_o = O.build_dsi(ev, pk, cam.lut, 40, depths, virt, dimX, dimY)
    assert np.array_equal(m.counts(), inb_o)
    np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
    conf, idx = m.dsi_.collapseMaxZSlice()
    conf_o, idx_o = O.collapse_max(m.dsi_.download())
    assert np.array_equal(conf, conf_o) and np.array_equal(idx, idx_o)
    m.close()


def test_inverse_depth_and_fov(ctx, O):
    """USE_INVERSE_DEPTH variant (plane 0 = farthest) and a DSI with its own field of view."""
    case = Case("esim_small", events_per_cam=20_000, seed=9)
    shape = api.ShapeDSI(120, 90, 32, 1.0, 5.0, 60.0, inverse_depth=True)
    cam = case.cams[0]
    m = api.MapperEMVS(ctx, cam, shape)
    depths = O.depth_vector(1.0, 5.0, 32, inverse=True)
    virt = O.virtual_camera(cam.fx, cam.cx, cam.cy, 120, 60.0)
    assert m.raw_depths_vec_.tobytes() == depths.tobytes() and m.virtual_cam_.tobytes() == virt.tobytes()
    assert depths[0] == 5.0 and depths[0] > depths[-1]
    traj = api.LinearTrajectory(case.trajs[0])
    pk = m.packetize(case.events[0], traj, case.T_rv_w)
    pk_o = O.packetize(case.events[0], case.trajs[0], case.T_rv_w, np.array([cam.fx, cam.fy, cam.cx, cam.cy], np.float32),
                       virt, depths[0])
    assert pk.tobytes() == pk_o.tobytes()
    m.build(case.events[0], pk)
    dsi_o, inb_o = O.build_dsi(case.events[0], pk_o, cam.lut, cam.width, depths, virt, 120, 90)
    assert np.array_equal(m.counts(), inb_o)
    np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
    m.close()


def test_u16_index_extension(ctx, O):
    """dimZ > 256 (beyond the reference's uchar index, main.cpp:156): uint16 indices."""
    rng = np.random.default_rng(8)
    dsi = rng.random((300, 6, 7)).astype(np.float32)
    g = api.Grid3D(ctx, 7, 6, 300)
    g.upload(dsi)
    conf, idx = g.collapseMaxZSlice()
    conf_o, idx_o = O.collapse_max(dsi)
    assert idx.dtype == np.uint16 and np.array_equal(idx, idx_o) and np.array_equal(conf, conf_o)
    assert idx.max() > 255


def test_errors_not_aborts(ctx, small_case):
    with pytest.raises(api.EmvsError):
        api.MapperEMVS(ctx, small_case.cams[0], api.ShapeDSI(0, 0, 10, 2.0, 1.0, 0.0))  # max < min (CHECK_GT)
    with pytest.raises(api.EmvsError):
        api.fuse_collapse([api.Grid3D(ctx, 4, 4, 4), api.Grid3D(ctx, 4, 4, 4)], 9)        # Improper fusion method
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    bad = small_case.packets[0][:2].copy()
    bad["first_event"][1] = len(small_case.events[0])  # packet reaches past the event list
    with pytest.raises(api.EmvsError):
        m.build(small_case.events[0], bad)
    m.close()


# ------------------------------------------------------------------------------------------
# BASELINE sizes: properties that do not need the oracle to finish
# ------------------------------------------------------------------------------------------
def test_full_size_properties(ctx, O):
    """640x480x256, DSEC-like rig with a real rectification LUT, 1M events: counts bit-exact vs the
    oracle on a plane subset, plane sums == counts, sharded == unsharded, linearity."""
    case = Case("dsec_stereo", events_per_cam=1_000_000, n_cams=1)
    m = api.MapperEMVS(ctx, case.cams[0], case.shape)
    m.build(case.events[0], case.packets[0])
    counts = m.counts()
    dsi = m.dsi_.download()
    sums = dsi.reshape(256, -1).sum(1, dtype=np.float64)
    np.testing.assert_allclose(sums, counts.astype(np.float64), rtol=2e-5)
    # oracle on 6 planes spread over the volume (per-plane work is independent)
    c = case.cams[0]
    for k in (0, 1, 77, 128, 200, 255):
        sel = np.array([case.depths[0], case.depths[k]], np.float32)  # depths[0] defines z0
        dsi_o, inb_o = O.build_dsi(case.events[0], case.packets[0], c.lut, c.width, sel, case.virts[0], 640, 480)
        assert inb_o[1] == counts[k]
        np.testing.assert_allclose(dsi[k], dsi_o[1], rtol=DSI_RTOL, atol=DSI_ATOL)
    # linearity / sharding: two halves accumulate to the whole
    pk = case.packets[0]
    m.build(case.events[0], pk[:len(pk) // 2])
    m.build(case.events[0], pk[len(pk) // 2:], accumulate=True)
    assert np.array_equal(m.counts(), counts)
    np.testing.assert_allclose(m.dsi_.download(), dsi, rtol=DSI_RTOL, atol=DSI_ATOL)
    # argmax is idempotent under collapse of a one-hot volume built from it
    conf, idx = m.dsi_.collapseMaxZSlice()
    assert conf.max() > 1.0 and idx.dtype == np.uint8
    m.close()


def _oracle_planes(O, case, cam_idx, planes):
    """Oracle DSI of selected planes only (per-plane work is independent; depths[0] defines z0)."""
    c = case.cams[cam_idx]
    out = {}
    for k in planes:
        sel = np.array([case.depths[0], case.depths[k]], np.float32)
        dsi_o, inb_o = O.build_dsi(case.events[cam_idx], case.packets[cam_idx], c.lut, c.width, sel, case.virts[cam_idx],
                                   case.dimX, case.dimY)
        out[k] = (dsi_o[1], int(inb_o[1]))
    return out


def test_config3_four_camera_geometric_mean(ctx, O):
    """BASELINE.json configs[2] at reduced event count: 4-camera bar, 640x480x256, n-ary GM (extension
    whose n = 2 case is the reference's sqrt(a*b)); oracle on a plane subset."""
    case = Case("bar4", events_per_cam=300_000)
    assert case.n_cams == 4 and case.method == 3
    mappers = make_mappers(ctx, case)
    planes = (0, 40, 128, 255)
    per_cam = []
    for i, m in enumerate(mappers):
        m.build(case.events[i], case.packets[i])
        counts = m.counts()
        ref = _oracle_planes(O, case, i, planes)
        per_cam.append(ref)
        dsi = m.dsi_.download()
        for k in planes:
            assert counts[k] == ref[k][1]
            np.testing.assert_allclose(dsi[k], ref[k][0], rtol=DSI_RTOL, atol=DSI_ATOL)
        del dsi
    fused = api.Grid3D(ctx, case.dimX, case.dimY, case.dimZ)
    conf, idx, depth = api.fuse_collapse([m.dsi_ for m in mappers], 3, case.depths, fused_out=fused)
    fz = fused.download()
    for k in planes:
        want = O.fuse_nary(3, [per_cam[i][k][0][None] for i in range(4)])[0]
        np.testing.assert_allclose(fz[k], want, rtol=MAP_RTOL, atol=1e-5)
    # the maps are consistent with the fused volume the GPU produced
    assert np.array_equal(conf, fz.max(axis=0)) and np.array_equal(idx, fz.argmax(axis=0).astype(np.uint8))
    assert np.array_equal(depth, case.depths[idx])
    # n = 2 of the extension is the reference op
    two = api.fuse_collapse([mappers[0].dsi_, mappers[1].dsi_], 3, fused_out=fused)
    a, b = mappers[0].dsi_.download(), mappers[1].dsi_.download()
    assert fused.download().tobytes() == O.fuse_reference(3, [a, b]).tobytes()
    for m in mappers:
        m.close()
    fused.close()


@pytest.mark.parametrize("dims,n_ev", [((256, 256, 128), 200_000), ((1024, 1024, 512), 150_000)])
def test_config5_sweep_corners(ctx, O, dims, n_ev):
    """Smallest and largest DSI of BASELINE.json configs[4] (1024^2 x 512 needs uint16 indices: extension)."""
    from dvs_mcemvs_b200 import synth
    W, H, Nz = dims
    cam = api.CameraModel(W, H, 0.8 * W, 0.8 * W, W / 2.0, H / 2.0)
    shape = api.ShapeDSI(0, 0, Nz, 1.0, 10.0, 0.0)
    sc = synth.Scene(synth.Rig([cam], [0.0]), shape, duration=0.2, translation=(0.2, 0.0, 0.0), rot_deg=1.0, seed=5)
    ev, traj = sc.events(0, n_ev), sc.trajectory(0)
    m = api.MapperEMVS(ctx, cam, shape)
    assert m.evaluateDSI(ev, api.LinearTrajectory(traj), sc.T_rv_w())
    depths, virt = m.raw_depths_vec_, m.virtual_cam_
    assert depths.tobytes() == O.depth_vector(1.0, 10.0, Nz).tobytes()
    pk = O.packetize(ev, traj, sc.T_rv_w(), np.array([cam.fx, cam.fy, cam.cx, cam.cy], np.float32), virt, depths[0])
    counts = m.counts()
    dsi = m.dsi_.download()
    np.testing.assert_allclose(dsi.reshape(Nz, -1).sum(1, dtype=np.float64), counts.astype(np.float64), rtol=2e-5)
    for k in (0, Nz // 3, Nz - 1):
        sel = np.array([depths[0], depths[k]], np.float32)
        dsi_o, inb_o = O.build_dsi(ev, pk, cam.lut, W, sel, virt, W, H)
        assert inb_o[1] == counts[k]
        np.testing.assert_allclose(dsi[k], dsi_o[1], rtol=DSI_RTOL, atol=DSI_ATOL)
    conf, idx, depth = m.dsi_.collapseMaxZSlice(depths)
    assert idx.dtype == (np.uint8 if Nz <= 256 else np.uint16)
    assert np.array_equal(conf, dsi.max(axis=0)) and np.array_equal(idx, dsi.argmax(axis=0).astype(idx.dtype))
    assert np.array_equal(depth, depths[idx])
    m.close()


@pytest.mark.parametrize("env,percent,pieces,deferred", [
    ({"EMVS_UPLOAD_PIECES": "3"}, 10, 3, False),                                                  # one slab: never deferred
    ({"EMVS_UPLOAD_PIECES": "4", "EMVS_UPLOAD_DEFER_MERGE": "0", "EMVS_SLAB": "24"}, 6, 4, False),
    ({"EMVS_SLAB": "24"}, 15, 4, True),                                                           # the defaults: 4 pieces from 6 %
    ({"EMVS_UPLOAD_DEFER_PIECES": "2", "EMVS_UPLOAD_DEFER_SPLIT": "15", "EMVS_SLAB": "24"}, 15, 2, True),
    ({"EMVS_UPLOAD_DEFER_PIECES": "3", "EMVS_UPLOAD_DEFER_SPLIT": "8", "EMVS_SLAB": "24"}, 15, 3, True),
    ({}, 15, 2, False),                                                                           # one slab, defaults: head + tail
], ids=["three", "four", "four_deferred", "two_deferred", "three_deferred", "default_single_slab"])
def test_evaluate_dsi_multi_piece_split(small_case, env, percent, pieces, deferred):
    """Split upload in more than two pieces (p %, 2.2 p %, 4.84 p %, rest).  As complete builds (EMVS_UPLOAD_PIECES, default 2)
    and in the deferred form that builds taking the multi-slab vote launch use by default (EMVS_UPLOAD_DEFER_MERGE /
    _DEFER_PIECES / _DEFER_SPLIT: the pieces before the last one only vote, into the per-slab scratch; the last one votes on
    top and merges once): the pieces add up to the one-piece DSI — counts bit-exact — and the number of kernel launches
    says which path ran.  The knobs are read when a context is created, hence a fresh process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from conftest import Case\n"
        "from dvs_mcemvs_b200 import api\n"
        "case = Case('esim_small')\n"
        "ctx = api.Context(0)\n"
        f"ctx.set_upload_split({percent}, 4096)\n"
        "m = api.MapperEMVS(ctx, case.cams[0], case.shape)\n"
        "tr = api.LinearTrajectory(case.trajs[0])\n"
        "dsi_o, inb_o = case.oracle_dsi(0)\n"
        "for rep in range(2):\n"                                      # twice: the scratch is clean again after a deferred sequence
        "    ctx.sync(); n0 = ctx.launch_count()\n"
        "    assert m.evaluateDSI(case.events[0], tr, case.T_rv_w)\n"
        "    n_split = ctx.launch_count() - n0\n"
        "    assert np.array_equal(m.counts(), inb_o)\n"
        "    np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=1e-5, atol=1e-5)\n"
        "ctx.set_upload_split(0)\n"
        "ctx.sync(); n0 = ctx.launch_count()\n"
        "assert m.evaluateDSI(case.events[0], tr, case.T_rv_w)\n"
        "n_one = ctx.launch_count() - n0\n"
        "assert np.array_equal(m.counts(), inb_o)\n"
        f"want = n_one + 2 * ({pieces} - 1) if {deferred} else {pieces} * n_one\n"   # a vote-only piece: event stage + one vote launch
        "assert n_split == want, (n_split, n_one)\n"
        # an accumulating evaluate on top of a split one (sub-interval style): counts and volume double
        f"ctx.set_upload_split({percent}, 4096)\n"
        "ctx.sync()\n"
        "assert m.evaluateDSI(case.events[0], tr, case.T_rv_w, accumulate=True)\n"
        "assert np.array_equal(m.counts(), 2 * inb_o)\n"
        "np.testing.assert_allclose(m.dsi_.download(), 2 * dsi_o, rtol=1e-5, atol=2e-5)\n"
        "print('pieces ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0 and "pieces ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
