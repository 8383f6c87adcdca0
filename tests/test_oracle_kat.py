"""Known-answer tests that pin the CPU oracle semantically — closed forms that
must hold whatever the floating-point details are — plus agreement of its two independent
statements (C++ loops vs numpy twin) and the host-side geometry of the product (C-ABI, no GPU)
against the oracle.  BASELINE.json configs[0] (stereo 50k events/cam, 240x180x64, harmonic fusion on
CPU) is exactly the `small_case` fixture used here."""
import ctypes as C

import numpy as np
import pytest

from dvs_mcemvs_b200 import _capi as capi
from dvs_mcemvs_b200 import api, synth

from conftest import Case


# ---- vote -------------------------------------------------------------------------------------
def test_vote_bounds_rule_and_weights(O):
    g = np.zeros((6, 8), np.float32)   # dimY=6, dimX=8
    p = g.ctypes.data_as(C.c_void_p)
    L = O.lib()
    assert L.oracle_vote(C.c_float(2.25), C.c_float(3.5), p, 8, 6) == 1
    np.testing.assert_array_equal(g[3:5, 2:4], np.array([[0.375, 0.125], [0.375, 0.125]], np.float32))
    assert g.sum() == 1.0
    for x, y in [(-0.01, 1), (1, -0.01), (7.0, 1), (6.999, 1), (1, 5.0), (np.nan, 1), (1, np.nan), (np.inf, 1), (3e9, 1)]:
        before = g.copy()
        ok = L.oracle_vote(C.c_float(x), C.c_float(y), p, 8, 6)
        assert ok == (1 if (0 <= x < 7 and 0 <= y < 5) else 0)      # accept iff 0 <= x < dimX-1 and 0 <= y < dimY-1
        if not ok:
            assert np.array_equal(before, g)
    g[:] = 0
    assert L.oracle_vote(C.c_float(0.0), C.c_float(0.0), p, 8, 6) == 1 and g[0, 0] == 1.0 and g.sum() == 1.0


# ---- config 0: the plumbing case ----------------------------------------------------------------
def test_config0_plane_sums_equal_accepted_votes_and_twin_agrees(O, small_case):
    for i in range(small_case.n_cams):
        dsi, inb = small_case.oracle_dsi(i)
        assert len(small_case.packets[i]) == 48                       # Np = ceil(50000/1024) - 1 (strict '<')
        sums = dsi.reshape(dsi.shape[0], -1).sum(1, dtype=np.float64)
        np.testing.assert_allclose(sums, inb.astype(np.float64), rtol=1e-5)   # bilinear weights sum to 1
        assert inb.max() <= 48 * 1024
        c = small_case.cams[i]
        xy0 = O.warp_events(small_case.events[i], small_case.packets[i], c.lut, c.width)
        dsi_np, inb_np = O.np_build_dsi(xy0, small_case.packets[i], small_case.depths, small_case.virts[i],
                                        small_case.dimX, small_case.dimY)
        assert np.array_equal(inb_np, inb)
        assert dsi_np.tobytes() == dsi.tobytes()                      # same float ops in the same order


def test_config0_harmonic_fusion_depth_is_meaningful(O, small_case):
    vols = [small_case.oracle_dsi(i)[0] for i in range(2)]
    fused = O.fuse_reference(2, vols)
    conf, idx, depth = O.collapse_max(fused, small_case.depths)
    # structured scene: confident pixels exist and their depth lies inside the DSI range
    strong = conf > 0.5 * conf.max()
    assert strong.sum() > 20
    assert (depth[strong] >= 1.0).all() and (depth[strong] < 5.0).all()
    assert np.array_equal(depth, small_case.depths[idx])
    assert O.mean_square(fused) == pytest.approx(float((fused.astype(np.float64) ** 2).mean()), rel=1e-12)


def test_z0_identity_plane(O, small_case):
    """Plane 0 has z_k == z0: bx = by = 0 and a == d, so the vote lands at (X0*a)/a ~= X0."""
    c = small_case.cams[0]
    pk = small_case.packets[0]
    xy0 = O.warp_events(small_case.events[0], pk, c.lut, c.width)
    dsi, inb = O.build_dsi(small_case.events[0], pk, c.lut, c.width, small_case.depths[:1], small_case.virts[0],
                           small_case.dimX, small_case.dimY)
    ok = (xy0[:, 0] >= 0) & (xy0[:, 1] >= 0) & (xy0[:, 0] < small_case.dimX - 1) & (xy0[:, 1] < small_case.dimY - 1)
    assert abs(int(inb[0]) - int(ok.sum())) <= 8      # only 1-ulp boundary flips of (X0*a)/a may differ


def test_fronto_parallel_plane_peaks_at_its_depth(O):
    """Points on the plane Z* = 2.5 m seen from a translating camera focus on the DSI plane containing Z*."""
    W, H, f, cx, cy, Z = 240, 180, 200.0, 120.0, 90.0, 2.5
    rng = np.random.default_rng(4)
    px, py = rng.uniform(-1.2, 1.6, 1500), rng.uniform(-0.9, 0.9, 1500)
    n = 49153                                                       # -> exactly 48 packets
    t = 1000.0 + 0.18 * np.arange(n) / n
    k = rng.integers(0, 1500, n)
    u, v = f * (px[k] - (t - 1000.0)) / Z + cx, f * py[k] / Z + cy
    ok = (np.rint(u) >= 0) & (np.rint(u) < W) & (np.rint(v) >= 0) & (np.rint(v) < H)
    ev = np.zeros(int(ok.sum()), O.EVENT_DTYPE)
    ev["x"], ev["y"] = np.rint(u[ok]), np.rint(v[ok])
    ev["sec"], ev["nsec"] = synth._split_time(t[ok])
    traj = np.zeros(11, capi.STAMPED_POSE_DTYPE)
    ts = 1000.0 + 0.02 * np.arange(11) - 0.01
    traj["sec"], traj["nsec"] = synth._split_time(ts)
    traj["T"]["q"][:, 0] = 1.0
    traj["T"]["t"][:, 0] = ts - 1000.0
    T_rv_w = O.pose_inv(O.pose_at(traj, 1000, 90_000_000))
    depths = O.depth_vector(1.0, 5.0, 64)
    virt = O.virtual_camera(f, cx, cy, W, 0.0)
    cam = api.CameraModel(W, H, f, f, cx, cy)
    pk = O.packetize(ev, traj, T_rv_w, np.array([f, f, cx, cy], np.float32), virt, depths[0])
    dsi, inb = O.build_dsi(ev, pk, cam.lut, W, depths, virt, W, H)
    conf, idx, depth = O.collapse_max(dsi, depths)
    k_star = int(round((Z - 1.0) * 64 / 4.0))                       # plane 24
    ci, ii = conf[10:-10, 10:-10].ravel(), idx[10:-10, 10:-10].ravel().astype(int)
    top = np.argsort(ci)[-300:]                                     # the 300 most confident pixels
    assert ci[top].min() > 5.0
    assert (np.abs(ii[top] - k_star) <= 2).mean() > 0.8             # events are rounded to pixels: +-2 cells
    assert np.median(ii[top]) == k_star


# ---- fusion closed forms ---------------------------------------------------------------------------
def test_fusion_closed_forms(O):
    a = np.array([[[0.0, 2.0, 3.0, 7.5, 4.0]]], np.float32)
    b = np.array([[[5.0, 2.0, 0.0, 1.5, 9.0]]], np.float32)
    f32 = np.float32
    hm = O.fuse_op(O.OP_HM, a.copy(), b, n=2, eps=0.1)
    assert hm[0, 0, 0] == 0 and hm[0, 0, 2] == 0                            # outlier rejection: HM(x, 0) = 0
    assert hm[0, 0, 1] == f32(2) * (f32(2) * f32(2)) / (f32(4) + f32(0.1))    # HM(a, a) = 2a^2 / (2a + 0.1)
    gm = O.fuse_op(O.OP_GM, a.copy(), b)
    assert gm[0, 0, 0] == 0 and gm[0, 0, 1] == 2 and gm[0, 0, 4] == 6
    assert np.array_equal(O.fuse_op(O.OP_AM, a.copy(), b), (a + b) * f32(0.5))
    assert np.array_equal(O.fuse_op(O.OP_MIN, a.copy(), b), np.minimum(a, b))
    assert np.array_equal(O.fuse_op(O.OP_MAX, a.copy(), b), np.maximum(a, b))
    rms = O.fuse_op(O.OP_RMS, a.copy(), b)
    want = np.sqrt((0.5 * (a.astype(np.float64) ** 2 + b.astype(np.float64) ** 2)).astype(np.float32))
    assert np.array_equal(rms, want)
    # HM-n recursion (cartesian3dgrid.h:130-139) equals 3abc/(ab+bc+ca) when eps -> 0
    x, y, z = (np.full((1, 1, 1), v, np.float32) for v in (2.0, 3.0, 6.0))
    h2 = O.fuse_op(O.OP_HM, x.copy(), y, n=2, eps=0.0)
    h3 = O.fuse_op(O.OP_HM_N, h2, z, n=3, eps=0.0)
    assert h3[0, 0, 0] == pytest.approx(3 * 2 * 3 * 6 / (2 * 3 + 3 * 6 + 6 * 2), rel=1e-6)
    # temporal: AM = sum / n, HM = n / sum(1 / (0.01 + x))  (process2.cpp:211-242)
    vols = [np.full((1, 1, 1), v, np.float32) for v in (1.0, 2.0, 4.0)]
    acc_am, acc_hm = np.zeros((1, 1, 1), np.float32), np.zeros((1, 1, 1), np.float32)
    for v in vols:
        O.fuse_op(O.OP_ADD, acc_am, v)
        O.fuse_op(O.OP_ADD_INV, acc_hm, v, eps=1e-2)
    assert O.fuse_op(O.OP_AM_FROM_SUM, acc_am, None, n=3)[0, 0, 0] == pytest.approx(7 / 3, rel=1e-6)
    assert O.fuse_op(O.OP_HM_FROM_SUMINV, acc_hm, None, n=3)[0, 0, 0] == pytest.approx(
        3 / (1 / 1.01 + 1 / 2.01 + 1 / 4.01), rel=1e-6)
    # third camera: ignored for GM / AM / RMS (process1.cpp:178-183), used for min / HM / max
    rng = np.random.default_rng(0)
    v3 = [rng.gamma(1, 3, (2, 3, 4)).astype(np.float32) for _ in range(3)]
    for m in (3, 4, 5):
        assert np.array_equal(O.fuse_reference(m, v3), O.fuse_reference(m, v3[:2]))
    for m in (1, 2, 6):
        assert not np.array_equal(O.fuse_reference(m, v3), O.fuse_reference(m, v3[:2]))
    # by-value argument copy changes nothing but the time
    assert np.array_equal(O.fuse_reference(2, v3, copy_arg=True), O.fuse_reference(2, v3))


def test_collapse_first_maximum_and_zero_column(O):
    dsi = np.zeros((5, 2, 2), np.float32)
    dsi[1, 0, 0] = dsi[3, 0, 0] = 2.0
    conf, idx = O.collapse_max(dsi)
    assert idx[0, 0] == 1 and conf[0, 0] == 2.0
    assert idx[1, 1] == 0 and conf[1, 1] == 0.0


# ---- packet stage / trajectory (restated minkindr + Eigen arithmetic; parity UNPINNED, semantics pinned) ----
def _traj(ts, positions, quats=None):
    tr = np.zeros(len(ts), capi.STAMPED_POSE_DTYPE)
    tr["sec"], tr["nsec"] = synth._split_time(np.asarray(ts, np.float64))
    tr["T"]["q"] = quats if quats is not None else [[1, 0, 0, 0]] * len(ts)
    tr["T"]["t"] = positions
    return tr


def test_pose_interpolation_semantics(O):
    half = np.deg2rad(40.0) / 2
    tr = _traj([10.0, 11.0], [[0, 0, 0], [2, 4, -6]], [[1, 0, 0, 0], [np.cos(half), 0, 0, np.sin(half)]])
    p = O.pose_at(tr, 10, 500_000_000)
    np.testing.assert_allclose(p["t"], [1, 2, -3], atol=1e-12)                 # translation lerp in the T0 frame
    np.testing.assert_allclose(p["q"], [np.cos(half / 2), 0, 0, np.sin(half / 2)], atol=1e-12)   # slerp: half the angle
    assert O.pose_at(tr, 10, 0)["t"].tolist() == [0, 0, 0]                      # t == first control pose is inside
    assert O.pose_at(tr, 9, 999_999_999) is None                               # no extrapolation in the past
    assert O.pose_at(tr, 11, 0) is None                                        # upper_bound == end: "future"
    a = O.pose_at(tr, 10, 250_000_000)
    np.testing.assert_allclose(O.pose_mul(O.pose_inv(a), a)["q"], [1, 0, 0, 0], atol=1e-15)
    np.testing.assert_allclose(O.pose_mul(O.pose_inv(a), a)["t"], [0, 0, 0], atol=1e-15)


def test_packetizer_drop_semantics(O):
    """Strict '<' loop bound, tail remainder dropped, skip-one-event on a pose miss (mapper_emvs_stereo.cpp:88-99)."""
    K = np.array([200, 200, 120, 90], np.float32)
    virt = K.copy()
    tr = _traj([100.0, 101.0], [[0, 0, 0], [0.1, 0, 0]])
    I = np.zeros((), capi.POSE_DTYPE); I["q"] = (1, 0, 0, 0)

    def events(ts):
        ev = np.zeros(len(ts), O.EVENT_DTYPE)
        ev["sec"], ev["nsec"] = synth._split_time(np.asarray(ts, np.float64))
        return ev
    assert len(O.packetize(events(np.linspace(100.1, 100.9, 1023)), tr, I, K, virt, 1.0)) == 0      # < 1024 events
    assert len(O.packetize(events(np.linspace(100.1, 100.9, 1024)), tr, I, K, virt, 1.0)) == 0      # 0 + 1024 < 1024 false
    assert len(O.packetize(events(np.linspace(100.1, 100.9, 1025)), tr, I, K, virt, 1.0)) == 1
    assert len(O.packetize(events(np.linspace(100.1, 100.9, 2048)), tr, I, K, virt, 1.0)) == 1      # tail of 1024 dropped
    pk = O.packetize(events(np.linspace(100.1, 100.9, 2049)), tr, I, K, virt, 1.0)
    assert pk["first_event"].tolist() == [0, 1024]
    # the first 100 events are before the trajectory: mid-event of the packet starting at i is event i+512, so the
    # packetiser slides one event at a time until event i+512 has a pose, i.e. first_event = 100 - ... >= 0
    ts = np.concatenate([np.linspace(99.0, 99.9, 600), np.linspace(100.1, 100.9, 3000)])
    pk = O.packetize(events(ts), tr, I, K, virt, 1.0)
    assert pk["first_event"][0] == 600 - 512 and np.all(np.diff(pk["first_event"].astype(np.int64)) == 1024)
    # identity pose, K == K_virtual: H is the identity homography (up to scale) and C = -R^T t
    pk = O.packetize(events(np.full(1025, 100.5)), tr, I, K, virt, 1.0)
    H = pk["H"][0].reshape(3, 3)
    np.testing.assert_allclose(H / H[2, 2], np.array([[1, 0, 0.05 * 200], [0, 1, 0], [0, 0, 1]]), atol=2e-4)
    np.testing.assert_allclose(pk["C"][0], [0.05, 0, 0], atol=1e-6)


def test_product_host_geometry_matches_oracle_bit_exact(O, small_case):
    """The product's own host stage (C-ABI: emvs_packetize / emvs_trajectory_pose_at / emvs_virtual_camera /
    emvs_depth_vector; no GPU needed) against the oracle: identical bits."""
    lib = capi.load()
    for i, c in enumerate(small_case.cams):
        cs = c.c_struct()
        sh = small_case.shape.c_struct()
        virt = np.zeros(4, np.float32)
        capi.check(lib.emvs_virtual_camera(C.byref(cs), C.byref(sh), capi.ptr(virt)))
        assert virt.tobytes() == small_case.virts[i].tobytes()
        ev, tr = small_case.events[i], small_case.trajs[i]
        out = np.zeros(len(ev) // 1024 + 1, capi.PACKET_DTYPE)
        n = C.c_size_t(0)
        capi.check(lib.emvs_packetize(capi.ptr(ev), len(ev), capi.ptr(tr), len(tr),
                                      capi.ptr(np.ascontiguousarray(small_case.T_rv_w)), C.byref(cs), capi.ptr(virt),
                                      float(small_case.depths[0]), capi.ptr(out), len(out), C.byref(n)))
        assert out[:n.value].tobytes() == small_case.packets[i].tobytes()
        for sec, nsec in ((1000, 0), (1000, 123_456_789), (999, 0), (1000, 199_999_999)):
            got = api.LinearTrajectory(tr).getPoseAt(sec, nsec)
            want = O.pose_at(tr, sec, nsec)
            assert (got is None) == (want is None)
            if got is not None:
                assert got.tobytes() == want.tobytes()
    # fov-based virtual focal (mapper_emvs_stereo.cpp:225-229)
    sh = api.ShapeDSI(120, 90, 8, 1.0, 5.0, 60.0).c_struct()
    cs = small_case.cams[0].c_struct()
    virt = np.zeros(4, np.float32)
    capi.check(lib.emvs_virtual_camera(C.byref(cs), C.byref(sh), capi.ptr(virt)))
    assert virt.tobytes() == O.virtual_camera(200.0, 120.0, 90.0, 120, 60.0).tobytes()
    assert virt[0] == pytest.approx(0.5 * 120 / np.tan(np.deg2rad(30.0)), rel=1e-6)


def test_packetize_range_equals_one_shot(O):
    """emvs_packetize_range with growing event limits (the split upload of evaluateDSI) yields exactly the packets of
    one pass over the whole list — also when a run of pose misses straddles a limit (no GPU needed)."""
    lib = capi.load()
    K = np.array([200, 200, 120, 90], np.float32)
    tr = _traj([100.0, 101.0], [[0, 0, 0], [0.1, 0, 0]])
    I = np.zeros((), capi.POSE_DTYPE); I["q"] = (1, 0, 0, 0)
    cam = capi.Camera(240, 180, 200, 200, 120, 90)
    # events before the trajectory, inside, a gap after its end, so that misses occur at the head and at the tail
    ts = np.concatenate([np.linspace(99.5, 99.99, 700), np.linspace(100.0, 100.99, 9000), np.linspace(101.01, 101.5, 2500)])
    ev = np.zeros(len(ts), O.EVENT_DTYPE)
    ev["sec"], ev["nsec"] = synth._split_time(ts)
    want = O.packetize(ev, tr, I, K, K.copy(), 1.0)
    assert len(want) >= 8
    for limits in ([len(ev)], [1024, len(ev)], [3000, 3001, 7777, len(ev)], [0, 500, 1023, 2048, 9700 + 512, len(ev)],
                   list(range(0, len(ev), 1000)) + [len(ev)]):
        cur = C.c_size_t(0)
        got = []
        for lim in limits:
            out = np.zeros(len(ev) // 1024 + 1, capi.PACKET_DTYPE)
            n = C.c_size_t(0)
            capi.check(lib.emvs_packetize_range(capi.ptr(ev), len(ev), capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam),
                                                capi.ptr(K), 1.0, C.byref(cur), lim, capi.ptr(out), len(out), C.byref(n)))
            assert np.all(out["first_event"][:n.value] + 1024 <= lim)
            got.append(out[:n.value].copy())
        got = np.concatenate(got)
        assert got.tobytes() == want.tobytes(), limits
    cur = C.c_size_t(len(ev) + 1)   # cursor past the list is rejected
    n = C.c_size_t(0)
    out = np.zeros(4, capi.PACKET_DTYPE)
    assert lib.emvs_packetize_range(capi.ptr(ev), len(ev), capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam), capi.ptr(K), 1.0,
                                    C.byref(cur), len(ev), capi.ptr(out), 4, C.byref(n)) == capi.EMVS_ERR_INVALID


def test_parallel_packetizer_equals_sequential_oracle(O):
    """Long lists take the speculative multi-threaded path of the product's packet stage (packets at cur + 1024*j
    computed by several threads, longest all-successful prefix kept).  It must reproduce the sequential loop bit for
    bit, including the one-event slides at both ends of the trajectory (no GPU needed)."""
    lib = capi.load()
    K = np.array([200, 200, 120, 90], np.float32)
    rng = np.random.default_rng(5)
    t_ctrl = np.linspace(100.0, 101.0, 21)
    pos = np.cumsum(rng.normal(0, 0.01, (21, 3)), axis=0)
    tr = _traj(t_ctrl.tolist(), pos.tolist())
    I = np.zeros((), capi.POSE_DTYPE); I["q"] = (1, 0, 0, 0)
    cam = capi.Camera(240, 180, 200, 200, 120, 90)
    n_in = 1024 * 1500 + 77
    ts = np.concatenate([np.linspace(99.9, 99.999, 3000), np.sort(rng.uniform(100.0, 100.999, n_in)),
                         np.linspace(101.001, 101.2, 2600)])
    ev = np.zeros(len(ts), O.EVENT_DTYPE)
    ev["sec"], ev["nsec"] = synth._split_time(ts)
    want = O.packetize(ev, tr, I, K, K.copy(), 1.0)
    assert len(want) > 1400 and want["first_event"][0] == 3000 - 512
    for limits in ([len(ev)], [700_000, len(ev)], [1024 * 600, 1024 * 600 + 5, 1_400_000, len(ev) - 1, len(ev)]):
        cur = C.c_size_t(0)
        got = []
        for lim in limits:
            out = np.zeros(len(ev) // 1024 + 1, capi.PACKET_DTYPE)
            n = C.c_size_t(0)
            capi.check(lib.emvs_packetize_range(capi.ptr(ev), len(ev), capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam),
                                                capi.ptr(K), 1.0, C.byref(cur), lim, capi.ptr(out), len(out), C.byref(n)))
            got.append(out[:n.value].copy())
        assert np.concatenate(got).tobytes() == want.tobytes(), limits
    # output capacity smaller than the list: exactly max_packets packets, the same ones, and the call reports that
    # the list was cut short instead of returning EMVS_OK with a silently truncated packet list
    out = np.zeros(1000, capi.PACKET_DTYPE)
    n = C.c_size_t(0)
    rc = lib.emvs_packetize(capi.ptr(ev), len(ev), capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam), capi.ptr(K), 1.0,
                            capi.ptr(out), 1000, C.byref(n))
    assert rc == capi.EMVS_ERR_INVALID and b"max_packets is too small" in lib.emvs_last_error()
    assert n.value == 1000 and out.tobytes() == want[:1000].tobytes()


def test_soa_packetizer_equals_aos(O, small_case):
    """emvs_packetize_soa (timestamps from an int64 nanosecond array, emvs_events_soa) produces the very packets of
    emvs_packetize on the dvs_msgs::Event structs with the same timestamps — and therefore the oracle's."""
    lib = capi.load()
    ev = small_case.events[0]
    cam = small_case.cams[0]
    tr = np.ascontiguousarray(small_case.trajs[0], capi.STAMPED_POSE_DTYPE)
    virt = small_case.virts[0]
    soa = api.EventsSoA.from_events(ev)
    assert len(soa) == len(ev) and soa.nbytes_device == 4 * len(ev)
    es = soa.c_struct()
    cs = cam.c_struct()
    out = np.zeros(len(ev) // 1024 + 1, capi.PACKET_DTYPE)
    n = C.c_size_t(0)
    capi.check(lib.emvs_packetize_soa(C.byref(es), capi.ptr(tr), len(tr), capi.ptr(np.ascontiguousarray(small_case.T_rv_w)),
                                      C.byref(cs), capi.ptr(virt), float(small_case.depths[0]), capi.ptr(out), len(out),
                                      C.byref(n)))
    assert out[:n.value].tobytes() == small_case.packets[0].tobytes()
    # too few events / missing arrays are errors, not crashes
    short = api.EventsSoA(soa.x[:100], soa.y[:100], soa.t_ns[:100]).c_struct()
    assert lib.emvs_packetize_soa(C.byref(short), capi.ptr(tr), len(tr), capi.ptr(np.ascontiguousarray(small_case.T_rv_w)),
                                  C.byref(cs), capi.ptr(virt), 1.0, capi.ptr(out), len(out), C.byref(n)) == capi.EMVS_ERR_TOO_FEW
    bad = capi.EventsSoA(None, soa.y.ctypes.data, soa.t_ns.ctypes.data, len(soa))
    assert lib.emvs_packetize_soa(C.byref(bad), capi.ptr(tr), len(tr), capi.ptr(np.ascontiguousarray(small_case.T_rv_w)),
                                  C.byref(cs), capi.ptr(virt), 1.0, capi.ptr(out), len(out), C.byref(n)) == capi.EMVS_ERR_INVALID
