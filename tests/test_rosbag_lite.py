"""Minimal rosbag v2 reader: round trip of synthetic poses / events, and — in the build
container only, where /root/reference exists — the pose bags the reference ships (data/DSEC/*/pose.bag)."""
import glob
import os

import numpy as np
import pytest

from dvs_mcemvs_b200 import _capi as capi
from dvs_mcemvs_b200 import rosbag_lite, synth


def test_round_trip_poses_and_events(tmp_path):
    sc, n, _, _ = synth.config("esim_small", events_per_cam=12_345)
    ev, traj = sc.events(0, n), sc.trajectory(0)
    p = str(tmp_path / "synthetic.bag")
    rosbag_lite.write_bag(p, poses=traj, events=ev, sensor=(240, 180), events_per_message=1000)
    assert rosbag_lite.read_poses(p, "/pose").tobytes() == traj.tobytes()
    back = rosbag_lite.read_events(p, "/dvs/events")
    assert back.tobytes() == ev.tobytes()                       # 13-byte wire events -> the 16-byte in-memory struct
    t = ev["sec"] + 1e-9 * ev["nsec"]
    window = rosbag_lite.read_events(p, t_min=float(t[3000]), t_max=float(t[9000]))
    assert 5900 <= len(window) <= 6100 and np.all(np.diff(window["sec"].astype(np.int64) * 10**9 + window["nsec"]) >= 0)
    assert len(rosbag_lite.read_poses(p, "/other_topic")) == 0
    with pytest.raises(ValueError):
        (tmp_path / "x.bag").write_bytes(b"not a bag")
        rosbag_lite.read_poses(str(tmp_path / "x.bag"))


BAGS = sorted(glob.glob("/root/reference/data/DSEC/*/pose.bag"))


@pytest.mark.skipif(not BAGS, reason="reference data not present (GPU box)")
def test_reads_the_reference_pose_bags():
    counts = {}
    for b in BAGS:
        poses = rosbag_lite.read_poses(b)
        counts[os.path.basename(os.path.dirname(b))] = len(poses)
        t = poses["sec"] + 1e-9 * poses["nsec"]
        assert np.all(np.diff(t) > 0) and 600 < t[-1] - t[0] < 700
        np.testing.assert_allclose(np.linalg.norm(poses["T"]["q"], axis=1), 1.0, atol=1e-4)
        # usable as a LinearTrajectory of the engine
        from dvs_mcemvs_b200 import api
        tr = api.LinearTrajectory(poses[:50])
        mid = tr.getPoseAt(int(poses["sec"][10]), int(poses["nsec"][10]) + 1000)
        assert mid is not None and tr.getPoseAt(int(poses["sec"][0]) - 1, 0) is None
    assert counts == {"interlaken_00-odometry": 13268, "zurich_city_02-odometry": 6750, "zurich_city_04-odometry": 6205}


def test_parse_rosbag_windows_and_retimes_like_data_loading(tmp_path):
    """data_loading::parse_rosbag semantics (data_loading.cpp:33-219): shared time origin, [tmin, tmax] window relative
    to it, stop only after the message that crossed tmax, re-timed stamps, events_offset, CameraInfo, final sort."""
    from dvs_mcemvs_b200 import _capi as capi
    t0 = 1000.25
    n = 10_000
    ts = t0 + np.sort(np.random.default_rng(0).uniform(0.0, 1.0, n))
    ts[0] = t0
    ev = np.zeros(n, capi.EVENT_DTYPE)
    ev["sec"], ev["nsec"] = synth._split_time(ts)
    ev["x"], ev["y"], ev["polarity"] = np.arange(n) % 240, np.arange(n) % 180, np.arange(n) % 2
    tp = t0 + np.linspace(0.0, 1.0, 21)
    poses = np.zeros(21, capi.STAMPED_POSE_DTYPE)
    poses["sec"], poses["nsec"] = synth._split_time(tp)
    poses["T"]["q"][:, 0] = 1.0
    poses["T"]["t"][:, 0] = np.linspace(0, 0.2, 21)
    info = dict(width=240, height=180, distortion_model="plumb_bob", D=[0.1, -0.2, 0.0, 0.001, 0.0],
                K=np.array([[200.0, 0, 120], [0, 200, 90], [0, 0, 1]]), R=np.eye(3),
                P=np.array([[190.0, 0, 121, 0], [0, 190, 91, 0], [0, 0, 1, 0]]))
    left, right = str(tmp_path / "left.bag"), str(tmp_path / "right.bag")
    rosbag_lite.write_bag(left, poses=poses, events=ev, events_per_message=1000, camera_info=info, sensor=(240, 180))
    shift = 0.5                                                   # the right bag starts later: it must reuse the origin
    ev_r = ev.copy()
    ev_r["sec"], ev_r["nsec"] = synth._split_time(ts + shift)
    rosbag_lite.write_bag(right, events=ev_r, events_per_message=1000, poses=poses, pose_with_covariance=True)

    origin = rosbag_lite.TimeOrigin()
    e, p, ci = rosbag_lite.parse_rosbag(left, "/dvs/events", "/dvs/camera_info", "/pose", tmin=0.2, tmax=0.6, origin=origin)
    assert origin.stamp == (int(ev["sec"][0]), int(ev["nsec"][0]))
    rel = (ts - t0)
    # replay of the loop: messages in bag-time order (an EventArray is stamped with its last event), the first message
    # of ANY topic that holds a stamp past tmax is consumed whole and ends the loop
    tp_rel = tp - t0
    stream = sorted([(ts[1000 * k + 999], "ev", k) for k in range(10)] + [(tp[j], "pose", j) for j in range(21)])
    want_idx, want_poses = [], []
    for _, kind, k in stream:
        if kind == "ev":
            idx = np.arange(1000 * k, 1000 * k + 1000)
            want_idx.extend(idx[rel[idx] >= 0.2])
            if (rel[idx] > 0.6).any():
                break
        elif tp_rel[k] >= 0.2 - 1e-12:
            want_poses.append(k)
            if tp_rel[k] > 0.6:
                break
    want_idx = np.array(want_idx)
    assert len(e) == len(want_idx) and len(e) > 3000
    t_new = e["sec"] + 1e-9 * e["nsec"]
    np.testing.assert_allclose(t_new, rel[want_idx], atol=2e-9)                 # re-timed to the origin
    assert np.array_equal(e["x"], ev["x"][want_idx]) and np.all(np.diff(t_new) >= 0)
    assert len(p) == len(want_poses) and (p["sec"][-1] + 1e-9 * p["nsec"][-1]) > 0.6   # the crossing message is consumed
    assert ci["width"] == 240 and ci["distortion_model"] == "plumb_bob" and ci["D"].tolist() == info["D"]
    assert np.array_equal(ci["K"], info["K"]) and np.array_equal(ci["P"], info["P"])
    assert abs((p["sec"][0] + 1e-9 * p["nsec"][0]) - 0.2) < 1e-6 and np.all(np.diff(p["sec"] + 1e-9 * p["nsec"]) > 0)
    # second bag, same origin: its stamps are 0.5 s later; events_offset shifts them back
    e2, p2, ci2 = rosbag_lite.parse_rosbag(right, "/dvs/events", None, "/pose", tmin=0.7, tmax=1.2, events_offset=0.5, origin=origin)
    assert ci2 is None and len(e2) > 3000 and len(p2) >= 5                      # PoseWithCovarianceStamped is read too
    t2 = e2["sec"] + 1e-9 * e2["nsec"]
    assert t2.min() == pytest.approx(0.2, abs=1e-3)                             # (0.7 rel) - 0.5 offset
    # without a pose topic (the MVSEC overload, data_loading.cpp:221-303)
    e3, p3, _ = rosbag_lite.parse_rosbag(left, "/dvs/events", tmin=0.0, tmax=0.05, origin=origin)
    assert len(p3) == 0 and 0 < len(e3) <= 1000
    # a fresh origin is taken from the first message of the first call
    o2 = rosbag_lite.TimeOrigin()
    rosbag_lite.parse_rosbag(right, "/dvs/events", origin=o2)
    assert o2.to_sec() == pytest.approx(min(t0 + shift, tp[0]), abs=1e-6) or o2.to_sec() == pytest.approx(t0 + shift, abs=1e-6)
    assert rosbag_lite._time_from_sec(1.9999999996) == (2, 0) and rosbag_lite._time_from_sec(0.25) == (0, 250000000)
    with pytest.raises(ValueError):
        rosbag_lite._time_from_sec(-1e-3)


@pytest.mark.parametrize("pose_type", ["geometry_msgs/PoseStamped", "geometry_msgs/PoseWithCovarianceStamped",
                                       "nav_msgs/Odometry", "vicon/Subject"])
def test_parse_rosbag_gt_reads_every_pose_message_type(tmp_path, pose_type):
    """data_loading::parse_rosbag_gt (data_loading.cpp:305-465) accepts four pose-carrying message types; every one
    round-trips through the writer, is re-timed to the first stamp and windowed with the stop-AFTER-the-message rule."""
    t0 = 1234.5
    tp = t0 + np.linspace(0.0, 2.0, 41)
    poses = np.zeros(41, capi.STAMPED_POSE_DTYPE)
    poses["sec"], poses["nsec"] = synth._split_time(tp)
    ang = np.linspace(0, 0.3, 41)
    poses["T"]["q"][:, 0], poses["T"]["q"][:, 2] = np.cos(ang / 2), np.sin(ang / 2)
    poses["T"]["t"] = np.stack([np.linspace(0, 1, 41), np.linspace(2, 3, 41), np.linspace(-1, 0, 41)], 1)
    bag = str(tmp_path / "poses.bag")
    rosbag_lite.write_bag(bag, poses=poses, pose_type=pose_type)
    got = rosbag_lite.parse_rosbag_gt(bag, "/pose", origin=rosbag_lite.TimeOrigin())
    assert len(got) == 41 and got["sec"][0] == 0 and got["nsec"][0] == 0            # re-timed to the first stamp
    np.testing.assert_allclose(got["sec"] + 1e-9 * got["nsec"], tp - t0, atol=2e-9)
    assert got["T"].tobytes() == poses["T"].tobytes()                                # position + quaternion bit for bit
    win = rosbag_lite.parse_rosbag_gt(bag, "/pose", tmin=0.5, tmax=1.0, origin=rosbag_lite.TimeOrigin())
    rel = win["sec"] + 1e-9 * win["nsec"]
    assert rel[0] == pytest.approx(0.5, abs=1e-6) and rel[-1] == pytest.approx(1.05, abs=1e-6)   # the first pose past tmax is kept
    assert len(rosbag_lite.parse_rosbag_gt(bag, "/other", origin=rosbag_lite.TimeOrigin())) == 0
    # the events + poses loader takes the same messages, except nav_msgs/Odometry (data_loading.cpp:111-219 has no such branch)
    _, p, _ = rosbag_lite.parse_rosbag(bag, "/dvs/events", None, "/pose", origin=rosbag_lite.TimeOrigin())
    assert len(p) == (0 if pose_type == "nav_msgs/Odometry" else 41)
