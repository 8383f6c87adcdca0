"""Minimal rosbag v2 reader: round trip of synthetic poses / events, and — in the build
container only, where /root/reference exists — the pose bags the reference ships (data/DSEC/*/pose.bag)."""
import glob
import os

import numpy as np
import pytest

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
