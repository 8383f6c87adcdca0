"""Calibration loaders (calib.cpp:31-138, 365-589): no GPU needed — the LUT comes from the
library's host-side emvs_rectify_lut."""
import json

import numpy as np
import pytest

from dvs_mcemvs_b200 import calib, synth

cv2 = pytest.importorskip("cv2")
yaml = pytest.importorskip("yaml")


def test_dsec_zurich04a_matches_the_bench_rig():
    c = calib.dsec_zurich04a()
    rig = synth.rig_dsec()
    for a, b in zip((c.cam0, c.cam1), rig.cams):
        assert (a.width, a.height) == (640, 480)
        assert (a.fx, a.fy, a.cx, a.cy) == (b.fx, b.fy, b.cx, b.cy)
        assert a.lut.tobytes() == b.lut.tobytes()
    # both cameras share camera 0's projection matrix; the event cameras sit ~0.6 m apart along x
    assert (c.cam0.fx, c.cam0.cx) == (c.cam1.fx, c.cam1.cx)
    assert c.mat_1_0[0, 3] == pytest.approx(-0.599, abs=2e-3) and abs(c.mat_1_0[1, 3]) < 5e-3
    np.testing.assert_allclose(c.mat_1_0[:3, :3] @ c.mat_1_0[:3, :3].T, np.eye(3), atol=1e-9)
    assert c.mat_hand_eye.shape == (4, 4) and c.mat_hand_eye[3].tolist() == [0, 0, 0, 1]
    # LUT against OpenCV for a few raw pixels (image_geometry::rectifyPoint == undistortPoints with R, P)
    info = c.info[1]
    pts = np.array([[[0.0, 0.0]], [[639.0, 479.0]], [[320.0, 240.0]], [[17.0, 401.0]]], np.float64)
    want = cv2.undistortPoints(pts, info["K"], info["D"], R=info["R"], P=info["P"]).reshape(-1, 2)
    got = np.array([c.cam1.lut[int(y) * 640 + int(x)] for x, y in pts.reshape(-1, 2)])
    np.testing.assert_allclose(got, want.astype(np.float32), rtol=0, atol=1e-4)


def _dsec_yaml_files(tmp_path, rig):
    r = calib._DSEC_RIGS[rig]

    def cam(i, kind="event"):
        return dict(camera_type=kind, camera_matrix=list(r["K"][i]), distortion_coeffs=list(r["D"][i]),
                    distortion_model="radtan", resolution=[640, 480])
    doc = dict(intrinsics=dict(cam0=cam(0), cam1=cam(0, "frame"), cam2=cam(1, "frame"), cam3=cam(1)),
               extrinsics=dict(T_10=r["T_10"], T_21=r["T_21"], T_32=r["T_32"], R_rect1=r["R_rect1"]))
    p = tmp_path / "cam_to_cam.yaml"
    p.write_text(yaml.safe_dump(doc))
    q = tmp_path / "cam_to_lidar.yaml"
    q.write_text(yaml.safe_dump(dict(T_lidar_camRect1=r["T_lidar_camRect1"])))
    return str(p), str(q)


@pytest.mark.parametrize("rig,builtin", [("zurich_city_04_a", calib.dsec_zurich04a), ("interlaken_00_b", calib.dsec_interlaken00b)])
def test_dsec_yaml_equals_hard_coded_rig(tmp_path, rig, builtin):
    p, q = _dsec_yaml_files(tmp_path, rig)
    a, b = calib.dsec_yaml(p, q), builtin()
    for x, y in ((a.cam0, b.cam0), (a.cam1, b.cam1)):
        assert (x.fx, x.fy, x.cx, x.cy) == (y.fx, y.fy, y.cx, y.cy) and x.lut.tobytes() == y.lut.tobytes()
    assert np.array_equal(a.mat_1_0, b.mat_1_0) and np.array_equal(a.mat_hand_eye, b.mat_hand_eye)
    assert calib.dsec_yaml(p).mat_hand_eye is None            # no cam_to_lidar file: extrinsics only


def test_kalibr_yaml_swap_inverse_fisheye_and_hand_eye(tmp_path):
    T = np.eye(4)
    T[:3, 3] = [-0.1, 0.002, 0.001]
    c, s = np.cos(0.01), np.sin(0.01)
    T[:3, :3] = [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    doc = dict(cam0=dict(intrinsics=[300.0, 301.0, 170.0, 130.0], distortion_model="equidistant",
                         distortion_coeffs=[-0.02, 0.01, -0.003, 0.001], resolution=[346, 260]),
               cam1=dict(intrinsics=[299.0, 300.0, 172.0, 128.0], distortion_model="none", resolution=[346, 260],
                         T_cn_cnm1=T.tolist()))
    p = tmp_path / "camchain.yaml"
    p.write_text(yaml.safe_dump(doc))
    h = tmp_path / "hand_eye.json"
    h.write_text(json.dumps(dict(rotation=dict(w="0.5", i="0.5", j="0.5", k="0.5"), translation=dict(x="0.1", y="0.2", z="0.3"))))
    r = calib.kalibr_yaml(str(p), str(h))
    # sic: cam1 is built from camera_info[0] (the fisheye one), cam0 from camera_info[1] (zero distortion -> identity LUT)
    ident = np.stack(np.meshgrid(np.arange(346, dtype=np.float32), np.arange(260, dtype=np.float32)), -1).reshape(-1, 2)
    info0 = r.info[0]
    assert info0["distortion_model"] == "fisheye" and r.info[1]["distortion_model"] == "plumb_bob"
    # camera 1 of the file has zero distortion: image_geometry's rectifyPoint returns the raw pixel (distortion state
    # NONE) although its K differs from the shared P — the LUT is the identity
    P = info0["P"]
    assert np.array_equal(r.cam0.lut, ident)
    pts = np.array([[[10.0, 20.0]], [[300.0, 200.0]], [[173.0, 130.0]]], np.float32)
    want1 = cv2.fisheye.undistortPoints(pts, info0["K"], info0["D"], R=np.eye(3), P=P).reshape(-1, 2)
    got1 = np.array([r.cam1.lut[int(y) * 346 + int(x)] for x, y in pts.reshape(-1, 2)])
    np.testing.assert_allclose(got1, want1, atol=1e-3)
    assert (r.cam0.fx, r.cam0.cx) == (r.cam1.fx, r.cam1.cx) == (P[0, 0], P[0, 2])
    np.testing.assert_allclose(r.mat_1_0 @ T, np.eye(4), atol=1e-12)
    # quaternion (0.5, 0.5, 0.5, 0.5) is the cyclic permutation x -> y -> z -> x
    np.testing.assert_allclose(r.mat_hand_eye[:3, :3], [[0, 0, 1], [1, 0, 0], [0, 1, 0]], atol=1e-12)
    assert r.mat_hand_eye[:3, 3].tolist() == [0.1, 0.2, 0.3]
    # a projection matrix in the file is used as it is
    doc["cam0"]["projection_matrix"] = [[250.0, 0, 173.0, 0], [0, 250.0, 130.0, 0], [0, 0, 1, 0]]
    p.write_text(yaml.safe_dump(doc))
    r2 = calib.kalibr_yaml(str(p))
    assert (r2.cam0.fx, r2.cam0.fy, r2.cam0.cx, r2.cam0.cy) == (250.0, 250.0, 173.0, 130.0) and r2.mat_hand_eye is None
    doc["cam0"]["distortion_model"] = "kannala_brandt_9"
    p.write_text(yaml.safe_dump(doc))
    with pytest.raises(ValueError):
        calib.kalibr_yaml(str(p))


def test_mvsec_esim_and_basalt_loaders(tmp_path):
    T = np.eye(4)
    T[:3, 3] = [-0.1, 0.0, 0.0]
    doc = dict(cam0=dict(intrinsics=[226.4, 226.1, 173.6, 133.7], distortion_model="equidistant",
                         distortion_coeffs=[-0.048, 0.011, -0.05, 0.02], resolution=[346, 260],
                         projection_matrix=[[199.7, 0, 177.6, 0], [0, 199.7, 126.9, 0], [0, 0, 1, 0]]),
               cam1=dict(intrinsics=[226.1, 226.0, 174.5, 124.2], distortion_model="equidistant",
                         distortion_coeffs=[-0.045, 0.01, -0.045, 0.018], resolution=[346, 260], T_cn_cnm1=T.tolist(),
                         projection_matrix=[[199.7, 0, 177.6, -19.9], [0, 199.7, 126.9, 0], [0, 0, 1, 0]]))
    p = tmp_path / "camchain.yaml"
    p.write_text(yaml.safe_dump(doc))
    r = calib.kalibr_yaml_mvsec(str(p))
    assert np.array_equal(r.mat_1_0, T) and np.array_equal(r.mat_hand_eye, np.eye(4))     # not inverted, identity hand-eye
    assert r.info[0]["D"][0] == -0.048 and r.info[1]["D"][0] == -0.045                      # cameras in file order
    assert (r.cam1.fx, r.cam1.cx) == (199.7, 177.6)                                         # camera 0's P for both
    assert calib.kalibr_yaml_m3ed is calib.kalibr_yaml_mvsec

    TL, TR = np.eye(4), np.eye(4)
    TR[0, 3] = 0.2
    esim = dict(cameras=[dict(camera=dict(image_width=240, image_height=180, intrinsics=dict(data=[200.0, 200.0, 120.0, 90.0]),
                                          distortion=dict(type="none")), T_B_C=dict(data=TL.reshape(-1).tolist())),
                         dict(camera=dict(image_width=240, image_height=180, intrinsics=dict(data=[201.0, 201.0, 121.0, 91.0]),
                                          distortion=dict(type="none")), T_B_C=dict(data=TR.reshape(-1).tolist()))])
    q = tmp_path / "esim.yaml"
    q.write_text(yaml.safe_dump(esim))
    e = calib.esim_yaml(str(q))
    assert (e.cam1.fx, e.cam1.cx) == (200.0, 120.0)                                         # both cameras are the LEFT one
    assert e.mat_1_0[0, 3] == pytest.approx(-0.2) and np.array_equal(e.mat_hand_eye, np.eye(4))
    assert e.cam0.lut[5 * 240 + 7].tolist() == [7.0, 5.0]

    def cam(fx, cx, px):
        return (dict(camera_type="kb4", intrinsics=dict(fx=fx, fy=fx + 1, cx=cx, cy=360.0, k1=0.01, k2=-0.02, k3=0.003, k4=-0.001)),
                dict(qw=1.0, qx=0.0, qy=0.0, qz=0.0, px=px, py=0.0, pz=0.0))
    intr, ext = zip(*[cam(700.0, 640.0, 0.0)] * 2 + [cam(1049.3, 634.0, 0.05), cam(1048.1, 641.0, 0.17)])
    basalt = dict(value0=dict(resolution=[[1024, 1024]] * 2 + [[1280, 720]] * 2, intrinsics=list(intr), T_imu_cam=list(ext)))
    j = tmp_path / "calib.json"
    j.write_text(json.dumps(basalt))
    m = tmp_path / "mocap.json"
    m.write_text(json.dumps(dict(value0=dict(T_imu_marker=dict(qw=1.0, qx=0.0, qy=0.0, qz=0.0, px=0.0, py=0.3, pz=0.0)))))
    b = calib.basalt_json(str(j), str(m))
    assert (b.cam0.width, b.cam0.height) == (1280, 720)
    assert b.cam0.fx == pytest.approx(0.8 * float(np.float32(1049.3)), rel=1e-7) and b.cam0.cx == pytest.approx(634.0)
    assert (b.cam1.fx, b.cam1.cx) == (b.cam0.fx, b.cam0.cx)
    assert b.mat_1_0[0, 3] == pytest.approx(0.05 - 0.17) and b.mat_hand_eye[:3, 3].tolist() == pytest.approx([0.05, -0.3, 0.0])
    info = b.info[1]
    pts = np.array([[[100.0, 50.0]], [[1200.0, 700.0]], [[641.0, 360.0]]], np.float32)
    want = cv2.fisheye.undistortPoints(pts, info["K"], info["D"], R=np.eye(3), P=info["P"]).reshape(-1, 2)
    got = np.array([b.cam1.lut[int(y) * 1280 + int(x)] for x, y in pts.reshape(-1, 2)])
    np.testing.assert_allclose(got, want, atol=2e-3)
    assert np.array_equal(calib.basalt_json(str(j)).mat_hand_eye[:3, 3], [0.05, 0.0, 0.0])


def test_hard_coded_rigs():
    e = calib.esim()
    assert (e.cam0.width, e.cam0.height, e.cam0.fx, e.cam0.cx, e.cam0.cy) == (240, 180, 200.0, 120.0, 90.0)
    assert e.mat_1_0[0, 3] == -0.2 and e.cam1.lut[0].tolist() == [0.0, 0.0]
    # the engine's synthetic ESIM rig (BASELINE configs[0]) is this calibration
    rig = synth.rig_esim()
    assert (rig.cams[0].fx, rig.cams[0].cx, rig.cams[0].cy) == (e.cam0.fx, e.cam0.cx, e.cam0.cy)

    s = calib.slider()                                           # rectification rotations R != I enter the LUT
    assert s.mat_1_0[0, 3] == -0.15 and (s.cam0.fx, s.cam0.cx) == (s.cam1.fx, s.cam1.cx) == (193.4488673170594, 137.1049880981445)
    for cam, info in zip((s.cam0, s.cam1), s.info):
        pts = np.array([[[0.0, 0.0]], [[239.0, 179.0]], [[120.0, 90.0]], [[33.0, 150.0]]], np.float64)
        want = cv2.undistortPoints(pts, info["K"], info["D"], R=info["R"], P=info["P"]).reshape(-1, 2)
        got = np.array([cam.lut[int(y) * 240 + int(x)] for x, y in pts.reshape(-1, 2)])
        np.testing.assert_allclose(got, want.astype(np.float32), atol=2e-4)

    h = calib.hkust()
    assert (h.cam1.width, h.cam1.height, h.cam1.fx) == (346, 260, 189.705) and h.mat_1_0[0, 3] == pytest.approx(-0.0730597639)
    info = h.info[1]
    Kstd = info["K"].copy()
    Kstd[2] = [0, 0, 1]                                          # cv::undistortPoints ignores K's third row
    pts = np.array([[[10.0, 10.0]], [[300.0, 200.0]]], np.float64)
    want = cv2.undistortPoints(pts, Kstd, info["D"], R=np.eye(3), P=info["P"]).reshape(-1, 2)
    got = np.array([h.cam1.lut[int(y) * 346 + int(x)] for x, y in pts.reshape(-1, 2)])
    np.testing.assert_allclose(got, want.astype(np.float32), atol=2e-4)

    v = calib.evimo2()
    assert (v.cam2.fx, v.cam2.cx) == (v.cam0.fx, v.cam0.cx) and v.cam2.lut.shape == (640 * 480, 2)
    for m in (v.mat_1_0, v.mat_2_0, v.mat_hand_eye):
        np.testing.assert_allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-12)
        assert np.linalg.det(m[:3, :3]) == pytest.approx(1.0)
    # T_B_0 = Rz(yaw) Ry(pitch) Rx(roll): yaw -1.79 rad maps the base x axis mostly onto -y
    assert v.mat_hand_eye[:3, 3].tolist() == [0.135419, -0.0214639, -0.0715952]
    assert v.mat_hand_eye[1, 0] == pytest.approx(np.sin(-1.79144) * np.cos(0.0496968))
    # cameras 0 and 1 are ~11 cm apart, cameras 0 and 2 ~11.8 cm
    assert 0.10 < np.linalg.norm(v.mat_1_0[:3, 3]) < 0.12 and 0.10 < np.linalg.norm(v.mat_2_0[:3, 3]) < 0.13
