"""precomputeRectifiedPoints (mapper_emvs_stereo.cpp:244-299): the product's host-side restatement of
image_geometry::rectifyPoint -> cv::undistortPoints (plumb_bob) and cv::fisheye::undistortPoints against
OpenCV itself (the Python cv2 of this image, 4.13).  Bit-exact: same double arithmetic, float result."""
import numpy as np
import pytest

from dvs_mcemvs_b200 import api

cv2 = pytest.importorskip("cv2")

W, H = 640, 480
# DSEC zurich_city_04_a left event camera (mapper_emvs_stereo/src/calib.cpp:466-488)
K = np.array([[553.4686750102932, 0, 346.65339162053317], [0, 553.3994078799127, 216.52092103243012], [0, 0, 1]])
D4 = np.array([-0.09356476362537607, 0.19445779814646236, 7.642434980998821e-05, 0.0019563864604273664])
TH = 0.01
R = np.array([[np.cos(TH), 0, np.sin(TH)], [0, 1, 0], [-np.sin(TH), 0, np.cos(TH)]])
P = np.array([[540.0, 0, 330.5, 0], [0, 541.0, 225.25, 0], [0, 0, 1, 0]])


def _pixels():
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    return np.stack([xs, ys], -1).reshape(-1, 1, 2)


@pytest.mark.parametrize("D", [D4, np.r_[D4, 0.01], np.r_[D4, 0.01, 0.002, -0.001, 0.0005]], ids=["k1k2p1p2", "k3", "rational"])
def test_plumb_bob_lut_equals_opencv(D):
    want = cv2.undistortPoints(_pixels(), K, D, R=R, P=P).reshape(-1, 2)
    cam = api.CameraModel.from_camera_info(W, H, K, D, R, P, "plumb_bob")
    assert cam.lut.tobytes() == want.tobytes()
    assert (cam.fx, cam.fy, cam.cx, cam.cy) == (540.0, 541.0, 330.5, 225.25)     # projection-matrix intrinsics (:46-48)


def test_fisheye_lut_equals_opencv():
    D = np.array([-0.03, 0.01, -0.004, 0.0007])
    want = cv2.fisheye.undistortPoints(_pixels(), K, D, R=R, P=P[:, :3]).reshape(-1, 2)
    cam = api.CameraModel.from_camera_info(W, H, K, D, R, P, "fisheye")
    assert cam.lut.tobytes() == want.tobytes()


def test_zero_distortion_is_identity_and_unknown_model_is_an_error():
    cam = api.CameraModel.from_camera_info(W, H, K, np.zeros(5), np.eye(3), P, "plumb_bob")
    assert np.array_equal(cam.lut, _pixels().reshape(-1, 2))          # image_geometry: distortion_state NONE
    with pytest.raises(api.EmvsError, match="Distortion model not set properly"):
        api.CameraModel.from_camera_info(W, H, K, D4, R, P, "equidistant")
