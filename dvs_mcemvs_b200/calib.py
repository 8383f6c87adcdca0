"""Stereo-rig calibration loaders of the mapping path's callers, ROS-free.

The reference fills two image_geometry::PinholeCameraModel objects plus the extrinsics `mat4_1_0` / `matR_L`
(right-from-left) and the hand-eye transform (mapper_emvs_stereo/src/calib.cpp); the mapper then reads the
projection-matrix intrinsics and builds its rectification LUT from K, D, R, P (mapper_emvs_stereo.cpp:29-64, 256-299).
Here the same inputs produce `api.CameraModel`s (projection intrinsics + the LUT, via `emvs_rectify_lut` — bit-exact
against OpenCV, tests/test_rectify_lut.py) and 4x4 float64 matrices.

  dsec_yaml(cam_to_cam.yaml, cam_to_lidar.yaml)    get_camera_calib_dsec_yaml           calib.cpp:365-457
  dsec_zurich04a(), dsec_interlaken00b()           the two hard-coded DSEC rigs         calib.cpp:459-589
  kalibr_yaml(camchain.yaml[, hand_eye.json])      get_camera_calib_sony                calib.cpp:31-138
  kalibr_yaml_mvsec / kalibr_yaml_m3ed             get_camera_calib_yaml_mvsec / _m3ed  calib.cpp:141-228, 811-885
  esim_yaml(rig.yaml)                              get_camera_calib_yaml                calib.cpp:231-267
  basalt_json(calib.json[, mocap.json])            get_camera_calib_json (TUM-VIE)      calib.cpp:271-361
  slider(), hkust(), evimo2(), esim()              the remaining hard-coded rigs        calib.cpp:591-806, 897-925

Conventions kept literally: the rectification rotation is ignored (R = I: "work on unrectified images"); when the file
has no projection matrix, P = cv::getOptimalNewCameraMatrix(K, D, size, alpha = 0); BOTH cameras get camera 0's P;
"none" distortion becomes plumb_bob with zero coefficients; get_camera_calib_sony hands camera_info[0] to cam1 and
camera_info[1] to cam0 (calib.cpp:106-108) and inverts T_cn_cnm1.
"""
import json

import numpy as np

from .api import CameraModel


class StereoCalib:
    """cam0 / cam1: api.CameraModel; mat_1_0: 4x4 right-from-left (the reference's mat4_1_0 / matR_L);
    mat_hand_eye: 4x4 or None; info: the sensor_msgs/CameraInfo-like dicts the cameras were built from."""

    def __init__(self, cam0, cam1, mat_1_0, mat_hand_eye, info):
        self.cam0, self.cam1, self.mat_1_0, self.mat_hand_eye, self.info = cam0, cam1, mat_1_0, mat_hand_eye, info


def _cv2():
    try:
        import cv2
    except ImportError as e:
        raise RuntimeError("calibration without a projection matrix needs OpenCV (cv2.getOptimalNewCameraMatrix)") from e
    return cv2


def optimal_projection(K, D, width, height):
    """P = [getOptimalNewCameraMatrix(K, D, (w, h), 0) | 0] — calib.cpp:92-99, 399-405, 475-481."""
    Kn = _cv2().getOptimalNewCameraMatrix(np.asarray(K, np.float64).reshape(3, 3), np.asarray(D, np.float64), (int(width), int(height)), 0)
    if isinstance(Kn, tuple):
        Kn = Kn[0]
    P = np.zeros((3, 4), np.float64)
    P[:, :3] = Kn
    return P


def _info(width, height, fx, fy, cx, cy, model, D):
    if model == "none":
        model, D = "plumb_bob", [0.0] * 5
    elif model == "radtan":
        model = "plumb_bob"
    elif model == "equidistant":
        model = "fisheye"
    elif model not in ("plumb_bob", "fisheye"):
        raise ValueError(f"unknown distortion model {model!r}")
    return dict(width=int(width), height=int(height), K=np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64),
                D=np.asarray(D, np.float64), R=np.eye(3), distortion_model=model, P=None)


def _camera(info):
    return CameraModel.from_camera_info(info["width"], info["height"], info["K"], info["D"], info["R"], info["P"],
                                        info["distortion_model"])


def _mat4(node):
    m = np.asarray(node, np.float64)
    if m.shape != (4, 4):
        raise ValueError(f"expected a 4x4 matrix, got shape {m.shape}")
    return m


def _rot4(node):
    m = np.eye(4)
    m[:3, :3] = np.asarray(node, np.float64).reshape(3, 3)
    return m


def _load_yaml(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def dsec_yaml(calib_path, mocap_calib_path=None, event_cam_ids=(0, 3)):
    """DSEC cam_to_cam.yaml (+ cam_to_lidar.yaml): event cameras 0 and 3, matR_L = T_32 T_21 T_10,
    hand-eye = T_lidar_camRect1 * R_rect1 * T_10 (calib.cpp:365-457)."""
    c = _load_yaml(calib_path)
    infos = []
    for cid in event_cam_ids:
        cam = c["intrinsics"][f"cam{cid}"]
        fx, fy, cx, cy = (float(v) for v in cam["camera_matrix"])
        model = cam["distortion_model"]
        D = [float(v) for v in cam.get("distortion_coeffs", [])][:4]
        infos.append(_info(cam["resolution"][0], cam["resolution"][1], fx, fy, cx, cy, model, D))
    infos[0]["P"] = optimal_projection(infos[0]["K"], infos[0]["D"], infos[0]["width"], infos[0]["height"])
    # the reference computes camera 1's own P and then overwrites it with camera 0's (calib.cpp:408-410)
    infos[1]["P"] = infos[0]["P"]
    ext = c["extrinsics"]
    T_10, T_21, T_32 = _mat4(ext["T_10"]), _mat4(ext["T_21"]), _mat4(ext["T_32"])
    mat_R_L = T_32 @ T_21 @ T_10
    hand_eye = None
    if mocap_calib_path:
        T_lidar_camRect1 = _mat4(_load_yaml(mocap_calib_path)["T_lidar_camRect1"])
        hand_eye = T_lidar_camRect1 @ _rot4(ext["R_rect1"]) @ T_10
    return StereoCalib(_camera(infos[0]), _camera(infos[1]), mat_R_L, hand_eye, infos)


_DSEC_RIGS = {
    # K (fx, fy, cx, cy) and D of event cameras 0 / 3, then T_10, T_21, T_32, T_lidar_camRect1, R_rect1 (calib.cpp:459-589)
    "zurich_city_04_a": dict(
        K=[(553.4686750102932, 553.3994078799127, 346.65339162053317, 216.52092103243012),
           (552.1819422959984, 551.4454720096484, 336.87432177064744, 226.32630571403274)],
        D=[(-0.09356476362537607, 0.19445779814646236, 7.642434980998821e-05, 0.0019563864604273664),
           (-0.09493681546997375, 0.2021148065491477, 0.0005821287651820125, 0.0014552921745527136)],
        T_10=[[0.9997329831508507, 0.00994674446197701, 0.020857245142004693, -0.043722240320426424],
              [-0.01003579267550241, 0.999940949009329, 0.004169095789442527, 0.0010155694745410755],
              [-0.020814544570561252, -0.004377301558648307, 0.9997737713930034, -0.013372668558381158], [0, 0, 0, 1]],
        T_21=[[0.9998379578286035, -0.017926384876108554, 0.0016440226264295469, -0.5092603987305321],
              [0.017914084504235202, 0.9998135043384297, 0.007214022378586629, -0.0022179629729152214],
              [-0.0017730373650056029, -0.007183402242479184, 0.9999726271607238, 0.0042971588717280644], [0, 0, 0, 1]],
        T_32=[[0.9999876185667624, -0.0034167786978265787, -0.0036177806040117192, -0.046041759529914676],
              [0.0033579259589126046, 0.9998639316478117, -0.016150619896091543, -0.0011068440180470077],
              [0.0036724714325840242, 0.01613827168886575, 0.9998630251891839, 0.012672727774474509], [0, 0, 0, 1]],
        T_lidar_camRect1=[[0.006502250714427837, 0.0016414391549515739, 0.9999775129537399, 0.448],
                          [-0.9996294044397522, 0.026445536238290795, 0.006456577459882262, 0.255],
                          [-0.026434343477244382, -0.999648908012493, 0.0018127863517872211, -0.215], [0, 0, 0, 1]],
        R_rect1=[[0.9998858610925897, -0.013510711178262034, -0.006762061119800281],
                 [0.013535205789223095, 0.9999019509726164, 0.0035897974036225495],
                 [0.00671289739037555, -0.0036809135568848755, 0.9999706935125713]]),
    "interlaken_00_b": dict(
        K=[(555.6627242364661, 555.8306341927942, 342.5725306057865, 215.26831427862848),
           (553.800041834315, 553.7026022383894, 333.21860953836267, 226.01033624096638)],
        D=[(-0.09094341408134071, 0.18339771556281387, -0.0006982341741678465, 0.00041396758898911876),
           (-0.09492592983896557, 0.20394312250370014, 0.00033282360055722797, -0.001101242451777801)],
        T_10=[[0.9996874046885865, 0.009652146488870916, 0.023063585478994113, -0.04410263392688484],
              [-0.009722042371104245, 0.9999484753460813, 0.0029203673010648615, 0.0005281285423087664],
              [-0.023034209322743096, -0.0031436795631953228, 0.9997297347181744, -0.01229891454144492], [0, 0, 0, 1]],
        T_21=[[0.9998543808844597, -0.01706309861700861, -0.00026017635946350924, -0.5094961871754736],
              [0.017064416377671962, 0.9998338346058513, 0.00641162000174109, -0.002022496204233391],
              [0.0001507310227716978, -0.006415126105036775, 0.9999794115066636, 0.005365297617411473], [0, 0, 0, 1]],
        T_32=[[0.9999880111304372, -0.003533401537847065, -0.003390083916194203, -0.04551026028184807],
              [0.003476600244706753, 0.9998558803824363, -0.016617211420558598, -0.001048727690114844],
              [0.0034483106189848347, 0.016605226232405814, 0.999856177465359, 0.013554100781902953], [0, 0, 0, 1]],
        T_lidar_camRect1=[[0.01539728189227399, -0.0012823052573279758, 0.9998806325774878, 0.448],
                          [-0.9996610000153124, 0.020978176075891836, 0.015420803380972237, 0.255],
                          [-0.02099544614233234, -0.9997791115150167, -0.0009588636652390625, -0.215], [0, 0, 0, 1]],
        R_rect1=[[0.9998572179847892, -0.013025778024398856, -0.010764420587133948],
                 [0.013060715513432202, 0.9999096430275752, 0.003181743349841093],
                 [0.01072200326407413, -0.0033218800890692088, 0.9999369998948329]]),
}


def _dsec_builtin(name):
    r = _DSEC_RIGS[name]
    infos = [_info(640, 480, *r["K"][i], "plumb_bob", r["D"][i]) for i in range(2)]
    infos[0]["P"] = optimal_projection(infos[0]["K"], infos[0]["D"], 640, 480)
    infos[1]["P"] = infos[0]["P"]          # camera_info.P is not recomputed for the second camera (calib.cpp:485-490)
    T_10, T_21, T_32 = (_mat4(r[k]) for k in ("T_10", "T_21", "T_32"))
    hand_eye = _mat4(r["T_lidar_camRect1"]) @ _rot4(r["R_rect1"]) @ T_10
    return StereoCalib(_camera(infos[0]), _camera(infos[1]), T_32 @ T_21 @ T_10, hand_eye, infos)


def dsec_zurich04a():
    """get_camera_calib_dsec_zurich04a (calib.cpp:459-521)."""
    return _dsec_builtin("zurich_city_04_a")


def dsec_interlaken00b():
    """get_camera_calib_dsec_interlaken00b (calib.cpp:525-587)."""
    return _dsec_builtin("interlaken_00_b")


def _quat_to_rot(w, x, y, z):
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def _kalibr_infos(c):
    infos = []
    for i in range(2):
        cam = c[f"cam{i}"]
        fx, fy, cx, cy = (float(v) for v in cam["intrinsics"])
        info = _info(cam["resolution"][0], cam["resolution"][1], fx, fy, cx, cy, cam["distortion_model"],
                     [float(v) for v in cam.get("distortion_coeffs", [])])
        if cam.get("projection_matrix") is not None:
            info["P"] = np.asarray(cam["projection_matrix"], np.float64).reshape(3, 4)
        else:
            info["P"] = optimal_projection(info["K"], info["D"], info["width"], info["height"])
        infos.append(info)
    infos[1]["P"] = infos[0]["P"]
    return infos


def _pose4(qw, qx, qy, qz, px, py, pz):
    m = np.eye(4)
    m[:3, :3] = _quat_to_rot(float(qw), float(qx), float(qy), float(qz))
    m[:3, 3] = [float(px), float(py), float(pz)]
    return m


def kalibr_yaml(calib_path, mocap_calib_path=None):
    """Kalibr camchain (cam0 / cam1 with `intrinsics`, `distortion_model`, `distortion_coeffs`, `resolution`, optional
    `projection_matrix`, cam1.T_cn_cnm1) + optional hand-eye JSON {rotation: {w,i,j,k}, translation: {x,y,z}} —
    get_camera_calib_sony (calib.cpp:31-138), including its camera swap and the inverted extrinsics."""
    c = _load_yaml(calib_path)
    infos = _kalibr_infos(c)
    cam1, cam0 = _camera(infos[0]), _camera(infos[1])          # sic: calib.cpp:106-108
    mat_1_0 = np.linalg.inv(_mat4(c["cam1"]["T_cn_cnm1"]))
    hand_eye = None
    if mocap_calib_path:
        with open(mocap_calib_path) as f:
            m = json.load(f)
        r, t = m["rotation"], m["translation"]
        hand_eye = _pose4(r["w"], r["i"], r["j"], r["k"], t["x"], t["y"], t["z"])
    return StereoCalib(cam0, cam1, mat_1_0, hand_eye, infos)


def kalibr_yaml_mvsec(calib_path):
    """The same camchain format read the way get_camera_calib_yaml_mvsec / get_camera_calib_yaml_m3ed do
    (calib.cpp:141-228, 811-885): cameras in file order, mat4_1_0 = T_cn_cnm1 as it is, identity hand-eye."""
    c = _load_yaml(calib_path)
    infos = _kalibr_infos(c)
    return StereoCalib(_camera(infos[0]), _camera(infos[1]), _mat4(c["cam1"]["T_cn_cnm1"]), np.eye(4), infos)


kalibr_yaml_m3ed = kalibr_yaml_mvsec


def esim_yaml(calib_path):
    """ESIM / rpg rig file: cameras[i].camera {image_width, image_height, intrinsics.data [fx,fy,cx,cy],
    distortion.type} and cameras[i].T_B_C.data (row-major 4x4) — get_camera_calib_yaml (calib.cpp:231-267): both
    cameras are the LEFT one with P = K, mat4_1_0 = T_B_right^-1 T_B_left, identity hand-eye."""
    c = _load_yaml(calib_path)
    cams = c["cameras"]
    left = cams[0]["camera"]
    fx, fy, cx, cy = (float(v) for v in left["intrinsics"]["data"])
    if left["distortion"]["type"] != "none":
        raise ValueError("get_camera_calib_yaml only handles distortion type 'none' (the reference leaves the model unset)")
    info = _info(left["image_width"], left["image_height"], fx, fy, cx, cy, "none", [])
    info["P"] = np.array([[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0]], np.float64)
    T_B_left = np.asarray(cams[0]["T_B_C"]["data"], np.float64).reshape(4, 4)
    T_B_right = np.asarray(cams[1]["T_B_C"]["data"], np.float64).reshape(4, 4)
    return StereoCalib(_camera(info), _camera(info), np.linalg.inv(T_B_right) @ T_B_left, np.eye(4), [info, dict(info)])


def basalt_json(camera_calib_path, mocap_calib_path=None):
    """TUM-VIE (basalt) calibration JSON: value0.{resolution, intrinsics, T_imu_cam}[2..3] are the two event cameras
    (kb4 = equidistant fisheye), P = diag(0.8 fx, 0.8 fy) around the raw principal point, mat4_1_0 =
    T_imu_cam1^-1 T_imu_cam0, hand-eye = T_imu_marker^-1 T_imu_cam0 (or T_imu_cam0) — get_camera_calib_json
    (calib.cpp:271-361)."""
    with open(camera_calib_path) as f:
        v = json.load(f)["value0"]
    infos, T_imu_cam = [], []
    for i in range(2):
        res = v["resolution"][i + 2]
        k = v["intrinsics"][i + 2]["intrinsics"]
        if v["intrinsics"][i + 2]["camera_type"] != "kb4":
            raise ValueError("get_camera_calib_json only handles camera_type 'kb4'")
        info = _info(res[0], res[1], k["fx"], k["fy"], k["cx"], k["cy"], "fisheye", [k["k1"], k["k2"], k["k3"], k["k4"]])
        # 0.8 * (float) fx: the reference scales the focal length in single precision
        info["P"] = np.array([[0.8 * float(np.float32(k["fx"])), 0, k["cx"], 0],
                              [0, 0.8 * float(np.float32(k["fy"])), k["cy"], 0], [0, 0, 1, 0]], np.float64)
        infos.append(info)
        e = v["T_imu_cam"][i + 2]
        T_imu_cam.append(_pose4(e["qw"], e["qx"], e["qy"], e["qz"], e["px"], e["py"], e["pz"]))
    infos[1]["P"] = infos[0]["P"]
    mat_1_0 = np.linalg.inv(T_imu_cam[1]) @ T_imu_cam[0]
    hand_eye = T_imu_cam[0]
    if mocap_calib_path:
        with open(mocap_calib_path) as f:
            e = json.load(f)["value0"]["T_imu_marker"]
        hand_eye = np.linalg.inv(_pose4(e["qw"], e["qx"], e["qy"], e["qz"], e["px"], e["py"], e["pz"])) @ T_imu_cam[0]
    return StereoCalib(_camera(infos[0]), _camera(infos[1]), mat_1_0, hand_eye, infos)


# ---- the remaining hard-coded rigs of calib.cpp ------------------------------------------------------------------
def _fixed(width, height, K, D, R, P):
    return dict(width=width, height=height, K=np.asarray(K, np.float64).reshape(3, 3), D=np.asarray(D, np.float64),
                R=np.asarray(R, np.float64).reshape(3, 3), P=np.asarray(P, np.float64).reshape(3, 4), distortion_model="plumb_bob")


def _translation(x, y=0.0, z=0.0):
    m = np.eye(4)
    m[:3, 3] = [x, y, z]
    return m


def esim():
    """get_camera_calib_ESIM (calib.cpp:897-925): 240x180, f = 200, no distortion, 0.2 m baseline."""
    info = _fixed(240, 180, [200, 0, 120, 0, 200, 90, 0, 0, 1], [0.0] * 5, np.eye(3), [200, 0, 120, 0, 0, 200, 90, 0, 0, 0, 1, 0])
    return StereoCalib(_camera(info), _camera(info), _translation(-0.2), np.eye(4), [info, dict(info)])


def slider():
    """get_camera_calib_slider (calib.cpp:591-631): the rpg stereo DAVIS slider, WITH rectification rotations, one
    projection matrix for both cameras, 0.15 m baseline."""
    P = [193.4488673170594, 0, 137.1049880981445, 0, 0, 193.4488673170594, 108.951057434082, 0, 0, 0, 1, 0]
    i0 = _fixed(240, 180, [198.9035679113487, 0, 139.8751842835105, 0, 198.8472302496314, 104.0170363461823, 0, 0, 1],
                [-0.3693817071651257, 0.1677750957297015, 0.0007676172676998043, -0.001200264930281811, 0],
                [0.9997156212398773, 0.02379292338064179, 0.001604196362382244,
                 -0.02378757584963585, 0.9997116745775861, -0.003273980524687744,
                 -0.001681631399562056, 0.003234889531517614, 0.9999933537806914], P)
    i1 = _fixed(240, 180, [198.1315372343827, 0, 132.4194623418875, 0, 198.0677328525099, 111.1773834719834, 0, 0, 1],
                [-0.3425648318682812, 0.1238467273033616, 0.0004063467878750188, 0.0004690582572504908, 0],
                [0.9999365173339012, 0.007076042854404519, 0.008768746756027635,
                 -0.007104545173989656, 0.999969566560146, 0.003223568113795293,
                 -0.008745669786783357, -0.003285661430544528, 0.9999563578921555], P)
    return StereoCalib(_camera(i0), _camera(i1), _translation(-0.15), np.eye(4), [i0, i1])


def hkust():
    """get_camera_calib_hkust (calib.cpp:635-674).  Camera 1's K carries a non-standard third row in the reference;
    only fx, fy, cx, cy of K enter the rectification (cv::undistortPoints), so it is kept as written."""
    P = [189.705, 0, 165.382, 0, 0, 189.705, 121.295, 0, 0, 0, 1, 0]
    i0 = _fixed(346, 260, [263.796, 0, 176.994, 0, 263.738, 124.373, 0, 0, 1],
                [-0.386589, 0.157241, 0.000322143, 6.13759e-06], np.eye(3), P)
    i1 = _fixed(346, 260, [263.485, 0, 162.942, 0, 263.276, 118.029, -0.0151344, 0.00133093, 0.999885],
                [-0.383425, 0.152823, -0.000257745, 0.000268432], np.eye(3), P)
    mat_1_0 = np.array([[9.99990798e-01, -6.32492385e-04, -4.24307214e-03, -7.30597639e-02],
                        [6.44736387e-04, 9.99995631e-01, 2.88489843e-03, -1.23275257e-03],
                        [4.24122892e-03, -2.88760755e-03, 9.99986837e-01, -1.10420407e-03], [0, 0, 0, 1.0]])
    return StereoCalib(_camera(i0), _camera(i1), mat_1_0, np.eye(4), [i0, i1])


def _rpy_pose(x, y, z, roll, pitch, yaw):
    """tf::Quaternion::setRPY + tf::Transform: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    m = np.eye(4)
    m[:3, :3] = [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                 [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                 [-sp, cp * sr, cp * cr]]
    m[:3, 3] = [x, y, z]
    return m


class TrinocularCalib(StereoCalib):
    """EVIMO2: three event cameras (cam2, mat_2_0 beside the stereo fields)."""

    def __init__(self, cam0, cam1, cam2, mat_1_0, mat_2_0, mat_hand_eye, info):
        super().__init__(cam0, cam1, mat_1_0, mat_hand_eye, info)
        self.cam2, self.mat_2_0 = cam2, mat_2_0


def evimo2():
    """get_camera_calib_evimo2 (calib.cpp:678-806): three 640x480 cameras sharing camera 0's optimal projection matrix,
    extrinsics given as x y z roll pitch yaw wrt the rig base."""
    Ks = [(519.638, 519.384, 321.661, 240.727), (558.417, 557.475, 324.905, 225.3), (556.184, 555.632, 326.875, 202.887)]
    Ds = [(0.108306, -0.154485, 0.00103538, -0.000401824), (-0.115993, 0.204851, -0.00217161, 0.000676025),
          (-0.110194, 0.205049, 0.00206719, -0.00040706)]
    infos = [_info(640, 480, *k, "plumb_bob", d) for k, d in zip(Ks, Ds)]
    infos[0]["P"] = optimal_projection(infos[0]["K"], infos[0]["D"], 640, 480)
    infos[1]["P"] = infos[2]["P"] = infos[0]["P"]
    ext = [(0.135419, -0.0214639, -0.0715952, -0.00748326, 0.0496968, -1.79144),
           (0.118804, 0.0850843, -0.0194297, 0.018838, 0.00459314, -0.195708),
           (0.0754507, -0.119035, -0.0336873, -0.0122178, -0.00473387, 2.93835)]
    T_B = [_rpy_pose(*e) for e in ext]
    return TrinocularCalib(_camera(infos[0]), _camera(infos[1]), _camera(infos[2]), np.linalg.inv(T_B[1]) @ T_B[0],
                           np.linalg.inv(T_B[2]) @ T_B[0], T_B[0], infos)
