"""Seeded synthetic rigs, trajectories and event streams.

The reference ships no event data; these generators produce the inputs of BASELINE.json's
configs.  Identical bytes are fed to the CUDA path and to the CPU oracle.

  structured  N_pts 3-D points uniform in the reference-view frustum; each event is a random
              point projected into the camera at the event's own time, rounded to the pixel.
              Gives a DSI with real ray intersections (meaningful argmax).
  uniform     x, y i.i.d. uniform over the sensor (worst-case locality, throughput only).

Rig constants: ESIM (mapper_emvs_stereo/src/calib.cpp:901-926) and DSEC zurich_city_04_a
(calib.cpp:466-488: intrinsics + plumb_bob distortion; the ~0.6 m stereo baseline).
"""
import numpy as np

from ._capi import EVENT_DTYPE, POSE_DTYPE, STAMPED_POSE_DTYPE
from .api import CameraModel, ShapeDSI

T_BASE = 1000.0  # seconds; all windows start here


# ------------------------------------------------------------------------------------------
# small SE(3) helpers (float64, numpy) — only used to *generate* data
# ------------------------------------------------------------------------------------------
def _axis_angle_quat(axis, angle):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[np.cos(angle / 2)], axis * np.sin(angle / 2)])


def _rotate(axis, angle, P):
    """Rodrigues rotation of points P [n,3] by per-point angles [n] about a fixed unit axis."""
    k = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    c, s = np.cos(angle)[:, None], np.sin(angle)[:, None]
    return P * c + np.cross(k[None, :], P) * s + k[None, :] * (P @ k)[:, None] * (1 - c)


def _split_time(t):
    sec = np.floor(t).astype(np.int64)
    nsec = np.rint((t - sec) * 1e9).astype(np.int64)
    carry = nsec >= 1000000000
    sec = sec + carry
    nsec = np.where(carry, nsec - 1000000000, nsec)
    return sec.astype(np.uint32), nsec.astype(np.uint32)


# ------------------------------------------------------------------------------------------
# forward plumb_bob distortion, used to place synthetic events on RAW (distorted) pixels; the inverse
# (the rectification LUT) comes from the engine's own precomputeRectifiedPoints restatement
# ------------------------------------------------------------------------------------------
def distort(xn, yn, D):
    k1, k2, p1, p2 = D[:4]
    k3 = D[4] if len(D) > 4 else 0.0
    r2 = xn * xn + yn * yn
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = xn * rad + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    yd = yn * rad + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
    return xd, yd


# ------------------------------------------------------------------------------------------
# rigs
# ------------------------------------------------------------------------------------------
class Rig:
    """cams: list of CameraModel; offsets: x position of camera i in the camera-0 frame (m);
    raw: optional per-camera (K_raw, D) to map rectified -> raw pixels when generating events."""

    def __init__(self, cams, offsets, raw=None):
        self.cams, self.offsets = cams, list(offsets)
        self.raw = raw or [None] * len(cams)


def rig_esim():
    cams = [CameraModel(240, 180, 200., 200., 120., 90.) for _ in range(2)]
    return Rig(cams, [0.0, 0.2])


# cv::getOptimalNewCameraMatrix(K_left, D_left, (640, 480), alpha=0) evaluated once with OpenCV 4.13; the
# reference computes it at start-up (calib.cpp:475-476) and reuses the LEFT camera's P for the right one (:482-488).
_DSEC_P = (557.9686136352767, 558.0457356705808, 345.95312896446654, 217.50102178331326)


# Who computes the DSEC rectification LUT: "engine" = the library's precomputeRectifiedPoints restatement
# (emvs_rectify_lut), "cv2" = OpenCV's cv2.undistortPoints, which it equals bit for bit (tests/test_rectify_lut.py).
# bench.py's CPU reference arm selects "cv2" so that it never maps libemvs_b200.so.
LUT_BACKEND = "engine"


def _plumb_bob_camera(width, height, Km, D, P):
    if LUT_BACKEND == "cv2":
        import cv2
        xs, ys = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(height, dtype=np.float32))
        px = np.stack([xs, ys], -1).reshape(-1, 1, 2)
        lut = cv2.undistortPoints(px, np.asarray(Km, np.float64), np.asarray(D, np.float64), R=np.eye(3),
                                  P=np.asarray(P, np.float64)).reshape(-1, 2)
        return CameraModel(width, height, P[0][0], P[1][1], P[0][2], P[1][2], lut=lut)
    return CameraModel.from_camera_info(width, height, Km, D, np.eye(3), P, "plumb_bob")


def rig_dsec():
    Ks = [(553.4686750102932, 553.3994078799127, 346.65339162053317, 216.52092103243012),
          (552.1819422959984, 551.4454720096484, 336.87432177064744, 226.32630571403274)]
    Ds = [(-0.09356476362537607, 0.19445779814646236, 7.642434980998821e-05, 0.0019563864604273664),
          (-0.09493681546997375, 0.2021148065491477, 0.0005821287651820125, 0.0014552921745527136)]
    fx, fy, cx, cy = _DSEC_P
    P = [[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0]]
    cams, raw = [], []
    for K, D in zip(Ks, Ds):
        Km = [[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]]
        # the LUT exactly as precomputeRectifiedPoints builds it (image_geometry::rectifyPoint -> cv::undistortPoints)
        cams.append(_plumb_bob_camera(640, 480, Km, D, P))
        raw.append((K, D))
    return Rig(cams, [0.0, 0.599], raw)


def rig_bar(n_cams=4, width=640, height=480, f=550.0, spacing=0.2):
    cams = [CameraModel(width, height, f, f, width / 2.0, height / 2.0) for _ in range(n_cams)]
    return Rig(cams, [spacing * i for i in range(n_cams)])


def rig_square(size, n_cams=1):
    """config-5 sweep rig: sensor == DSI x-y size, f = 0.8 W, identity LUT."""
    cams = [CameraModel(size, size, 0.8 * size, 0.8 * size, size / 2.0, size / 2.0) for _ in range(n_cams)]
    return Rig(cams, [0.2 * i for i in range(n_cams)])


# ------------------------------------------------------------------------------------------
# scene = rig + motion + reference view
# ------------------------------------------------------------------------------------------
class Scene:
    """Constant-velocity motion of camera 0: p(t) = v (t - t0), rotation about `axis` at rate w.
    Camera i sits at offsets[i] along x of camera 0.  Reference view = camera 0 at t_ref shifted
    by rv_pos along its x axis (process1.cpp:60-68)."""

    def __init__(self, rig, shape, duration=0.2, translation=(0.2, 0.0, 0.0), rot_axis=(0.0, 1.0, 0.0),
                 rot_deg=1.0, t_ref_frac=0.5, rv_pos=0.0, n_ctrl=64, seed=0, n_points=20000):
        self.rig, self.shape = rig, shape
        self.t0, self.T = T_BASE, float(duration)
        self.v = np.asarray(translation, np.float64) / self.T
        self.axis = np.asarray(rot_axis, np.float64)
        self.w = np.deg2rad(rot_deg) / self.T
        self.t_ref = self.t0 + t_ref_frac * self.T
        self.rv_pos = rv_pos
        self.n_ctrl = n_ctrl
        self.rng = np.random.default_rng(seed)
        cam0 = rig.cams[0]
        dimX = shape.dimX_ or cam0.width
        dimY = shape.dimY_ or cam0.height
        f = cam0.fx if shape.fov_ < 10 else 0.5 * dimX / np.tan(0.5 * np.deg2rad(shape.fov_))
        # points uniform in the reference-view frustum
        z = self.rng.uniform(shape.min_depth_, shape.max_depth_, n_points)
        u = self.rng.uniform(0, dimX - 1, n_points)
        v = self.rng.uniform(0, dimY - 1, n_points)
        P_rv = np.stack([(u - cam0.cx) / f * z, (v - cam0.cy) / f * z, z], axis=1)
        self.point_depth = z
        # reference view pose in the world: T_w_rv = T_w_c0(t_ref) * trans(rv_pos, 0, 0)
        ang = np.full(n_points, self.w * (self.t_ref - self.t0))
        self.P_w = _rotate(self.axis, ang, P_rv + np.array([rv_pos, 0, 0])) + self.v * (self.t_ref - self.t0)

    # --- poses ----------------------------------------------------------------------------
    def _pose_w_cam(self, cam_idx, t):
        q = _axis_angle_quat(self.axis, self.w * (t - self.t0))
        off = np.array([self.rig.offsets[cam_idx], 0.0, 0.0])
        p = self.v * (t - self.t0) + _rotate(self.axis, np.array([self.w * (t - self.t0)]), off[None, :])[0]
        return q, p

    def trajectory(self, cam_idx):
        """Control poses T_w_cam of camera cam_idx, padded beyond the window on both sides."""
        ts = np.linspace(self.t0 - 0.05 * self.T, self.t0 + 1.05 * self.T, self.n_ctrl)
        out = np.zeros(self.n_ctrl, dtype=STAMPED_POSE_DTYPE)
        sec, nsec = _split_time(ts)
        out["sec"], out["nsec"] = sec, nsec
        for i, t in enumerate(ts):
            q, p = self._pose_w_cam(cam_idx, t)
            out["T"]["q"][i] = q
            out["T"]["t"][i] = p
        return out

    def T_rv_w(self):
        q, p = self._pose_w_cam(0, self.t_ref)
        # T_w_rv = T_w_c0 * trans(rv_pos); invert analytically
        shift = _rotate(self.axis, np.array([self.w * (self.t_ref - self.t0)]), np.array([[self.rv_pos, 0., 0.]]))[0]
        pw = p + shift
        qi = q * np.array([1, -1, -1, -1.0])
        ang = np.array([-self.w * (self.t_ref - self.t0)])
        ti = -_rotate(self.axis, ang, pw[None, :])[0]
        out = np.zeros((), dtype=POSE_DTYPE)
        out["q"], out["t"] = qi, ti
        return out

    # --- events ---------------------------------------------------------------------------
    def _project(self, cam_idx, t, pts):
        cam = self.rig.cams[cam_idx]
        ang = self.w * (t - self.t0)
        off = np.array([self.rig.offsets[cam_idx], 0.0, 0.0])
        centre = self.v[None, :] * (t - self.t0)[:, None] + _rotate(self.axis, ang, np.broadcast_to(off, pts.shape))
        Pc = _rotate(self.axis, -ang, pts - centre)
        Z = Pc[:, 2]
        ok = Z > 1e-3
        Zs = np.where(ok, Z, 1.0)
        xn, yn = Pc[:, 0] / Zs, Pc[:, 1] / Zs
        raw = self.rig.raw[cam_idx]
        if raw is None:
            x, y = cam.fx * xn + cam.cx, cam.fy * yn + cam.cy
        else:
            K, D = raw
            xd, yd = distort(xn, yn, D)
            x, y = K[0] * xd + K[2], K[1] * yd + K[3]
        xi, yi = np.rint(x), np.rint(y)
        ok &= (xi >= 0) & (xi < cam.width) & (yi >= 0) & (yi < cam.height)
        return xi, yi, ok

    def skip_events(self):
        """Advance the scene's seed sequence exactly as one events() call does, without generating anything: processes
        that generate different subsets of the cameras' lists stay in step (the k-th events() call of a Scene seeds its
        generator from the k-th draw of the scene's own generator)."""
        self.rng.integers(1 << 62)

    def events(self, cam_idx, n_events, kind="structured", chunk=2_000_000, stream=0):
        """`stream` selects an independent event sample of the same scene, window and trajectory
        (a shard of a denser stream: multi-GPU sub-interval sharding)."""
        cam = self.rig.cams[cam_idx]
        rng = np.random.default_rng([int(self.rng.integers(1 << 62)), cam_idx, stream])
        ev = np.zeros(n_events, dtype=EVENT_DTYPE)
        t = np.sort(rng.uniform(self.t0, self.t0 + self.T, n_events))
        ev["sec"], ev["nsec"] = _split_time(t)
        ev["polarity"] = rng.integers(0, 2, n_events, dtype=np.uint8)
        if kind == "uniform":
            ev["x"] = rng.integers(0, cam.width, n_events, dtype=np.uint16)
            ev["y"] = rng.integers(0, cam.height, n_events, dtype=np.uint16)
            return ev
        for s in range(0, n_events, chunk):
            e = min(s + chunk, n_events)
            tt = t[s:e]
            x = np.zeros(e - s)
            y = np.zeros(e - s)
            todo = np.arange(e - s)
            for _ in range(6):
                ids = rng.integers(0, self.P_w.shape[0], todo.shape[0])
                xi, yi, ok = self._project(cam_idx, tt[todo], self.P_w[ids])
                x[todo[ok]], y[todo[ok]] = xi[ok], yi[ok]
                todo = todo[~ok]
                if todo.size == 0:
                    break
            if todo.size:  # leftover: noise events
                x[todo] = rng.integers(0, cam.width, todo.size)
                y[todo] = rng.integers(0, cam.height, todo.size)
            ev["x"][s:e], ev["y"][s:e] = x.astype(np.uint16), y.astype(np.uint16)
        return ev


# ------------------------------------------------------------------------------------------
# BASELINE.json configs
# ------------------------------------------------------------------------------------------
def config(name, events_per_cam=None, seed=None):
    """-> (scene, n_events_per_cam, fusion_method, description)."""
    if name == "esim_small":       # configs[0]: 240x180x64, 50k events/cam, HM
        shape = ShapeDSI(0, 0, 64, 1.0, 5.0, 0.0)
        sc = Scene(rig_esim(), shape, duration=0.2, translation=(0.15, 0.05, 0.0), rot_deg=1.0, seed=seed or 1)
        return sc, events_per_cam or 50_000, 2, "stereo 50k ev/cam 240x180x64 HM"
    if name == "dsec_stereo":      # configs[1]: 640x480x256, 5M (bench: 10M) events/cam, HM
        shape = ShapeDSI(0, 0, 256, 4.0, 200.0, 0.0)
        sc = Scene(rig_dsec(), shape, duration=0.2, translation=(0.0, 0.0, 2.0), rot_axis=(0, 1, 0), rot_deg=1.0,
                   t_ref_frac=0.97, seed=seed or 2, n_ctrl=8)
        return sc, events_per_cam or 5_000_000, 2, "DSEC-like stereo 640x480x256 HM"
    if name == "bar4":             # configs[2]: 4 cameras, 10M ev/cam, GM (n-ary extension)
        shape = ShapeDSI(0, 0, 256, 1.0, 10.0, 0.0)
        sc = Scene(rig_bar(4), shape, duration=0.2, translation=(0.2, 0.0, 0.0), rot_deg=2.0, seed=seed or 3)
        return sc, events_per_cam or 10_000_000, 3, "4-camera bar 640x480x256 GM"
    raise KeyError(name)
