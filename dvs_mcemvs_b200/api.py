"""Python mirror of the reference's mapping-path classes over the C-ABI.

Names and argument meaning follow the reference so that tests read like the reference's
callers (paths relative to the reference root):

    ShapeDSI, MapperEMVS      mapper_emvs_stereo/include/mapper_emvs_stereo/mapper_emvs_stereo.hpp:40-155
    Grid3D                    cartesian3dgrid/include/cartesian3dgrid/cartesian3dgrid.h:22-247
    LinearTrajectory          mapper_emvs_stereo/include/mapper_emvs_stereo/trajectory.hpp:80-127
    process_1                 mapper_emvs_stereo/src/process1.cpp:28-224 (steps 1-3, no file output)

All compute happens in libemvs_b200.so (hand-written sm_100a CUDA).  There is no CPU path:
constructing a Context without a B200-class GPU raises EmvsError.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import (EVENT_DTYPE, PACKET_DTYPE, POSE_DTYPE, STAMPED_POSE_DTYPE, Camera, EmvsError, Shape, check, ptr)


def _lib():
    return capi.load()


# --------------------------------------------------------------------------------------------
# plain data
# --------------------------------------------------------------------------------------------
class ShapeDSI:
    """EMVS::ShapeDSI(dimX, dimY, dimZ, min_depth, max_depth, fov) — mapper_emvs_stereo.hpp:40-65.

    ``inverse_depth`` selects the compile-time USE_INVERSE_DEPTH variant (CMakeLists.txt:41-44)."""

    def __init__(self, dimX, dimY, dimZ, min_depth, max_depth, fov, inverse_depth=False):
        self.dimX_, self.dimY_, self.dimZ_ = int(dimX), int(dimY), int(dimZ)
        self.min_depth_, self.max_depth_, self.fov_ = float(min_depth), float(max_depth), float(fov)
        self.inverse_depth = bool(inverse_depth)

    def c_struct(self):
        return Shape(self.dimX_, self.dimY_, self.dimZ_, self.min_depth_, self.max_depth_, self.fov_,
                     1 if self.inverse_depth else 0)


class CameraModel:
    """The part of image_geometry::PinholeCameraModel the mapper reads (mapper_emvs_stereo.cpp:34-48):
    fullResolution() and fx(), fy(), cx(), cy() of the projection matrix, plus the rectification
    LUT that precomputeRectifiedPoints (:256-299) derives from it with OpenCV — here an input."""

    def __init__(self, width, height, fx, fy, cx, cy, lut=None):
        self.width, self.height = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        if lut is None:  # zero distortion: rectified pixel == raw pixel
            xs, ys = np.meshgrid(np.arange(self.width, dtype=np.float32), np.arange(self.height, dtype=np.float32))
            lut = np.stack([xs, ys], axis=-1)
        self.lut = np.ascontiguousarray(lut, dtype=np.float32).reshape(self.height * self.width, 2)

    def c_struct(self):
        return Camera(self.width, self.height, self.fx, self.fy, self.cx, self.cy)

    @classmethod
    def from_camera_info(cls, width, height, K, D, R, P, distortion_model):
        """What MapperEMVS derives from a sensor_msgs/CameraInfo-backed PinholeCameraModel: projection-matrix
        intrinsics (mapper_emvs_stereo.cpp:46-48) and the LUT of precomputeRectifiedPoints (:256-299).
        distortion_model: "plumb_bob" or "fisheye" (anything else raises, like the reference's LOG(ERROR))."""
        model = {"plumb_bob": capi.DISTORTION_PLUMB_BOB, "fisheye": capi.DISTORTION_FISHEYE}.get(distortion_model, -1)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        P = np.ascontiguousarray(P, np.float64).reshape(12)
        D = np.ascontiguousarray(D, np.float64).reshape(-1)
        lut = np.empty((int(height) * int(width), 2), np.float32)
        check(_lib().emvs_rectify_lut(model, ptr(K), ptr(D), D.shape[0], ptr(R), ptr(P), int(width), int(height), ptr(lut)))
        return cls(width, height, P[0], P[5], P[2], P[6], lut=lut)


class EventsSoA:
    """Structure-of-arrays event list (emvs_events_soa): x, y uint16 and t_ns int64 (ros::Time::toNSec()), the
    layout of the DSEC / TUM-VIE HDF5 event files.  Only x and y travel to the GPU (4 bytes per event); the
    packet stage reads one timestamp per 1024 events on the host.  The arrays are kept alive by this object."""

    def __init__(self, x, y, t_ns):
        self.x = np.ascontiguousarray(x, dtype=np.uint16)
        self.y = np.ascontiguousarray(y, dtype=np.uint16)
        self.t_ns = np.ascontiguousarray(t_ns, dtype=np.int64)
        if not (self.x.shape == self.y.shape == self.t_ns.shape and self.x.ndim == 1):
            raise ValueError("x, y, t_ns must be 1-D arrays of equal length")

    @classmethod
    def from_events(cls, events, pinned=False):
        """From a dvs_msgs::Event-layout array (EVENT_DTYPE)."""
        t = events["sec"].astype(np.int64) * 1_000_000_000 + events["nsec"].astype(np.int64)
        if not pinned:
            return cls(events["x"], events["y"], t)
        out = cls.__new__(cls)
        out.x, out.y, out.t_ns = (pinned_empty(len(events), d) for d in (np.uint16, np.uint16, np.int64))
        out.x[...], out.y[...], out.t_ns[...] = events["x"], events["y"], t
        return out

    def __len__(self):
        return self.x.shape[0]

    @property
    def nbytes_device(self):
        return self.x.nbytes + self.y.nbytes

    def c_struct(self):
        return capi.EventsSoA(self.x.ctypes.data, self.y.ctypes.data, self.t_ns.ctypes.data, self.x.shape[0])


def make_pose(q=(1.0, 0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)):
    p = np.zeros((), dtype=POSE_DTYPE)
    p["q"] = q
    p["t"] = t
    return p


def pose_compose(a, b):
    out = np.zeros((), dtype=POSE_DTYPE)
    check(_lib().emvs_pose_compose(ptr(np.ascontiguousarray(a)), ptr(np.ascontiguousarray(b)), ptr(out)))
    return out


def pose_inverse(a):
    out = np.zeros((), dtype=POSE_DTYPE)
    check(_lib().emvs_pose_inverse(ptr(np.ascontiguousarray(a)), ptr(out)))
    return out


class LinearTrajectory:
    """LinearTrajectory(poses) — trajectory.hpp:80-127.  ``poses``: array of STAMPED_POSE_DTYPE
    sorted by strictly increasing time (the std::map ordering)."""

    def __init__(self, poses):
        poses = np.ascontiguousarray(poses, dtype=STAMPED_POSE_DTYPE)
        if poses.shape[0] < 2:
            raise ValueError("At least two poses need to be provided")  # trajectory.hpp:89
        self.poses = poses

    def getPoseAt(self, sec, nsec):
        """Returns the pose or None when t is outside the control poses (getPoseAt -> false)."""
        out = np.zeros((), dtype=POSE_DTYPE)
        found = C.c_int(0)
        check(_lib().emvs_trajectory_pose_at(ptr(self.poses), self.poses.shape[0], int(sec), int(nsec), ptr(out),
                                             C.byref(found)))
        return out if found.value else None

    def getNumControlPoses(self):
        return self.poses.shape[0]


# --------------------------------------------------------------------------------------------
# device objects
# --------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + stream + scratch (emvs_context)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(_lib().emvs_context_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._prefetch_ref = None   # the arrays of the pending prefetch: they must outlive it (the engine matches by address)

    def close(self):
        if getattr(self, "_h", None):
            _lib().emvs_context_destroy(self._h)
            self._h = None
            self._prefetch_ref = None

    __del__ = close

    def sync(self):
        check(_lib().emvs_context_sync(self._h))

    def set_slab(self, planes):
        check(_lib().emvs_context_set_slab(self._h, int(planes)))

    def set_upload_split(self, percent, min_events=1 << 20):
        """evaluateDSI on an idle pipeline votes the first `percent` % of the events while the rest is uploaded
        (head + tail as two builds).  Builds on the single multi-slab vote launch use the deferred form with its own
        geometry instead (emvs_b200.h: EMVS_UPLOAD_DEFER_*); percent = 0 turns both off, min_events applies to both."""
        check(_lib().emvs_context_set_upload_split(self._h, int(percent), int(min_events)))

    def prefetch_events(self, events):
        """Start uploading the event list of a LATER evaluateDSI / build call (same array object) now, under the
        current work.  `events` must already be contiguous EVENT_DTYPE (ideally pinned) and stay unchanged."""
        if events.dtype != EVENT_DTYPE or not events.flags["C_CONTIGUOUS"]:
            raise ValueError("prefetch_events needs a contiguous EVENT_DTYPE array (the later call must see the same buffer)")
        check(_lib().emvs_context_prefetch_events(self._h, ptr(events), events.shape[0]))
        self._prefetch_ref = events

    def prefetch_pending(self):
        """Generation number of the pending prefetch, 0 when none is pending."""
        g = C.c_uint64(0)
        check(_lib().emvs_context_prefetch_pending(self._h, C.byref(g)))
        if not g.value:
            self._prefetch_ref = None
        return g.value

    def prefetch_cancel(self):
        """Withdraw the pending prefetch (waits for its copies); its arrays may be freed or rewritten afterwards."""
        check(_lib().emvs_context_prefetch_cancel(self._h))
        self._prefetch_ref = None

    def selftest_division(self, n_pairs=1 << 28, seed=1):
        """Mismatches between the vote kernel's prepared division and __fdiv_rn over ~n_pairs operand pairs."""
        bad = C.c_uint64(0)
        check(_lib().emvs_selftest_division(self._h, int(n_pairs), int(seed), C.byref(bad)))
        return bad.value

    def launch_count(self):
        n = C.c_uint64(0)
        check(_lib().emvs_context_launch_count(self._h, C.byref(n)))
        return n.value

    def timer(self):
        return Timer(self)

    def profile_vote(self, enable=True):
        check(_lib().emvs_context_profile_vote(self._h, 1 if enable else 0))

    def vote_time(self):
        """-> (summed device ms of the vote-kernel launches since profile_vote(True), launch count)"""
        ms, n = C.c_double(0), C.c_uint64(0)
        check(_lib().emvs_context_vote_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # multi-GPU
    def comm_init(self, unique_id, n_ranks, rank):
        buf = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        check(_lib().emvs_comm_init(self._h, ptr(buf), int(n_ranks), int(rank)))

    def comm_destroy(self):
        check(_lib().emvs_comm_destroy(self._h))


def comm_unique_id():
    buf = np.zeros(128, dtype=np.uint8)
    check(_lib().emvs_comm_unique_id(ptr(buf)))
    return buf.tobytes()


class Timer:
    def __init__(self, ctx):
        h = C.c_void_p()
        check(_lib().emvs_timer_create(ctx._h, C.byref(h)))
        self._h = h

    def start(self):
        check(_lib().emvs_timer_start(self._h))

    def stop(self):
        check(_lib().emvs_timer_stop(self._h))

    def elapsed_ms(self):
        ms = C.c_float(0)
        check(_lib().emvs_timer_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().emvs_timer_destroy(self._h)
            self._h = None


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (cudaHostAlloc); keep the returned array alive."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(_lib().emvs_host_alloc(n, C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = (buf, p)
    return arr


_PINNED = {}


class Grid3D:
    """Grid3D(dimX, dimY, dimZ) — cartesian3dgrid.h:22-247; the volume lives in HBM."""

    def __init__(self, ctx, dimX, dimY, dimZ, _handle=None, _owner=None):
        self.ctx = ctx
        self._owner = _owner  # a MapperEMVS keeps its dsi_ alive
        if _handle is None:
            h = C.c_void_p()
            check(_lib().emvs_grid_create(ctx._h, int(dimX), int(dimY), int(dimZ), C.byref(h)))
            self._h, self._owned = h, True
        else:
            self._h, self._owned = _handle, False
        self.size_ = (int(dimX), int(dimY), int(dimZ))

    def close(self):
        if getattr(self, "_owned", False) and self._h:
            _lib().emvs_grid_destroy(self._h)
        self._h = None

    __del__ = close

    def getDimensions(self):
        return self.size_

    def resetGrid(self):
        check(_lib().emvs_grid_reset(self._h))

    def _op(self, other, op, n=0, eps=0.0):
        check(_lib().emvs_grid_op(self._h, other._h if other is not None else None, op, int(n), float(eps)))

    # cartesian3dgrid.h:64-192 — same names, same defaults
    def addTwoGrids(self, g): self._op(g, capi.OP_ADD)
    def addInverseOfTwoGrids(self, g, eps=1e-2): self._op(g, capi.OP_ADD_INV, 0, eps)
    def computeHMfromSumOfInv(self, n): self._op(None, capi.OP_HM_FROM_SUMINV, n)
    def computeAMfromSum(self, n): self._op(None, capi.OP_AM_FROM_SUM, n)
    def minTwoGrids(self, g): self._op(g, capi.OP_MIN)
    def geometricMeanTwoGrids(self, g): self._op(g, capi.OP_GM)
    def arithmeticMeanTwoGrids(self, g): self._op(g, capi.OP_AM)
    def rmsTwoGrids(self, g): self._op(g, capi.OP_RMS)
    def maxTwoGrids(self, g): self._op(g, capi.OP_MAX)

    def harmonicMeanTwoGrids(self, g, n=None, eps=1e-1):
        if n is None:
            self._op(g, capi.OP_HM, 2, eps)
        else:
            self._op(g, capi.OP_HM_N, n, eps)

    def copyFrom(self, g):
        """resetGrid(); addTwoGrids(g) — the initialisation idiom of process1.cpp:126-127."""
        check(_lib().emvs_grid_copy(self._h, g._h))

    def computeMeanSquare(self):
        out = C.c_double(0)
        check(_lib().emvs_grid_mean_square(self._h, C.byref(out)))
        return out.value

    def collapseMaxZSlice(self, depths=None):
        """-> (max_val float32 [dimY, dimX], max_pos_idx uint8|uint16 [dimY, dimX][, depth])."""
        dimX, dimY, dimZ = self.size_
        conf = np.empty((dimY, dimX), np.float32)
        idx = np.empty((dimY, dimX), np.uint8 if dimZ <= 256 else np.uint16)
        depth = np.empty((dimY, dimX), np.float32) if depths is not None else None
        d = np.ascontiguousarray(depths, np.float32) if depths is not None else None
        check(_lib().emvs_grid_collapse_max(self._h, ptr(d), ptr(conf), ptr(idx), ptr(depth)))
        return (conf, idx) if depth is None else (conf, idx, depth)

    def download(self):
        """Host copy, shape [dimZ, dimY, dimX] (the layout writeGridNpy dumps, cartesian3dgrid_IO.cpp:30-36)."""
        dimX, dimY, dimZ = self.size_
        out = np.empty((dimZ, dimY, dimX), np.float32)
        check(_lib().emvs_grid_download(self._h, ptr(out)))
        return out

    def upload(self, vol):
        dimX, dimY, dimZ = self.size_
        vol = np.ascontiguousarray(vol, np.float32)
        assert vol.size == dimX * dimY * dimZ
        check(_lib().emvs_grid_upload(self._h, ptr(vol)))

    def device_ptr(self):
        p = C.c_void_p()
        check(_lib().emvs_grid_device_ptr(self._h, C.byref(p)))
        return p.value

    def allreduce(self):
        check(_lib().emvs_grid_allreduce(self._h))

    def allreduce_async(self):
        check(_lib().emvs_grid_allreduce_async(self._h))


def fuse_collapse(grids, method, depths=None, fused_out=None):
    """Fast path of process_1 steps 2+3: n-ary fusion + Z-argmax without materialising the fused DSI."""
    g0 = grids[0]
    dimX, dimY, dimZ = g0.size_
    arr = (C.c_void_p * len(grids))(*[g._h for g in grids])
    conf = np.empty((dimY, dimX), np.float32)
    idx = np.empty((dimY, dimX), np.uint8 if dimZ <= 256 else np.uint16)
    depth = np.empty((dimY, dimX), np.float32) if depths is not None else None
    d = np.ascontiguousarray(depths, np.float32) if depths is not None else None
    check(_lib().emvs_fuse_collapse(arr, len(grids), int(method), ptr(d), fused_out._h if fused_out else None,
                                    ptr(conf), ptr(idx), ptr(depth)))
    return (conf, idx) if depth is None else (conf, idx, depth)


def fuse_collapse_device(grids, method, d_depths, d_conf, d_idx, d_depth, fused_out=None):
    """fuse_collapse with DEVICE pointers (ints) for the depth table and the outputs; asynchronous on
    the context's stream, nothing travels to the host."""
    arr = (C.c_void_p * len(grids))(*[g._h for g in grids])
    check(_lib().emvs_fuse_collapse_device(arr, len(grids), int(method), C.c_void_p(d_depths),
                                           fused_out._h if fused_out else None, C.c_void_p(d_conf),
                                           C.c_void_p(d_idx), C.c_void_p(d_depth) if d_depth else None))


class OptionsDepthMap:
    """The fields of EMVS::OptionsDepthMap (mapper_emvs_stereo.hpp:68-82) that getDepthMapFromDSI reads;
    defaults of main.cpp:73-75,97."""

    def __init__(self, adaptive_threshold_kernel_size=5, adaptive_threshold_c=5.0, median_filter_size=5, max_confidence=0.0):
        self.adaptive_threshold_kernel_size_ = int(adaptive_threshold_kernel_size)
        self.adaptive_threshold_c_ = float(adaptive_threshold_c)
        self.median_filter_size_ = int(median_filter_size)
        self.max_confidence = float(max_confidence)

    def c_struct(self):
        return capi.DepthMapOptions(self.adaptive_threshold_kernel_size_, self.adaptive_threshold_c_, self.max_confidence,
                                    self.median_filter_size_)


def depth_map_from_dsi(grids, method, depths, options):
    """getDepthMapFromDSI for method = -1 without the inpainting (mapper_emvs_stereo.cpp:339-436) on the fusion of
    `grids` (one grid: no fusion) -> (depth_map, confidence_map, mask, depth_cell_indices_filtered)."""
    g0 = grids[0]
    dimX, dimY, dimZ = g0.size_
    arr = (C.c_void_p * len(grids))(*[g._h for g in grids])
    depth = np.empty((dimY, dimX), np.float32)
    conf = np.empty((dimY, dimX), np.float32)
    mask = np.empty((dimY, dimX), np.uint8)
    idx_f = np.empty((dimY, dimX), np.uint8)
    d = np.ascontiguousarray(depths, np.float32)
    opt = options.c_struct()
    check(_lib().emvs_depth_map_from_dsi(arr, len(grids), int(method), ptr(d), C.byref(opt), ptr(depth), ptr(conf), ptr(mask),
                                         ptr(idx_f)))
    return depth, conf, mask, idx_f


def depth_map_postprocess(ctx, conf, idx, depths, ks=5, c=5.0, max_confidence=0.0, median_size=5):
    """The post-processing alone on host maps -> dict(conf, conf8, mask, idx_filtered, depth)."""
    conf = np.ascontiguousarray(conf, np.float32)
    idx = np.ascontiguousarray(idx, np.uint8)
    rows, cols = conf.shape
    d = np.ascontiguousarray(depths, np.float32)
    out = dict(conf=np.empty((rows, cols), np.float32), conf8=np.empty((rows, cols), np.uint8),
               mask=np.empty((rows, cols), np.uint8), idx_filtered=np.empty((rows, cols), np.uint8),
               depth=np.empty((rows, cols), np.float32))
    opt = capi.DepthMapOptions(int(ks), float(c), float(max_confidence), int(median_size))
    check(_lib().emvs_depth_map_postprocess(ctx._h, ptr(conf), ptr(idx), rows, cols, ptr(d), d.shape[0], C.byref(opt),
                                            ptr(out["depth"]), ptr(out["conf"]), ptr(out["mask"]), ptr(out["idx_filtered"]),
                                            ptr(out["conf8"])))
    return out


class PeerExchange:
    """Fused multi-GPU reduce + fuse + argmax over NVLink peer memory (emvs_exchange_*).

    grids: this rank's partial DSIs, one per camera (same order on every rank).
    allgather: callable(bytes) -> list of every rank's bytes in rank order (e.g. built on
    torch.distributed.all_gather_object); used once to exchange the CUDA IPC handles."""

    def __init__(self, ctx, grids, n_ranks, rank, allgather, participants=None):
        """participants: optional [n_cams][n_ranks] table (shard.participants) for camera x sub-interval sharding:
        rank r builds camera c iff participants[c][r]; default: every rank builds every camera."""
        self.ctx, self.grids = ctx, list(grids)
        arr = (C.c_void_p * len(grids))(*[g._h for g in grids])
        h = C.c_void_p()
        check(_lib().emvs_exchange_create(ctx._h, arr, len(grids), int(n_ranks), int(rank), C.byref(h)))
        self._h = h
        self._n_ranks = int(n_ranks)
        self.size_ = grids[0].size_
        n = C.c_size_t(0)
        check(_lib().emvs_exchange_blob_bytes(self._h, C.byref(n)))
        blob = np.zeros(n.value, np.uint8)
        check(_lib().emvs_exchange_export(self._h, ptr(blob)))
        blobs = allgather(blob.tobytes())
        assert len(blobs) == n_ranks and all(len(b) == n.value for b in blobs)
        cat = np.frombuffer(b"".join(blobs), np.uint8).copy()
        check(_lib().emvs_exchange_import(self._h, ptr(cat)))
        if participants is not None:
            self.set_participants(participants)

    def set_participants(self, participants):
        tab = np.ascontiguousarray(participants, np.uint8)
        assert tab.shape == (len(self.grids), self._n_ranks)
        check(_lib().emvs_exchange_set_participants(self._h, ptr(tab)))

    def close(self):
        if getattr(self, "_h", None):
            _lib().emvs_exchange_destroy(self._h)
            self._h = None

    __del__ = close

    def begin(self):
        """Start an overlapped round: builds issued with peer_reduce=True until fuse_collapse() reduce every
        Z-slab of this rank's row band over NVLink while the next slab is voted."""
        check(_lib().emvs_exchange_begin(self._h))

    def fuse_collapse(self, method, d_depths=None):
        """Collective and asynchronous: call on every rank after its builds were issued."""
        check(_lib().emvs_exchange_fuse_collapse(self._h, int(method), C.c_void_p(d_depths) if d_depths else None))

    def download(self, with_depth=True):
        dimX, dimY, dimZ = self.size_
        conf = np.empty((dimY, dimX), np.float32)
        idx = np.empty((dimY, dimX), np.uint8 if dimZ <= 256 else np.uint16)
        depth = np.empty((dimY, dimX), np.float32) if with_depth else None
        check(_lib().emvs_exchange_download(self._h, ptr(conf), ptr(idx), ptr(depth)))
        return (conf, idx, depth) if with_depth else (conf, idx)


class MapperEMVS:
    """EMVS::MapperEMVS(cam, dsi_shape) — mapper_emvs_stereo.hpp:94-155."""

    def __init__(self, ctx, cam, dsi_shape):
        self.ctx, self.cam, self.name = ctx, cam, ""
        h = C.c_void_p()
        cs, ss = cam.c_struct(), dsi_shape.c_struct()
        check(_lib().emvs_mapper_create(ctx._h, C.byref(cs), C.byref(ss), C.byref(h)))
        self._h = h
        shape = Shape()
        virt = np.zeros(4, np.float32)
        check(_lib().emvs_mapper_shape(self._h, C.byref(shape), ptr(virt)))
        self.dsi_shape_ = ShapeDSI(shape.dimX, shape.dimY, shape.dimZ, shape.min_depth, shape.max_depth,
                                   shape.fov_deg, bool(shape.inverse_depth))
        self.virtual_cam_ = virt  # fx, fy, cx, cy
        self.raw_depths_vec_ = np.zeros(shape.dimZ, np.float32)
        check(_lib().emvs_mapper_depths(self._h, ptr(self.raw_depths_vec_)))
        check(_lib().emvs_mapper_set_lut(self._h, ptr(cam.lut), cam.lut.shape[0]))
        gh = C.c_void_p()
        check(_lib().emvs_mapper_grid(self._h, C.byref(gh)))
        self.dsi_ = Grid3D(ctx, shape.dimX, shape.dimY, shape.dimZ, _handle=gh, _owner=self)

    def close(self):
        if getattr(self, "_h", None):
            _lib().emvs_mapper_destroy(self._h)
            self._h = None
            self.dsi_._h = None

    __del__ = close

    def packetize(self, events, trajectory, T_rv_w):
        """Packet stage of evaluateDSI (mapper_emvs_stereo.cpp:86-126) -> packets, or None if < 1024 events.
        `events`: EVENT_DTYPE array or EventsSoA."""
        n_pk = C.c_size_t(0)
        cs = self.cam.c_struct()
        T = ptr(np.ascontiguousarray(T_rv_w, dtype=POSE_DTYPE))
        if isinstance(events, EventsSoA):
            n = len(events)
            out = np.zeros(n // capi.PACKET_SIZE + 1, dtype=PACKET_DTYPE)
            es = events.c_struct()
            rc = _lib().emvs_packetize_soa(C.byref(es), ptr(trajectory.poses), trajectory.poses.shape[0], T, C.byref(cs),
                                           ptr(self.virtual_cam_), float(self.raw_depths_vec_[0]), ptr(out), out.shape[0],
                                           C.byref(n_pk))
        else:
            events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
            n = events.shape[0]
            out = np.zeros(n // capi.PACKET_SIZE + 1, dtype=PACKET_DTYPE)
            rc = _lib().emvs_packetize(ptr(events), n, ptr(trajectory.poses), trajectory.poses.shape[0], T, C.byref(cs),
                                       ptr(self.virtual_cam_), float(self.raw_depths_vec_[0]), ptr(out), out.shape[0],
                                       C.byref(n_pk))
        if rc == capi.EMVS_ERR_TOO_FEW:
            return None
        check(rc)
        return out[:n_pk.value]

    def packetize_range(self, events, trajectory, T_rv_w, cursor, event_limit):
        """Resumable packet stage: packets that fit in events[:event_limit], continuing at event `cursor`.
        Returns (packets, new cursor)."""
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        n = events.shape[0]
        out = np.zeros(n // capi.PACKET_SIZE + 1, dtype=PACKET_DTYPE)
        n_pk, cur = C.c_size_t(0), C.c_size_t(int(cursor))
        cs = self.cam.c_struct()
        check(_lib().emvs_packetize_range(ptr(events), n, ptr(trajectory.poses), trajectory.poses.shape[0],
                                          ptr(np.ascontiguousarray(T_rv_w, dtype=POSE_DTYPE)), C.byref(cs),
                                          ptr(self.virtual_cam_), float(self.raw_depths_vec_[0]), C.byref(cur),
                                          int(event_limit), ptr(out), out.shape[0], C.byref(n_pk)))
        return out[:n_pk.value], cur.value

    def prefetch(self, events, trajectory, T_rv_w):
        """Streaming callers: start the upload AND the host packet stage of a LATER evaluateDSI(events, trajectory,
        T_rv_w) on this mapper now, under the current device work; that call then only launches kernels.  `events`
        must be the same contiguous EVENT_DTYPE array object (ideally pinned), `trajectory` the same object."""
        T = ptr(np.ascontiguousarray(T_rv_w, dtype=POSE_DTYPE))
        if isinstance(events, EventsSoA):
            es = events.c_struct()
            check(_lib().emvs_mapper_prefetch_dsi_soa(self._h, C.byref(es), ptr(trajectory.poses), trajectory.poses.shape[0], T))
        else:
            if events.dtype != EVENT_DTYPE or not events.flags["C_CONTIGUOUS"]:
                raise ValueError("prefetch needs a contiguous EVENT_DTYPE array (the later call must see the same buffer)")
            check(_lib().emvs_mapper_prefetch_dsi(self._h, ptr(events), events.shape[0], ptr(trajectory.poses),
                                                  trajectory.poses.shape[0], T))
        # the engine recognises the later call by the arrays' addresses: keep them (and the poses) alive until then
        self.ctx._prefetch_ref = (events, trajectory)

    def evaluateDSI(self, events, trajectory, T_rv_w, allreduce=False, peer_reduce=False, accumulate=False):
        """bool evaluateDSI(events, trajectory, T_rv_w) — mapper_emvs_stereo.cpp:67-148.
        `events`: EVENT_DTYPE array (std::vector<dvs_msgs::Event>) or EventsSoA.
        allreduce / peer_reduce: the events are this rank's shard of a multi-GPU build; accumulate: vote on top of the
        current DSI instead of resetting it (EMVS_BUILD_ACCUMULATE, sub-interval style)."""
        T = ptr(np.ascontiguousarray(T_rv_w, dtype=POSE_DTYPE))
        flags = self._flags(accumulate, allreduce, peer_reduce)
        if isinstance(events, EventsSoA):
            es = events.c_struct()
            rc = _lib().emvs_mapper_evaluate_dsi_soa(self._h, C.byref(es), ptr(trajectory.poses), trajectory.poses.shape[0], T, flags)
        else:
            events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
            rc = _lib().emvs_mapper_evaluate_dsi_flags(self._h, ptr(events), events.shape[0], ptr(trajectory.poses),
                                                       trajectory.poses.shape[0], T, flags)
        self.ctx._prefetch_ref = None   # consumed or dropped by this call (emvs_b200.h: prefetch lifetime rule)
        if rc == capi.EMVS_ERR_TOO_FEW:
            return False
        check(rc)
        return True

    @staticmethod
    def _flags(accumulate, allreduce, peer_reduce=False):
        return ((capi.BUILD_ACCUMULATE if accumulate else capi.BUILD_RESET) | (capi.BUILD_ALLREDUCE if allreduce else 0)
                | (capi.BUILD_PEER_REDUCE if peer_reduce else 0))

    def build(self, events, packets, accumulate=False, allreduce=False, peer_reduce=False):
        """Event stage + reset + fillVoxelGrid (mapper_emvs_stereo.cpp:129-205) for given packets.
        allreduce=True: this is one rank's shard; Z-slabs are summed over the ranks while voting goes on."""
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        packets = np.ascontiguousarray(packets, dtype=PACKET_DTYPE)
        check(_lib().emvs_mapper_build(self._h, ptr(events), events.shape[0], ptr(packets), packets.shape[0],
                                       self._flags(accumulate, allreduce, peer_reduce)))
        self.ctx._prefetch_ref = None

    def build_device(self, d_events, n_events, d_packets, n_packets, accumulate=False, allreduce=False, peer_reduce=False):
        """Same with device pointers (ints); asynchronous on the context's stream."""
        check(_lib().emvs_mapper_build_device(self._h, C.c_void_p(d_events), int(n_events), C.c_void_p(d_packets),
                                              int(n_packets), self._flags(accumulate, allreduce, peer_reduce)))

    def depths_device_ptr(self):
        p = C.c_void_p()
        check(_lib().emvs_mapper_depths_device(self._h, C.byref(p)))
        return p.value

    def counts(self):
        out = np.zeros(self.dsi_shape_.dimZ_, np.uint64)
        check(_lib().emvs_mapper_counts(self._h, ptr(out)))
        return out

    def counts_allreduce(self):
        check(_lib().emvs_mapper_counts_allreduce(self._h))

    def getDepthMapFromDSI(self, options_depth_map=None):
        """Without options: the hot part of getDepthMapFromDSI (mapper_emvs_stereo.cpp:344-375 + :302-313,
        method=-1) -> (depth_map, confidence_map, depth_cell_indices) with depth = depths[raw argmax].
        With an OptionsDepthMap: the whole function minus the inpainting (:339-436)
        -> (depth_map, confidence_map, mask, depth_cell_indices_filtered)."""
        if options_depth_map is not None:
            return depth_map_from_dsi([self.dsi_], 6, self.raw_depths_vec_, options_depth_map)
        conf, idx, depth = self.dsi_.collapseMaxZSlice(self.raw_depths_vec_)
        return depth, conf, idx


# --------------------------------------------------------------------------------------------
# process_1 (Alg. 1: fusion across cameras) — process1.cpp:28-224, compute steps only
# --------------------------------------------------------------------------------------------
def process_1(mappers, events, trajectories, T_rv_w, fusion_method, mapper_fused=None):
    """Back-project every camera's events, fuse the DSIs, extract depth + confidence.

    mappers/events/trajectories: one entry per camera (2 or 3 in the reference; more is an
    extension).  Returns (depth_map, confidence_map, depth_cell_indices).  If ``mapper_fused`` is
    given its dsi_ receives the fused volume (the reference always materialises it)."""
    if fusion_method not in (1, 2, 3, 4, 5, 6):
        raise ValueError("Improper fusion method selected")  # process1.cpp:155-157
    for m, ev, tr in zip(mappers, events, trajectories):
        m.evaluateDSI(ev, tr, T_rv_w)
    grids = [m.dsi_ for m in mappers]
    if fusion_method in (3, 4, 5) and len(grids) == 3:
        grids = grids[:2]  # reference quirk: the third camera is ignored for GM/AM/RMS (process1.cpp:178-183)
    conf, idx, depth = fuse_collapse(grids, fusion_method, mappers[0].raw_depths_vec_,
                                     mapper_fused.dsi_ if mapper_fused is not None else None)
    return depth, conf, idx


# --------------------------------------------------------------------------------------------
# process_2 / process_5 (Alg. 2: fusion across cameras and time) — process2.cpp:28-302,
# process5.cpp:27-260, compute steps only
# --------------------------------------------------------------------------------------------
_PAIR_METHOD = {1: "minTwoGrids", 2: "harmonicMeanTwoGrids", 3: "geometricMeanTwoGrids", 4: "arithmeticMeanTwoGrids",
                5: "rmsTwoGrids", 6: "maxTwoGrids"}


def subinterval_slices(n_events, num_subintervals, shift=0):
    """Event ranges of the sub-intervals: split BY COUNT (process2.cpp:46-47, 105-107), the remainder
    n_events % num_subintervals is dropped.  shift > 0 (process_5, process5.cpp:89-150): start at
    sub-interval `shift` and wrap around the end of the stream.  -> list of (list of (lo, hi))."""
    per = n_events // num_subintervals
    out, first = [], shift * per
    for _ in range(num_subintervals):
        if shift and first + per >= n_events:                 # process5.cpp:136-143
            out.append([(first, n_events), (0, first + per - n_events)])
            first = first + per - n_events
        else:
            out.append([(first, first + per)])
            first += per
    return out


def process_2(ctx, cams, trajectories, events, dsi_shape, num_subintervals, T_rv_w, stereo_fusion, temporal_fusion,
              shuffle=False, replicate_camera_time_id_swap=False):
    """Alg. 2 on two cameras.  Per sub-interval k: evaluateDSI left / right, stereo-fuse, accumulate over
    time (temporal_fusion 2 = HM: sum of 1/(0.01+x) then n/sum; 4 = AM: sum then /n; any other id
    accumulates nothing, like the reference's empty switch cases).  Returns the four volumes the reference
    produces: dict(fused [camera then time], left, right [time per camera], camera_time [time then camera]).

    shuffle=True is process_5: the right camera starts at sub-interval n/2 and wraps around.
    The reference's time-then-camera switch maps ids 3 -> AM and 4 -> GM, swapped relative to every other
    place (process2.cpp:274-279); that is replicated only on request."""
    if stereo_fusion not in _PAIR_METHOD:
        raise ValueError("Improper fusion method selected")
    cam0, cam1 = cams
    mapper0, mapper1 = MapperEMVS(ctx, cam0, dsi_shape), MapperEMVS(ctx, cam1, dsi_shape)
    dimX, dimY, dimZ = mapper0.dsi_.size_
    fused_sub, fused, left, right, camera_time = (Grid3D(ctx, dimX, dimY, dimZ) for _ in range(5))
    ev = [np.ascontiguousarray(e, dtype=EVENT_DTYPE) for e in events]
    sl0 = subinterval_slices(len(ev[0]), num_subintervals)
    sl1 = subinterval_slices(len(ev[1]), num_subintervals, num_subintervals // 2 if shuffle else 0)
    for k in range(num_subintervals):
        for m, e, sl, tr in ((mapper0, ev[0], sl0[k], trajectories[0]), (mapper1, ev[1], sl1[k], trajectories[1])):
            subset = np.concatenate([e[lo:hi] for lo, hi in sl]) if len(sl) > 1 else e[sl[0][0]:sl[0][1]]
            if not m.evaluateDSI(subset, tr, T_rv_w):
                m.dsi_.resetGrid()      # the reference resets both DSIs at the top of every iteration (process2.cpp:100-101)
        fused_sub.copyFrom(mapper0.dsi_)                                   # resetGrid + addTwoGrids (process2.cpp:159-160)
        getattr(fused_sub, _PAIR_METHOD[stereo_fusion])(mapper1.dsi_)
        if temporal_fusion == 2:
            left.addInverseOfTwoGrids(mapper0.dsi_)
            right.addInverseOfTwoGrids(mapper1.dsi_)
            fused.addInverseOfTwoGrids(fused_sub)
        elif temporal_fusion == 4:
            left.addTwoGrids(mapper0.dsi_)
            right.addTwoGrids(mapper1.dsi_)
            fused.addTwoGrids(fused_sub)
    if temporal_fusion == 2:
        for g in (left, right, fused):
            g.computeHMfromSumOfInv(num_subintervals)
    elif temporal_fusion == 4:
        for g in (left, right, fused):
            g.computeAMfromSum(num_subintervals)
    camera_time.addTwoGrids(left)                                          # process2.cpp:267
    m = stereo_fusion
    if replicate_camera_time_id_swap and m in (3, 4):
        m = 7 - m
    getattr(camera_time, _PAIR_METHOD[m])(right)
    mapper0.close()
    mapper1.close()
    fused_sub.close()
    return dict(fused=fused, left=left, right=right, camera_time=camera_time)


def process_2_sharded(ctx, cams, trajectories, events, dsi_shape, T_rv_w, stereo_fusion, temporal_fusion, rank, world):
    """Alg. 2 with ONE sub-interval per GPU (BASELINE.json configs[3], the "AtHc" ordering): rank r builds the DSIs of
    both cameras for sub-interval r of `world`, fuses them across cameras locally, and the fusion across time is the
    only exchange — ONE ncclAllReduce(sum) of the camera-fused volume (of 1/(0.01 + x) for the harmonic temporal
    mean), followed by the local finalisation (/world, or world/sum).  That is exactly the `fused` output of
    process_2(num_subintervals=world) (process2.cpp:98-247) with the time loop spread over the ranks; float sums
    differ by the allreduce order only.  Needs ctx.comm_init; every rank returns the same Grid3D.

    Checked against process_2 by tests/mgpu_check_alg2.py (torchrun, 2 and 8 ranks)."""
    if stereo_fusion not in _PAIR_METHOD:
        raise ValueError("Improper fusion method selected")
    if temporal_fusion not in (2, 4):
        raise ValueError("process_2_sharded: temporal fusion 2 (harmonic) or 4 (arithmetic) — the other ids accumulate nothing")
    ev = [np.ascontiguousarray(e, dtype=EVENT_DTYPE) for e in events]
    mappers = [MapperEMVS(ctx, c, dsi_shape) for c in cams]
    for m, e, tr in zip(mappers, ev, trajectories):
        (lo, hi), = subinterval_slices(len(e), world)[rank]
        if not m.evaluateDSI(e[lo:hi], tr, T_rv_w):
            m.dsi_.resetGrid()
    dimX, dimY, dimZ = mappers[0].dsi_.size_
    fused_sub, fused = Grid3D(ctx, dimX, dimY, dimZ), Grid3D(ctx, dimX, dimY, dimZ)
    fused_sub.copyFrom(mappers[0].dsi_)
    getattr(fused_sub, _PAIR_METHOD[stereo_fusion])(mappers[1].dsi_)
    if temporal_fusion == 2:
        fused.addInverseOfTwoGrids(fused_sub)
        fused.allreduce()
        fused.computeHMfromSumOfInv(world)
    else:
        fused.addTwoGrids(fused_sub)
        fused.allreduce()
        fused.computeAMfromSum(world)
    for m in mappers:
        m.close()
    fused_sub.close()
    return fused
