"""ctypes front-end of the CPU oracle (oracle/emvs_oracle.cpp) + a numpy twin for small cases.

TEST INFRASTRUCTURE ONLY — see the header of emvs_oracle.cpp.  May be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
package.  Pinning: the Grid3D layer (bilinear vote, all voxel ops, collapseMaxZSlice,
computeMeanSquare), the depth tables, the masked median and the OpenCV post-processing arithmetic are
pinned bit-exactly against the reference's own sources compiled in place (oracle/_ref) and against
cv2 (tests/golden/*.npz, tests/test_oracle_pinned.py, tests/test_depthmap_post.py).  PARITY UNPINNED
for the packet / event stages of mapper_emvs_stereo.cpp (Eigen + minkindr arithmetic: the mapper
cannot be compiled here); those are held by the known-answer tests in tests/test_oracle_kat.py and by
agreement between the two independent statements (C++ loops here vs. numpy below).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libemvs_oracle.so")

EVENT_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"),
                        ("polarity", "u1"), ("pad", "u1", (3,))])
POSE_DTYPE = np.dtype([("q", "<f8", (4,)), ("t", "<f8", (3,))])
PACKET_DTYPE = np.dtype([("H", "<f4", (9,)), ("C", "<f4", (3,)), ("first_event", "<u8")])
PACKET_SIZE = 1024

(OP_ADD, OP_MIN, OP_HM, OP_GM, OP_AM, OP_RMS, OP_MAX, OP_HM_N, OP_ADD_INV, OP_HM_FROM_SUMINV,
 OP_AM_FROM_SUM) = range(11)

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "emvs_oracle.cpp")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_float
        L.oracle_depth_vector.argtypes = [C.c_int, f32, f32, u64, vp]
        L.oracle_virtual_focal.argtypes = [f32, u64, f32]
        L.oracle_virtual_focal.restype = f32
        L.oracle_pinhole_kinv.argtypes = [f32, f32, f32, f32, vp]
        L.oracle_mat3_inv.argtypes = [vp, vp]
        L.oracle_mat3_mul.argtypes = [vp, vp, vp]
        L.oracle_pose_mul.argtypes = [vp, vp, vp]
        L.oracle_pose_inv.argtypes = [vp, vp]
        L.oracle_pose_at.argtypes = [vp, vp, vp, u64, u32, u32, vp]
        L.oracle_pose_at.restype = C.c_int
        L.oracle_packetize.argtypes = [vp, u64, vp, vp, vp, u64, vp, vp, vp, f32, vp, u64]
        L.oracle_packetize.restype = u64
        L.oracle_warp_events.argtypes = [vp, vp, u64, vp, C.c_int, vp]
        L.oracle_fill_voxel_grid.argtypes = [vp, vp, u64, vp, u64, vp, u32, u32, vp, vp]
        L.oracle_build_dsi.argtypes = [vp, vp, u64, vp, C.c_int, vp, u64, vp, u32, u32, vp, vp]
        L.oracle_vote.argtypes = [f32, f32, vp, u32, u32]
        L.oracle_vote.restype = C.c_int
        L.oracle_fuse_op.argtypes = [C.c_int, vp, vp, u64, C.c_int, f32, C.c_int]
        L.oracle_fuse_nary.argtypes = [C.c_int, vp, C.c_int, u64, vp]
        L.oracle_collapse_max.argtypes = [vp, u32, u32, u32, vp, vp, vp, vp]
        L.oracle_mean_square.argtypes = [vp, u64]
        L.oracle_mean_square.restype = C.c_double
        L.oracle_num_threads.restype = C.c_int
        L.oracle_depth_map_post.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp, vp, vp, vp]
        L.oracle_depth_map_post.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---- thin wrappers ---------------------------------------------------------------------------
def depth_vector(zmin, zmax, nz, inverse=False):
    out = np.zeros(nz, np.float32)
    lib().oracle_depth_vector(int(inverse), zmin, zmax, nz, _p(out))
    return out


def virtual_camera(cam_fx, cam_cx, cam_cy, dimX, fov_deg):
    f = lib().oracle_virtual_focal(cam_fx, dimX, fov_deg)
    return np.array([f, f, cam_cx, cam_cy], np.float32)


def pose_at(traj, sec, nsec):
    """traj: structured array with fields sec, nsec, T (pose)."""
    tsec, tnsec = _c(traj["sec"], np.uint32), _c(traj["nsec"], np.uint32)
    poses = _c(traj["T"], POSE_DTYPE)
    out = np.zeros((), POSE_DTYPE)
    ok = lib().oracle_pose_at(_p(tsec), _p(tnsec), _p(poses), poses.shape[0], int(sec), int(nsec), _p(out))
    return out if ok else None


def pose_mul(a, b):
    out = np.zeros((), POSE_DTYPE)
    lib().oracle_pose_mul(_p(_c(a, POSE_DTYPE)), _p(_c(b, POSE_DTYPE)), _p(out))
    return out


def pose_inv(a):
    out = np.zeros((), POSE_DTYPE)
    lib().oracle_pose_inv(_p(_c(a, POSE_DTYPE)), _p(out))
    return out


def packetize(events, traj, T_rv_w, K4, virt4, z0):
    events = _c(events, EVENT_DTYPE)
    tsec, tnsec = _c(traj["sec"], np.uint32), _c(traj["nsec"], np.uint32)
    poses = _c(traj["T"], POSE_DTYPE)
    out = np.zeros(events.shape[0] // PACKET_SIZE + 1, PACKET_DTYPE)
    n = lib().oracle_packetize(_p(events), events.shape[0], _p(tsec), _p(tnsec), _p(poses), poses.shape[0],
                               _p(_c(T_rv_w, POSE_DTYPE)), _p(_c(K4, np.float32)), _p(_c(virt4, np.float32)),
                               float(z0), _p(out), out.shape[0])
    return out[:n]


def warp_events(events, packets, lut, W):
    events, packets = _c(events, EVENT_DTYPE), _c(packets, PACKET_DTYPE)
    out = np.zeros((packets.shape[0] * PACKET_SIZE, 2), np.float32)
    lib().oracle_warp_events(_p(events), _p(packets), packets.shape[0], _p(_c(lut, np.float32)), int(W), _p(out))
    return out


def build_dsi(events, packets, lut, W, depths, virt4, dimX, dimY):
    """-> (dsi [nz, dimY, dimX] float32, inb [nz] uint64)"""
    events, packets = _c(events, EVENT_DTYPE), _c(packets, PACKET_DTYPE)
    depths = _c(depths, np.float32)
    nz = depths.shape[0]
    dsi = np.zeros((nz, dimY, dimX), np.float32)
    inb = np.zeros(nz, np.uint64)
    lib().oracle_build_dsi(_p(events), _p(packets), packets.shape[0], _p(_c(lut, np.float32)), int(W), _p(depths), nz,
                           _p(_c(virt4, np.float32)), dimX, dimY, _p(dsi), _p(inb))
    return dsi, inb


def fuse_op(op, a, b=None, n=0, eps=0.0, copy_arg=False):
    """In-place a = op(a, b) on a float32 array; returns a."""
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    bb = None if b is None else _c(b, np.float32)
    lib().oracle_fuse_op(int(op), _p(a), _p(bb), a.size, int(n), float(eps), int(copy_arg))
    return a


def fuse_reference(method, vols, copy_arg=False):
    """process_1 step 2 (process1.cpp:126-191): fused = 0 + dsi0; op(fused, dsi1); [op3(fused, dsi2)]."""
    fused = np.zeros_like(vols[0])
    fuse_op(OP_ADD, fused, vols[0], copy_arg=copy_arg)
    pair = {1: OP_MIN, 2: OP_HM, 3: OP_GM, 4: OP_AM, 5: OP_RMS, 6: OP_MAX}[method]
    fuse_op(pair, fused, vols[1], n=2, eps=0.1, copy_arg=copy_arg)
    if len(vols) > 2:
        if method == 1:
            fuse_op(OP_MIN, fused, vols[2], copy_arg=copy_arg)
        elif method == 2:
            fuse_op(OP_HM_N, fused, vols[2], n=3, eps=0.1, copy_arg=copy_arg)
        elif method == 6:
            fuse_op(OP_MAX, fused, vols[2], copy_arg=copy_arg)
        # 3, 4, 5: the reference ignores the third camera (process1.cpp:178-183)
    return fused


def fuse_nary(method, vols):
    """EXTENSION (not in the reference): n-ary fusion as defined in emvs_kernels.cuh::fuse_voxel."""
    vols = [_c(v, np.float32) for v in vols]
    ptrs = (C.c_void_p * len(vols))(*[v.ctypes.data for v in vols])
    out = np.zeros_like(vols[0])
    lib().oracle_fuse_nary(int(method), ptrs, len(vols), out.size, _p(out))
    return out


def collapse_max(dsi, depths=None):
    dsi = _c(dsi, np.float32)
    nz, dimY, dimX = dsi.shape
    conf = np.zeros((dimY, dimX), np.float32)
    idx = np.zeros((dimY, dimX), np.uint16)
    depth = np.zeros((dimY, dimX), np.float32) if depths is not None else None
    d = _c(depths, np.float32) if depths is not None else None
    lib().oracle_collapse_max(_p(dsi), dimX, dimY, nz, _p(d), _p(conf), _p(idx), _p(depth))
    return (conf, idx) if depth is None else (conf, idx, depth)


def depth_map_post(conf, idx, depths, ks=5, c=5.0, max_confidence=0.0, median_size=5):
    """getDepthMapFromDSI after the collapse, without inpainting (mapper_emvs_stereo.cpp:393-436).
    -> dict(conf [with conf(0,0) = max_confidence, as the reference leaves it], conf8, mask, idx_filtered, depth)"""
    conf = np.array(conf, np.float32, copy=True, order="C")
    idx = _c(idx, np.uint8)
    rows, cols = conf.shape
    depths = _c(depths, np.float32)
    out = dict(conf=conf, conf8=np.zeros((rows, cols), np.uint8), mask=np.zeros((rows, cols), np.uint8),
               idx_filtered=np.zeros((rows, cols), np.uint8), depth=np.zeros((rows, cols), np.float32))
    rc = lib().oracle_depth_map_post(_p(conf), _p(idx), rows, cols, int(ks), float(c), float(max_confidence), int(median_size),
                                     _p(depths), _p(out["conf8"]), _p(out["mask"]), _p(out["idx_filtered"]), _p(out["depth"]))
    if rc != 0:
        raise ValueError("unsupported adaptive_threshold_kernel_size / median_filter_size")
    return out


def mean_square(dsi):
    dsi = _c(dsi, np.float32)
    return lib().oracle_mean_square(_p(dsi), dsi.size)


def num_threads():
    return lib().oracle_num_threads()


def use_all_host_threads():
    """OpenMP team size = the CPUs this process may run on (overrides an OMP_NUM_THREADS=1 set by torchrun)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(int(n))
    return num_threads()


# ---- numpy twin (independent second statement, small cases only) -------------------------------
def np_build_dsi(xy0, packets, depths, virt4, dimX, dimY):
    """fillVoxelGrid (mapper_emvs_stereo.cpp:151-205) + vote (cartesian3dgrid.h:253-273) with numpy
    float32 scalar semantics.  Votes are accumulated in event order per voxel (np.add.at is
    sequential), so the result equals the C++ loops bit for bit."""
    f32 = np.float32
    depths = np.asarray(depths, f32)
    nz = depths.shape[0]
    z0 = depths[0]
    vfx, vfy, vcx, vcy = [f32(v) for v in virt4]
    n_pk = packets.shape[0]
    X0 = xy0[:, 0].reshape(n_pk, PACKET_SIZE)
    Y0 = xy0[:, 1].reshape(n_pk, PACKET_SIZE)
    Cx, Cy, Cz = [packets["C"][:, i].astype(f32)[:, None] for i in range(3)]
    dsi = np.zeros((nz, dimY * dimX), f32)
    inb = np.zeros(nz, np.uint64)
    with np.errstate(all="ignore"):
        for k in range(nz):
            zi = depths[k]
            a = z0 * (zi - Cz)
            bx = (z0 - zi) * (Cx * vfx + Cz * vcx)
            by = (z0 - zi) * (Cy * vfy + Cz * vcy)
            d = zi * (z0 - Cz)
            X = ((X0 * a + bx) / d).ravel()
            Y = ((Y0 * a + by) / d).ravel()
            ok = (X >= 0) & (Y >= 0) & (X < 2147483648.0) & (Y < 2147483648.0)
            xi = np.where(ok, X, 0).astype(np.int64)
            yi = np.where(ok, Y, 0).astype(np.int64)
            ok &= (xi + 1 < dimX) & (yi + 1 < dimY)
            X, Y, xi, yi = X[ok], Y[ok], xi[ok], yi[ok]
            fx = X - xi.astype(f32)
            fy = Y - yi.astype(f32)
            fx1, fy1 = f32(1) - fx, f32(1) - fy
            base = xi + yi * dimX
            # one voxel can receive several of an event's four contributions only across
            # events, never within one event, so four sequential add.at passes reorder
            # nothing per voxel relative to the scalar loop ... except between the 4 taps of
            # different events; do it strictly in event order instead:
            idx = np.stack([base, base + 1, base + dimX, base + dimX + 1], axis=1).ravel()
            w = np.stack([fx1 * fy1, fx * fy1, fx1 * fy, fx * fy], axis=1).ravel().astype(f32)
            np.add.at(dsi[k], idx, w)
            inb[k] = ok.sum()
    return dsi.reshape(nz, dimY, dimX), inb


def np_collapse_max(dsi):
    idx = np.argmax(dsi, axis=0)  # first maximum
    conf = np.take_along_axis(dsi, idx[None], axis=0)[0]
    return conf, idx.astype(np.uint16)
