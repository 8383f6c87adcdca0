"""ctypes front-end of oracle/_ref/libgrid3d_ref.so — the REFERENCE's own Grid3D
(cartesian3dgrid.h/.cpp) and depth_vector.hpp compiled in place from /root/reference by
`make -C oracle ref` (only a cv::Mat container and a glog stand-in are ours, oracle/shim/).

TEST INFRASTRUCTURE ONLY: used to pin the restated oracle (tests/test_oracle_pinned.py) and to
generate tests/golden/ (tests/golden/make_golden.py).  The .so is git-ignored and travels to the
GPU box with the snapshot; /root/reference itself is never read at test time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgrid3d_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def build():
    """(Re)build when the reference sources are present; otherwise keep the prebuilt library."""
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return available()


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float
        L.ref_vote.argtypes = [u32, u32, u32, vp, u32, vp, vp, u64]
        L.ref_grid_op.argtypes = [C.c_int, u32, u32, u32, vp, vp, C.c_int, f32]
        L.ref_collapse_max.argtypes = [u32, u32, u32, vp, vp, vp]
        L.ref_mean_square.argtypes = [u32, u32, u32, vp]
        L.ref_mean_square.restype = C.c_double
        L.ref_depth_vector.argtypes = [C.c_int, f32, f32, u64, vp]
        L.ref_huang_median.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.ref_time_fuse_collapse.argtypes = [u32, u32, u32, vp, vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp]
        L.ref_time_fuse_collapse.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def vote(vol, k, x, y):
    """accumulateGridValueAt for every (x[i], y[i]) on slice k; vol [dimZ, dimY, dimX] float32, in place."""
    assert vol.dtype == np.float32 and vol.flags["C_CONTIGUOUS"]
    dimZ, dimY, dimX = vol.shape
    x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
    lib().ref_vote(dimX, dimY, dimZ, _p(vol), k, _p(x), _p(y), x.shape[0])
    return vol


def grid_op(op, a, b=None, n=0, eps=0.0):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    dimZ, dimY, dimX = a.shape
    bb = None if b is None else np.ascontiguousarray(b, np.float32)
    lib().ref_grid_op(int(op), dimX, dimY, dimZ, _p(a), _p(bb), int(n), float(eps))
    return a


def collapse_max(vol):
    vol = np.ascontiguousarray(vol, np.float32)
    dimZ, dimY, dimX = vol.shape
    assert dimZ <= 256
    conf = np.zeros((dimY, dimX), np.float32)
    idx = np.zeros((dimY, dimX), np.uint8)
    lib().ref_collapse_max(dimX, dimY, dimZ, _p(vol), _p(conf), _p(idx))
    return conf, idx


def mean_square(vol):
    vol = np.ascontiguousarray(vol, np.float32)
    dimZ, dimY, dimX = vol.shape
    return lib().ref_mean_square(dimX, dimY, dimZ, _p(vol))


def depth_vector(zmin, zmax, nz, inverse=False):
    out = np.zeros(nz, np.float32)
    lib().ref_depth_vector(int(inverse), zmin, zmax, nz, _p(out))
    return out


def huang_median(img, mask, patch_size):
    """The reference's huangMedianFilter (median_filtering.cpp:33-158) on uint8 images."""
    img, mask = np.ascontiguousarray(img, np.uint8), np.ascontiguousarray(mask, np.uint8)
    out = np.zeros_like(img)
    lib().ref_huang_median(_p(img), _p(mask), img.shape[0], img.shape[1], int(patch_size), _p(out))
    return out


def time_fuse_collapse(a, b, method):
    """Wall-clock ms of the reference's own fusion (process1.cpp:126-166) and collapseMaxZSlice on two volumes
    -> (fuse_ms, argmax_ms, conf, idx)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    dimZ, dimY, dimX = a.shape
    conf = np.zeros((dimY, dimX), np.float32)
    idx = np.zeros((dimY, dimX), np.uint8)
    f, m = C.c_double(0), C.c_double(0)
    rc = lib().ref_time_fuse_collapse(dimX, dimY, dimZ, _p(a), _p(b), int(method), C.byref(f), C.byref(m), _p(conf), _p(idx))
    if rc:
        raise ValueError("Improper fusion method selected")
    return f.value, m.value, conf, idx
