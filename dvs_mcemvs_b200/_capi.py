"""ctypes binding of include/emvs_b200.h (the C-ABI of libemvs_b200.so).

This is the same stub a maintainer of the reference would write to call the engine from
Python; the C++ host classes in ``dvs_mcemvs_b200/host`` bind the same symbols.  The library
is loaded from ``dvs_mcemvs_b200/lib`` (built in-tree by ``__graft_entry__.build()`` /
``make -C dvs_mcemvs_b200/csrc``).  There is no fallback: a missing library raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libemvs_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "emvs_b200.h")

EMVS_OK, EMVS_ERR_INVALID, EMVS_ERR_CUDA, EMVS_ERR_TOO_FEW, EMVS_ERR_NCCL, EMVS_ERR_STATE = range(6)
PACKET_SIZE = 1024

FUSE_MIN, FUSE_HM, FUSE_GM, FUSE_AM, FUSE_RMS, FUSE_MAX = 1, 2, 3, 4, 5, 6
(OP_ADD, OP_MIN, OP_HM, OP_GM, OP_AM, OP_RMS, OP_MAX, OP_HM_N, OP_ADD_INV, OP_HM_FROM_SUMINV,
 OP_AM_FROM_SUM) = range(11)
BUILD_RESET, BUILD_ACCUMULATE, BUILD_ALLREDUCE, BUILD_PEER_REDUCE = 0, 1, 2, 4
DISTORTION_NONE, DISTORTION_PLUMB_BOB, DISTORTION_FISHEYE = 0, 1, 2

# numpy views of the POD structs (layouts asserted against the C side in tests/test_abi.py)
EVENT_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"),
                        ("polarity", "u1"), ("pad", "u1", (3,))], align=False)
POSE_DTYPE = np.dtype([("q", "<f8", (4,)), ("t", "<f8", (3,))], align=False)
STAMPED_POSE_DTYPE = np.dtype([("sec", "<u4"), ("nsec", "<u4"), ("T", POSE_DTYPE)], align=False)
PACKET_DTYPE = np.dtype([("H", "<f4", (9,)), ("C", "<f4", (3,)), ("first_event", "<u8")], align=False)
assert EVENT_DTYPE.itemsize == 16 and POSE_DTYPE.itemsize == 56
assert STAMPED_POSE_DTYPE.itemsize == 64 and PACKET_DTYPE.itemsize == 56


class Shape(C.Structure):
    _fields_ = [("dimX", C.c_uint32), ("dimY", C.c_uint32), ("dimZ", C.c_uint32),
                ("min_depth", C.c_float), ("max_depth", C.c_float), ("fov_deg", C.c_float),
                ("inverse_depth", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class EventsSoA(C.Structure):
    """emvs_events_soa: separate x / y (uint16) and t_ns (int64) arrays of one event list."""
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("t_ns", C.c_void_p), ("n", C.c_size_t)]


class DepthMapOptions(C.Structure):
    _fields_ = [("adaptive_threshold_kernel_size", C.c_int32), ("adaptive_threshold_c", C.c_double),
                ("max_confidence", C.c_double), ("median_filter_size", C.c_int32)]


class EmvsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"emvs error {code}: {msg}")
        self.code = code


_vp, _sz = C.c_void_p, C.c_size_t
_PROTOTYPES = {
    "emvs_abi_version": (C.c_int, []),
    "emvs_last_error": (C.c_char_p, []),
    "emvs_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "emvs_context_destroy": (C.c_int, [_vp]),
    "emvs_context_sync": (C.c_int, [_vp]),
    "emvs_context_set_slab": (C.c_int, [_vp, C.c_uint32]),
    "emvs_context_set_upload_split": (C.c_int, [_vp, C.c_uint32, C.c_uint64]),
    "emvs_context_prefetch_events": (C.c_int, [_vp, _vp, _sz]),
    "emvs_context_prefetch_pending": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "emvs_context_prefetch_cancel": (C.c_int, [_vp]),
    "emvs_mapper_prefetch_dsi": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "emvs_mapper_prefetch_dsi_soa": (C.c_int, [_vp, C.POINTER(EventsSoA), _vp, _sz, _vp]),
    "emvs_mapper_evaluate_dsi_soa": (C.c_int, [_vp, C.POINTER(EventsSoA), _vp, _sz, _vp, C.c_int]),
    "emvs_packetize_soa": (C.c_int, [C.POINTER(EventsSoA), _vp, _sz, _vp, C.POINTER(Camera), _vp, C.c_float, _vp, _sz,
                                     C.POINTER(_sz)]),
    "emvs_selftest_division": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "emvs_context_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "emvs_context_profile_vote": (C.c_int, [_vp, C.c_int]),
    "emvs_context_vote_time": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "emvs_context_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "emvs_timer_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "emvs_timer_destroy": (C.c_int, [_vp]),
    "emvs_timer_start": (C.c_int, [_vp]),
    "emvs_timer_stop": (C.c_int, [_vp]),
    "emvs_timer_elapsed_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "emvs_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "emvs_host_free": (C.c_int, [_vp]),
    "emvs_depth_vector": (C.c_int, [C.POINTER(Shape), _vp]),
    "emvs_virtual_camera": (C.c_int, [C.POINTER(Camera), C.POINTER(Shape), _vp]),
    "emvs_rectify_lut": (C.c_int, [C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "emvs_trajectory_pose_at": (C.c_int, [_vp, _sz, C.c_uint32, C.c_uint32, _vp, C.POINTER(C.c_int)]),
    "emvs_pose_compose": (C.c_int, [_vp, _vp, _vp]),
    "emvs_pose_inverse": (C.c_int, [_vp, _vp]),
    "emvs_packetize": (C.c_int, [_vp, _sz, _vp, _sz, _vp, C.POINTER(Camera), _vp, C.c_float, _vp, _sz,
                                 C.POINTER(_sz)]),
    "emvs_packetize_range": (C.c_int, [_vp, _sz, _vp, _sz, _vp, C.POINTER(Camera), _vp, C.c_float, C.POINTER(_sz), _sz,
                                       _vp, _sz, C.POINTER(_sz)]),
    "emvs_grid_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "emvs_grid_destroy": (C.c_int, [_vp]),
    "emvs_grid_dims": (C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "emvs_grid_reset": (C.c_int, [_vp]),
    "emvs_grid_op": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float]),
    "emvs_grid_copy": (C.c_int, [_vp, _vp]),
    "emvs_grid_download": (C.c_int, [_vp, _vp]),
    "emvs_grid_upload": (C.c_int, [_vp, _vp]),
    "emvs_grid_mean_square": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "emvs_grid_collapse_max": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "emvs_fuse_collapse": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "emvs_fuse_collapse_device": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "emvs_depth_map_from_dsi": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.POINTER(DepthMapOptions), _vp, _vp, _vp, _vp]),
    "emvs_depth_map_postprocess": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32,
                                             C.POINTER(DepthMapOptions), _vp, _vp, _vp, _vp, _vp]),
    "emvs_grid_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "emvs_mapper_create": (C.c_int, [_vp, C.POINTER(Camera), C.POINTER(Shape), C.POINTER(_vp)]),
    "emvs_mapper_destroy": (C.c_int, [_vp]),
    "emvs_mapper_set_lut": (C.c_int, [_vp, _vp, _sz]),
    "emvs_mapper_shape": (C.c_int, [_vp, C.POINTER(Shape), _vp]),
    "emvs_mapper_depths": (C.c_int, [_vp, _vp]),
    "emvs_mapper_depths_device": (C.c_int, [_vp, C.POINTER(_vp)]),
    "emvs_mapper_grid": (C.c_int, [_vp, C.POINTER(_vp)]),
    "emvs_mapper_build": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.c_int]),
    "emvs_mapper_build_device": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.c_int]),
    "emvs_mapper_evaluate_dsi": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "emvs_mapper_evaluate_dsi_flags": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, C.c_int]),
    "emvs_mapper_counts": (C.c_int, [_vp, _vp]),
    "emvs_comm_unique_id": (C.c_int, [_vp]),
    "emvs_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "emvs_comm_destroy": (C.c_int, [_vp]),
    "emvs_grid_allreduce": (C.c_int, [_vp]),
    "emvs_grid_allreduce_async": (C.c_int, [_vp]),
    "emvs_mapper_counts_allreduce": (C.c_int, [_vp]),
    "emvs_exchange_create": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "emvs_exchange_destroy": (C.c_int, [_vp]),
    "emvs_exchange_blob_bytes": (C.c_int, [_vp, C.POINTER(_sz)]),
    "emvs_exchange_export": (C.c_int, [_vp, _vp]),
    "emvs_exchange_import": (C.c_int, [_vp, _vp]),
    "emvs_exchange_set_participants": (C.c_int, [_vp, _vp]),
    "emvs_exchange_begin": (C.c_int, [_vp]),
    "emvs_exchange_fuse_collapse": (C.c_int, [_vp, C.c_int, _vp]),
    "emvs_exchange_maps": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "emvs_exchange_download": (C.c_int, [_vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load libemvs_b200.so and declare every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("EMVS_B200_FORBID_LOAD"):   # set by tests to prove that a code path (bench.py's CPU arm) never maps the library
        raise ImportError("libemvs_b200.so must not be loaded in this process (EMVS_B200_FORBID_LOAD is set)")
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dvs_mcemvs_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != EMVS_OK:
        raise EmvsError(code, load().emvs_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
