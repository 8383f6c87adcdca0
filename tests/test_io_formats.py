"""On-disk formats: the DSI .npy layout the reference's viewers expect and the
depth_points text file."""
import numpy as np

from dvs_mcemvs_b200 import io


def test_grid_npy_matches_cnpy_layout(tmp_path):
    rng = np.random.default_rng(0)
    vol = rng.random((5, 4, 3)).astype(np.float32)            # [Z, Y, X]
    p = tmp_path / "dsi_fused.npy"
    io.write_grid_npy(vol, str(p))
    raw = p.read_bytes()
    assert raw[:6] == b"\x93NUMPY" and b"'descr': '<f4'" in raw[:128] and b"'fortran_order': False" in raw[:128]
    assert b"(5, 4, 3)" in raw[:128]
    back = np.load(str(p))
    assert back.dtype == np.float32 and np.array_equal(back, vol)
    # element (x, y, z) sits at x + dimX*(y + dimY*z) of the flat payload (cartesian3dgrid.h:34-35)
    flat = np.frombuffer(raw[-vol.nbytes:], np.float32)
    assert flat[1 + 3 * (2 + 4 * 3)] == vol[3, 2, 1]


def test_depth_points_txt(tmp_path):
    depth = np.array([[1.5, 2.25, 3.0], [4.123456789, 100.0, 1e-3]], np.float32)
    mask = np.array([[1, 0, 1], [1, 0, 0]], np.uint8)
    p = tmp_path / "depth_points_0.txt"
    assert io.save_depth_points(depth, mask, str(p)) == 3
    assert p.read_text() == "0 0 1.5\n2 0 3\n0 1 4.12346\n"
