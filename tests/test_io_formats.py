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


def test_save_depth_maps_files(tmp_path):
    cv2 = __import__("pytest").importorskip("cv2")
    rng = np.random.default_rng(3)
    H, W = 24, 32
    depth = (1.0 + 4.0 * rng.random((H, W))).astype(np.float32)
    conf = (50 * rng.random((H, W))).astype(np.float32)
    mask = np.zeros((H, W), np.uint8)
    mask[10, 12] = 1                      # one isolated pixel: the 3x3 ellipse dilates it to a plus
    mask[3:5, 20:22] = 1
    prefix = str(tmp_path) + "/000.1_"
    p_txt, p_conf, p_inv = io.save_depth_maps(depth, conf, mask, 1.0, 5.0, "fused_2", prefix)
    assert p_txt.endswith("000.1_depth_points_fused_2.txt") and p_conf.endswith("000.1_confidence_map_negated_fused_2.png")
    assert p_inv.endswith("000.1_inv_depth_colored_dilated_fused_2.png")
    assert len(open(p_txt).read().splitlines()) == 5
    neg = cv2.imread(p_conf, cv2.IMREAD_UNCHANGED)
    assert neg.shape == (H, W) and neg.dtype == np.uint8
    want = 255.0 - (conf - conf.min()) * (255.0 / (conf.max() - conf.min()))
    assert np.abs(neg.astype(np.float64) - want).max() <= 0.51           # 8-bit rounding of the float image
    assert neg[np.unravel_index(conf.argmax(), conf.shape)] == 0 and neg[np.unravel_index(conf.argmin(), conf.shape)] == 255
    col = cv2.imread(p_inv, cv2.IMREAD_UNCHANGED)
    assert col.shape == (H, W, 3)
    lit = col.any(axis=2)
    plus = np.zeros((H, W), bool)
    plus[10, 11:14] = plus[9:12, 12] = True
    assert np.array_equal(lit[8:13, 10:15], plus[8:13, 10:15])            # MORPH_ELLIPSE 3x3 is a plus
    assert not lit[0, 0] and lit.sum() == 5 + 12                          # plus (5) + dilated 2x2 block (12)
    # the colour of the isolated pixel is JET of its inverse-depth byte
    u8 = io.inverse_depth_u8(depth, 1.0, 5.0)
    inv = (1.0 / depth[10, 12] - 1.0 / 5.0) / (1.0 - 1.0 / 5.0) * 255.0
    assert abs(int(u8[10, 12]) - inv) <= 0.51
    assert np.array_equal(col[10, 12], cv2.applyColorMap(u8[10:11, 12:13], cv2.COLORMAP_JET)[0, 0])
    assert io.inverse_depth_u8(np.array([[1.0, 5.0]], np.float32), 1.0, 5.0).tolist() == [[255, 0]]


def test_accumulate_events():
    from dvs_mcemvs_b200 import _capi as capi
    ev = np.zeros(7, capi.EVENT_DTYPE)
    ev["x"] = [1, 1, 1, 2, 2, 0, 3]
    ev["y"] = [0, 0, 0, 1, 1, 1, 1]
    ev["polarity"] = [1, 1, 1, 0, 0, 1, 0]
    img = io.accumulate_events(ev, True, 2, 4)
    # sums: (0,1)=+3, (1,2)=-2, (1,0)=+1, (1,3)=-1; half range 3 -> value*128/3 + 128, rounded half to even
    assert img.tolist() == [[128, 255, 128, 128], [171, 128, 43, 85]]
    assert io.accumulate_events(ev[:0], True, 2, 4).tolist() == [[128] * 4] * 2
    cnt = io.accumulate_events(ev, False, 2, 4)
    assert cnt.dtype == np.uint8 and cnt[0, 1] == 255 and cnt[0, 0] == 0 and cnt[1, 2] == 170 and cnt[1, 0] == 85
