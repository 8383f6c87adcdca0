"""On-disk formats of the mapping path, so that the reference's own scripts
(scripts/visualize_dsi_*.py, evaluate_mcemvs_dsec.py) read this engine's output unchanged.

  write_grid_npy      Grid3D::writeGridNpy — cartesian3dgrid_IO.cpp:30-36 (cnpy: float32, C order, shape {Z, Y, X})
  save_depth_points   the `depth_points_<suffix>.txt` part of saveDepthMaps — utils.cpp:29-44
  save_depth_maps     saveDepthMaps — utils.cpp:22-117 (txt + negated-confidence PNG + JET inverse-depth PNG, via cv2)
  accumulate_events   accumulateEvents — utils.cpp:180-216 (the per-sub-interval event images of process_2 / process_5)
"""
import numpy as np


def write_grid_npy(grid_or_volume, filename):
    """grid_or_volume: api.Grid3D (downloaded here) or an ndarray [dimZ, dimY, dimX]."""
    vol = grid_or_volume.download() if hasattr(grid_or_volume, "download") else np.asarray(grid_or_volume)
    vol = np.ascontiguousarray(vol, dtype="<f4")
    assert vol.ndim == 3
    with open(filename, "wb") as f:      # not np.save(path): cnpy never appends ".npy" to the name
        np.save(f, vol, allow_pickle=False)
    return 0


def _ostream_float(v):
    """operator<<(std::ostream&, float) with default flags: %g with 6 significant digits."""
    return "%g" % float(v)


def save_depth_points(depth_map, semidense_mask, filename):
    """One line `col row depth` per pixel of the mask, row-major order (utils.cpp:35-42)."""
    depth_map = np.asarray(depth_map, np.float32)
    rows, cols = np.nonzero(np.asarray(semidense_mask) > 0)
    with open(filename, "w") as f:
        for r, c in zip(rows, cols):
            f.write(f"{c} {r} {_ostream_float(depth_map[r, c])}\n")
    return len(rows)


def _cv2():
    try:
        import cv2
    except ImportError as e:   # the PNG outputs are produced with the very OpenCV calls the reference makes
        raise RuntimeError("save_depth_maps / accumulate_events need OpenCV (cv2) for the PNG encoders and colour maps") from e
    return cv2


def inverse_depth_u8(depth_map, min_depth, max_depth):
    """The 8-bit inverse-depth image of utils.cpp:84-89: (1/d - 1/max) / (1/min - 1/max) * 255 evaluated as ONE
    scaled conversion in float32 (that is what the cv::MatExpr chain collapses to), then cvRound + saturate."""
    inv = np.float32(1.0) / np.asarray(depth_map, np.float32)
    mod_max = float(np.float32(max_depth))
    alpha = 255.0 / (1.0 / float(np.float32(min_depth)) - 1.0 / mod_max)
    beta = -(1.0 / mod_max) * alpha
    v = inv * np.float32(alpha) + np.float32(beta)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def save_depth_maps(depth_map, confidence_map, semidense_mask, min_depth, max_depth, suffix, out_path):
    """saveDepthMaps (utils.cpp:22-117): the three files the reference writes per depth map,
         <out_path>depth_points_<suffix>.txt               `col row depth` per mask pixel
         <out_path>confidence_map_negated_<suffix>.png     255 - minmax-normalised confidence
         <out_path>inv_depth_colored_dilated_<suffix>.png  JET-coloured inverse depth on black, dilated 3x3 ellipse
    out_path is a PREFIX (the reference concatenates strings, process1.cpp:193-200).  Returns the three paths."""
    cv2 = _cv2()
    depth_map = np.ascontiguousarray(depth_map, np.float32)
    confidence_map = np.ascontiguousarray(confidence_map, np.float32)
    mask = np.ascontiguousarray(semidense_mask, np.uint8)
    p_txt = f"{out_path}depth_points_{suffix}.txt"
    save_depth_points(depth_map, mask, p_txt)
    conf255 = cv2.normalize(confidence_map, None, 0, 255.0, cv2.NORM_MINMAX, cv2.CV_32FC1)
    p_conf = f"{out_path}confidence_map_negated_{suffix}.png"
    cv2.imwrite(p_conf, 255 - conf255)               # float image: imwrite converts to 8 bit (cvRound, saturate)
    color = cv2.applyColorMap(inverse_depth_u8(depth_map, min_depth, max_depth), cv2.COLORMAP_JET)
    canvas = np.zeros(depth_map.shape + (3,), np.uint8)
    canvas[mask > 0] = color[mask > 0]
    canvas = cv2.dilate(canvas, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3)))
    p_inv = f"{out_path}inv_depth_colored_dilated_{suffix}.png"
    cv2.imwrite(p_inv, canvas)
    return p_txt, p_conf, p_inv


def accumulate_events(events, use_polarity, height, width):
    """accumulateEvents (utils.cpp:180-216): the event image process_2 / process_5 save per sub-interval.
    With polarity: sum of +-1 per pixel scaled so that the largest magnitude maps to 128 around a mid-grey of 128;
    without: per-pixel counts (8-bit wrap-around like the reference's uchar +=) min-max normalised to [0, 255]."""
    ev = np.asarray(events)
    x, y = ev["x"].astype(np.int64), ev["y"].astype(np.int64)
    if use_polarity:
        imgf = np.zeros((height, width), np.float32)
        np.add.at(imgf, (y, x), np.where(ev["polarity"] != 0, 1.0, -1.0).astype(np.float32))
        half_range = max(abs(float(imgf.min())), abs(float(imgf.max())))
        if half_range <= 0:
            return np.full((height, width), 128, np.uint8)
        v = imgf * np.float32(128 / half_range) + np.float32(128)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    img = np.zeros((height, width), np.uint8)
    np.add.at(img, (y, x), np.uint8(1))
    return _cv2().normalize(img, None, 0, 255, _cv2().NORM_MINMAX, _cv2().CV_8UC1)
