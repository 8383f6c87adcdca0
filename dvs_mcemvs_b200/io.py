"""On-disk formats of the mapping path, so that the reference's own scripts
(scripts/visualize_dsi_*.py, evaluate_mcemvs_dsec.py) read this engine's output unchanged.

  write_grid_npy      Grid3D::writeGridNpy — cartesian3dgrid_IO.cpp:30-36 (cnpy: float32, C order, shape {Z, Y, X})
  save_depth_points   the `depth_points_<suffix>.txt` part of saveDepthMaps — utils.cpp:29-44
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
