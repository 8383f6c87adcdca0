"""Generates tests/golden/grid3d_ref.npz from the REFERENCE's own code (oracle/_ref/libgrid3d_ref.so,
built in place from /root/reference by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are seeded; outputs are whatever the reference's Grid3D / depth_vector.hpp produce.  The
fixtures pin (a) the restated oracle on CPU and (b) the CUDA path on the GPU box, where
/root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OPS = {"add": 0, "min": 1, "hm": 2, "gm": 3, "am": 4, "rms": 5, "max": 6, "hm_n": 7, "add_inv": 8,
       "hm_from_suminv": 9, "am_from_sum": 10}


def vote_points(rng, dimX, dimY, n):
    """Random sub-pixel positions plus every boundary class of accumulateGridValueAt."""
    x = rng.uniform(-3, dimX + 3, n).astype(np.float32)
    y = rng.uniform(-3, dimY + 3, n).astype(np.float32)
    special = [(0.0, 0.0), (dimX - 1.0, 5.0), (dimX - 2.0, 5.0), (np.nextafter(np.float32(dimX - 1), np.float32(0)), 4.5),
               (5.0, dimY - 1.0), (5.0, dimY - 2.0), (4.25, np.nextafter(np.float32(dimY - 1), np.float32(0))),
               (-0.0, 3.0), (-1e-30, 3.0), (np.nan, 1.0), (1.0, np.nan), (np.inf, 1.0), (1.0, -np.inf),
               (3e9, 2.0), (2.0, 3e9), (2147483648.0, 1.0), (1e-40, 1e-40), (dimX - 1.5, dimY - 1.5),
               (0.5, 0.5), (7.0, 9.0), (7.999999, 9.000001)]
    sx = np.array([s[0] for s in special], np.float32)
    sy = np.array([s[1] for s in special], np.float32)
    return np.concatenate([x, sx]), np.concatenate([y, sy])


def main():
    assert ref.build(), "reference library could not be built (is /root/reference present?)"
    rng = np.random.default_rng(20221001)
    out = {}
    # --- bilinear vote (cartesian3dgrid.h:253-273) ------------------------------------------
    dimX, dimY, dimZ = 37, 23, 3
    x, y = vote_points(rng, dimX, dimY, 6000)
    vol = np.zeros((dimZ, dimY, dimX), np.float32)
    with np.errstate(all="ignore"):
        ref.vote(vol, 1, x, y)
    out["vote_x"], out["vote_y"], out["vote_dims"], out["vote_out"] = x, y, np.array([dimX, dimY, dimZ]), vol
    # --- voxel-wise ops (cartesian3dgrid.h:64-192) -------------------------------------------
    shape = (4, 21, 29)
    a = rng.gamma(1.0, 3.0, shape).astype(np.float32)
    b = rng.gamma(1.0, 3.0, shape).astype(np.float32)
    a[rng.random(shape) < 0.3] = 0.0
    b[rng.random(shape) < 0.3] = 0.0
    a[0, 0, :4] = [0.0, 1e-20, 3e19, 65504.0]
    b[0, 0, :4] = [0.0, 1e-20, 3e19, 1e-3]
    out["op_a"], out["op_b"] = a, b
    for name, op in OPS.items():
        kw = {"hm": dict(eps=0.1), "hm_n": dict(n=3, eps=0.1), "add_inv": dict(eps=1e-2), "hm_from_suminv": dict(n=5),
              "am_from_sum": dict(n=5)}.get(name, {})
        with np.errstate(all="ignore"):
            out["op_" + name] = ref.grid_op(op, a.copy(), None if op >= 9 else b, **kw)
    # --- collapseMaxZSlice (cartesian3dgrid.cpp:115-137) ---------------------------------------
    c = np.where(rng.random((40, 9, 11)) < 0.6, 0, rng.gamma(1, 4, (40, 9, 11))).astype(np.float32)
    c[:, 0, 0] = 0.0                      # all-zero column -> (0, index 0)
    c[5, 1, 1] = c[17, 1, 1] = 99.0       # tie -> first
    c[:, 2, 2] = -1.0
    c[33, 2, 2] = -0.5                    # negative column
    c[39, 3, 3] = 1e9                     # last plane
    conf, idx = ref.collapse_max(c)
    out["collapse_in"], out["collapse_conf"], out["collapse_idx"] = c, conf, idx
    out["mean_square"] = np.array([ref.mean_square(c), ref.mean_square(a)])
    # --- depth tables (depth_vector.hpp:88-103, 131-148) ---------------------------------------
    cases = [(0, 1.0, 5.0, 64), (0, 4.0, 200.0, 256), (0, 0.45, 3.1, 100), (0, 10.0, 1.0, 7), (1, 1.0, 5.0, 64),
             (1, 4.0, 200.0, 256), (1, 0.3, 12.5, 100), (0, 1.0, 10.0, 512), (0, 2.0, 2.5, 1)]
    out["depth_cases"] = np.array(cases, np.float64)
    for i, (inv, zmin, zmax, nz) in enumerate(cases):
        out[f"depth_{i}"] = ref.depth_vector(zmin, zmax, nz, bool(inv))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grid3d_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {len(out)} arrays")


if __name__ == "__main__":
    main()
