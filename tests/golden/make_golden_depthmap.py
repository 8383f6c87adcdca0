"""Generates tests/golden/depthmap_ref.npz: the depth-map post-processing of
MapperEMVS::getDepthMapFromDSI (mapper_emvs_stereo.cpp:393-436, without the Telea inpainting)
computed with the SAME library calls as the reference — OpenCV's normalize / adaptiveThreshold
(Python cv2 of this image) and the reference's own huangMedianFilter compiled in place
(oracle/_ref, `make -C oracle ref`).  Inputs: confidence / index maps of the seeded esim_small
case (harmonic fusion of the oracle DSIs).

    python tests/golden/make_golden_depthmap.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [dict(ks=5, c=5.0, max_confidence=0.0, median_size=5),      # main.cpp defaults
         dict(ks=5, c=14.0, max_confidence=33.4, median_size=9),     # typical cfg/ values
         dict(ks=3, c=4.5, max_confidence=0.0, median_size=7),
         dict(ks=7, c=7.0, max_confidence=2.0, median_size=3)]       # max_confidence below the map maximum


def reference_pipeline(conf, idx, depths, ks, c, max_confidence, median_size):
    conf = conf.copy()
    conf[0, 0] = max_confidence                                               # :396
    conf_norm = cv2.normalize(conf, None, 0.0, 255.0, cv2.NORM_MINMAX)        # :397
    conf_norm[0, 0] = 0                                                       # :399
    conf8 = cv2.convertScaleAbs(conf_norm)                                    # :400 convertTo(CV_8U); values are >= 0
    mask = cv2.adaptiveThreshold(conf8, 1, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, ks, -c)   # :405-411
    idx_f = ref.huang_median(idx, mask, median_size)                          # :419-423
    b = max(ks // 2, 1)                                                       # :426-427
    rows, cols = mask.shape
    yy, xx = np.mgrid[0:rows, 0:cols]
    mask = mask.copy()
    mask[(xx <= b) | (xx >= cols - b) | (yy <= b) | (yy >= rows - b)] = 0
    return dict(conf=conf, conf8=conf8, mask=mask, idx_filtered=idx_f, depth=depths[idx_f])   # :435


def main():
    from conftest import Case
    assert ref.build()
    case = Case("esim_small")
    vols = [case.oracle_dsi(i)[0] for i in range(2)]
    conf, idx, _ = O.collapse_max(O.fuse_reference(2, vols), case.depths)
    idx = idx.astype(np.uint8)
    out = dict(conf=conf, idx=idx, depths=case.depths, n_cases=np.array(len(CASES)))
    for i, kw in enumerate(CASES):
        r = reference_pipeline(conf, idx, case.depths, **kw)
        out[f"case{i}_params"] = np.array([kw["ks"], kw["c"], kw["max_confidence"], kw["median_size"]], np.float64)
        for k in ("conf8", "mask", "idx_filtered"):   # conf and depth follow from the inputs: conf(0,0) = max_confidence, depth = depths[idx_filtered]
            out[f"case{i}_{k}"] = r[k]
        print(i, kw, "mask density", float(r["mask"].mean()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "depthmap_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
