"""Map / count comparison used by bench.py's `parity` block and by the full-size GPU tests.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the CUDA path's outputs go IN, a verdict comes out;
nothing here is on the product path.

What "parity" means for this path:
  * per-plane accepted-vote counts are integers and must be bit-exact;
  * the DSI is a float sum whose order differs (CPU: event order; GPU: atomic order), so fused confidence agrees to
    the north-star tolerance (1e-4 relative, plus 1e-6 absolute for values near zero);
  * the arg-max index may differ only where the oracle's own fused column has a near-tie (two planes within the
    tolerance of each other); where the indices agree the depth is the same table entry, hence bit-equal;
  * computeMeanSquare (the reference's run-time checksum, process1.cpp:86, cartesian3dgrid.cpp:164-174) agrees to
    float-sum tolerance.
"""
import numpy as np

CONF_RTOL, CONF_ATOL = 1e-4, 1e-6
MS_RTOL = 1e-5


def compare_maps(conf, idx, depth, conf_o, idx_o, depth_o, fused_o=None):
    """GPU maps (conf, idx, depth) against reference maps (conf_o, idx_o, depth_o).  fused_o: the reference's fused
    volume [Z, Y, X] (optional) to classify index mismatches as near-ties."""
    conf, conf_o = np.asarray(conf, np.float32), np.asarray(conf_o, np.float32)
    idx, idx_o = np.asarray(idx).astype(np.int64), np.asarray(idx_o).astype(np.int64)
    err = np.abs(conf.astype(np.float64) - conf_o.astype(np.float64))
    tol = CONF_ATOL + CONF_RTOL * np.abs(conf_o.astype(np.float64))
    big = np.abs(conf_o) > 1e-3
    same = idx == idx_o
    out = {
        "conf_max_rel": float((err[big] / np.abs(conf_o[big])).max()) if big.any() else 0.0,
        "conf_max_err_over_tol": float((err / tol).max()),
        "idx_agreement": float(same.mean()),
        "idx_mismatches": int((~same).sum()),
        "depth_exact_where_idx_agrees": bool(np.array_equal(np.asarray(depth)[same], np.asarray(depth_o)[same])),
    }
    if fused_o is not None and (~same).any():
        ys, xs = np.nonzero(~same)
        at_gpu = fused_o[idx[ys, xs], ys, xs].astype(np.float64)
        top = conf_o[ys, xs].astype(np.float64)
        out["idx_mismatches_are_near_ties"] = bool(np.all(top - at_gpu <= 2 * (CONF_ATOL + CONF_RTOL * np.abs(top))))
    else:
        out["idx_mismatches_are_near_ties"] = True if fused_o is not None or same.all() else None
    return out


def verdict(p):
    """ok iff every criterion present in the block passes."""
    ok = p.get("counts_exact", True) is True
    ok = ok and p.get("conf_max_err_over_tol", 0.0) <= 1.0
    ok = ok and p.get("depth_exact_where_idx_agrees", True) is True
    nt = p.get("idx_mismatches_are_near_ties", True)
    if nt is None:                       # no fused reference volume: fall back to an agreement threshold
        nt = p.get("idx_agreement", 1.0) > 0.995
    ok = ok and bool(nt)
    ok = ok and p.get("mean_square_rel", 0.0) <= MS_RTOL
    return bool(ok)
