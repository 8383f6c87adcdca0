"""oracle/parity.py — the verdict bench.py's `parity` block and the full-size GPU tests rely on — on hand-made maps."""
import numpy as np

from oracle import parity as P


def _case():
    rng = np.random.default_rng(0)
    fused = rng.random((16, 6, 7)).astype(np.float32)
    idx = fused.argmax(0)
    conf = fused.max(0)
    depths = np.linspace(1, 2, 16, dtype=np.float32)
    return fused, conf, idx, depths[idx]


def test_identical_maps_pass():
    fused, conf, idx, depth = _case()
    p = P.compare_maps(conf, idx, depth, conf, idx, depth, fused)
    p["counts_exact"] = True
    assert p["idx_agreement"] == 1.0 and p["conf_max_err_over_tol"] == 0.0 and P.verdict(p)


def test_confidence_outside_tolerance_fails():
    fused, conf, idx, depth = _case()
    bad = conf.copy()
    bad[2, 3] *= np.float32(1.001)
    p = P.compare_maps(bad, idx, depth, conf, idx, depth, fused)
    assert p["conf_max_err_over_tol"] > 1 and not P.verdict(p)


def test_index_flip_only_allowed_at_near_ties():
    fused, conf, idx, depth = _case()
    # a genuine near-tie: second-best plane within 1e-5 relative of the best -> a flip there is float-sum noise
    k2 = (idx[1, 1] + 3) % 16
    fused[k2, 1, 1] = conf[1, 1] * np.float32(1 - 1e-5)
    flip = idx.copy()
    flip[1, 1] = k2
    conf2 = conf.copy()
    conf2[1, 1] = fused[k2, 1, 1]
    dep2 = depth.copy()
    dep2[1, 1] = 0
    p = P.compare_maps(conf2, flip, dep2, conf, idx, depth, fused)
    assert p["idx_mismatches"] == 1 and p["idx_mismatches_are_near_ties"] and P.verdict(p)
    # a flip to a plane that is NOT close to the maximum is a real error
    flip[4, 4] = int(fused[:, 4, 4].argmin())
    p = P.compare_maps(conf2, flip, dep2, conf, idx, depth, fused)
    assert not p["idx_mismatches_are_near_ties"] and not P.verdict(p)


def test_counts_depth_and_checksum_criteria():
    fused, conf, idx, depth = _case()
    p = P.compare_maps(conf, idx, depth, conf, idx, depth, fused)
    assert not P.verdict(dict(p, counts_exact=False))
    assert not P.verdict(dict(p, mean_square_rel=1e-3))
    d2 = depth.copy()
    d2[0, 0] += 1
    assert not P.verdict(P.compare_maps(conf, idx, d2, conf, idx, depth, fused))
