#!/usr/bin/env python
"""Large DSIs: one vote launch per build (scratch for every slab, EMVS_MULTISLAB_BUDGET_MB) against the per-slab fallback.

Same events on two contexts of one process (the knobs are read when a context is created): accepted-vote counts must be
identical, the volumes equal up to float summation order; build time of each.

    python tools/multislab_large_check.py [--sizes 1024x1024x256,1024x1024x512] [--events 5000000] [--budget-mb 16384]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024x1024x256,1024x1024x512")
    ap.add_argument("--events", type=int, default=5_000_000)
    ap.add_argument("--budget-mb", type=int, default=16384)
    ap.add_argument("--deadline-s", type=float, default=80.0, help="do not start another size after this many seconds")
    a = ap.parse_args()
    t_start = time.time()
    import torch
    from dvs_mcemvs_b200 import api
    from sweep import scene_for
    os.environ["EMVS_MULTISLAB_BUDGET_MB"] = str(a.budget_mb)
    ctx_ms = api.Context(0)
    os.environ["EMVS_MULTISLAB_BUDGET_MB"] = "1"
    ctx_ps = api.Context(0)
    del os.environ["EMVS_MULTISLAB_BUDGET_MB"]
    for size in a.sizes.split(","):
        if time.time() - t_start > a.deadline_s:
            break
        W, H, Nz = (int(v) for v in size.split("x"))
        cam, shape, sc = scene_for(W, H, Nz)
        ev = sc.events(0, a.events)
        traj = api.LinearTrajectory(sc.trajectory(0))
        out = {"dsi": size, "events": a.events}
        vols = {}
        for name, ctx in (("multislab", ctx_ms), ("per_slab", ctx_ps)):
            m = api.MapperEMVS(ctx, cam, shape)
            pk = m.packetize(ev, traj, sc.T_rv_w())
            d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
            d_pk = torch.from_numpy(pk.view(np.uint8).reshape(-1)).cuda()
            torch.cuda.synchronize()
            t = ctx.timer()
            best = float("inf")
            for it in range(4):
                t.start()
                m.build_device(d_ev.data_ptr(), len(ev), d_pk.data_ptr(), len(pk))
                t.stop()
                ms = t.elapsed_ms()
                if it:
                    best = min(best, ms)
            ctx.sync()
            out[name + "_build_ms"] = round(best, 3)
            out[name + "_launches_per_build"] = None
            vols[name] = (m.counts().copy(), m.dsi_.computeMeanSquare(), m.dsi_.download() if Nz <= 256 else None)
            m.close()
            del d_ev, d_pk
        (c0, ms0, v0), (c1, ms1, v1) = vols["multislab"], vols["per_slab"]
        out["accepted_votes"] = int(c0.sum())
        out["counts_identical"] = bool(np.array_equal(c0, c1))
        out["mean_square_rel_diff"] = abs(ms0 - ms1) / max(abs(ms1), 1e-30)
        if v0 is not None:
            out["dsi_allclose_1e-5"] = bool(np.allclose(v0, v1, rtol=1e-5, atol=1e-5))
            out["dsi_max_abs_diff"] = float(np.abs(v0 - v1).max())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
