#!/usr/bin/env python
"""does the HOST packet stage stay off the critical path at 100 M events without the prefetch API?

One camera, 640x480x256, `--events` uniform events over a 0.2 s window.  Measures (host wall clock, best of 3):
  packetize_ms        emvs_packetize alone (one pose interpolation + one 3x3 inverse per 1024 events, multi-threaded)
  upload_ms           cudaMemcpy of the event list from pinned memory (what the packet stage runs under)
  evaluate_stock_ms   MapperEMVS.evaluateDSI(events, trajectory, T_rv_w) + sync: upload, packet stage, event stage, votes
  build_resident_ms   the same build with events and packets already in HBM (device-timed)
so that   evaluate_stock - build_resident   is everything the host side adds, to be compared with upload_ms."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=int, default=100_000_000)
    a = ap.parse_args()
    import torch
    from dvs_mcemvs_b200 import api, synth
    sc, _, _, _ = synth.config("dsec_stereo", events_per_cam=1024)
    cam = sc.rig.cams[0]
    n = a.events
    rng = np.random.default_rng(7)
    ev = api.pinned_empty(n, api.EVENT_DTYPE)
    t = np.linspace(sc.t0, sc.t0 + sc.T, n)          # sorted timestamps over the window
    ev["sec"], ev["nsec"] = synth._split_time(t)
    del t
    ev["x"] = rng.integers(0, cam.width, n, dtype=np.uint16)
    ev["y"] = rng.integers(0, cam.height, n, dtype=np.uint16)
    traj = api.LinearTrajectory(sc.trajectory(0))
    T = sc.T_rv_w()
    ctx = api.Context(0)
    ctx.set_upload_split(0)
    m = api.MapperEMVS(ctx, cam, sc.shape)

    def best(f, reps=3):
        out = []
        for _ in range(reps):
            t0 = time.perf_counter()
            f()
            out.append((time.perf_counter() - t0) * 1e3)
        return min(out)

    pk = m.packetize(ev, traj, T)
    packetize_ms = best(lambda: m.packetize(ev, traj, T))
    d_ev = torch.empty(ev.nbytes, dtype=torch.uint8, device="cuda")
    h_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1))

    def upload():
        d_ev.copy_(h_ev, non_blocking=True)
        torch.cuda.synchronize()
    upload()
    upload_ms = best(upload)

    def stock():
        assert m.evaluateDSI(ev, traj, T)
        ctx.sync()
    stock()
    evaluate_stock_ms = best(stock)
    ctx.set_upload_split(15)
    stock()
    evaluate_split_ms = best(stock)
    d_pk = torch.from_numpy(pk.view(np.uint8).reshape(-1).copy()).cuda()
    torch.cuda.synchronize()
    tm = ctx.timer()
    res = []
    for _ in range(3):
        tm.start()
        m.build_device(d_ev.data_ptr(), n, d_pk.data_ptr(), len(pk))
        tm.stop()
        res.append(tm.elapsed_ms())
    build_resident_ms = min(res)
    print(json.dumps({"events": n, "packets": int(len(pk)), "host_threads": os.environ.get("EMVS_HOST_THREADS", "default (<= 4)"),
                      "packetize_ms": round(packetize_ms, 2), "upload_ms": round(upload_ms, 2),
                      "evaluate_stock_ms": round(evaluate_stock_ms, 2), "evaluate_stock_split15_ms": round(evaluate_split_ms, 2),
                      "build_resident_ms": round(build_resident_ms, 2),
                      "host_side_adds_ms": round(evaluate_stock_ms - build_resident_ms, 2)}))


if __name__ == "__main__":
    main()
