#!/usr/bin/env python
"""BASELINE.json configs[4]: DSI-size x event-count sweep on one GPU (per-GPU numbers; multi-GPU
scaling is bench.py --gpus N).  Sensor == DSI x-y size, f = 0.8 W, identity LUT, one camera,
structured events.  For every point: build Mevents/s (event stage +
reset + votes, inputs in HBM), Z-argmax ms, accepted votes, and the k_vote algorithmic-bandwidth
fraction; the CPU oracle is timed beside the points that finish in a few seconds and extrapolated
linearly in the event count otherwise (marked with *).

    python tools/sweep.py [--quick] > gpurun_out/sweep.md
"""
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
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu-events", type=int, default=1_000_000)
    ap.add_argument("--sizes", default="", help="comma-separated WxHxZ list (default: the five sizes of configs[4])")
    ap.add_argument("--counts", default="", help="comma-separated event counts (default 1M,10M,100M)")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from dvs_mcemvs_b200 import api, synth
    from oracle import oracle as O

    sizes = [(256, 256, 128), (512, 512, 256), (640, 480, 256), (1024, 1024, 256), (1024, 1024, 512)]
    counts = [1_000_000, 10_000_000, 100_000_000]
    if a.quick:
        sizes, counts = sizes[:3], counts[:2]
    if a.sizes:
        sizes = [tuple(int(v) for v in t.split("x")) for t in a.sizes.split(",")]
    if a.counts:
        counts = [int(v) for v in a.counts.split(",")]
    peak = 6548.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    ctx = api.Context(0)
    print("| DSI | events | build ms | build Mev/s | argmax ms | accepted votes | k_vote alg. GB/s | frac of %.0f GB/s | CPU oracle Mev/s (%d thr) |"
          % (peak, O.num_threads()))
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (W, H, Nz) in sizes:
        cam = api.CameraModel(W, H, 0.8 * W, 0.8 * W, W / 2.0, H / 2.0)
        rig = synth.Rig([cam], [0.0])
        shape = api.ShapeDSI(0, 0, Nz, 1.0, 10.0, 0.0)
        sc = synth.Scene(rig, shape, duration=0.2, translation=(0.2, 0.0, 0.0), rot_deg=1.0, seed=5)
        base_n = min(max(counts), 10_000_000)
        ev_base = sc.events(0, base_n)
        traj = api.LinearTrajectory(sc.trajectory(0))
        m = api.MapperEMVS(ctx, cam, shape)
        d_conf = torch.empty(W * H, dtype=torch.float32, device="cuda")
        d_depth = torch.empty(W * H, dtype=torch.float32, device="cuda")
        d_idx = torch.empty(W * H, dtype=torch.uint8 if Nz <= 256 else torch.int16, device="cuda")
        # CPU oracle on a bounded sample (linear in the event count)
        n_cpu = min(a.cpu_events, base_n)
        ev_cpu = ev_base[:: base_n // n_cpu][:n_cpu].copy()
        pk_cpu = m.packetize(ev_cpu, traj, sc.T_rv_w())
        cpu_mevs = float("nan")
        if not a.no_cpu:
            t0 = time.perf_counter()
            O.build_dsi(ev_cpu, pk_cpu, cam.lut, W, m.raw_depths_vec_, m.virtual_cam_, W, H)
            cpu_mevs = n_cpu / (time.perf_counter() - t0) / 1e6
        for n in counts:
            if n <= base_n:
                ev = ev_base[:: base_n // n][:n].copy()   # same window, thinned
                pk = m.packetize(ev, traj, sc.T_rv_w())
                reps = 1
            else:                                        # tile the 10 M stream: same per-event work
                ev, pk, reps = ev_base, m.packetize(ev_base, traj, sc.T_rv_w()), n // base_n
            d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
            d_pk = torch.from_numpy(pk.view(np.uint8).reshape(-1)).cuda()
            torch.cuda.synchronize()
            t = ctx.timer()
            for it in range(2):                          # warm-up + timed
                ctx.profile_vote(True)
                t.start()
                for r in range(reps):
                    m.build_device(d_ev.data_ptr(), len(ev), d_pk.data_ptr(), len(pk), accumulate=r > 0)
                t.stop()
                build_ms = t.elapsed_ms()
                vote_ms, _ = ctx.vote_time()
                ctx.profile_vote(False)
            votes = int(m.counts().sum())
            for it in range(2):                          # warm-up (first call allocates the chunk scratch) + timed
                t.start()
                api.fuse_collapse_device([m.dsi_], 6, m.depths_device_ptr(), d_conf.data_ptr(), d_idx.data_ptr(), d_depth.data_ptr())
                t.stop()
                argmax_ms = t.elapsed_ms()
            alg = (votes * 32.0 + len(pk) * 1024 * reps * 8.0) / (vote_ms * 1e-3) / 1e9
            star = "" if n <= n_cpu else "*"
            print(f"| {W}x{H}x{Nz} | {n:,} | {build_ms:.2f} | {n / build_ms / 1e3:.1f} | {argmax_ms:.3f} | {votes:,} | "
                  f"{alg:.0f} | {alg / peak:.2f} | {cpu_mevs:.2f}{star} |", flush=True)
            del d_ev, d_pk
        m.close()
    print("\n`*` = CPU figure measured on %d events and constant in the event count (the vote loop is linear)." % a.cpu_events)


if __name__ == "__main__":
    main()
