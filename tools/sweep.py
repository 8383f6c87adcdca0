#!/usr/bin/env python
"""BASELINE.json configs[4]: DSI-size x event-count sweep, on one GPU or — under torchrun — on N GPUs of one box.

Sensor == DSI x-y size, f = 0.8 W, identity LUT, one camera, structured events.

    python tools/sweep.py [--quick] > gpurun_out/sweep.md                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/sweep.py --no-cpu > gpurun_out/sweep_nN.md                 # N GPUs (strong scaling)

1 GPU: for every point build Mevents/s (event stage + reset + votes, inputs in HBM), Z-argmax ms, accepted votes and
the vote kernel's RED payload rate; the CPU oracle is timed beside the points that finish in a few seconds and
extrapolated linearly in the event count otherwise (marked with *).
N GPUs: the point's event list is split into N packet-aligned sub-intervals (one per GPU, shard.plan); a step is
{build of the rank's shard with the slab-wise NVLink peer reduce under the votes, arg-max of the rank's row band,
distribution of the maps} and is timed as the max over ranks — so the column is directly comparable with the 1-GPU
`build ms + argmax ms`.  Points above 10 M events tile the rank's 10 M-event sample with EMVS_BUILD_ACCUMULATE and use
the one-sweep peer exchange after the builds (the slab-wise reduce needs final slabs).  Rank 0 checks every point's
vote counts (sum over ranks) and maps against an unsharded build of the same events on its own GPU.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [(256, 256, 128), (512, 512, 256), (640, 480, 256), (1024, 1024, 256), (1024, 1024, 512)]
COUNTS = [1_000_000, 10_000_000, 100_000_000]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu-events", type=int, default=1_000_000)
    ap.add_argument("--sizes", default="", help="comma-separated WxHxZ list (default: the five sizes of configs[4])")
    ap.add_argument("--counts", default="", help="comma-separated event counts (default 1M,10M,100M)")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    sizes, counts = SIZES, COUNTS
    if a.quick:
        sizes, counts = sizes[:3], counts[:2]
    if a.sizes:
        sizes = [tuple(int(v) for v in t.split("x")) for t in a.sizes.split(",")]
    if a.counts:
        counts = [int(v) for v in a.counts.split(",")]
    return a, sizes, counts


def scene_for(W, H, Nz):
    from dvs_mcemvs_b200 import api, synth
    cam = api.CameraModel(W, H, 0.8 * W, 0.8 * W, W / 2.0, H / 2.0)
    shape = api.ShapeDSI(0, 0, Nz, 1.0, 10.0, 0.0)
    sc = synth.Scene(synth.Rig([cam], [0.0]), shape, duration=0.2, translation=(0.2, 0.0, 0.0), rot_deg=1.0, seed=5)
    return cam, shape, sc


def red_ceiling():
    p = os.path.join(ROOT, "profiles", "red_port_ceiling.json")
    return float(json.load(open(p))["payload_gb_per_s"]) if os.path.exists(p) else float("nan")


def single_gpu(a, sizes, counts):
    import torch
    from dvs_mcemvs_b200 import api
    from oracle import oracle as O
    ceil = red_ceiling()
    ctx = api.Context(0)
    print("| DSI | events | build ms | build Mev/s | argmax ms | accepted votes | RED payload GB/s | frac of %.0f GB/s RED ceiling | CPU oracle Mev/s (%d thr) |"
          % (ceil, O.num_threads()))
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (W, H, Nz) in sizes:
        cam, shape, sc = scene_for(W, H, Nz)
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
            payload = votes * 16.0 / (vote_ms * 1e-3) / 1e9
            star = "" if n <= n_cpu else "*"
            print(f"| {W}x{H}x{Nz} | {n:,} | {build_ms:.2f} | {n / build_ms / 1e3:.1f} | {argmax_ms:.3f} | {votes:,} | "
                  f"{payload:.0f} | {payload / ceil:.2f} | {cpu_mevs:.2f}{star} |", flush=True)
            del d_ev, d_pk
        m.close()
    print("\n`*` = CPU figure measured on %d events and constant in the event count (the vote loop is linear)." % a.cpu_events)


def multi_gpu(a, sizes, counts):
    import torch
    import torch.distributed as dist
    from dvs_mcemvs_b200 import api, shard
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)

    def allgather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def say(s):
        if rank == 0:
            print(s, flush=True)

    say(f"| DSI | events | GPUs | step ms (build + exchange + argmax, max over ranks) | Mev/s | accepted votes | sharded == unsharded |")
    say("|---|---:|---:|---:|---:|---:|---|")
    for (W, H, Nz) in sizes:
        cam, shape, sc = scene_for(W, H, Nz)
        base_n = min(max(counts), 10_000_000)
        ev_base = sc.events(0, base_n)                   # same seed on every rank: identical lists
        traj = api.LinearTrajectory(sc.trajectory(0))
        m = api.MapperEMVS(ctx, cam, shape)
        ex = api.PeerExchange(ctx, [m.dsi_], world, rank, allgather)
        for n in counts:
            if n <= base_n:
                ev = ev_base[:: base_n // n][:n].copy()
                reps = 1
            else:
                ev, reps = ev_base, n // base_n
            pk = m.packetize(ev, traj, sc.T_rv_w())
            # this rank's share: packet range `rank` of the list (tiled lists: of every tile)
            (_, lo, hi), = shard.plan([len(pk)], world, rank)
            mine = pk[lo:hi].copy()
            e_lo = int(mine["first_event"][0]) if len(mine) else 0
            e_hi = int(mine["first_event"][-1]) + 1024 if len(mine) else 0
            mine["first_event"] -= e_lo
            d_ev = torch.from_numpy(ev[e_lo:e_hi].view(np.uint8).reshape(-1).copy()).cuda()
            d_pk = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).cuda()
            torch.cuda.synchronize()
            t = ctx.timer()
            step_ms = 0.0
            for it in range(2):
                dist.barrier()
                torch.cuda.synchronize()
                t.start()
                if reps == 1:
                    ex.begin()
                    m.build_device(d_ev.data_ptr(), e_hi - e_lo, d_pk.data_ptr(), len(mine), peer_reduce=True)
                else:
                    for r in range(reps):
                        m.build_device(d_ev.data_ptr(), e_hi - e_lo, d_pk.data_ptr(), len(mine), accumulate=r > 0)
                ex.fuse_collapse(6, m.depths_device_ptr())
                t.stop()
                ctx.sync()
                step_ms = t.elapsed_ms()
            tt = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            step_ms = float(tt.item())
            conf, idx, depth = ex.download()
            cnt = torch.from_numpy(m.counts().astype(np.int64)).cuda()
            dist.all_reduce(cnt)
            ok = ""
            if rank == 0:
                full = api.MapperEMVS(ctx, cam, shape)
                d_ev_f = torch.from_numpy(ev.view(np.uint8).reshape(-1).copy()).cuda()
                d_pk_f = torch.from_numpy(pk.view(np.uint8).reshape(-1).copy()).cuda()
                torch.cuda.synchronize()
                for r in range(reps):
                    full.build_device(d_ev_f.data_ptr(), len(ev), d_pk_f.data_ptr(), len(pk), accumulate=r > 0)
                conf_f, idx_f, depth_f = full.dsi_.collapseMaxZSlice(full.raw_depths_vec_)
                counts_ok = np.array_equal(cnt.cpu().numpy().astype(np.uint64), full.counts())
                conf_ok = np.allclose(conf, conf_f, rtol=1e-4, atol=1e-6)
                agree = float((idx == idx_f).mean())
                ok = f"counts {'exact' if counts_ok else 'DIFFER'}, conf {'<= 1e-4' if conf_ok else 'DIFFERS'}, idx {agree:.5f}"
                votes = int(full.counts().sum())
                full.close()
                del d_ev_f, d_pk_f
                print(f"| {W}x{H}x{Nz} | {n:,} | {world} | {step_ms:.2f} | {n / step_ms / 1e3:.1f} | {votes:,} | {ok} |", flush=True)
            del d_ev, d_pk
            dist.barrier()
        ctx.sync()
        dist.barrier()
        ex.close()
        m.close()
    ctx.sync()
    dist.barrier()
    dist.destroy_process_group()


def main():
    a, sizes, counts = parse()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        multi_gpu(a, sizes, counts)
    else:
        single_gpu(a, sizes, counts)


if __name__ == "__main__":
    main()
