#!/usr/bin/env python
"""A/B of the engine's tuning / experiment knobs on ONE generated workload (BASELINE.json configs[1] by default).

Every variant is a set of EMVS_* environment variables read when a context is created; per variant a fresh context
runs `--warmup` + `--steps` device-resident steps (both cameras' builds + fuse + argmax, the `value` loop of bench.py)
and one JSON line is printed: ms per step, Mevents/s, average vote-launch ms, depth-map ms.

    python tools/ab_bench.py [--events-per-cam N] [--steps K] [--warmup W] [--variants name=K1:V1,K2:V2;name2=...]

The EMVS_DEBUG_SKIP_* variants produce WRONG volumes: they exist to measure what the merge / re-zero traffic that runs
beside the votes costs (profiles/r2_interference.md)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_VARIANTS = [
    ("default", {}),
    ("classic", {"EMVS_VOTE_KERNEL": "classic"}),
    ("classic_memset", {"EMVS_VOTE_KERNEL": "classic", "EMVS_ZERO_CTAS": "0"}),
    ("tma_8cta", {"EMVS_VOTE_CTAS_PER_SM": "8"}),
    ("tma_6cta", {"EMVS_VOTE_CTAS_PER_SM": "6"}),
    ("tma_5cta", {"EMVS_VOTE_CTAS_PER_SM": "5"}),
    ("zero_memset", {"EMVS_ZERO_CTAS": "0"}),
    ("zero_148", {"EMVS_ZERO_CTAS": "148"}),
    ("zero_592", {"EMVS_ZERO_CTAS": "592"}),
    ("slab8", {"EMVS_SLAB": "8"}),
    ("slab24", {"EMVS_SLAB": "24"}),
    ("no_overlap", {"EMVS_OVERLAP": "0"}),
    ("skip_zero", {"EMVS_DEBUG_SKIP_ZERO": "1"}),
    ("skip_merge", {"EMVS_DEBUG_SKIP_MERGE": "1"}),
    ("skip_both", {"EMVS_DEBUG_SKIP_ZERO": "1", "EMVS_DEBUG_SKIP_MERGE": "1"}),
    ("fc_v4", {"EMVS_FC_V4": "1"}),
    ("fc_v4_z8", {"EMVS_FC_V4": "1", "EMVS_FC_ZSPLIT": "8"}),
    ("fc_z8", {"EMVS_FC_ZSPLIT": "8"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dsec_stereo")
    ap.add_argument("--events-per-cam", type=int, default=5_000_000)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", default="")
    a = ap.parse_args()
    import torch
    from dvs_mcemvs_b200 import api, synth

    variants = DEFAULT_VARIANTS
    if a.variants:
        variants = []
        for item in a.variants.split(";"):
            name, _, kv = item.partition("=")
            variants.append((name, dict(p.split(":") for p in kv.split(",") if p)))
    sc, _, method, _ = synth.config(a.workload, events_per_cam=a.events_per_cam)
    cams = sc.rig.cams
    events = [sc.events(i, a.events_per_cam) for i in range(len(cams))]
    trajs = [api.LinearTrajectory(sc.trajectory(i)) for i in range(len(cams))]
    T = sc.T_rv_w()
    d_events = [torch.from_numpy(ev.view(np.uint8).reshape(-1).copy()).cuda() for ev in events]
    base_env = {k: os.environ.get(k) for _, env in variants for k in env}
    for name, env in variants:
        for k, v in base_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        os.environ.update(env)
        ctx = api.Context(0)
        mappers = [api.MapperEMVS(ctx, c, sc.shape) for c in cams]
        packets = [m.packetize(ev, tr, T) for m, ev, tr in zip(mappers, events, trajs)]
        d_packets = [torch.from_numpy(pk.view(np.uint8).reshape(-1).copy()).cuda() for pk in packets]
        dimX, dimY, dimZ = mappers[0].dsi_.size_
        d_conf = torch.empty(dimX * dimY, dtype=torch.float32, device="cuda")
        d_depth = torch.empty(dimX * dimY, dtype=torch.float32, device="cuda")
        d_idx = torch.empty(dimX * dimY, dtype=torch.uint8 if dimZ <= 256 else torch.int16, device="cuda")
        torch.cuda.synchronize()
        grids = [m.dsi_ for m in mappers]
        t_all, t_fc = ctx.timer(), ctx.timer()

        def step(time_fc=False):
            for m, de, dp, pk, ev in zip(mappers, d_events, d_packets, packets, events):
                m.build_device(de.data_ptr(), len(ev), dp.data_ptr(), len(pk))
            if time_fc:
                t_fc.start()
            api.fuse_collapse_device(grids, method, mappers[0].depths_device_ptr(), d_conf.data_ptr(), d_idx.data_ptr(),
                                     d_depth.data_ptr())
            if time_fc:
                t_fc.stop()

        for _ in range(a.warmup):
            step()
        ctx.sync()
        ctx.profile_vote(True)
        t_all.start()
        for _ in range(a.steps):
            step()
        t_all.stop()
        ctx.sync()
        ms = t_all.elapsed_ms() / a.steps
        vote_ms, n_vote = ctx.vote_time()
        ctx.profile_vote(False)
        step(time_fc=True)
        ctx.sync()
        votes = int(sum(int(m.counts().sum()) for m in mappers))
        print(json.dumps({"variant": name, "env": env, "ms_per_step": round(ms, 4),
                          "mevents_per_s": round(len(cams) * a.events_per_cam / ms / 1e3, 1),
                          "vote_ms_per_launch": round(vote_ms / max(n_vote, 1), 5), "vote_launches_per_step": n_vote / a.steps,
                          "vote_share": round(vote_ms / a.steps / ms, 4), "depth_map_ms": round(t_fc.elapsed_ms(), 4),
                          "accepted_votes": votes}), flush=True)
        for m in mappers:
            m.close()
        del d_packets
        ctx.close()


if __name__ == "__main__":
    main()
