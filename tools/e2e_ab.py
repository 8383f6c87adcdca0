#!/usr/bin/env python
"""A/B of the end-to-end (host buffers) call sequence on ONE generated workload: per variant (EMVS_* environment read
when the context is created) the stock sequence `evaluateDSI per camera -> fuse_collapse into host maps` and the
streaming one (next step's first list announced with prefetch) are timed with the host wall clock.

    python tools/e2e_ab.py [--events-per-cam N] [--steps K] [--variants name=K1:V1,K2:V2;...]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_VARIANTS = [
    ("default", {}),
    ("pieces2_split15", {"EMVS_UPLOAD_PIECES": "2", "EMVS_UPLOAD_SPLIT": "15"}),
    ("pieces3_split8", {"EMVS_UPLOAD_PIECES": "3", "EMVS_UPLOAD_SPLIT": "8"}),
    ("pieces3_split10", {"EMVS_UPLOAD_PIECES": "3", "EMVS_UPLOAD_SPLIT": "10"}),
    ("pieces3_split12", {"EMVS_UPLOAD_PIECES": "3", "EMVS_UPLOAD_SPLIT": "12"}),
    ("pieces3_split15", {"EMVS_UPLOAD_PIECES": "3", "EMVS_UPLOAD_SPLIT": "15"}),
    ("classic", {"EMVS_VOTE_KERNEL": "classic"}),
    ("vote_split0", {"EMVS_VOTE_SPLIT": "0"}),
    ("vote_split2", {"EMVS_VOTE_SPLIT": "2"}),
    ("upload_split0", {"EMVS_UPLOAD_SPLIT": "0"}),
    ("upload_split15", {"EMVS_UPLOAD_SPLIT": "15"}),
    ("upload_split40", {"EMVS_UPLOAD_SPLIT": "40"}),
    ("upload_split60", {"EMVS_UPLOAD_SPLIT": "60"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dsec_stereo")
    ap.add_argument("--events-per-cam", type=int, default=5_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="")
    a = ap.parse_args()
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
    pinned = []
    for ev in events:
        buf = api.pinned_empty(ev.shape, ev.dtype)
        buf[...] = ev
        pinned.append(buf)
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
        depths = mappers[0].raw_depths_vec_

        calls = {}

        def timed(label, f):
            t0 = time.perf_counter()
            r = f()
            calls[label] = calls.get(label, 0.0) + time.perf_counter() - t0
            return r

        def step(prefetch):
            for i, (m, ev, tr) in enumerate(zip(mappers, pinned, trajs)):
                assert timed(f"evaluateDSI{i}", lambda: m.evaluateDSI(ev, tr, T))
            if prefetch:
                timed("prefetch", lambda: mappers[0].prefetch(pinned[0], trajs[0], T))
            return timed("fuse_collapse", lambda: api.fuse_collapse([m.dsi_ for m in mappers], method, depths))

        out = {"variant": name, "env": env}
        for label, pf in (("stock_ms", False), ("streaming_ms", True)):
            for _ in range(2):
                step(pf)
            ctx.sync()
            calls.clear()
            ctx.profile_vote(True)
            t0 = time.perf_counter()
            for _ in range(a.steps):
                step(pf)
            ctx.sync()
            out[label] = round((time.perf_counter() - t0) / a.steps * 1e3, 3)
            vote_ms, n_vote = ctx.vote_time()
            ctx.profile_vote(False)
            out[label.replace("_ms", "_vote_ms_per_launch")] = round(vote_ms / max(n_vote, 1), 5)
            out[label.replace("_ms", "_vote_ms_per_step")] = round(vote_ms / a.steps, 3)
            out[label.replace("_ms", "_calls_ms")] = {k: round(v / a.steps * 1e3, 3) for k, v in calls.items()}
            ctx.prefetch_cancel()
        print(json.dumps(out), flush=True)
        for m in mappers:
            m.close()
        ctx.close()


if __name__ == "__main__":
    main()
