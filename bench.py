#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the DSI ray-voting path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]

One "step" = one pass of the hot path over one batch of synthetic input (BASELINE.json
configs[1]: DSEC-like stereo pair, 5 M events per camera, 640x480x256 DSI, harmonic fusion):

    for each camera: event warp + resetGrid + fillVoxelGrid        (MapperEMVS::evaluateDSI)
    [N > 1: sum of the ranks' partial DSIs (NVLink peer reduce under the votes, or ncclAllReduce)]
    harmonic-mean fusion + Z-argmax + index->depth                 (process_1 step 2-3)

Metric (BASELINE.json): Mevents/s = events of all cameras of all ranks / step time; the depth-map
milliseconds (fuse + argmax) and the build-only Mevents/s are reported beside it.

  value          inputs (events, packets) already resident in HBM, timed with CUDA events on the
                 engine's stream, max over ranks.
  e2e            the reference's own call sequence (MapperEMVS.evaluateDSI per camera with HOST
                 dvs_msgs::Event buffers in pinned memory -> fuse_collapse into HOST maps): host packet stage,
                 H2D of the events, build, fuse/argmax, D2H of depth/confidence/index all inside the timed
                 region, all `steps` timed.  No call the reference does not have.
  e2e_streaming  the same with the next step's first list announced by emvs_mapper_prefetch_dsi (a streaming
                 caller: its upload and packet stage run under the current votes).
  e2e_soa        e2e_streaming with structure-of-arrays event lists (x, y, t separate: 4 bytes per event cross
                 PCIe instead of 16).
  parity         the maps / counts / checksums of the timed configuration against the CPU oracle (N = 1) and
                 against an unsharded build on rank 0 (N > 1); outside every timed region.

Multi-GPU: the GPUs are divided over the cameras and each GPU builds one packet-aligned
sub-interval of ITS camera's event stream (--sharding 2d, the default when the GPU count is a multiple of the camera
count); --sharding interval makes every GPU build a sub-interval of EVERY camera (round 1).
  --scaling weak   (default) n_cameras x 5 M events per GPU — the streams get denser with N;
  --scaling strong --events-per-cam E: E events/camera in total, split over the N GPUs (BASELINE.json configs[3] is
                   E = 20 M on 8 GPUs).
Launched by torchrun; torch.distributed is only the rendezvous / barrier plumbing.

`--impl reference` times the CPU oracle (the restated reference loops, oracle/) on a bounded
sample of the same workload with all host threads; it never touches the GPU or libemvs_b200.so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mevents/s DSI build + fuse + argmax @ 640x480x256 (depth-map ms reported beside)"
UNIT = "Mevents/s"
WORKLOAD = "dsec_stereo"          # BASELINE.json configs[1]; --workload bar4 selects configs[2] (4 cameras, 10 M events each, GM)
FUSION_NAMES = {1: "min", 2: "harmonic", 3: "geometric", 4: "arithmetic", 5: "rms", 6: "max"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dsec_stereo", choices=["dsec_stereo", "bar4"],
                    help="dsec_stereo = BASELINE.json configs[1] (the metric's configuration); bar4 = configs[2]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --events-per-cam events per camera PER GPU; strong: in total, split over the GPUs")
    ap.add_argument("--events-per-cam", type=int, default=0, help="default: 5 M (dsec_stereo), 10 M (bar4)")
    ap.add_argument("--kind", default="structured", choices=["structured", "uniform"])
    ap.add_argument("--cpu-sample-events", type=int, default=5_000_000,
                    help="events per camera of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sharding", default="2d", choices=["2d", "interval"],
                    help="multi-GPU work split: 2d = camera x event sub-interval (one camera per GPU when the GPU count is a "
                         "multiple of the camera count; needs the peer exchange), interval = every GPU builds its sub-interval "
                         "of every camera")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU DSI exchange: fused reduce+fuse+argmax over NVLink peer memory, or "
                         "slab-wise ncclAllReduce overlapped with voting followed by a local sweep")
    a = ap.parse_args()
    global WORKLOAD
    WORKLOAD = a.workload
    if not a.events_per_cam:
        a.events_per_cam = 5_000_000 if a.workload == "dsec_stereo" else 10_000_000
    a.cpu_sample_events = min(a.cpu_sample_events, a.events_per_cam)
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_block(args, n_cams, dims, method, world):
    """The keys that name the workload — identical in the B200 arm and in the reference arm, so the driver can see
    that both lines are about the same configuration."""
    return {"workload": WORKLOAD, "cameras": n_cams, "dsi": [int(d) for d in dims], "fusion": FUSION_NAMES[method],
            "event_distribution": args.kind, "scaling": args.scaling, "gpus": world,
            "events_per_camera": int(args.events_per_cam * (world if args.scaling == "weak" else 1)),
            "events_per_camera_per_gpu": (int(args.events_per_cam) if args.scaling == "weak"
                                          else int(args.events_per_cam // world)),
            "l2": "inputs + DSIs (> 700 MB per step) exceed the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_workload(n_events, stream, kind):
    """Seeded synthetic stereo streams.  `stream` selects an independent event sample of the same scene, rig and
    trajectory (weak scaling: rank r owns sample r of a denser stream, the per-GPU event count is fixed)."""
    from dvs_mcemvs_b200 import synth
    sc, _, method, desc = synth.config(WORKLOAD, events_per_cam=n_events)
    cams = sc.rig.cams
    events = [sc.events(i, n_events, kind, stream=stream) for i in range(len(cams))]
    trajs = [sc.trajectory(i) for i in range(len(cams))]
    return sc, cams, events, trajs, sc.T_rv_w(), method, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU reference leg (oracle): bounded sample, extrapolated linearly in the event count
# ----------------------------------------------------------------------------------------------
_CPU_WORKLOAD = {}


def cpu_reference(args, n_full, keep=False):
    """Times the restated reference loops on `--cpu-sample-events` events per camera (the vote
    loop is linear in the event count; fusion and argmax do not depend on it) and converts to the
    full workload: t_full = t_build * n_full / n_sample + t_fuse + t_argmax.
    keep=True also returns what the run computed (volumes, counts, maps) for the parity block."""
    from oracle import oracle as O
    O.use_all_host_threads()   # torchrun sets OMP_NUM_THREADS=1 for its workers; the baseline gets every host thread
    n_s = min(args.cpu_sample_events, n_full)
    if (n_s, args.kind) not in _CPU_WORKLOAD:      # synthetic input generation is not part of any timed scope
        _CPU_WORKLOAD[(n_s, args.kind)] = make_workload(n_s, 0, args.kind)
    sc, cams, events, trajs, T_rv_w, method, _ = _CPU_WORKLOAD[(n_s, args.kind)]
    sh = sc.shape
    dimX, dimY, dimZ = cams[0].width, cams[0].height, sh.dimZ_
    depths = O.depth_vector(sh.min_depth_, sh.max_depth_, dimZ, sh.inverse_depth)
    t_build, vols, counts = 0.0, [], []
    for cam, ev, tr in zip(cams, events, trajs):
        virt = O.virtual_camera(cam.fx, cam.cx, cam.cy, dimX, sh.fov_)
        t0 = time.perf_counter()
        pk = O.packetize(ev, tr, T_rv_w, np.array([cam.fx, cam.fy, cam.cx, cam.cy], np.float32), virt, depths[0])
        dsi, inb = O.build_dsi(ev, pk, cam.lut, cam.width, depths, virt, dimX, dimY)
        t_build += time.perf_counter() - t0
        vols.append(dsi)
        counts.append(inb)
    n_cams = len(cams)
    t0 = time.perf_counter()
    if n_cams <= 3:
        fused = O.fuse_reference(method, vols, copy_arg=True)   # by-value argument copy like the reference
    else:
        fused = O.fuse_nary(method, vols)                       # > 3 cameras: the n-ary extension
    t_fuse = time.perf_counter() - t0
    t0 = time.perf_counter()
    conf, idx, depth = O.collapse_max(fused, depths)
    t_argmax = time.perf_counter() - t0
    t_full = t_build * (n_full / n_s) + t_fuse + t_argmax
    ref_code = None
    try:   # the reference's OWN Grid3D code (compiled in place into oracle/_ref) for the fusion + argmax stage
        from oracle import ref as R
        if R.available() and n_cams == 2:
            f_ms, a_ms, conf_r, _ = R.time_fuse_collapse(vols[0], vols[1], method)
            ref_code = {"fuse_ms": f_ms, "argmax_ms": a_ms, "depth_map_ms": f_ms + a_ms,
                        "what": "cartesian3dgrid.h fusion ops (by-value argument, .at()) + Grid3D::collapseMaxZSlice compiled "
                                "from the reference sources (oracle/_ref); the build stage has no compilable reference"}
    except Exception as e:   # never let the extra baseline break the bench line
        ref_code = {"unavailable": str(e)[:200]}
    block = {
        "value": n_cams * n_full / t_full / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
        "sample": (f"{n_s} events/camera x {n_cams} cameras voted into the full {dimX}x{dimY}x{dimZ} DSI by the oracle "
                   f"(g++ -O3 -fopenmp, OMP over planes like mapper_emvs_stereo.cpp:168), fusion (single thread, "
                   f"by-value copy) and argmax at full size"
                   + ("" if n_s == n_full else f"; build time scaled x{n_full / n_s:.1f} to {n_full} events/camera")),
        "build_mevents_per_s": n_cams * n_s / t_build / 1e6,
        "depth_map_ms": (t_fuse + t_argmax) * 1e3, "fuse_ms": t_fuse * 1e3, "argmax_ms": t_argmax * 1e3,
        "sample_build_s": t_build, "host_cpus": os.cpu_count(), "reference_code_depth_map": ref_code,
    }
    art = None
    if keep and n_s == n_full:
        art = {"vols": vols, "counts": counts, "fused": fused, "conf": conf, "idx": idx, "depth": depth,
               "mean_square": [O.mean_square(v) for v in vols]}
    return block, t_full, art


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dvs_mcemvs_b200 import synth
    synth.LUT_BACKEND = "cv2"      # the DSEC rectification LUT from OpenCV: this arm never maps libemvs_b200.so
    world = max(1, args.gpus)
    # the arm's workload is the B200 arm's at the same N: `events_per_camera` of config_block
    n_full = args.events_per_cam * (world if args.scaling == "weak" else 1)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        base, t_full, _ = cpu_reference(args, n_full)
        if i >= args.warmup:
            vals.append(t_full)
        last = base
        if i >= args.warmup and sum(vals) > 240:   # keep the whole run within a few minutes
            break
    t = float(np.mean(vals))
    wl = _CPU_WORKLOAD[next(iter(_CPU_WORKLOAD))]
    cams, method = wl[1], wl[5]
    n_cams = len(cams)
    value = n_cams * n_full / t / 1e6
    last["value"] = value
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args, n_cams, (cams[0].width, cams[0].height, wl[0].shape.dimZ_), method, world),
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from dvs_mcemvs_b200 import _capi as capi
    from dvs_mcemvs_b200 import api, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = api.Context(local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(idt.cpu().numpy().tobytes(), world, rank)

    strong = args.scaling == "strong" and world > 1
    n_ev = args.events_per_cam
    sc, _, method, desc = synth.config(WORKLOAD, events_per_cam=n_ev)
    cams = sc.rig.cams
    n_cams = len(cams)
    trajs = [sc.trajectory(i) for i in range(n_cams)]
    T_rv_w = sc.T_rv_w()
    # Sharding.  2d: camera x sub-interval — the GPUs are divided over the cameras, a rank builds ONE
    # camera's sub-interval (needs the peer exchange and a GPU count that is a multiple of the camera count).
    # interval: every rank builds a sub-interval of EVERY camera (round 1; the only form the NCCL exchange supports).
    group = shard.camera_groups(n_cams, world) if (world > 1 and args.sharding == "2d" and args.exchange == "peer") else 0
    if group:
        my_cam, part = divmod(rank, group)
        local = [my_cam]
    else:
        part, local = rank, list(range(n_cams))
    need_all = strong and rank == 0 and not args.no_parity          # rank 0 builds every camera unsharded for the parity block
    mappers = [api.MapperEMVS(ctx, c, sc.shape) for c in cams]
    ltrajs = [api.LinearTrajectory(t) for t in trajs]
    depths = mappers[0].raw_depths_vec_
    dimX, dimY, dimZ = mappers[0].dsi_.size_

    # ---- this rank's event lists ---------------------------------------------------------------
    events, packets = {}, {}            # camera -> this rank's list / packets (first_event relative to the list)
    events_all, packets_all = {}, {}    # strong scaling: the whole list of a camera and its packets
    if strong:
        # a camera's list is the same on every rank; a rank builds packet range `part` of its group
        for c in range(n_cams):
            if not (need_all or c in local):
                sc.skip_events()        # keep the scene's seed sequence in step with the ranks that do generate this list
                continue
            events_all[c] = sc.events(c, n_ev, args.kind, stream=0)
            packets_all[c] = mappers[c].packetize(events_all[c], ltrajs[c], T_rv_w)
        for c in local:
            n_parts = group if group else world
            lo, hi = shard.split_range(len(packets_all[c]), n_parts)[part]
            pk = packets_all[c][lo:hi].copy()
            if len(pk):
                e_lo = int(pk["first_event"][0])
                # one event past the last packet: the packet loop's strict '<' (mapper_emvs_stereo.cpp:88) needs it to
                # form that packet when this slice is packetised on its own (the e2e leg)
                e_hi = min(len(events_all[c]), int(pk["first_event"][-1]) + capi.PACKET_SIZE + 1)
                pk["first_event"] -= e_lo
                events[c] = np.ascontiguousarray(events_all[c][e_lo:e_hi])
            else:
                events[c] = events_all[c][:0].copy()
            packets[c] = pk
    else:
        # weak scaling: n_cams * n_ev events per GPU.  2d: all of them belong to the rank's camera (sample `part` of that
        # camera's stream); interval: n_ev of every camera (sample `rank`)
        for c in range(n_cams):
            if c not in local:
                sc.skip_events()
                continue
            n_local = n_ev * n_cams if group else n_ev
            events[c] = sc.events(c, n_local, args.kind, stream=part)
            packets[c] = mappers[c].packetize(events[c], ltrajs[c], T_rv_w)
        if world == 1:
            events_all, packets_all = events, packets

    # ---- device-resident inputs for `value` -------------------------------------------------
    d_events = {c: torch.from_numpy(events[c].view(np.uint8).reshape(-1).copy()).cuda() for c in local}
    d_packets = {c: torch.from_numpy(packets[c].view(np.uint8).reshape(-1).copy()).cuda() for c in local}
    torch.cuda.synchronize()
    d_conf = torch.empty(dimY * dimX, dtype=torch.float32, device="cuda")
    d_depth = torch.empty(dimY * dimX, dtype=torch.float32, device="cuda")
    d_idx = torch.empty(dimY * dimX, dtype=torch.uint8 if dimZ <= 256 else torch.int16, device="cuda")
    d_tab = mappers[0].depths_device_ptr()
    torch.cuda.synchronize()
    grids = [m.dsi_ for m in mappers]

    def collapse_device():
        api.fuse_collapse_device(grids, method, d_tab, d_conf.data_ptr(), d_idx.data_ptr(), d_depth.data_ptr())

    peer = None
    if world > 1 and args.exchange == "peer":
        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        peer = api.PeerExchange(ctx, grids, world, rank, allgather,
                                participants=shard.participants(n_cams, world, two_d=bool(group)))

    def builds_device():
        for c in local:
            mappers[c].build_device(d_events[c].data_ptr(), len(events[c]), d_packets[c].data_ptr(), len(packets[c]),
                                    allreduce=world > 1 and peer is None, peer_reduce=peer is not None)

    def step_device():
        if peer is not None:
            peer.begin()      # slab-wise band reduce over NVLink, overlapped with voting
        builds_device()
        if peer is not None:
            peer.fuse_collapse(method, d_tab)
        else:
            collapse_device()

    t_all, t_build, t_depth = ctx.timer(), ctx.timer(), ctx.timer()

    def timed_device(k):
        """K steps bracketed by barrier + synchronize; returns the total ms."""
        barrier()
        t_all.start()
        for _ in range(k):
            step_device()
        t_all.stop()
        ctx.sync()
        barrier()
        return t_all.elapsed_ms()

    for _ in range(args.warmup):
        step_device()
    ctx.sync()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ms_total = timed_device(args.steps)
    n_launch = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-launch timing of the vote kernel (roofline): a second pass of the same steps with an event pair around every
    # vote launch — kept out of the `value` pass, where those extra stream operations would be inside the timed region
    ctx.profile_vote(True)
    timed_device(args.steps)
    vote_ms, vote_launches = ctx.vote_time()
    ctx.profile_vote(False)

    # stage split (a separate pass, not part of `value`): build vs depth-map (vs multi-GPU exchange)
    barrier()
    if peer is not None:
        peer.begin()
    t_build.start()
    builds_device()
    t_build.stop()
    t_depth.start()
    if peer is not None:
        peer.fuse_collapse(method, d_tab)    # fuse + argmax of this rank's band + distribution of the maps (+ the epoch waits)
    else:
        collapse_device()
    t_depth.stop()
    ctx.sync()
    build_ms, depth_ms = t_build.elapsed_ms(), t_depth.elapsed_ms()
    # accepted (event, plane) votes of this rank's builds, per camera (zero rows for the cameras it does not build)
    counts_local = np.zeros((n_cams, dimZ), np.uint64)
    for c in local:
        counts_local[c] = mappers[c].counts()
    votes = int(sum(int(counts_local[c].sum()) for c in local))
    if world > 1 and peer is None:
        votes //= world                                # the NCCL path has already summed the counters over the ranks
    n_voted_events = sum(len(packets[c]) for c in local) * capi.PACKET_SIZE
    if peer is not None:
        maps_dev = peer.download()                     # the exchanged maps of the timed configuration (every rank has them)
    else:
        maps_dev = api.fuse_collapse(grids, method, depths)
    mean_square = [g.computeMeanSquare() for g in grids] if world == 1 or peer is None else None

    # ---- e2e: host buffers through the reference-facing calls ---------------------------------
    e2e_runs = {}
    if not args.no_e2e:
        h_events = {}
        for c in local:
            buf = api.pinned_empty(events[c].shape, events[c].dtype)
            buf[...] = events[c]
            h_events[c] = buf
        h_soa = {c: api.EventsSoA.from_events(events[c], pinned=True) for c in local}
        if strong:   # the slice of a rank, packetised on its own, must give that rank's packets of the global list
            for c in local:
                mine = mappers[c].packetize(h_events[c], ltrajs[c], T_rv_w)
                assert mine is not None and mine.tobytes() == packets[c].tobytes(), "sub-interval packets differ from the global packet list"

        def step_host(lists, prefetch):
            if peer is not None:
                peer.begin()
            for c in local:
                assert mappers[c].evaluateDSI(lists[c], ltrajs[c], T_rv_w, allreduce=world > 1 and peer is None,
                                              peer_reduce=peer is not None)
            if prefetch:
                # streaming caller: the NEXT step's first event list starts crossing PCIe now, under this step's
                # votes (every step still uploads every list once, inside the timed region)
                mappers[local[0]].prefetch(lists[local[0]], ltrajs[local[0]], T_rv_w)
            if peer is not None:
                peer.fuse_collapse(method, d_tab)
                return peer.download()
            return api.fuse_collapse([m.dsi_ for m in mappers], method, depths)

        def time_host(lists, prefetch, bytes_per_event):
            for _ in range(max(1, min(args.warmup, 2))):
                step_host(lists, prefetch)
            ctx.prefetch_cancel()
            if prefetch:
                mappers[local[0]].prefetch(lists[local[0]], ltrajs[local[0]], T_rv_w)   # what the previous window's step would have announced
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                conf, idx, depth = step_host(lists, prefetch)
            ctx.sync()
            barrier()
            e2e_s = (time.perf_counter() - t0) / args.steps
            ctx.prefetch_cancel()
            h2d = sum(len(events[c]) for c in local) * bytes_per_event + sum(packets[c].nbytes for c in local) + depths.nbytes
            d2h = conf.nbytes + idx.nbytes + depth.nbytes
            return e2e_s, h2d, d2h, (conf, idx, depth)

        e2e_runs["e2e"] = time_host(h_events, False, 16)
        e2e_runs["e2e_streaming"] = time_host(h_events, True, 16)
        e2e_runs["e2e_soa"] = time_host(h_soa, True, 4)

    # ---- parity (outside every timed region) ---------------------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_block(args, world, rank, ctx, api, dist, torch, sc, cams, trajs, T_rv_w, packets_all, events_all, events,
                              packets, local, group, method, depths, maps_dev, counts_local, mean_square, strong, peer,
                              {k: v[3] for k, v in e2e_runs.items()})

    # ---- reduce over ranks ------------------------------------------------------------------
    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    total_events = float(n_cams * n_ev) if strong else sum_over_ranks(float(sum(len(events[c]) for c in local)))
    value = total_events / (ms_step * 1e-3) / 1e6
    votes_all = sum_over_ranks(float(votes))
    vote_ms_max = max_over_ranks(vote_ms)
    build_ms, depth_ms = max_over_ranks(build_ms), max_over_ranks(depth_ms)
    e2e_out = {}
    notes = {
        "e2e": "the reference's own call sequence, nothing else: evaluateDSI per camera on pinned dvs_msgs::Event lists, "
               "fuse_collapse into host maps",
        "e2e_streaming": "+ the next step's first evaluateDSI is announced with emvs_mapper_prefetch_dsi after the last "
                         "evaluateDSI of a step (its upload and host packet stage run under the current votes); every list "
                         "is uploaded and packetised once per step inside the timed region",
        "e2e_soa": "e2e_streaming with structure-of-arrays event lists (emvs_mapper_evaluate_dsi_soa): x, y cross PCIe, "
                   "timestamps stay on the host",
    }
    for name, (e2e_s, h2d, d2h, _) in e2e_runs.items():
        e2e_s = max_over_ranks(e2e_s)
        e2e_out[name] = {"value": total_events / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                         "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": args.steps,
                         "timer": "host wall clock around the blocking public calls (each call syncs its stream), max over ranks",
                         "calls": notes[name]}

    if rank == 0:
        peak, peak_src = load_peaks()
        n_builds = max(len(local), 1)
        roof = roofline_block(peak, peak_src, votes, n_voted_events, vote_ms / args.steps, vote_launches / args.steps, ms_step,
                              n_builds, sum(len(events[c]) for c in local) // n_builds, dimX, dimY, dimZ)
        sharding = ("none" if world == 1 else
                    (f"camera x event sub-interval: {group} GPU(s) per camera, each builds one camera's sub-interval" if group else
                     "event sub-interval: every GPU builds its sub-interval of every camera")
                    + ("; slab-wise reduce of each GPU's row band over NVLink peer memory under the votes, then fuse+argmax of "
                       "the band and peer stores of the maps" if peer is not None else
                       "; ncclAllReduce(sum) per Z-slab of each camera DSI, overlapped with voting"))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(args, n_cams, (dimX, dimY, dimZ), method, world),
            "config_detail": {"description": desc, "sharding": sharding},
            "build_mevents_per_s": total_events / (build_ms * 1e-3) / 1e6, "build_ms": build_ms,
            "depth_map_ms": depth_ms, "accepted_votes_per_step": int(votes_all),
            "vote_ms_per_launch_max_over_ranks": vote_ms_max / max(vote_launches, 1),
            "stage_note": ("build_ms = event stage + reset + votes + merges of all cameras; depth_map_ms = fuse + argmax + index->depth"
                           + ("" if world == 1 else
                              " of this GPU's row band + distribution of the maps; the slab-wise peer reduce runs inside build_ms" if peer is not None else
                              "; the slab-wise ncclAllReduce runs inside build_ms, overlapped with voting")),
            "mean_square": mean_square,
            "clocks": clocks, "e2e": e2e_out.get("e2e"), "e2e_streaming": e2e_out.get("e2e_streaming"),
            "e2e_soa": e2e_out.get("e2e_soa"), "gpu_launches": int(n_launch), "roofline": roof, "parity": parity,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = parity.get("_cpu_baseline") if parity else None
            if cb is None:
                _CPU_WORKLOAD.setdefault((min(args.cpu_sample_events, n_ev), args.kind),
                                         (sc, cams, [events[c] for c in range(n_cams)], trajs, T_rv_w, method, desc))
                cb, _, _ = cpu_reference(args, n_ev)
            out["cpu_baseline"] = cb
        if parity:
            parity.pop("_cpu_baseline", None)
        _emit(json.dumps(out))
    if world > 1:
        if peer is not None:
            ctx.sync()
            dist.barrier()
            peer.close()
        ctx.comm_destroy()
        dist.destroy_process_group()


def parity_block(args, world, rank, ctx, api, dist, torch, sc, cams, trajs, T_rv_w, packets_all, events_all, events, packets,
                 local, group, method, depths, maps_dev, counts_local, mean_square, strong, peer, e2e_maps):
    """N = 1: the timed configuration's counts / maps / checksums against the CPU oracle on the same events (the run
    that is also the cpu_baseline).  N > 1: the exchanged maps and the summed counts against an UNSHARDED build of all
    ranks' events on rank 0's GPU (itself oracle-checked at N = 1).  Returns a dict on rank 0, None elsewhere."""
    from oracle import parity as P
    conf, idx, depth = maps_dev
    n_cams = len(cams)
    if world == 1:
        n_ev = args.events_per_cam
        if min(args.cpu_sample_events, n_ev) != n_ev:
            return {"skipped": "the CPU sample is smaller than the workload (--cpu-sample-events): no full-size oracle run"}
        _CPU_WORKLOAD[(n_ev, args.kind)] = (sc, cams, [events[c] for c in range(n_cams)], trajs, T_rv_w, method, "")
        cb, _, art = cpu_reference(args, n_ev, keep=True)
        p = P.compare_maps(conf, idx, depth, art["conf"], art["idx"], art["depth"], art["fused"])
        p["counts_exact"] = bool(all(np.array_equal(counts_local[c], art["counts"][c]) for c in range(n_cams)))
        p["mean_square_gpu"] = mean_square
        p["mean_square_oracle"] = art["mean_square"]
        p["mean_square_rel"] = float(max(abs(g - o) / max(abs(o), 1e-300) for g, o in zip(mean_square, art["mean_square"])))
        for name, (c2, i2, d2) in e2e_maps.items():   # the e2e legs (host buffers, split upload, prefetch, SoA) produce the same maps
            q = P.compare_maps(c2, i2, d2, art["conf"], art["idx"], art["depth"], art["fused"])
            p[name + "_ok"] = P.verdict(q)
        p["against"] = "CPU oracle (oracle/emvs_oracle.cpp) on the same events, full size, all planes"
        p["ok"] = P.verdict(p) and all(p.get(k + "_ok", True) for k in e2e_maps)
        p["_cpu_baseline"] = cb
        return p
    # ---- N > 1 ------------------------------------------------------------------------------
    cnt = torch.from_numpy(counts_local.astype(np.int64)).cuda()
    if peer is not None:
        dist.all_reduce(cnt)          # peer mode keeps per-rank counts: the job's counts are their sum
    cnt = cnt.cpu().numpy().astype(np.uint64)
    full = None
    if strong:
        if rank == 0:
            full = [api.MapperEMVS(ctx, c, sc.shape) for c in cams]
            for c, m in enumerate(full):
                m.build(events_all[c], packets_all[c])
    else:
        # weak scaling: rank 0 accumulates every rank's list(s) into one DSI per camera (votes add)
        seen = set()
        for slot in range(len(local)):            # the same number of lists on every rank
            c = local[slot]
            ev_t = torch.from_numpy(events[c].view(np.uint8).reshape(-1).copy()).cuda()
            meta = torch.tensor([len(packets[c]), c, len(events[c])], dtype=torch.int64, device="cuda")
            metas = [torch.zeros_like(meta) for _ in range(world)]
            dist.all_gather(metas, meta)
            metas = [[int(v) for v in t.tolist()] for t in metas]
            max_pk = max(mm[0] for mm in metas)
            assert all(mm[2] == len(events[c]) for mm in metas), "ranks hold lists of different lengths"
            pk_pad = np.zeros(max_pk, packets[c].dtype)
            pk_pad[:len(packets[c])] = packets[c]
            pk_t = torch.from_numpy(pk_pad.view(np.uint8).reshape(-1).copy()).cuda()
            ev_list = [torch.empty_like(ev_t) for _ in range(world)] if rank == 0 else None
            pk_list = [torch.empty_like(pk_t) for _ in range(world)] if rank == 0 else None
            dist.gather(ev_t, ev_list, dst=0)
            dist.gather(pk_t, pk_list, dst=0)
            torch.cuda.synchronize()     # the engine reads these tensors on its own stream
            if rank == 0:
                if full is None:
                    full = [api.MapperEMVS(ctx, cc, sc.shape) for cc in cams]
                for r in range(world):
                    n_pk_r, cam_r, n_ev_r = metas[r]
                    full[cam_r].build_device(ev_list[r].data_ptr(), n_ev_r, pk_list[r].data_ptr(), n_pk_r, accumulate=cam_r in seen)
                    seen.add(cam_r)
                ctx.sync()
            del ev_list, pk_list
    p = None
    if rank == 0:
        conf_f, idx_f, depth_f = api.fuse_collapse([m.dsi_ for m in full], method, depths)
        fused_g = api.Grid3D(ctx, *full[0].dsi_.size_)
        api.fuse_collapse([m.dsi_ for m in full], method, depths, fused_out=fused_g)
        p = P.compare_maps(conf, idx, depth, conf_f, idx_f, depth_f, fused_g.download())
        fused_g.close()
        p["counts_exact"] = bool(all(np.array_equal(cnt[i], full[i].counts()) for i in range(n_cams)))
        for name, (c2, i2, d2) in e2e_maps.items():
            q = P.compare_maps(c2, i2, d2, conf_f, idx_f, depth_f)
            q["idx_mismatches_are_near_ties"] = None
            p[name + "_ok"] = P.verdict(q)
        p["against"] = ("an unsharded build of all ranks' events on rank 0's GPU (that single-GPU path is what the N = 1 "
                        "bench line and tests/test_gpu_fullsize.py check against the CPU oracle)")
        p["ok"] = P.verdict(p) and all(p.get(k + "_ok", True) for k in e2e_maps)
        for m in full:
            m.close()
    # every rank holds identical maps
    t = torch.from_numpy(np.ascontiguousarray(conf).copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([1 if torch.equal(t, ref) else 0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        p["maps_identical_on_all_ranks"] = bool(same.item())
        p["ok"] = bool(p["ok"] and same.item())
    return p


def roofline_block(peak, peak_src, votes, n_voted_events, vote_ms_per_step, launches_per_step, ms_step, n_cams, n_ev_per_cam,
                   dimX, dimY, dimZ):
    """The dominant kernel (k_vote_tma) against what limits it.  SURVEY.md §8(d) defines the algorithmic bytes
    (32 B per accepted vote: 4 voxels x read + write, + 8 B per warped event); physically every vote is ONE 16-byte
    red.global.add.v4.f32 that resolves in an L2-resident slab, so DRAM is idle and the limiter is the path that carries
    RED payload from the SMs into the L2 atomic units.  Both views are printed; `frac` is the physical one."""
    red = load_json("profiles/red_port_ceiling.json") or {}
    payload = votes * 16.0 / (vote_ms_per_step * 1e-3) / 1e9            # GB/s of RED payload delivered to L2
    alg_bytes_step = votes * 32.0 + n_voted_events * 8.0
    alg = alg_bytes_step / (vote_ms_per_step * 1e-3) / 1e9
    ceil = red.get("payload_gb_per_s")
    slab = 16
    compulsory = n_cams * ((dimX * dimY * dimZ) * 4.0 * 2 + n_ev_per_cam * 16.0 * ((dimZ + slab - 1) // slab))
    return {
        "bound": "l2_red_port", "kernel": "k_vote_tma<8>", "achieved": payload, "peak": ceil, "unit": "GB/s",
        "frac": (payload / ceil) if ceil else None,
        "peak_source": red.get("source", "profiles/red_port_ceiling.json missing"),
        "what": "RED payload (16 B per accepted vote) per second delivered by the vote kernel, against the ceiling a "
                "micro-benchmark reaches with the same access pattern (8 lanes per 128-byte line, full sectors) and no "
                "arithmetic; port theory: 148 SMs x 32 B/clk x 1.965 GHz = 9307 GB/s",
        "traffic": (load_json("profiles/vote_traffic.json") or {}).get("dram_bytes_per_launch"),
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per vote launch, ncu --set full (profiles/vote_traffic.json)",
        "algorithmic_bytes_per_launch": alg_bytes_step / launches_per_step,
        "avg_launch_ms": vote_ms_per_step / launches_per_step, "launches_per_step": launches_per_step,
        "kernel_share_of_step": vote_ms_per_step / ms_step,
        "hbm_algorithmic": {"achieved": alg, "peak": peak, "unit": "GB/s", "frac": alg / peak, "peak_source": peak_src,
                            "label": "cache-served: SURVEY.md §8(d)'s uncached-scatter bytes (32 B per vote) over the kernel "
                                     "time; the read-modify-write happens in L2, so this exceeds 1 while DRAM is ~5 % busy"},
        "compulsory": {"bytes_per_step": compulsory, "time_at_hbm_peak_ms": compulsory / (peak * 1e9) * 1e3,
                       "frac_of_step": compulsory / (peak * 1e9) * 1e3 / ms_step,
                       "label": "SURVEY.md §8(d)(ii): Nvox*4*2 + Ne*16*ceil(Nz/slab) per camera — the DRAM traffic a perfect "
                                "implementation of this slab scheme cannot avoid, as a share of the measured step"},
    }


def load_json(rel):
    p = os.path.join(ROOT, rel)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def _emit(line):
    os.write(_REAL_STDOUT, (line + "\n").encode())


if __name__ == "__main__":
    # Exactly ONE line may reach stdout (the driver parses it): route everything libraries print
    # (NCCL banners, torchrun notices) to stderr and keep the real stdout for the JSON line.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
