"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank builds its packet sub-interval of every camera (dvs_mcemvs_b200.shard.plan), the partial
DSIs are summed with the engine's ncclAllReduce, then fused + collapsed.  Rank 0 also builds the
unsharded DSIs on its own GPU and checks: vote counts bit-exact, DSI / confidence within float-sum
tolerance, argmax identical away from near-ties.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dvs_mcemvs_b200 import api, shard, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(idt.cpu().numpy().tobytes(), world, rank)

    def allgather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    # (2 000 events = 1 packet per camera: every rank but the first has NO packets and must still follow the
    #  same sequence of slab exchanges)
    for name, n_ev, overlapped in (("esim_small", 50_000, False), ("esim_small", 50_000, True),
                                   ("esim_small", 2_000, True), ("dsec_stereo", 400_000, True)):
        sc, _, method, _ = synth.config(name, events_per_cam=n_ev)
        cams = sc.rig.cams
        events = [sc.events(i, n_ev) for i in range(len(cams))]     # same seed on every rank: identical streams
        trajs = [api.LinearTrajectory(sc.trajectory(i)) for i in range(len(cams))]
        T = sc.T_rv_w()
        mappers = [api.MapperEMVS(ctx, c, sc.shape) for c in cams]
        packets = [m.packetize(ev, tr, T) for m, ev, tr in zip(mappers, events, trajs)]
        # (a) fused path: partial DSIs stay partial, one sweep over peer memory produces the maps
        ex = api.PeerExchange(ctx, [m.dsi_ for m in mappers], world, rank, allgather)
        peer_results = []
        two_d_possible = shard.camera_groups(len(cams), world) > 0
        for rep in range(8 if two_d_possible else 4):   # repeated rounds: the epoch flags must keep working; even = one exposed
            banded = rep % 2 == 1                       # sweep, odd = slab-wise band reduce overlapped with voting
            two_d = rep >= 4                            # rounds 4-7: camera x sub-interval sharding (one camera per rank)
            if rep in (0, 4):
                ex.set_participants(shard.participants(len(cams), world, two_d=two_d))
            if banded:
                ex.begin()
            plan = shard.plan2d if two_d else shard.plan
            for cam, lo, hi in plan([len(p) for p in packets], world, rank):
                mappers[cam].build(events[cam], packets[cam][lo:hi], peer_reduce=banded)
            ex.fuse_collapse(method, mappers[0].depths_device_ptr())
            peer_results.append(ex.download())
        # the host-buffer call with the split upload (head voted while the tail crosses PCIe, only the tail build
        # announces its slabs to the peers): every rank hands evaluateDSI the slice of events its packets cover
        ex.set_participants(shard.participants(len(cams), world, two_d=two_d_possible))
        ctx.set_upload_split(30, 4096)
        for rep in range(2):
            ctx.sync()                                  # idle pipeline: the split path is taken
            ex.begin()
            plan = shard.plan2d if two_d_possible else shard.plan
            for cam, lo, hi in plan([len(p) for p in packets], world, rank):
                pk = packets[cam][lo:hi]
                if len(pk) == 0:
                    mappers[cam].build(events[cam], pk, peer_reduce=True)
                    continue
                e_lo, e_hi = int(pk["first_event"][0]), min(len(events[cam]), int(pk["first_event"][-1]) + 1025)
                assert mappers[cam].evaluateDSI(events[cam][e_lo:e_hi], trajs[cam], T, peer_reduce=True)
            ex.fuse_collapse(method, mappers[0].depths_device_ptr())
            peer_results.append(ex.download())
        ctx.set_upload_split(15)
        ex.close()
        for r in peer_results[1:]:   # every round re-votes with atomics, so rounds agree to float-sum tolerance only
            np.testing.assert_allclose(r[0], peer_results[0][0], rtol=1e-4, atol=1e-6)
            assert float((r[1] == peer_results[0][1]).mean()) > 0.995, "peer sweep forms disagree"
        conf_p, idx_p, depth_p = peer_results[-1]
        # (b) allreduce paths
        for cam, lo, hi in shard.plan([len(p) for p in packets], world, rank):
            if overlapped:   # EMVS_BUILD_ALLREDUCE: every Z-slab is summed while the next one is voted
                mappers[cam].build(events[cam], packets[cam][lo:hi], allreduce=True)
            else:            # build, then one allreduce of the whole DSI
                mappers[cam].build(events[cam], packets[cam][lo:hi])
                mappers[cam].dsi_.allreduce()
                mappers[cam].counts_allreduce()
        conf, idx, depth = api.fuse_collapse([m.dsi_ for m in mappers], method, mappers[0].raw_depths_vec_)
        if rank == 0:
            full = [api.MapperEMVS(ctx, c, sc.shape) for c in cams]
            for m, ev, pk in zip(full, events, packets):
                m.build(ev, pk)
            for m, f in zip(mappers, full):
                assert np.array_equal(m.counts(), f.counts()), "sharded vote counts differ from unsharded"
                np.testing.assert_allclose(m.dsi_.download(), f.dsi_.download(), rtol=1e-5, atol=1e-5)
            conf_f, idx_f, depth_f = api.fuse_collapse([m.dsi_ for m in full], method, full[0].raw_depths_vec_)
            np.testing.assert_allclose(conf, conf_f, rtol=1e-4, atol=1e-6)
            agree = float((idx == idx_f).mean())
            assert agree > 0.995, agree
            np.testing.assert_allclose(conf_p, conf_f, rtol=1e-4, atol=1e-6)   # fused peer sweep
            agree_p = float((idx_p == idx_f).mean())
            assert agree_p > 0.995, agree_p
            same = idx_p == idx_f
            assert np.array_equal(depth_p[same], depth_f[same])
            print(f"mgpu_check {name} overlapped={overlapped}: world={world} counts exact, conf within 1e-4, index agreement {agree:.5f}")
        # all ranks hold identical maps after the allreduce (same summands, same NCCL reduction order)
        for arr, what in ((conf, "allreduce"), (conf_p, "peer sweep")):
            t = torch.from_numpy(arr.copy()).cuda()
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(t, ref), f"ranks disagree on the confidence map ({what})"
        for m in mappers:
            m.close()
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_check ok")


if __name__ == "__main__":
    main()
