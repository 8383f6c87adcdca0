"""Alg. 2 spread over the GPUs (api.process_2_sharded, BASELINE.json configs[3] "AtHc" ordering) against the
single-GPU api.process_2 with num_subintervals = world, run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/mgpu_check_alg2.py

Run by tests/test_gpu_multi.py::test_two_rank_alg2_sharded_matches_process_2 on boxes with >= 2 GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dvs_mcemvs_b200 import api, synth  # noqa: E402


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
    sc, n_ev, _, _ = synth.config("esim_small", events_per_cam=60_000)
    cams = sc.rig.cams[:2]
    events = [sc.events(i, n_ev) for i in range(2)]
    trajs = [api.LinearTrajectory(sc.trajectory(i)) for i in range(2)]
    T = sc.T_rv_w()
    for stereo, temporal in ((2, 4), (2, 2), (4, 4), (1, 2)):
        fused = api.process_2_sharded(ctx, cams, trajs, events, sc.shape, T, stereo, temporal, rank, world)
        got = fused.download()
        fused.close()
        if rank == 0:
            ref = api.process_2(ctx, cams, trajs, events, sc.shape, world, T, stereo, temporal)
            want = ref["fused"].download()
            for g in ref.values():
                g.close()
            np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-5)
            print(f"mgpu_check_alg2 stereo={stereo} temporal={temporal}: world={world} fused volume matches process_2")
        t = torch.from_numpy(got.copy()).cuda()
        ref_t = t.clone()
        dist.broadcast(ref_t, 0)
        assert torch.equal(t, ref_t), "ranks disagree on the fused volume"
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_check_alg2 ok")


if __name__ == "__main__":
    main()
