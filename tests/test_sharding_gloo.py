"""Multi-process (world_size 2, gloo, CPU) test of the sharding logic: packet-aligned sub-interval
shards built independently and summed with one all_reduce equal the unsharded build — integer vote
counts exactly, float DSI to summation-order tolerance.  The oracle stands in for the GPU build here
(no GPU in this container); the same plan drives the CUDA path under torchrun (tests/mgpu_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dvs_mcemvs_b200 import shard


def test_split_range_properties():
    for n in (0, 1, 7, 48, 4882, 9765):
        for parts in (1, 2, 3, 8):
            r = shard.split_range(n, parts)
            assert len(r) == parts and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard.plan([48, 47], 2, 1) == [(0, 24, 48), (1, 24, 47)]
    assert shard.row_bands(480, 8)[3] == (180, 240)
    with pytest.raises(ValueError):
        shard.plan([4], 2, 2)


def test_camera_x_subinterval_plan():
    """with G GPUs and C cameras, G a multiple of C, each camera gets G/C GPUs and its packet list is
    split over them; otherwise every GPU builds a sub-interval of every camera."""
    n = [4882, 4881]
    for world in (2, 4, 8):
        g = shard.camera_groups(2, world)
        assert g == world // 2
        covered = {0: [], 1: []}
        for r in range(world):
            (cam, lo, hi), = shard.plan2d(n, world, r)
            assert cam == r // g
            covered[cam].append((lo, hi))
        for cam in (0, 1):     # the group's ranges tile the camera's packet list exactly once
            assert covered[cam][0][0] == 0 and covered[cam][-1][1] == n[cam]
            assert all(a[1] == b[0] for a, b in zip(covered[cam], covered[cam][1:]))
        tab = shard.participants(2, world)
        assert [sum(row) for row in tab] == [g, g] and all(tab[0][r] + tab[1][r] == 1 for r in range(world))
    # not divisible / fewer GPUs than cameras: the sub-interval-only plan, everybody participates everywhere
    assert shard.camera_groups(2, 3) == 0 and shard.camera_groups(4, 2) == 0 and shard.camera_groups(2, 1) == 0
    assert shard.plan2d(n, 3, 1) == shard.plan(n, 3, 1)
    assert shard.participants(2, 3) == [[1, 1, 1], [1, 1, 1]] and shard.participants(2, 4, two_d=False) == [[1] * 4] * 2
    assert shard.plan2d([10, 10, 10, 10], 8, 7) == [(3, 5, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, two_d=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan_fn = shard.plan2d if two_d else shard.plan
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import Case
        from oracle import oracle as O
        case = Case("esim_small", events_per_cam=30_000)
        vols, counts = [], []
        my = {cam: (lo, hi) for cam, lo, hi in plan_fn([len(p) for p in case.packets], world, rank)}
        for cam in range(case.n_cams):
            c = case.cams[cam]
            lo, hi = my.get(cam, (0, 0))      # camera x sub-interval sharding: nothing to vote for the other camera
            dsi, inb = O.build_dsi(case.events[cam], case.packets[cam][lo:hi], c.lut, c.width, case.depths,
                                   case.virts[cam], case.dimX, case.dimY)
            t, ti = torch.from_numpy(dsi), torch.from_numpy(inb.astype(np.int64))
            dist.all_reduce(t)        # the DSI sum-exchange (peer reduce / NCCL on the GPU path)
            dist.all_reduce(ti)
            vols.append(t.numpy())
            counts.append(ti.numpy())
        fused = O.fuse_reference(case.method, vols)
        conf, idx, depth = O.collapse_max(fused, case.depths)
        if rank == 0:
            full = [case.oracle_dsi(i) for i in range(case.n_cams)]
            for v, n, (dsi_o, inb_o) in zip(vols, counts, full):
                assert np.array_equal(n.astype(np.uint64), inb_o)
                np.testing.assert_allclose(v, dsi_o, rtol=1e-5, atol=1e-5)
            conf_o, idx_o, _ = O.collapse_max(O.fuse_reference(case.method, [f[0] for f in full]), case.depths)
            np.testing.assert_allclose(conf, conf_o, rtol=1e-4, atol=1e-6)
            assert (idx == idx_o).mean() > 0.999
        # every rank ends with the same maps
        t = torch.from_numpy(conf.copy())
        dist.broadcast(t, 0)
        assert np.array_equal(t.numpy(), conf)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("two_d", [False, True], ids=["interval", "camera_x_interval"])
def test_sharded_equals_unsharded_world2(tmp_path, two_d):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), two_d), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
