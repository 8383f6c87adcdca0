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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import Case
        from oracle import oracle as O
        case = Case("esim_small", events_per_cam=30_000)
        vols, counts = [], []
        for cam, lo, hi in shard.plan([len(p) for p in case.packets], world, rank):
            c = case.cams[cam]
            dsi, inb = O.build_dsi(case.events[cam], case.packets[cam][lo:hi], c.lut, c.width, case.depths,
                                   case.virts[cam], case.dimX, case.dimY)
            t, ti = torch.from_numpy(dsi), torch.from_numpy(inb.astype(np.int64))
            dist.all_reduce(t)        # the DSI sum-exchange (NCCL on the GPU path)
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


def test_sharded_equals_unsharded_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
