"""Launches tests/mgpu_check.py under torchrun when the box has >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharded_build_matches_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mgpu_check ok" in r.stdout


def test_two_rank_alg2_sharded_matches_process_2():
    """BASELINE.json configs[3], second ordering: one sub-interval per GPU, cameras fused locally, ONE allreduce of the
    fused volume across time (api.process_2_sharded) == the single-GPU process_2 with num_subintervals = world
    (process2.cpp:98-249)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29518", os.path.join(ROOT, "tests", "mgpu_check_alg2.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mgpu_check_alg2 ok" in r.stdout


def test_two_rank_bench_parity_weak_and_strong():
    """bench.py's own parity block at N = 2 (exchanged maps and summed counts vs an unsharded build on rank 0), for
    the weak-scaling default and for the fixed-total (strong) split of BASELINE.json configs[3] under camera x
    sub-interval sharding (the default), and for sub-interval-only sharding, at a reduced size."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for extra in (["--scaling", "weak", "--events-per-cam", "300000"], ["--scaling", "strong", "--events-per-cam", "600000"],
                  ["--scaling", "weak", "--events-per-cam", "300000", "--sharding", "interval"]):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
               "--warmup", "1", "--no-cpu-baseline"] + extra
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        assert d["n_gpus"] == 2 and d["scaling"] == extra[1]
        p = d["parity"]
        assert p["ok"] and p["counts_exact"] and p["maps_identical_on_all_ranks"], p
        assert p["e2e_ok"] and p["e2e_streaming_ok"] and p["e2e_soa_ok"], p


def test_host_cpp_mirror_example():
    """The C++ caller written against the reference's class API (Grid3D, EMVS::MapperEMVS, LinearTrajectory, process_1,
    getDepthMapFromDSI with options, writeGridNpy) passes its known-answer checks."""
    import numpy as np
    exe = os.path.join(ROOT, "dvs_mcemvs_b200", "host", "example_process1")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "example_process1 ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    vol = np.load("/tmp/emvs_example_dsi.npy")                     # what the reference's viewers read
    assert vol.shape == (64, 180, 240) and vol.dtype == np.float32 and vol.sum() > 1e4


def test_host_cpp_process_2_and_5_example():
    """process_2 / process_5 (Alg. 2 and its shuffled variant) through the C++ mirror: AM commutes across the two
    fusion orders, the reference's id swap in the time-then-camera switch, process_5 == process_2 for one
    sub-interval, error returns."""
    exe = os.path.join(ROOT, "dvs_mcemvs_b200", "host", "example_process2")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "example_process2 ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
