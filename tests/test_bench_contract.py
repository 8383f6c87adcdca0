"""bench.py's JSON contract, checked on the CPU arm (`--impl reference` never touches the GPU): one line on stdout with
every key the driver reads; the B200 arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # EMVS_B200_FORBID_LOAD: the CPU arm must not map libemvs_b200.so (its DSEC rectification LUT comes from cv2)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--events-per-cam", "60000"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, EMVS_B200_FORBID_LOAD="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mevents/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32" and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "dsec_stereo" and d["config"]["dsi"] == [640, 480, 256] and d["config"]["fusion"] == "harmonic"
    # the keys both arms print (the driver compares the two `config` dicts)
    assert sorted(d["config"]) == sorted(["workload", "cameras", "dsi", "fusion", "event_distribution", "scaling", "gpus",
                                          "events_per_camera", "events_per_camera_per_gpu", "l2"])
    assert d["config"]["events_per_camera"] == 60000 and d["config"]["gpus"] == 1 and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "Mevents/s" and "sample" in cb and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout) and r.stdout.strip() == ""
