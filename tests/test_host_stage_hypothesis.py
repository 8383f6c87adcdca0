"""Runs tests/host_stage_hypothesis_impl.py twice in fresh processes: with the library's defaults (short lists stay on the
sequential packet loop) and with the speculative multi-threaded path forced onto short lists."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"EMVS_PACKET_MIN_BATCH": "2", "EMVS_HOST_THREADS": "3"},
                                 {"EMVS_PACKET_MIN_BATCH": "1", "EMVS_HOST_THREADS": "8"}],
                         ids=["defaults", "speculative-3-threads", "speculative-8-threads"])
def test_packet_stage_properties(env):
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "host_stage_hypothesis_impl.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "1 passed" in r.stdout
