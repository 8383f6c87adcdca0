import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    from dvs_mcemvs_b200 import api
    c = api.Context(0)
    yield c
    c.close()


class Case:
    """A synthetic multi-camera case with everything the oracle and the engine need."""

    def __init__(self, name, events_per_cam=None, seed=None, kind="structured", n_cams=None):
        from dvs_mcemvs_b200 import synth
        from oracle import oracle as Or
        self.scene, n_ev, self.method, self.desc = synth.config(name, events_per_cam, seed)
        sc = self.scene
        self.n_cams = n_cams or len(sc.rig.cams)
        self.cams = sc.rig.cams[:self.n_cams]
        self.events = [sc.events(i, n_ev, kind) for i in range(self.n_cams)]
        self.trajs = [sc.trajectory(i) for i in range(self.n_cams)]
        self.T_rv_w = sc.T_rv_w()
        sh, c0 = sc.shape, self.cams[0]
        self.shape = sh
        self.dimX, self.dimY, self.dimZ = sh.dimX_ or c0.width, sh.dimY_ or c0.height, sh.dimZ_
        self.depths = Or.depth_vector(sh.min_depth_, sh.max_depth_, sh.dimZ_, sh.inverse_depth)
        self.virts = [Or.virtual_camera(c.fx, c.cx, c.cy, self.dimX, sh.fov_) for c in self.cams]
        self.packets = [Or.packetize(self.events[i], self.trajs[i], self.T_rv_w,
                                     np.array([c.fx, c.fy, c.cx, c.cy], np.float32), self.virts[i], self.depths[0])
                        for i, c in enumerate(self.cams)]

    def oracle_dsi(self, i):
        from oracle import oracle as Or
        c = self.cams[i]
        return Or.build_dsi(self.events[i], self.packets[i], c.lut, c.width, self.depths, self.virts[i], self.dimX,
                            self.dimY)


@pytest.fixture(scope="session")
def small_case():
    return Case("esim_small")
