"""Property-based check of the product's host packet stage (no GPU): for arbitrary event counts, trajectory coverage
(pose misses at the head, at the tail, everywhere) and arbitrary sequences of event limits, emvs_packetize_range
reproduces the oracle's sequential packet loop bit for bit (mapper_emvs_stereo.cpp:86-126).

Run by tests/test_host_stage_hypothesis.py in a subprocess with EMVS_PACKET_MIN_BATCH=2 / EMVS_HOST_THREADS=3, so that
these short lists take the speculative multi-threaded path (the library reads both once per process)."""
import ctypes as C

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from dvs_mcemvs_b200 import _capi as capi
from dvs_mcemvs_b200 import synth

COMMON = dict(deadline=None, max_examples=40, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _traj(times, xs):
    tr = np.zeros(len(times), capi.STAMPED_POSE_DTYPE)
    tr["sec"], tr["nsec"] = synth._split_time(np.asarray(times, np.float64))
    tr["T"]["q"][:, 0] = 1.0
    tr["T"]["t"][:, 0] = xs
    return tr


@settings(**COMMON)
@given(n=st.integers(1024, 7000), t_lo=st.floats(99.0, 100.6), span=st.floats(0.05, 2.0), n_ctrl=st.integers(2, 9),
       cuts=st.lists(st.integers(0, 7000), min_size=0, max_size=6), seed=st.integers(0, 2**31 - 1),
       max_packets=st.integers(0, 8))
def test_packetize_range_any_limits(O, n, t_lo, span, n_ctrl, cuts, seed, max_packets):
    rng = np.random.default_rng(seed)
    ts = np.sort(rng.uniform(100.0, 101.0, n))
    ev = np.zeros(n, capi.EVENT_DTYPE)
    ev["sec"], ev["nsec"] = synth._split_time(ts)
    tr = _traj(np.linspace(t_lo, t_lo + span, n_ctrl), np.linspace(0.0, 0.1, n_ctrl))
    K = np.array([200, 200, 120, 90], np.float32)
    I = np.zeros((), capi.POSE_DTYPE)
    I["q"] = (1, 0, 0, 0)
    cam = capi.Camera(240, 180, 200, 200, 120, 90)
    want = O.packetize(ev, tr, I, K, K.copy(), 1.0)
    lib = capi.load()
    limits = sorted(min(c, n) for c in cuts) + [n]
    cur, got = C.c_size_t(0), []
    for lim in limits:
        out = np.zeros(n // 1024 + 1, capi.PACKET_DTYPE)
        k = C.c_size_t(0)
        capi.check(lib.emvs_packetize_range(capi.ptr(ev), n, capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam), capi.ptr(K), 1.0,
                                            C.byref(cur), lim, capi.ptr(out), len(out), C.byref(k)))
        got.append(out[:k.value].copy())
    assert np.concatenate(got).tobytes() == want.tobytes()
    # a bounded output buffer returns the first max_packets packets of the same sequence
    out = np.zeros(max(max_packets, 1), capi.PACKET_DTYPE)
    k = C.c_size_t(0)
    rc = lib.emvs_packetize(capi.ptr(ev), n, capi.ptr(tr), len(tr), capi.ptr(I), C.byref(cam), capi.ptr(K), 1.0,
                            capi.ptr(out), max_packets, C.byref(k))
    assert k.value == min(max_packets, len(want)) and out[:k.value].tobytes() == want[:k.value].tobytes()
    # ... and says so: a list that did not fit is an error (never a silently shortened packet list); when the buffer
    # was big enough for every packet AND for noticing that nothing is left, the call succeeds
    if max_packets < len(want):
        assert rc == capi.EMVS_ERR_INVALID and b"max_packets" in lib.emvs_last_error()
    elif max_packets > len(want):
        assert rc == capi.EMVS_OK
