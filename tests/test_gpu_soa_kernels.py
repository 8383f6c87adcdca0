"""Round-2 entry points and kernel variants through the C-ABI (pytest -m gpu): structure-of-arrays event lists,
the TMA-staged persistent vote kernel against the one-CTA-per-packet kernel it replaced, the persistent-grid size."""
import os

import numpy as np
import pytest

from dvs_mcemvs_b200 import api

pytestmark = pytest.mark.gpu

DSI_RTOL, DSI_ATOL = 1e-5, 1e-5


def _oracle(case, i):
    return case.oracle_dsi(i)


def test_soa_equals_aos_and_oracle(ctx, small_case):
    """evaluateDSI on (x, y, t_ns) arrays == evaluateDSI on dvs_msgs::Event structs == the oracle; the prefetched
    form too (events + packets ahead of time, host arrays scribbled afterwards)."""
    import time
    tr = api.LinearTrajectory(small_case.trajs[0])
    dsi_o, inb_o = _oracle(small_case, 0)
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    try:
        soa = api.EventsSoA.from_events(small_case.events[0], pinned=True)
        assert m.evaluateDSI(soa, tr, small_case.T_rv_w) is True
        assert np.array_equal(m.counts(), inb_o)
        np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
        pk = m.packetize(soa, tr, small_case.T_rv_w)
        assert pk.tobytes() == small_case.packets[0].tobytes()
        # split upload on the SoA path (idle pipeline, tiny threshold)
        ctx.sync()
        ctx.set_upload_split(30, 4096)
        assert m.evaluateDSI(soa, tr, small_case.T_rv_w) is True
        ctx.set_upload_split(15)
        assert np.array_equal(m.counts(), inb_o)
        np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
        # prefetch: after it landed neither x, y nor t on the host matter any more
        m.prefetch(soa, tr, small_case.T_rv_w)
        assert ctx.prefetch_pending() > 0
        time.sleep(0.2)
        keep = soa.x.copy(), soa.t_ns.copy()
        soa.x[...] = 0
        soa.t_ns[...] = 0
        assert m.evaluateDSI(soa, tr, small_case.T_rv_w) is True
        soa.x[...], soa.t_ns[...] = keep
        assert np.array_equal(m.counts(), inb_o)
        np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=DSI_RTOL, atol=DSI_ATOL)
        # an AoS call after an SoA announcement of other arrays drops it
        m.prefetch(soa, tr, small_case.T_rv_w)
        assert m.evaluateDSI(small_case.events[0], tr, small_case.T_rv_w) is True
        assert ctx.prefetch_pending() == 0
        assert np.array_equal(m.counts(), inb_o)
        # too few events -> False like the AoS path
        short = api.EventsSoA(soa.x[:500], soa.y[:500], soa.t_ns[:500])
        assert m.evaluateDSI(short, tr, small_case.T_rv_w) is False
    finally:
        ctx.prefetch_cancel()
        m.close()


@pytest.mark.parametrize("env", [{"EMVS_VOTE_MULTISLAB": "1"}, {"EMVS_VOTE_MULTISLAB": "0"}, {"EMVS_SLAB": "24"}, {"EMVS_MULTISLAB_BUDGET_MB": "1"},
                                 {"EMVS_VOTE_KERNEL": "classic"}, {"EMVS_VOTE_CTAS_PER_SM": "1"}, {"EMVS_VOTE_CTAS_PER_SM": "8"},
                                 {"EMVS_ZERO_CTAS": "0", "EMVS_HINT_RED": "0", "EMVS_HINT_XY0": "0", "EMVS_HINT_DSI": "0"},
                                 {"EMVS_HINT_XY0": "2", "EMVS_HINT_ZERO": "1"}, {"EMVS_VOTE_GROUP": "4"}, {"EMVS_VOTE_GROUP": "16"},
                                 {"EMVS_VOTE_SPLIT": "0"}, {"EMVS_VOTE_SPLIT": "2"}, {"EMVS_FC_V4": "0", "EMVS_FC_ZSPLIT": "4"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_vote_kernel_variants_agree(small_case, env):
    """The TMA-staged persistent kernel (default), the classic one-CTA-per-packet kernel and other grid / group /
    re-zero settings produce the same counts bit for bit and the same DSI up to float summation order.  (The
    settings are read when a context is created / when the library first builds, so each runs in a fresh process.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from conftest import Case\n"
        "from dvs_mcemvs_b200 import api\n"
        "case = Case('esim_small')\n"
        "ctx = api.Context(0)\n"
        "for i in range(2):\n"
        "    m = api.MapperEMVS(ctx, case.cams[i], case.shape)\n"
        "    m.build(case.events[i], case.packets[i])\n"
        "    dsi_o, inb_o = case.oracle_dsi(i)\n"
        "    assert np.array_equal(m.counts(), inb_o)\n"
        "    np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=1e-5, atol=1e-5)\n"
        "    m.build(case.events[i], case.packets[i][:7])\n"       # fewer packets than resident CTAs
        "    m.build(case.events[i], case.packets[i])\n"           # and the scratch is clean again afterwards
        "    assert np.array_equal(m.counts(), inb_o)\n"
        "    np.testing.assert_allclose(m.dsi_.download(), dsi_o, rtol=1e-5, atol=1e-5)\n"
        "    conf, idx, depth = m.dsi_.collapseMaxZSlice(m.raw_depths_vec_)\n"
        "    vol = m.dsi_.download()\n"
        "    assert np.array_equal(conf, vol.max(0)) and np.array_equal(idx, vol.argmax(0)) and np.array_equal(depth, m.raw_depths_vec_[idx])\n"
        "    m.close()\n"
        "print('variant ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_new_entry_points_report_errors(ctx, small_case):
    """Round-2 entry points keep the C-ABI's conventions: bad arguments are status codes with a message, never aborts."""
    import ctypes as C
    from dvs_mcemvs_b200 import _capi as capi
    lib = capi.load()
    m = api.MapperEMVS(ctx, small_case.cams[0], small_case.shape)
    tr = api.LinearTrajectory(small_case.trajs[0])
    T = np.ascontiguousarray(small_case.T_rv_w, dtype=capi.POSE_DTYPE)
    soa = api.EventsSoA.from_events(small_case.events[0])
    try:
        bad = capi.EventsSoA(soa.x.ctypes.data, None, soa.t_ns.ctypes.data, len(soa))
        assert lib.emvs_mapper_evaluate_dsi_soa(m._h, C.byref(bad), capi.ptr(tr.poses), len(tr.poses), capi.ptr(T), 0) == capi.EMVS_ERR_INVALID
        assert b"NULL event arrays" in lib.emvs_last_error()
        es = soa.c_struct()
        assert lib.emvs_mapper_evaluate_dsi_soa(m._h, C.byref(es), capi.ptr(tr.poses), 1, capi.ptr(T), 0) == capi.EMVS_ERR_INVALID
        assert lib.emvs_mapper_prefetch_dsi_soa(m._h, None, capi.ptr(tr.poses), len(tr.poses), capi.ptr(T)) == capi.EMVS_ERR_INVALID
        gen = C.c_uint64(7)
        assert lib.emvs_context_prefetch_pending(ctx._h, C.byref(gen)) == capi.EMVS_OK and gen.value == 0
        assert lib.emvs_context_prefetch_pending(ctx._h, None) == capi.EMVS_ERR_INVALID
        assert lib.emvs_context_prefetch_cancel(ctx._h) == capi.EMVS_OK          # nothing pending: a no-op
        # a one-rank exchange: the participants table is validated
        arr = (C.c_void_p * 1)(m.dsi_._h)
        ex = C.c_void_p()
        assert lib.emvs_exchange_create(ctx._h, arr, 1, 1, 0, C.byref(ex)) == capi.EMVS_OK
        none = np.zeros(1, np.uint8)
        assert lib.emvs_exchange_set_participants(ex, capi.ptr(none)) == capi.EMVS_ERR_INVALID   # a camera nobody builds
        assert lib.emvs_exchange_set_participants(ex, None) == capi.EMVS_ERR_INVALID
        one = np.ones(1, np.uint8)
        assert lib.emvs_exchange_set_participants(ex, capi.ptr(one)) == capi.EMVS_OK
        assert lib.emvs_exchange_destroy(ex) == capi.EMVS_OK
        # PEER_REDUCE without an active exchange round is a state error, not a crash
        with pytest.raises(api.EmvsError):
            m.build(small_case.events[0], small_case.packets[0], peer_reduce=True)
    finally:
        m.close()
