"""Static properties of the built vote kernel, read from the cubin inside libemvs_b200.so (no GPU needed).

The vote loop sits at the RED-path ceiling only while it stays lean: a refactoring of the persistent kernel once took it
from 65 to 85 SASS instructions per vote without any source change in the loop itself (ptxas re-materialised thread
indices at 32 registers, profiles/r2_multislab.md).  These checks catch that class of regression at build time:
registers, no local-memory spills, the TMA / mbarrier / vector-RED instructions DESIGN.md §4.2 names, and the
instruction distance between consecutive votes of the unrolled loop."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dvs_mcemvs_b200", "lib", "libemvs_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)),
                                reason="needs the built library and cuobjdump")

VOTE = "k_vote_tmaILi8ELi2E"      # k_vote_tma<8, 2>: 8 planes per instruction, evict_last REDs — the default build path


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=300).stdout
    funcs, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
            if m:
                funcs[name].append(m.group(1).strip())
    return funcs


def _one(funcs, key):
    names = [n for n in funcs if key in n]
    assert len(names) == 1, names
    return funcs[names[0]]


def test_vote_kernel_resources():
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, timeout=120).stdout
    lines = out.splitlines()
    usage = {}
    for i, ln in enumerate(lines):
        m = re.search(r"Function (\S+):", ln)
        if m and i + 1 < len(lines):
            usage[m.group(1)] = dict(kv.split(":") for kv in re.findall(r"[A-Z]+(?:\[\d\])?:\d+", lines[i + 1]))
    votes = {n: u for n, u in usage.items() if "k_vote_tmaILi" in n}
    assert len(votes) >= 10                                   # G = 2..32 x two RED policies
    for n, u in votes.items():
        # 7 CTAs of 256 threads per SM + one slot for the merge / exchange kernels: 8 x 256 x 32 = the register file
        assert int(u["REG"]) <= 32, (n, u)
        assert int(u["LOCAL"]) == 0, (n, u)
    assert int(votes[[n for n in votes if VOTE in n][0]]["STACK"]) == 0     # no spills around the out-of-line vote loop


def test_vote_kernel_uses_tma_mbarrier_and_vector_reds(sass):
    ins = _one(sass, VOTE)
    ops = [i.split()[0] if not i.startswith("@") else i.split()[1] for i in ins]
    assert any(o.startswith("UBLKCP") for o in ops)                                   # cp.async.bulk (TMA 1-D) event tiles
    assert any(o.startswith("SYNCS.ARRIVE.TRANS64") for o in ops)                     # mbarrier expect-tx
    assert any(o.startswith("SYNCS.PHASECHK.TRANS64.TRYWAIT") for o in ops)           # mbarrier wait
    reds = [o for o in ops if o.startswith("REDG.E.ADD.F32x4")]
    assert len(reds) >= 8                                                             # one 16-byte RED per vote, unrolled by 8
    assert not any(o.startswith(("LDL", "STL")) for o in ops)                         # no local memory at all


def test_vote_loop_instruction_budget(sass):
    ins = _one(sass, VOTE)
    pos = [k for k, i in enumerate(ins) if "REDG.E.ADD.F32x4" in i]
    gaps = sorted(b - a for a, b in zip(pos, pos[1:]))
    main = gaps[: max(1, len(gaps) // 2)]                     # the unrolled main loop (the remainder loop's gaps are larger)
    per_vote = sum(main) / len(main)
    # 66 today (two IEEE divisions, range test, floor, four weights, index, RED); 78-85 was the regression
    assert per_vote <= 70, (per_vote, gaps)


def test_classic_and_merge_kernels_present(sass):
    assert any("k_vote_groupedILi8ELb0E" in n for n in sass)
    assert any("k_merge_quads_groupedILi8E" in n for n in sass)
    assert any("k_wait_count" in n for n in sass)
    assert any("k_peer_reduce_band_v4" in n for n in sass)
    fc = [n for n in sass if "k_fuse_collapse_zsplit_v4" in n]
    assert fc
    ins = sass[fc[0]]
    assert any(re.search(r"LDG\.E(\.[A-Z]+)*\.128", i) for i in ins)   # four pixels per thread: 16-byte loads
