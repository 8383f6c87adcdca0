"""The C-ABI library loads and exports every symbol include/emvs_b200.h declares (no GPU)."""
import ctypes as C
import re
import subprocess

import numpy as np

from dvs_mcemvs_b200 import _capi as capi


def _declared():
    src = open(capi.HEADER_PATH).read()
    return sorted(set(re.findall(r"EMVS_API\s+[\w\s\*]+?\b(emvs_\w+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (emvs_\w+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared but not exported: {missing}"
    extra = sorted(exported - set(names))
    assert not extra, f"exported but not declared in the header: {extra}"


def test_python_binding_covers_header():
    assert sorted(capi._PROTOTYPES) == _declared()
    lib = capi.load()
    assert lib.emvs_abi_version() == 2


def test_pod_layouts():
    assert C.sizeof(capi.Shape) == 28 and C.sizeof(capi.Camera) == 24
    assert capi.EVENT_DTYPE.itemsize == 16
    assert capi.EVENT_DTYPE.fields["sec"][1] == 4 and capi.EVENT_DTYPE.fields["polarity"][1] == 12
    assert capi.PACKET_DTYPE.fields["C"][1] == 36 and capi.PACKET_DTYPE.fields["first_event"][1] == 48
    assert capi.STAMPED_POSE_DTYPE.fields["T"][1] == 8


def test_no_gpu_fails_loudly():
    """Without a usable device the product refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        return
    lib = capi.load()
    h = C.c_void_p()
    rc = lib.emvs_context_create(0, C.byref(h))
    assert rc == capi.EMVS_ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.emvs_last_error()


def test_product_never_touches_oracle():
    import os
    root = os.path.dirname(os.path.dirname(capi.HEADER_PATH))
    pkg = os.path.join(root, "dvs_mcemvs_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|oracle/|emvs_oracle|libemvs_oracle", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, f"product files reference the oracle: {bad}"
