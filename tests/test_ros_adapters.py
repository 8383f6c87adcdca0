"""dvs_mcemvs_b200/host/ros_adapters.hpp (the conversions a maintainer uses on the reference side) is
inert in this image — no ROS, minkindr or image_geometry headers.  Test shims with the same type and member names
(tests/shims/ros/) let it compile here: the event layout, the pose and camera conversions run on the CPU, and the
adapter's evaluateDSI is type-checked against the host mirror.  (`adapters_check gpu` also calls it; that leg is a manual
check on a GPU box, not part of the suite.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "dvs_mcemvs_b200", "lib")


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(LIBDIR, "libemvs_b200.so")),
                    reason="needs g++ and the built library")
def test_ros_adapters_compile_and_convert(tmp_path):
    exe = str(tmp_path / "adapters_check")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "dvs_mcemvs_b200", "host"), "-I", os.path.join(ROOT, "tests", "shims", "ros"),
           "-o", exe, os.path.join(ROOT, "tests", "shims", "ros", "adapters_check.cpp"),
           "-L", LIBDIR, "-lemvs_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ros adapters ok" in r.stdout, r.stdout + r.stderr


def test_without_the_shims_the_header_is_inert(tmp_path):
    """In a tree without the ROS headers the adapter header must compile to nothing (it is included unconditionally by
    maintainers' translation units that are built in both environments)."""
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    src = tmp_path / "inert.cpp"
    src.write_text('#include "ros_adapters.hpp"\n#ifdef EMVS_HOST_HAVE_ROS\n#error "ROS headers unexpectedly found"\n#endif\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "dvs_mcemvs_b200", "host"), str(src)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
