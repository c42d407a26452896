"""The C ABI is usable from a plain-C host (no Python, no torch types at the boundary): tools/c_host/schedule_demo.c is
compiled with gcc against include/dit4sr_b200.h, linked to the in-tree library and run (host-only entry points)."""
import os
import shutil
import subprocess

import pytest

from dit4sr_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_host_links_and_runs(tmp_path):
    build.build_library()
    libdir = os.path.join(ROOT, "dit4sr_b200")
    exe = str(tmp_path / "demo")
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "c_host", "schedule_demo.c"), "-L", libdir, "-ldit4sr_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.check_output([exe], text=True)
    assert out.startswith("abi 5 items 984 sliced 96 "), out      # 888 whole tiles + 24 remainder tiles in 4 slices each
