"""The ctypes structures of dit4sr_b200/_lib.py must have exactly the layout of the C structs in include/dit4sr_b200.h:
a plain-C translation unit (gcc, no CUDA) prints sizeof / offsetof of every field and the test compares them with ctypes.
(A mismatch would not fail to load - it would silently pass garbage pointers to the kernels.)"""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

from dit4sr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = {
    "dit4sr_ln_modulate_params": _lib.LnModulateParams,
    "dit4sr_dwconv3x3_params": _lib.DwconvParams,
    "dit4sr_gemm_params": _lib.GemmParams,
    "dit4sr_gemm_sp": _lib.GemmSp,
    "dit4sr_attn_params": _lib.AttnParams,
    "dit4sr_attn_sp_out": _lib.AttnSpOut,
    "dit4sr_sp_comm": _lib.SpComm,
    "dit4sr_a2a_params": _lib.A2aParams,
    "dit4sr_halo_params": _lib.HaloParams,
}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ctypes_structs_match_the_c_header(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dit4sr_b200.h"', "int main(void) {"]
    for cname, cls in STRUCTS.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    seen = set()
    for cname, field, value in re.findall(r"(\w+) (\w+) (\d+)", out):
        cls = STRUCTS[cname]
        if field == "size":
            assert ctypes.sizeof(cls) == int(value), (cname, ctypes.sizeof(cls), value)
        else:
            assert getattr(cls, field).offset == int(value), (cname, field, getattr(cls, field).offset, value)
        seen.add((cname, field))
    assert len(seen) == sum(len(c._fields_) + 1 for c in STRUCTS.values())


def test_header_struct_fields_are_all_bound():
    """Every field the header declares in these structs exists in the ctypes class, in the same order."""
    hdr = open(os.path.join(ROOT, "include", "dit4sr_b200.h")).read()
    for cname, cls in STRUCTS.items():
        end = hdr.index("} " + cname + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [re.search(r"(\w+)\s*(\[\w+\])?\s*$", d.strip()).group(1) for d in body.split(";") if d.strip()]
        assert names == [f for f, _ in cls._fields_], (cname, names)
