"""The drop-in boundary: libc5gpu.so loads, exports every symbol include/c5gpu.h declares, the
ctypes mirror has the C layout, and without a GPU the library fails loudly (no CPU path)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from course5_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "c5gpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c5_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpu_lib):
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(gpu_lib, name), f"{name} is declared in c5gpu.h but not exported"
    assert sorted(api.SYMBOLS) == names, "course5_b200.api.SYMBOLS is out of sync with include/c5gpu.h"
    header_version = int(re.search(r"#define\s+C5_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert gpu_lib.c5_abi_version() == header_version == api.ABI_VERSION


def test_build_entry_point_checks_the_same_version():
    """__graft_entry__.build() must compare against api.ABI_VERSION, not a literal that goes stale."""
    text = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "api.ABI_VERSION" in text and not re.search(r"c5_abi_version\(\)\s*==\s*\d", text)


def test_ctypes_structs_match_the_c_layout(tmp_path):
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "c5gpu.h"\n'
        "int main(void){\n"
        'printf("%zu %zu %zu %zu %zu\\n", sizeof(c5_view), sizeof(c5_stats), sizeof(c5_mesh_info), '
        "sizeof(c5_rotation), offsetof(c5_view, alpha_limit));\n"
        'printf("%zu %zu %zu\\n", offsetof(c5_view, rot), offsetof(c5_view, row_begin), offsetof(c5_stats, ms_rotate));\n'
        'printf("%zu\\n", offsetof(c5_stats, ms_graze));\n'
        "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(api.View), C.sizeof(api.Stats), C.sizeof(api.MeshInfo), C.sizeof(api.Rotation),
            api.View.alpha_limit.offset, api.View.rot.offset, api.View.row_begin.offset,
            api.Stats.ms_rotate.offset, api.Stats.ms_graze.offset]
    assert got == want


def test_view_from_flags_restates_main_cpp(gpu_lib):
    v = api.make_view(1200, 900, X=0.5, Y=0.25, I=-0.03, alpha_limit=3.0, lib=gpu_lib)
    assert list(v.window) == [2.2, -0.2, 0.9, -0.9]                      # main.cpp:83
    a0 = 0.03 * api.PI + api.PI / 2.0                                    # main.cpp:96
    assert v.n_rot == 3
    assert (v.rot[0].axis, v.rot[0].angle) == (0, a0)
    assert (v.rot[1].axis, v.rot[1].angle, v.rot[1].x0) == (1, 0.25 * api.PI, 1.0)
    assert (v.rot[2].axis, v.rot[2].angle) == (0, -a0 + 0.5 * api.PI)
    assert v.alpha_limit == 3.0 and v.round_through_float == 1 and v.use_solids == 1 and v.precision == 64


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_gpu_is_a_loud_error_not_a_fallback(gpu_lib):
    with pytest.raises(api.C5Error) as e:
        api.Context(devices=(0,), lib=gpu_lib)
    assert e.value.code == api.E_CUDA
    assert "no CPU path" in str(e.value)


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        api.load_library(str(tmp_path / "libc5gpu.so"))


def test_balanced_bands_cover_rows_contiguously():
    cost = np.zeros(900)
    cost[300:600] = 1000.0
    bands = api.balanced_bands(cost, 8, base_cost=1.0)
    assert bands[0][0] == 0 and bands[-1][1] == 900
    assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    assert all(hi > lo for lo, hi in bands)
    per_band = [cost[lo:hi].sum() + (hi - lo) for lo, hi in bands]
    assert max(per_band) < 1.2 * (sum(per_band) / 8)
    assert api.balanced_bands(cost, 1) == [(0, 900)]
