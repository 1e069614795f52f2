"""The C-ABI library builds, loads and exports every symbol include/wtracker_b200.h declares
(no compute calls: there is no GPU on the CPU test box)."""
import ctypes
import os
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_are_exported(native_lib):
    header = (ROOT / "include" / "wtracker_b200.h").read_text()
    declared = set(re.findall(r"\b(wt_[a-z0-9_]+)\s*\(", header))
    declared -= {"wt_engine"}
    assert len(declared) == 26
    from wtracker_b200 import _lib

    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(native_lib, name), name


def test_abi_version_and_error_string(native_lib):
    import re

    from wtracker_b200 import _lib as L

    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "wtracker_b200.h")).read()
    assert native_lib.wt_abi_version() == L.ABI_VERSION == int(re.search(r"#define WT_ABI_VERSION (\d+)", header).group(1))
    assert isinstance(native_lib.wt_last_error(), bytes)
    assert native_lib.wt_launch_count() >= 0


def test_struct_sizes_match_header(native_lib):
    from wtracker_b200 import _lib as L

    # sizes printed by a C program compiled against include/wtracker_b200.h (x86-64 SysV)
    expect = {"WtLetterbox": 64, "WtBuf": 16, "WtOp": 120, "WtPostParams": 40, "WtHeadLevel": 96, "WtResmlpDesc": 72,
              "WtTailArgs": 280}
    for name, size in expect.items():
        assert ctypes.sizeof(getattr(L, name)) == size, name


def test_struct_sizes_against_compiled_header(tmp_path):
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        return
    src = tmp_path / "sz.c"
    src.write_text(f'#include "{ROOT}/include/wtracker_b200.h"\n#include <stdio.h>\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu", sizeof(wt_letterbox), sizeof(wt_buf), sizeof(wt_op),'
                   'sizeof(wt_post_params), sizeof(wt_head_level), sizeof(wt_resmlp_desc), sizeof(wt_tail_args)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from wtracker_b200 import _lib as L

    got = [ctypes.sizeof(c) for c in (L.WtLetterbox, L.WtBuf, L.WtOp, L.WtPostParams, L.WtHeadLevel, L.WtResmlpDesc,
                                      L.WtTailArgs)]
    assert sizes == got


def test_compute_calls_fail_loudly_without_gpu(native_lib):
    import torch

    if torch.cuda.is_available():
        return
    from wtracker_b200.eval.error_calculator import ErrorCalculator
    import numpy as np
    import pytest

    with pytest.raises(RuntimeError):
        ErrorCalculator.calculate_bbox_error(np.zeros((2, 4)), np.zeros((2, 4)))


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from wtracker_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NativeLibraryError):
        _lib.lib()
