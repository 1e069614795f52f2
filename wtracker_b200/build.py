"""Builds the C-ABI shared library (libwtracker_b200.so) in-tree with nvcc for sm_100a.

Run as ``python -m wtracker_b200.build`` (or through ``__graft_entry__.build()``).  Objects are
compiled in parallel and re-used when the source is older than the object.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OUT_DIR = PKG_DIR / "_native"
LIB_PATH = OUT_DIR / "libwtracker_b200.so"

SOURCES = [
    "common.cu",
    "conv_tcgen05.cu",
    "ops_simt.cu",
    "pre.cu",
    "post.cu",
    "resmlp.cu",
    "tail.cu",
    "metrics.cu",
    "log.cu",
    "precise.cu",
    "analysis.cu",
    "engine.cu",
]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the wtracker_b200 native library cannot be built")
    return nvcc


def _newer(src: Path, obj: Path, deps: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *deps])


def build(verbose: bool = False, force: bool = False, tuning: bool = False) -> Path:
    """``tuning``: a second library, ``_native/tuning/libwtracker_b200.so``, compiled with -DWT_TUNING_KNOBS so that the
    WT_* environment variables select kernel variants (A/B runs: tools/gpu_r2c.sh points WTRACKER_B200_LIB at it).
    The product library never reads the environment."""
    nvcc = find_nvcc()
    out_dir = OUT_DIR / "tuning" if tuning else OUT_DIR
    out_dir.mkdir(parents=True, exist_ok=True)
    lib_path = out_dir / LIB_PATH.name
    headers = sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "wtracker_b200.h"]

    def compile_one(name: str) -> Path:
        src = CSRC / name
        obj = out_dir / (src.stem + ".o")
        if force or _newer(src, obj, headers):
            cmd = [nvcc, *NVCC_FLAGS, *(["-DWT_TUNING_KNOBS"] if tuning else []), "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}:\n{res.stdout}\n{res.stderr}")
            if verbose and res.stderr:
                print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))

    if force or not lib_path.exists() or any(o.stat().st_mtime > lib_path.stat().st_mtime for o in objs):
        cmd = [nvcc, "-shared", "-o", str(lib_path), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib_path


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv, tuning="--tuning" in sys.argv)
    print(path)
