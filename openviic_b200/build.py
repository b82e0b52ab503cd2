"""Builds libopenviic_cap.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

The shared object lands in ``openviic_b200/lib/`` so that it travels to the GPU box with the
repository snapshot.  No torch headers are involved: the library is a plain C-ABI .so that the
Python side binds with ctypes (``openviic_b200/cabi.py``).
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libopenviic_cap.so"
STAMP = LIB_DIR / "libopenviic_cap.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


HOST_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread"]   # csrc/*.cpp: host-only code, compiled by g++


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for path in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "openviic_cap.h"]):
        h.update(path.name.encode())
        h.update(path.read_bytes())
    h.update(" ".join(NVCC_FLAGS + HOST_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libopenviic_cap.so")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu into one shared library; no-op when sources are unchanged."""
    LIB_DIR.mkdir(exist_ok=True)
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = obj_dir / (src.stem + ".o")
        if src.suffix == ".cpp":
            cuda_include = str(Path(nvcc).resolve().parent.parent / "include")
            cmd = [shutil.which("g++") or "g++", *HOST_FLAGS, "-I", cuda_include, "-c", str(src), "-o", str(obj)]
        else:
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(str(obj))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-cudart", "static"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
