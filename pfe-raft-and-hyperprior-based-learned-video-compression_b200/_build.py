"""Builds ``lib/librdvc_corr.so`` from ``csrc/`` with nvcc for sm_100a, in-tree.

The shared object is git-ignored but travels to the GPU box with the gpurun
snapshot.  nvcc cross-compiles without a GPU, so this runs in the CPU-only
build container too (``__graft_entry__.build()``).

Staleness is decided by CONTENT, not by mtime: a SHA-256 over every file under
``csrc/`` and ``include/`` is compiled into the library (``rdvc_corr_build_info()``,
marker ``RDVC_SRC_HASH=``); ``is_stale()`` reads the marker out of the file and
compares it with the tree, ``_cabi.load()`` refuses a library whose hash differs.

``python -m ..._build --experiments`` (or ``build(experiments=True)``) compiles a
SEPARATE library, ``lib/librdvc_corr_exp.so``, with ``-DRDVC_EXPERIMENTS``: the
timing knobs that skip work and the build variants that lost their measurements
(fused pooling epilogue, CTA-pair kernel).  Point ``RDVC_CORR_LIB`` at it to run
the tests that cover them; the product library contains none of that.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librdvc_corr.so")
EXP_LIB_PATH = os.path.join(LIB_DIR, "librdvc_corr_exp.so")
INCLUDE = os.path.abspath(os.path.join(PKG_DIR, "..", "include"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xfatbin", "-compress-all",      # 10 MB -> 4 MB of device code in the .so (it travels to the GPU box with every snapshot)
]

_MARKER = b"RDVC_SRC_HASH="


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(INCLUDE, "rdvc_corr.h"))
    return sorted(out)


def source_hash() -> str:
    """SHA-256 over (relative name, content) of every source of the library."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())          # a change of compiler flags is a change of the binary
    for path in sources():
        h.update(os.path.relpath(path, os.path.join(PKG_DIR, "..")).encode())
        h.update(b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    return h.hexdigest()


def embedded_hash(lib_path: str):
    """The source hash compiled into a built library (read from the file, without loading it), or None."""
    try:
        with open(lib_path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    m = re.search(re.escape(_MARKER) + rb"([0-9a-f]{64})", blob)
    return m.group(1).decode() if m else None


def is_stale(lib_path: str = LIB_PATH) -> bool:
    return embedded_hash(lib_path) != source_hash()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build librdvc_corr.so")


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """Compile the CUDA library unless a library built from exactly these sources is already there."""
    out = EXP_LIB_PATH if experiments else LIB_PATH
    if not force and not is_stale(out):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = out + ".tmp"
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        f'-DRDVC_SRC_HASH="{source_hash()}"',
    ] + (["-DRDVC_EXPERIMENTS"] if experiments else []) + [
        "-I", INCLUDE, "-o", tmp, os.path.join(CSRC, "rdvc_corr_abi.cu"),
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, out)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
