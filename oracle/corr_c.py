"""TEST INFRASTRUCTURE ONLY -- ctypes binding of ``oracle/corr_oracle.c``.

``build()`` compiles the scalar C restatement with gcc into
``oracle/_build/libcorr_oracle.so`` (git-ignored, travels to the GPU box with
the gpurun snapshot).  See corr_oracle.c for the reference file:line each
function follows.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "corr_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libcorr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _SO, "-lm"]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        L.oracle_pyramid_elems.restype = ctypes.c_size_t
        L.oracle_pyramid_elems.argtypes = [ctypes.c_int] * 4
        L.oracle_level_offset.restype = ctypes.c_size_t
        L.oracle_level_offset.argtypes = [ctypes.c_int] * 4
        L.oracle_build_pyramid.restype = ctypes.c_int
        L.oracle_build_pyramid.argtypes = [fp, fp] + [ctypes.c_int] * 5 + [fp]
        L.oracle_index_pyramid.restype = ctypes.c_int
        L.oracle_index_pyramid.argtypes = [fp, fp] + [ctypes.c_int] * 5 + [fp]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def pyramid_elems(B, h, w, num_levels):
    return int(lib().oracle_pyramid_elems(B, h, w, num_levels))


def level_offset(B, h, w, level):
    return int(lib().oracle_level_offset(B, h, w, level))


def build_pyramid(fmap1: np.ndarray, fmap2: np.ndarray, num_levels: int = 4) -> np.ndarray:
    """Returns the flat pyramid buffer (levels back to back), float32."""
    f1 = np.ascontiguousarray(fmap1, dtype=np.float32)
    f2 = np.ascontiguousarray(fmap2, dtype=np.float32)
    B, C, h, w = f1.shape
    out = np.empty(pyramid_elems(B, h, w, num_levels), dtype=np.float32)
    rc = lib().oracle_build_pyramid(_fp(f1), _fp(f2), B, C, h, w, num_levels, _fp(out))
    if rc != 0:
        raise ValueError(f"oracle_build_pyramid failed rc={rc}")
    return out


def split_levels(flat: np.ndarray, B, h, w, num_levels):
    out = []
    N = h * w
    for l in range(num_levels):
        hl, wl = h >> l, w >> l
        off = level_offset(B, h, w, l)
        out.append(flat[off: off + B * N * hl * wl].reshape(B * N, hl, wl))
    return out


def index_pyramid(flat_pyramid: np.ndarray, coords: np.ndarray, num_levels: int = 4,
                  radius: int = 4) -> np.ndarray:
    c = np.ascontiguousarray(coords, dtype=np.float32)
    B, _, h, w = c.shape
    S = 2 * radius + 1
    out = np.empty((B, num_levels * S * S, h, w), dtype=np.float32)
    p = np.ascontiguousarray(flat_pyramid, dtype=np.float32)
    rc = lib().oracle_index_pyramid(_fp(p), _fp(c), B, h, w, num_levels, radius, _fp(out))
    if rc != 0:
        raise ValueError(f"oracle_index_pyramid failed rc={rc}")
    return out
