"""ctypes front end of oracle/msda_oracle.c (TEST INFRASTRUCTURE -- see oracle/__init__.py).

All functions take / return numpy arrays laid out exactly like the reference operator's tensors
(/root/reference/detrex/layers/multi_scale_deform_attn.py:44-54):
    value [B,S,H,D], spatial_shapes [L,2] int64 (H_l, W_l), level_start_index [L] int64,
    sampling_locations [B,Q,H,L,P,2] (x, y) normalised, attention_weights [B,Q,H,L,P].
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")

COORD_REFERENCE = 0    # fp32 rounding order of the reference CUDA kernel (cuh:284-285)
COORD_COMPENSATED = 1  # exact-floor fp32 split used by the sm_100a kernels

_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (a second or two). Returns the .so path."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _LIB_PATH, src, "-lm"]
        )
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, loc):
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    return [ctypes.c_int(int(x)) for x in (B, S, H, D, L, Q, P)]


def _prep(value, shapes, lsi, loc, w, dtype):
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    w = np.ascontiguousarray(w, dtype=dtype)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    assert value.ndim == 4 and loc.ndim == 6 and w.ndim == 5
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == value.shape[1]
    return value, shapes, lsi, loc, w


def forward(value, shapes, lsi, loc, w, dtype=np.float64, coord_mode=COORD_COMPENSATED):
    value, shapes, lsi, loc, w = _prep(value, shapes, lsi, loc, w, dtype)
    B, S, H, D = value.shape
    Q = loc.shape[1]
    out = np.zeros((B, Q, H * D), dtype=dtype)
    d = _dims(value, loc)
    if dtype == np.float64:
        lib().msda_oracle_forward_f64(_p(value), _p(shapes), _p(lsi), _p(loc), _p(w), _p(out), *d)
    else:
        lib().msda_oracle_forward_f32(_p(value), _p(shapes), _p(lsi), _p(loc), _p(w), _p(out), *d,
                                      ctypes.c_int(coord_mode))
    return out


def backward(grad_out, value, shapes, lsi, loc, w, dtype=np.float64, coord_mode=COORD_COMPENSATED):
    value, shapes, lsi, loc, w = _prep(value, shapes, lsi, loc, w, dtype)
    grad_out = np.ascontiguousarray(grad_out, dtype=dtype)
    gv = np.zeros_like(value)
    gl = np.zeros_like(loc)
    gw = np.zeros_like(w)
    d = _dims(value, loc)
    if dtype == np.float64:
        lib().msda_oracle_backward_f64(_p(grad_out), _p(value), _p(shapes), _p(lsi), _p(loc), _p(w),
                                       _p(gv), _p(gl), _p(gw), *d)
    else:
        lib().msda_oracle_backward_f32(_p(grad_out), _p(value), _p(shapes), _p(lsi), _p(loc), _p(w),
                                       _p(gv), _p(gl), _p(gw), *d, ctypes.c_int(coord_mode))
    return gv, gl, gw


def bookkeeping(loc, shapes, lsi, B, S, H, D, is_f32=True, coord_mode=COORD_COMPENSATED):
    """Integer bookkeeping of every point: (corner_offsets [N,4] int64, frac [N,2] float32)."""
    loc = np.ascontiguousarray(loc, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    _, Q, _, L, P, _ = loc.shape
    n = B * Q * H * L * P
    offs = np.full((n, 4), -7, dtype=np.int64)
    frac = np.zeros((n, 2), dtype=np.float32)
    lib().msda_oracle_bookkeeping(_p(loc), _p(shapes), _p(lsi), *[ctypes.c_int(int(x)) for x in (B, S, H, D, L, Q, P)],
                                  ctypes.c_int(1 if is_f32 else 0), ctypes.c_int(coord_mode), _p(offs), _p(frac))
    return offs, frac
