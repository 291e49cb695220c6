"""ctypes wrapper over oracle/_ref/libmsda_refcuda.so -- the REFERENCE's own CUDA kernels
(/root/reference/detrex/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh) compiled unmodified
for sm_100a by oracle/Makefile.  TEST / BASELINE INFRASTRUCTURE ONLY: a second GPU-side checker and
the "reference kernels on the same B200" line of bench.py.  GPU only.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmsda_refcuda.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.msda_ref_forward.restype = ctypes.c_int
        _lib.msda_ref_backward.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _dims(value, loc):
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    return [ctypes.c_int(int(x)) for x in (B, S, H, D, L, Q, P)]


def forward(value, shapes, lsi, loc, w, out=None):
    """float32 / float64 CUDA tensors, contiguous (ms_deform_attn_cuda.cu:29-39)."""
    tag = 0 if value.dtype == torch.float32 else 1
    B, S, H, D = value.shape
    Q = loc.shape[1]
    if out is None:
        out = torch.zeros((B, Q, H * D), dtype=value.dtype, device=value.device)   # at::zeros, cu:55
    st = ctypes.c_void_p(torch.cuda.current_stream(value.device).cuda_stream)
    err = lib().msda_ref_forward(ctypes.c_int(tag), _p(value), _p(shapes), _p(lsi), _p(loc), _p(w),
                                 *_dims(value, loc), _p(out), st)
    if err:
        raise RuntimeError(f"reference forward kernel: cuda error {err}")
    return out


def backward(grad_out, value, shapes, lsi, loc, w, bufs=None):
    tag = 0 if value.dtype == torch.float32 else 1
    if bufs is None:
        gv, gl, gw = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(w)        # cu:122-124
    else:
        gv, gl, gw = bufs
        gv.zero_(); gl.zero_(); gw.zero_()
    st = ctypes.c_void_p(torch.cuda.current_stream(value.device).cuda_stream)
    err = lib().msda_ref_backward(ctypes.c_int(tag), _p(grad_out), _p(value), _p(shapes), _p(lsi), _p(loc), _p(w),
                                  *_dims(value, loc), _p(gv), _p(gl), _p(gw), st)
    if err:
        raise RuntimeError(f"reference backward kernel: cuda error {err}")
    return gv, gl, gw


def forward_backward(value, shapes, lsi, loc, w, grad_out):
    out = forward(value, shapes, lsi, loc, w)
    gv, gl, gw = backward(grad_out.contiguous(), value, shapes, lsi, loc, w)
    return out, gv, gl, gw


# ---------------------------------------------------------------------------------------------
# DCNv3: the reference's own kernels (oracle/_ref/libdcnv3_refcuda.so), float32 only
# ---------------------------------------------------------------------------------------------
DCN_LIB_PATH = os.path.join(_HERE, "_ref", "libdcnv3_refcuda.so")
_dcn = None


def dcn_available() -> bool:
    return os.path.exists(DCN_LIB_PATH)


def _dcn_lib():
    global _dcn
    if _dcn is None:
        _dcn = ctypes.CDLL(DCN_LIB_PATH)
        _dcn.dcnv3_ref_forward.restype = ctypes.c_int
        _dcn.dcnv3_ref_backward.restype = ctypes.c_int
    return _dcn


def dcnv3_forward_backward(input, offset, mask, grad_out, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                           dilation_h, dilation_w, group, group_channels, offset_scale, backward=True):
    N, H, W, _ = input.shape
    _, Ho, Wo, _ = offset.shape
    geom = [ctypes.c_int(int(x)) for x in (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                                           group, group_channels, N, H, W, Ho, Wo)]
    st = ctypes.c_void_p(torch.cuda.current_stream(input.device).cuda_stream)
    out = torch.zeros((N, Ho, Wo, group * group_channels), dtype=torch.float32, device=input.device)
    err = _dcn_lib().dcnv3_ref_forward(_p(input), _p(offset), _p(mask), _p(out), *geom, ctypes.c_float(offset_scale), st)
    if err:
        raise RuntimeError(f"reference dcnv3 forward: cuda error {err}")
    if not backward:
        return out
    gi, go, gm = torch.zeros_like(input), torch.zeros_like(offset), torch.zeros_like(mask)
    err = _dcn_lib().dcnv3_ref_backward(_p(grad_out.contiguous()), _p(input), _p(offset), _p(mask), *geom,
                                        ctypes.c_float(offset_scale), _p(gi), _p(go), _p(gm), st)
    if err:
        raise RuntimeError(f"reference dcnv3 backward: cuda error {err}")
    return out, gi, go, gm
