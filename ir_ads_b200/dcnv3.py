"""DCNv3 core op on the sm_100a gather/scatter kernels.

Mirrors the reference's ``DCNv3Function`` and the ``detrex._C.dcnv3_forward / dcnv3_backward`` pair
(/root/reference/detrex/layers/dcn_v3.py:21-65, detrex/layers/csrc/vision.cpp:57-58,
detrex/layers/csrc/DCNv3/dcnv3_cuda.cu): same argument order, channels-last tensors, autograd for
input / offset / mask.  DCNv3 is MSDeformAttn with one level, K = kernel_h*kernel_w points whose
positions come from a convolution-style grid, and groups in the role of heads -- so it runs on the
same kernels with a different record-building phase (ir_ads_b200/csrc/msda_fast.cuh, PRE == 2).

Supported here: group_channels in {16, 32, 64, 128}, K <= 64, float32 / bfloat16 input (float16 is
widened like the reference's custom_fwd would keep it); anything else raises (no fallback).
"""
from __future__ import annotations

import ctypes
from typing import List

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _DTYPE_TAG, _ptr, _require


def _declare(handle):
    if getattr(handle, "_dcn_declared", False):
        return handle
    vp, i, u, sz, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_float
    handle.msda_dcnv3_forward.restype = i
    handle.msda_dcnv3_forward.argtypes = [vp, vp, vp, vp] + [i] * 10 + [f] + [i] * 5 + [vp, i, u]
    handle.msda_dcnv3_backward.restype = i
    handle.msda_dcnv3_backward.argtypes = [vp, vp, vp, vp, vp] + [i] * 10 + [f] + [i] * 5 + [vp, vp, vp, vp, sz, i, u]
    handle._dcn_declared = True
    return handle


def _geometry(input, offset, mask, kernel_h, kernel_w, group, group_channels):
    _require(input.is_cuda, "Not implemented on the CPU")
    for name, t in (("input", input), ("offset", offset), ("mask", mask)):
        _require(t.is_cuda and t.device == input.device, f"{name} must be a CUDA tensor on {input.device}")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    _require(input.dim() == 4 and offset.dim() == 4 and mask.dim() == 4, "input / offset / mask must be 4-d (N,H,W,C)")
    N, H_in, W_in, C = input.shape
    _, H_out, W_out, _ = offset.shape
    K = kernel_h * kernel_w
    _require(C == group * group_channels, f"input channels {C} != group*group_channels {group * group_channels}")
    _require(tuple(offset.shape) == (N, H_out, W_out, group * K * 2), "offset must be [N, H_out, W_out, group*K*2]")
    _require(tuple(mask.shape) == (N, H_out, W_out, group * K), "mask must be [N, H_out, W_out, group*K]")
    _require(input.dtype in (torch.float32, torch.bfloat16), f"unsupported input dtype {input.dtype}")
    _require(offset.dtype == torch.float32 and mask.dtype == torch.float32, "offset / mask must be float32")
    return N, H_in, W_in, H_out, W_out


def dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                  group, group_channels, offset_scale, im2col_step=256) -> torch.Tensor:
    """``detrex._C.dcnv3_forward``: returns ``[N, H_out, W_out, group*group_channels]``."""
    N, H_in, W_in, H_out, W_out = _geometry(input, offset, mask, kernel_h, kernel_w, group, group_channels)
    out = torch.empty((N, H_out, W_out, group * group_channels), dtype=input.dtype, device=input.device)
    stream = torch.cuda.current_stream(input.device).cuda_stream
    status = _declare(_lib.lib()).msda_dcnv3_forward(
        ctypes.c_void_p(stream), _ptr(input), _ptr(offset), _ptr(mask), kernel_h, kernel_w, stride_h, stride_w, pad_h,
        pad_w, dilation_h, dilation_w, group, group_channels, float(offset_scale), N, H_in, W_in, H_out, W_out,
        _ptr(out), _DTYPE_TAG[input.dtype], 0)
    _lib.check(status, "dcnv3_forward")
    return out


def dcnv3_backward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                   group, group_channels, offset_scale, grad_output, im2col_step=256) -> List[torch.Tensor]:
    """``detrex._C.dcnv3_backward``: ``[grad_input, grad_offset, grad_mask]``."""
    N, H_in, W_in, H_out, W_out = _geometry(input, offset, mask, kernel_h, kernel_w, group, group_channels)
    _require(grad_output.is_cuda and grad_output.dtype == input.dtype, "grad_output must match input's device / dtype")
    grad_output = grad_output.contiguous()
    grad_input = torch.empty_like(input)
    grad_offset = torch.empty_like(offset)
    grad_mask = torch.empty_like(mask)
    ws_bytes = input.numel() * 4 if input.dtype == torch.bfloat16 else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=input.device) if ws_bytes else None
    stream = torch.cuda.current_stream(input.device).cuda_stream
    status = _declare(_lib.lib()).msda_dcnv3_backward(
        ctypes.c_void_p(stream), _ptr(grad_output), _ptr(input), _ptr(offset), _ptr(mask), kernel_h, kernel_w,
        stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, group_channels, float(offset_scale), N, H_in,
        W_in, H_out, W_out, _ptr(grad_input), _ptr(grad_offset), _ptr(grad_mask),
        _ptr(ws) if ws is not None else ctypes.c_void_p(0), ws_bytes, _DTYPE_TAG[input.dtype], 0)
    _lib.check(status, "dcnv3_backward")
    return [grad_input, grad_offset, grad_mask]


class DCNv3Function(Function):
    """Same 15-argument ``apply`` as the reference (dcn_v3.py:21-65)."""

    @staticmethod
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                group, group_channels, offset_scale, im2col_step):
        ctx.args = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, group_channels,
                    offset_scale)
        ctx.im2col_step = im2col_step
        ctx.in_dtype = input.dtype
        if input.dtype == torch.float16:
            input = input.float()
        output = dcnv3_forward(input, offset.float(), mask.float(), *ctx.args, im2col_step)
        ctx.save_for_backward(input, offset, mask)
        return output.to(ctx.in_dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        gi, go, gm = dcnv3_backward(input, offset.float(), mask.float(), *ctx.args, grad_output.to(input.dtype),
                                    ctx.im2col_step)
        return (gi.to(ctx.in_dtype), go.to(offset.dtype), gm.to(mask.dtype)) + (None,) * 12
