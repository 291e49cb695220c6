"""Torch restatement of the reference's DCNv3 CPU path, ``dcnv3_core_pytorch``
(/root/reference/detrex/layers/dcn_v3.py:121-166 with its helpers :68-118) -- TEST INFRASTRUCTURE.
Pads the input, builds convolution-style reference points and the dilation grid in normalised
coordinates of the PADDED map, adds offset*offset_scale/size, and gathers with F.grid_sample
(bilinear, zeros, align_corners=False); pinned to the reference function by tests/golden/dcnv3_golden.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _reference_points(H_, W_, kernel_h, kernel_w, dilation_h, dilation_w, stride_h, stride_w, device, dtype):
    H_out = (H_ - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    W_out = (W_ - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    y0 = (dilation_h * (kernel_h - 1)) // 2 + 0.5
    x0 = (dilation_w * (kernel_w - 1)) // 2 + 0.5
    ys = torch.linspace(y0, y0 + (H_out - 1) * stride_h, H_out, dtype=torch.float32, device=device)
    xs = torch.linspace(x0, x0 + (W_out - 1) * stride_w, W_out, dtype=torch.float32, device=device)
    ref_y, ref_x = torch.meshgrid(ys, xs, indexing="ij")
    ref = torch.stack((ref_x.reshape(-1)[None] / W_, ref_y.reshape(-1)[None] / H_), -1)
    return ref.reshape(1, H_out, W_out, 1, 2)            # float32 on purpose, like the reference (dcn_v3.py:74-95)


def _dilation_grid(H_, W_, kernel_h, kernel_w, dilation_h, dilation_w, group, device, dtype):
    xs = torch.linspace(-((dilation_w * (kernel_w - 1)) // 2), -((dilation_w * (kernel_w - 1)) // 2) + (kernel_w - 1) * dilation_w,
                        kernel_w, dtype=torch.float32, device=device)
    ys = torch.linspace(-((dilation_h * (kernel_h - 1)) // 2), -((dilation_h * (kernel_h - 1)) // 2) + (kernel_h - 1) * dilation_h,
                        kernel_h, dtype=torch.float32, device=device)
    x, y = torch.meshgrid(xs, ys, indexing="ij")        # x varies along dim 0: point index = i*kernel_h + j
    grid = torch.stack([x / W_, y / H_], -1).reshape(-1, 1, 2).repeat(1, group, 1).permute(1, 0, 2)
    return grid.reshape(1, 1, 1, group * kernel_h * kernel_w, 2)   # float32, like the reference


def forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
            group_channels, offset_scale):
    input = F.pad(input, [0, 0, pad_h, pad_h, pad_w, pad_w])
    N_, H_in, W_in, _ = input.shape
    _, H_out, W_out, _ = offset.shape
    dt = input.dtype
    ref = _reference_points(H_in, W_in, kernel_h, kernel_w, dilation_h, dilation_w, stride_h, stride_w, input.device, dt)
    grid = _dilation_grid(H_in, W_in, kernel_h, kernel_w, dilation_h, dilation_w, group, input.device, dt)
    P_ = kernel_h * kernel_w
    norm = torch.tensor([W_in, H_in]).reshape(1, 1, 1, 2).repeat(1, 1, 1, group * P_).to(input.device)   # int64, py:139-140
    loc = (ref + grid * offset_scale).repeat(N_, 1, 1, 1, 1).flatten(3, 4) + offset * offset_scale / norm
    grids = 2 * loc - 1
    image = input.view(N_, H_in * W_in, group * group_channels).transpose(1, 2).reshape(N_ * group, group_channels, H_in, W_in)
    g = grids.view(N_, H_out * W_out, group, P_, 2).transpose(1, 2).flatten(0, 1)
    sampled = F.grid_sample(image, g, mode="bilinear", padding_mode="zeros", align_corners=False)
    m = mask.view(N_, H_out * W_out, group, P_).transpose(1, 2).reshape(N_ * group, 1, H_out * W_out, P_)
    out = (sampled * m).sum(-1).view(N_, group * group_channels, H_out * W_out)
    return out.transpose(1, 2).reshape(N_, H_out, W_out, -1).contiguous()


def forward_backward(input, offset, mask, grad_output, *geom):
    i = input.detach().clone().requires_grad_(True)
    o = offset.detach().clone().requires_grad_(True)
    m = mask.detach().clone().requires_grad_(True)
    out = forward(i, o, m, *geom)
    out.backward(grad_output)
    return out.detach(), i.grad, o.grad, m.grad
