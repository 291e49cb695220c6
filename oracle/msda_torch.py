"""Torch restatement of the reference's CPU path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference computes MSDeformAttn on the CPU with
``multi_scale_deformable_attn_pytorch`` (/root/reference/detrex/layers/multi_scale_deform_attn.py:96-136):
per level, the value slab is viewed as an image batch ``[B*H, D, H_l, W_l]``, the normalised
locations become a grid ``2*loc-1`` of shape ``[B*H, Q, P, 2]`` and ``F.grid_sample(bilinear,
zeros, align_corners=False)`` does the gather; the per-level samples are stacked, multiplied by
the attention weights and summed over ``L*P``.  The arithmetic itself is third-party (ATen
``grid_sampler_2d``; the reference pins torch==1.10.0 in requirements.txt:132, this image has
torch 2.11) -- this module restates the call sequence around it, and oracle/make_golden.py pins
it to the reference function's actual outputs.

``forward`` is what bench.py times as the CPU baseline / ``--impl reference`` arm
(kind = "port"), with all host threads.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn.functional as F


def _level_sizes(spatial_shapes) -> Sequence[Tuple[int, int]]:
    if isinstance(spatial_shapes, torch.Tensor):
        return [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    return [(int(h), int(w)) for h, w in spatial_shapes]


def forward(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
            attention_weights: torch.Tensor) -> torch.Tensor:
    """value [B,S,H,D], sampling_locations [B,Q,H,L,P,2], attention_weights [B,Q,H,L,P] -> [B,Q,H*D]."""
    B, S, H, D = value.shape
    Q, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    sizes = _level_sizes(spatial_shapes)
    assert len(sizes) == L and sum(h * w for h, w in sizes) == S
    grids = sampling_locations * 2 - 1                       # py:106
    per_level = []
    begin = 0
    for lvl, (h_l, w_l) in enumerate(sizes):
        slab = value[:, begin:begin + h_l * w_l]            # [B, h*w, H, D]      (py:105 split)
        begin += h_l * w_l
        image = slab.reshape(B, h_l * w_l, H * D).permute(0, 2, 1).reshape(B * H, D, h_l, w_l)  # py:113-115
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * H, Q, P, 2)                # py:119
        per_level.append(F.grid_sample(image, grid, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))  # [B*H, D, Q, P]          (py:121-123)
    sampled = torch.stack(per_level, dim=-2).reshape(B * H, D, Q, L * P)                         # py:132
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(B * H, 1, Q, L * P)               # py:128-130
    out = (sampled * weights).sum(dim=-1)                    # [B*H, D, Q]
    return out.reshape(B, H * D, Q).permute(0, 2, 1).contiguous()                                # py:134-136


def forward_backward(value, spatial_shapes, sampling_locations, attention_weights, grad_output=None):
    """Forward + autograd backward. Returns (out, grad_value, grad_loc, grad_w)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    w = attention_weights.detach().clone().requires_grad_(True)
    out = forward(v, spatial_shapes, loc, w)
    if grad_output is None:
        grad_output = torch.ones_like(out)
    out.backward(grad_output)
    return out.detach(), v.grad, loc.grad, w.grad


def forward_backward_f64(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """The same inputs evaluated in float64 -- the yardstick the fp32 / bf16 kernels are judged by."""
    return forward_backward(value.double(), spatial_shapes, sampling_locations.double(),
                            attention_weights.double(), grad_output.double())
