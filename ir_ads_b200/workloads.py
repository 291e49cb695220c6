"""Named MSDeformAttn workloads (BASELINE.json configs) and synthetic input generators.

Shapes and distributions follow SURVEY.md section 8(d):
  cfg1  CPU case                 levels (100x150, 50x75, 25x38, 13x19), B=2, Q=300
  cfg2  DINO-R50 encoder @800x1333 levels (100x167, 50x84, 25x42, 13x21), B=8, Q=S=22223
  cfg3  vCLR decoder cross-attn  same levels, B=8, Q=2000, bf16 value
  cfg4  VOC-shape encoder        levels (100x134, 50x67, 25x34, 13x17), B=2 per GPU
  cfg5  stress                   levels 128^2..8^2 (5), P=8, B=8 per GPU, deterministic backward

Distributions (primary = "model"):
  model  encoder: reference points = pixel centres of every level (what get_reference_points
         builds, /root/reference/projects/vCLR_deformable_mask/modeling/dino_transformer.py:322-351,
         valid_ratios = 1), offsets = the module's initial bias pattern (head direction x (p+1) px,
         /root/reference/detrex/layers/multi_scale_deform_attn.py:205-217) + N(0, 2 px) jitter,
         normalised by (W_l, H_l) (py:320-324).  decoder: boxes cx,cy~U(0,1), w,h~U(0.02,0.5),
         loc = c + off/P*wh*0.5 (py:326-332).  value ~ N(0,1), weights = softmax(N(0,1)).
  test   the reference unit test's inputs (tests/test_ms_deform_attn.py:78-81): loc~U[0,1),
         value = U*0.01, w = (U+1e-5) normalised -- no spatial locality at all.
  edge   loc~U[-0.3,1.3] with some locations snapped to pixel centres / integer coordinates,
         some rows fully out of range, some weights exactly zero.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

LEVELS = {
    "cfg1": [(100, 150), (50, 75), (25, 38), (13, 19)],
    "dino_r50": [(100, 167), (50, 84), (25, 42), (13, 21)],
    "voc": [(100, 134), (50, 67), (25, 34), (13, 17)],
    "stress": [(128, 128), (64, 64), (32, 32), (16, 16), (8, 8)],
}


@dataclass(frozen=True)
class Workload:
    name: str
    levels: Tuple[Tuple[int, int], ...]
    batch: int          # per GPU
    num_query: int      # 0 => Q = S (encoder self-attention)
    num_heads: int = 8
    head_dim: int = 32
    num_points: int = 4
    kind: str = "encoder"      # "encoder" | "decoder"
    value_dtype: str = "f32"   # "f32" | "bf16"
    deterministic: bool = False

    @property
    def spatial_size(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def queries(self) -> int:
        return self.num_query if self.num_query > 0 else self.spatial_size

    @property
    def num_levels(self) -> int:
        return len(self.levels)

    @property
    def points(self) -> int:
        """Sampled points of one forward = B*Q*H*L*P (the unit of the metric)."""
        return self.batch * self.queries * self.num_heads * self.num_levels * self.num_points

    def algorithmic_bytes(self) -> Tuple[int, int]:
        """(fwd_bytes, bwd_bytes): compulsory traffic, SURVEY.md section 8(d)."""
        sv = 2 if self.value_dtype == "bf16" else 4
        so = sg = sv
        n_val = self.batch * self.spatial_size * self.num_heads * self.head_dim
        n_out = self.batch * self.queries * self.num_heads * self.head_dim
        n_pts = self.points
        fwd = n_val * sv + n_pts * 12 + n_out * so
        bwd = n_out * so + n_val * sv + n_pts * 12 + n_val * sg + n_pts * 12
        return fwd, bwd


WORKLOADS = {
    "cfg1": Workload("cfg1", tuple(LEVELS["cfg1"]), batch=2, num_query=300, kind="decoder"),
    "cfg2": Workload("cfg2", tuple(LEVELS["dino_r50"]), batch=8, num_query=0, kind="encoder"),
    "cfg2_bf16": Workload("cfg2_bf16", tuple(LEVELS["dino_r50"]), batch=8, num_query=0, kind="encoder", value_dtype="bf16"),
    "cfg3": Workload("cfg3", tuple(LEVELS["dino_r50"]), batch=8, num_query=2000, kind="decoder", value_dtype="bf16"),
    "cfg3_f32": Workload("cfg3_f32", tuple(LEVELS["dino_r50"]), batch=8, num_query=2000, kind="decoder"),
    "cfg4": Workload("cfg4", tuple(LEVELS["voc"]), batch=2, num_query=0, kind="encoder"),
    "cfg5": Workload("cfg5", tuple(LEVELS["stress"]), batch=8, num_query=0, num_points=8, kind="encoder",
                     deterministic=True),
}


def level_tensors(levels: Sequence[Tuple[int, int]], device="cpu"):
    """spatial_shapes [L,2] int64 (H_l, W_l) and level_start_index [L] int64, as
    dino_transformer.py:393-398 builds them."""
    shapes = torch.as_tensor(list(levels), dtype=torch.long, device=device)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, lsi


def _head_directions(num_heads: int, device) -> torch.Tensor:
    thetas = torch.arange(num_heads, dtype=torch.float32, device=device) * (2.0 * math.pi / num_heads)
    d = torch.stack([thetas.cos(), thetas.sin()], -1)
    return d / d.abs().max(-1, keepdim=True)[0]          # [H,2]


def _pixel_centres(levels, device) -> torch.Tensor:
    pts: List[torch.Tensor] = []
    for h, w in levels:
        ys = (torch.arange(h, dtype=torch.float32, device=device) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32, device=device) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    return torch.cat(pts, 0)                              # [S,2] (x,y)


def make_inputs(levels, batch, num_query=0, num_heads=8, head_dim=32, num_points=4, kind="encoder",
                dist="model", seed=0, device="cpu", value_dtype=torch.float32):
    """Returns (value [B,S,H,D], spatial_shapes, level_start_index, loc [B,Q,H,L,P,2] fp32,
    w [B,Q,H,L,P] fp32).  loc and w are float32 (float64 if value_dtype is float64)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    Q = num_query if num_query > 0 else S
    H, D, P = num_heads, head_dim, num_points
    shapes, lsi = level_tensors(levels, device)
    rt = torch.float64 if value_dtype == torch.float64 else torch.float32

    def randn(*s):
        return torch.randn(*s, generator=g, device=device, dtype=torch.float32)

    def rand(*s):
        return torch.rand(*s, generator=g, device=device, dtype=torch.float32)

    wh = torch.as_tensor([[w, h] for h, w in levels], dtype=torch.float32, device=device)  # (W_l, H_l)
    if dist == "model":
        value = randn(batch, S, H, D)
        weights = torch.softmax(randn(batch, Q, H, L * P), -1).view(batch, Q, H, L, P)
        steps = torch.arange(1, P + 1, dtype=torch.float32, device=device)
        if kind == "encoder" and Q == S:
            ref = _pixel_centres(levels, device)[None, :, None, None, None, :]             # [1,Q,1,1,1,2]
            off = _head_directions(H, device)[None, None, :, None, None, :] * steps[None, None, None, None, :, None]
            off = off + 2.0 * randn(batch, Q, H, L, P, 2)
            loc = ref + off / wh[None, None, None, :, None, :]
        else:
            cxy = rand(batch, Q, 1, 1, 1, 2)
            bwh = 0.02 + 0.48 * rand(batch, Q, 1, 1, 1, 2)
            off = randn(batch, Q, H, L, P, 2) * steps[None, None, None, None, :, None]
            loc = cxy + off / P * bwh * 0.5
    elif dist == "test":
        value = rand(batch, S, H, D) * 0.01
        loc = rand(batch, Q, H, L, P, 2)
        weights = rand(batch, Q, H, L, P) + 1e-5
        weights = weights / weights.sum(-1, keepdim=True).sum(-2, keepdim=True)
    elif dist == "edge":
        value = randn(batch, S, H, D)
        loc = rand(batch, Q, H, L, P, 2) * 1.6 - 0.3
        # snap a third of the points to exact pixel centres / integer pixel coordinates
        snap = rand(batch, Q, H, L, P, 1)
        whb = wh[None, None, None, :, None, :]
        centre = (torch.floor(loc * whb) + 0.5) / whb
        corner = torch.round(loc * whb) / whb
        loc = torch.where(snap < 0.15, centre, loc)
        loc = torch.where((snap >= 0.15) & (snap < 0.30), corner, loc)
        if Q >= 2:
            loc[:, 0] = 1.5          # a fully out-of-range query
            loc[:, 1] = -0.7
        weights = torch.softmax(randn(batch, Q, H, L * P), -1).view(batch, Q, H, L, P)
        weights = torch.where(rand(batch, Q, H, L, P) < 0.1, torch.zeros_like(weights), weights)
    else:
        raise ValueError(f"unknown dist {dist!r}")
    value = value.to(value_dtype)
    return value.contiguous(), shapes, lsi, loc.to(rt).contiguous(), weights.to(rt).contiguous()


def make_workload_inputs(wl: Workload, dist="model", seed=0, device="cpu", batch=None):
    vd = torch.bfloat16 if wl.value_dtype == "bf16" else torch.float32
    return make_inputs(wl.levels, wl.batch if batch is None else batch, wl.num_query, wl.num_heads,
                       wl.head_dim, wl.num_points, wl.kind, dist, seed, device, vd)


def valid_corner_fraction(loc: torch.Tensor, levels) -> float:
    """Share of the 4 * N bilinear corners of `loc` [B,Q,H,L,P,2] that lie inside their level (the rest is zero padding
    or belongs to gated-out points, ms_deform_im2col_cuda.cuh:288, :56-80): the rows the kernels actually load / scatter.
    Bench bookkeeping only; evaluated in float64 like the oracle's exact cell (the kernels' compensated fp32 split picks
    the same cell, a plain fp32 product does not at exact-integer coordinates)."""
    wh = torch.as_tensor([[w, h] for h, w in levels], dtype=torch.float64, device=loc.device)
    size = wh[None, None, None, :, None, :]
    c = loc.double() * size - 0.5
    lo = torch.floor(c)
    gate = ((c > -1.0) & (c < size)).all(-1)
    n_axis = ((lo >= 0).to(torch.int32) + (lo + 1 <= size - 1).to(torch.int32))        # valid corners per axis: 0..2
    valid = n_axis[..., 0] * n_axis[..., 1] * gate
    return float(valid.sum().item()) / float(4 * valid.numel())
