"""Module-level parity on the GPU: MultiScaleDeformableAttention.forward (projections + softmax +
sampling-location arithmetic + core op + output projection + residual) against the same
computation assembled from torch ops and the CPU oracle's core
(/root/reference/detrex/layers/multi_scale_deform_attn.py:270-363)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_module_forward(m, query, value, identity, query_pos, key_padding_mask, reference_points, spatial_shapes):
    """py:270-363 re-assembled on the CPU in float64 around the oracle's core."""
    from oracle import msda_torch
    d = torch.float64
    sd = {k: v.detach().cpu().to(d) for k, v in m.state_dict().items()}
    lin = lambda x, n: x @ sd[n + ".weight"].T + sd[n + ".bias"]
    query, value, identity = query.cpu().to(d), value.cpu().to(d), identity.cpu().to(d)
    if query_pos is not None:
        query = query + query_pos.cpu().to(d)
    if not m.batch_first:
        query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    bs, nq, _ = query.shape
    nv = value.shape[1]
    H, L, P = m.num_heads, m.num_levels, m.num_points
    v = lin(value, "value_proj")
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask.cpu()[..., None], 0.0)
    v = v.view(bs, nv, H, -1)
    off = lin(query, "sampling_offsets").view(bs, nq, H, L, P, 2)
    aw = lin(query, "attention_weights").view(bs, nq, H, L * P).softmax(-1).view(bs, nq, H, L, P)
    ref = reference_points.cpu().to(d)
    ss = spatial_shapes.cpu()
    if ref.shape[-1] == 2:
        norm = torch.stack([ss[..., 1], ss[..., 0]], -1).to(d)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    out = msda_torch.forward(v, ss, loc, aw)
    out = lin(out, "output_proj")
    if not m.batch_first:
        out = out.permute(1, 0, 2)
    return out + identity


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("batch_first", [False, True])
def test_module_forward_backward_matches_reference_assembly(ref_dim, batch_first):
    from ir_ads_b200 import MultiScaleDeformableAttention
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(ref_dim * 2 + int(batch_first))
    levels = [(12, 17), (6, 9), (3, 5), (2, 3)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q = 2, (S if ref_dim == 2 else 23)
    m = MultiScaleDeformableAttention(dropout=0.0, batch_first=batch_first).to(DEV)
    with torch.no_grad():   # make the data-dependent branches non-trivial
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.attention_weights.weight.normal_(0, 0.2)
    q = torch.randn(B, Q, 256, device=DEV)
    val = q if ref_dim == 2 else torch.randn(B, S, 256, device=DEV)
    pos = torch.randn(B, Q, 256, device=DEV) * 0.1
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    mask[1, -7:] = True
    ref_pts = torch.rand(B, Q, 4, ref_dim, device=DEV)
    if ref_dim == 4:
        ref_pts[..., 2:] = ref_pts[..., 2:] * 0.4 + 0.05
    if not batch_first:
        q, val, pos = q.transpose(0, 1).contiguous(), val.transpose(0, 1).contiguous(), pos.transpose(0, 1).contiguous()
    q = q.requires_grad_(True)
    out = m(q, value=None if ref_dim == 2 else val, query_pos=pos, key_padding_mask=mask, reference_points=ref_pts,
            spatial_shapes=shapes, level_start_index=lsi, attn_mask=None, key_pos=None)   # extra kwargs swallowed
    want = reference_module_forward(m, q.detach(), (q.detach() if ref_dim == 2 else val), q.detach(), pos, mask,
                                    ref_pts, shapes)
    err = (out.detach().cpu().double() - want).abs().max().item()
    assert err <= 2e-5 * want.abs().max().item() + 1e-5, err
    out.square().mean().backward()
    assert q.grad is not None and torch.isfinite(q.grad).all()
    for p in m.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert m.sampling_offsets.weight.grad.abs().sum() > 0 and m.value_proj.weight.grad.abs().sum() > 0


def test_module_autocast_bf16_and_fp16():
    from ir_ads_b200 import MultiScaleDeformableAttention
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(0)
    levels = [(12, 17), (6, 9), (3, 5), (2, 3)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    m = MultiScaleDeformableAttention(dropout=0.0, batch_first=True).to(DEV)
    q = torch.randn(2, S, 256, device=DEV)
    ref_pts = torch.rand(2, S, 4, 2, device=DEV)
    base = m(q, reference_points=ref_pts, spatial_shapes=shapes, level_start_index=lsi)
    for dt in (torch.bfloat16, torch.float16):
        with torch.autocast("cuda", dtype=dt):
            out = m(q, reference_points=ref_pts, spatial_shapes=shapes, level_start_index=lsi)
        assert torch.isfinite(out).all()
        rel = (out.float() - base).abs().max() / base.abs().max()
        assert rel < 3e-2, (dt, float(rel))


def test_module_rejects_wrong_value_length():
    from ir_ads_b200 import MultiScaleDeformableAttention
    from ir_ads_b200.workloads import level_tensors

    shapes, lsi = level_tensors([(4, 4), (2, 2)], DEV)
    m = MultiScaleDeformableAttention(embed_dim=64, num_heads=4, num_levels=2, dropout=0.0, batch_first=True).to(DEV)
    q = torch.randn(1, 19, 64, device=DEV)                      # 19 != 20
    with pytest.raises(AssertionError):
        m(q, reference_points=torch.rand(1, 19, 2, 2, device=DEV), spatial_shapes=shapes, level_start_index=lsi)
