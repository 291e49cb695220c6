"""Encoder-side callers of the op: glue functions on CPU, the stack on the GPU."""
import numpy as np
import pytest
import torch

from ir_ads_b200 import encoder


def test_reference_points_are_pixel_centres_scaled_by_valid_ratio():
    shapes = [(3, 4), (2, 2)]
    vr = torch.ones(2, 2, 2)
    vr[1, 0] = torch.tensor([0.5, 0.75])           # image 1, level 0: half valid in x, 3/4 in y
    ref = encoder.get_reference_points(shapes, vr, "cpu")
    assert ref.shape == (2, 16, 2, 2)
    # image 0 (no padding): the centre of pixel (y=1, x=2) of level 0 is ((2+.5)/4, (1+.5)/3) for every level
    assert torch.allclose(ref[0, 1 * 4 + 2], torch.tensor([[2.5 / 4, 1.5 / 3]] * 2))
    # level-1 pixels come after the 12 level-0 pixels
    assert torch.allclose(ref[0, 12 + 3, 0], torch.tensor([1.5 / 2, 1.5 / 2]))
    # image 1: divided by its own level's valid ratio, then multiplied by each target level's ratio
    x = 2.5 / (0.5 * 4)
    y = 1.5 / (0.75 * 3)
    assert torch.allclose(ref[1, 1 * 4 + 2, 0], torch.tensor([x * 0.5, y * 0.75]))
    assert torch.allclose(ref[1, 1 * 4 + 2, 1], torch.tensor([x * 1.0, y * 1.0]))


def test_flatten_levels_and_valid_ratio():
    B, C = 2, 8
    feats = [torch.randn(B, C, 3, 4), torch.randn(B, C, 2, 2)]
    masks = [torch.zeros(B, 3, 4, dtype=torch.bool), torch.zeros(B, 2, 2, dtype=torch.bool)]
    masks[1][1, :, 1:] = True                       # level 1 of image 1: only the first column is valid
    pos = [torch.randn(B, C, 3, 4), torch.randn(B, C, 2, 2)]
    lvl = torch.randn(2, C)
    feat, mask, p, ss, lsi, host, vr = encoder.flatten_levels(feats, masks, pos, lvl)
    assert feat.shape == (B, 16, C) and mask.shape == (B, 16) and p.shape == (B, 16, C)
    assert ss.tolist() == [[3, 4], [2, 2]] and lsi.tolist() == [0, 12] and host == [(3, 4), (2, 2)]
    assert torch.equal(feat[:, 5], feats[0][:, :, 1, 1]) and torch.equal(feat[:, 12 + 3], feats[1][:, :, 1, 1])
    assert torch.allclose(p[:, 12], pos[1][:, :, 0, 0] + lvl[1])
    assert torch.allclose(vr[1, 1], torch.tensor([0.5, 1.0])) and torch.allclose(vr[0], torch.ones(2, 2))
    assert mask[1, 12:].tolist() == [False, True, False, True]


def test_encoder_state_dict_names_match_reference_layout():
    enc = encoder.DeformableEncoder(embed_dim=64, num_heads=4, feedforward_dim=128, num_layers=2, post_norm=True)
    keys = set(enc.state_dict())
    for k in ("layers.0.attentions.0.sampling_offsets.weight", "layers.0.attentions.0.output_proj.bias",
              "layers.1.ffns.0.layers.0.0.weight", "layers.1.ffns.0.layers.1.bias", "layers.0.norms.1.weight",
              "post_norm_layer.weight"):
        assert k in keys, k


@pytest.mark.gpu
def test_encoder_stack_matches_fp64_cpu_restatement():
    """2-layer encoder on the GPU vs the same parameters evaluated on the CPU in float64 around the oracle's
    core op (grid_sample formulation)."""
    from oracle import msda_torch
    torch.manual_seed(0)
    dev = "cuda:0"
    B, C = 2, 64
    feats = [torch.randn(B, C, 7, 9), torch.randn(B, C, 4, 5), torch.randn(B, C, 2, 3)]
    masks = [torch.zeros(B, f.shape[2], f.shape[3], dtype=torch.bool) for f in feats]
    masks[0][1, :, 6:] = True
    masks[1][1, :, 3:] = True
    masks[2][1, :, 2:] = True
    pos = [torch.randn_like(f) * 0.1 for f in feats]
    enc = encoder.DeformableEncoder(embed_dim=C, num_heads=4, feedforward_dim=96, attn_dropout=0.0, ffn_dropout=0.0,
                                    num_layers=2, num_feature_levels=3, num_points=2)
    for layer in enc.layers:
        with torch.no_grad():
            layer.attentions[0].sampling_offsets.weight.normal_(0, 0.05)
            layer.attentions[0].attention_weights.weight.normal_(0, 0.2)

    def run(module, device, dtype, op):
        f = [t.to(device, dtype) for t in feats]
        p = [t.to(device, dtype) for t in pos]
        m = [t.to(device) for t in masks]
        feat, mask, posf, ss, lsi, host, vr = encoder.flatten_levels(f, m, p)
        ref = encoder.get_reference_points(host, vr.to(dtype), device).to(dtype)
        if op is None:
            return module(feat, query_pos=posf, query_key_padding_mask=mask, reference_points=ref,
                          spatial_shapes=ss, level_start_index=lsi)
        x = feat
        for layer in module.layers:                     # float64 restatement around the oracle op
            a = layer.attentions[0]
            q = x + posf
            H, L, P = a.num_heads, a.num_levels, a.num_points
            v = a.value_proj(x).masked_fill(mask[..., None], 0.0).view(B, x.shape[1], H, -1)
            off = a.sampling_offsets(q).view(B, -1, H, L, P, 2)
            w = a.attention_weights(q).view(B, -1, H, L * P).softmax(-1).view(B, -1, H, L, P)
            norm = torch.stack([ss[..., 1], ss[..., 0]], -1).to(dtype)
            loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
            x = layer.norms[0](a.output_proj(op(v, host, loc, w)) + x)
            x = layer.norms[1](layer.ffns[0](x))
        return x

    want = run(enc.double(), "cpu", torch.float64, msda_torch.forward)
    got = run(enc.float().to(dev), dev, torch.float32, None)
    err = (got.detach().cpu().double() - want).abs().max().item()
    assert err <= 5e-5 * want.abs().max().item() + 1e-5, err


def test_decoder_reference_points_input():
    ref4 = torch.tensor([[[0.5, 0.5, 0.2, 0.4]]])
    vr = torch.tensor([[[1.0, 1.0], [0.5, 0.25]]])
    out = encoder.decoder_reference_points_input(ref4, vr)
    assert out.shape == (1, 1, 2, 4)
    assert torch.allclose(out[0, 0, 1], torch.tensor([0.25, 0.125, 0.1, 0.1]))
    out2 = encoder.decoder_reference_points_input(ref4[..., :2], vr)
    assert torch.allclose(out2[0, 0, 1], torch.tensor([0.25, 0.125]))
    with pytest.raises(ValueError):
        encoder.decoder_reference_points_input(torch.zeros(1, 1, 3), vr)


@pytest.mark.gpu
def test_decoder_cross_attention_block_matches_fp64_restatement():
    from oracle import msda_torch
    torch.manual_seed(2)
    dev = "cuda:0"
    levels = [(9, 12), (5, 6), (3, 3)]
    B, Q, C = 2, 17, 64
    S = sum(h * w for h, w in levels)
    blk = encoder.DeformableCrossAttentionBlock(embed_dim=C, num_heads=4, attn_dropout=0.0, num_feature_levels=3,
                                                num_points=2)
    with torch.no_grad():
        blk.attn.sampling_offsets.weight.normal_(0, 0.05)
        blk.attn.attention_weights.weight.normal_(0, 0.2)
    q, mem, pos = torch.randn(B, Q, C), torch.randn(B, S, C), torch.randn(B, Q, C) * 0.1
    ref = torch.rand(B, Q, 4) * torch.tensor([1, 1, 0.4, 0.4]) + torch.tensor([0, 0, 0.05, 0.05])
    vr = torch.rand(B, 3, 2) * 0.3 + 0.7
    mask = torch.zeros(B, S, dtype=torch.bool)
    mask[1, -4:] = True
    from ir_ads_b200.workloads import level_tensors
    ss, lsi = level_tensors(levels, "cpu")
    # float64 restatement around the oracle op
    d = blk.double()
    a = d.attn
    H, L, P = 4, 3, 2
    refin = encoder.decoder_reference_points_input(ref.double(), vr.double())
    qq = q.double() + pos.double()
    v = a.value_proj(mem.double()).masked_fill(mask[..., None], 0.0).view(B, S, H, -1)
    off = a.sampling_offsets(qq).view(B, Q, H, L, P, 2)
    w = a.attention_weights(qq).view(B, Q, H, L * P).softmax(-1).view(B, Q, H, L, P)
    loc = refin[:, :, None, :, None, :2] + off / P * refin[:, :, None, :, None, 2:] * 0.5
    want = d.norm(a.output_proj(msda_torch.forward(v, levels, loc, w)) + q.double()).detach()
    blk = blk.float().to(dev)
    got = blk(q.to(dev), mem.to(dev), pos.to(dev), ref.to(dev), vr.to(dev), ss.to(dev), lsi.to(dev), mask.to(dev))
    err = (got.detach().cpu().double() - want).abs().max().item()
    assert err <= 5e-5 * want.abs().max().item() + 1e-5, err
