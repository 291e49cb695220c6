"""Module-level parity on the GPU: MultiScaleDeformableAttention.forward (projections + softmax +
sampling-location arithmetic + core op + output projection + residual) against the same
computation assembled from torch ops and the CPU oracle's core
(/root/reference/detrex/layers/multi_scale_deform_attn.py:270-363)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_module_forward(m, query, value, identity, query_pos, key_padding_mask, reference_points, spatial_shapes,
                             params=None):
    """py:270-363 re-assembled on the CPU in float64 around the oracle's core (differentiable torch ops:
    pass float64 leaf tensors as `params` / `query` / `value` to get the reference gradients by autograd)."""
    from oracle import msda_torch
    d = torch.float64
    sd = params if params is not None else {k: v.detach().cpu().to(d) for k, v in m.state_dict().items()}
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
    # backward: input and parameter gradients against autograd through the float64 reference assembly
    params = {k: v.detach().cpu().double().requires_grad_(True) for k, v in m.state_dict().items()}
    q64 = q.detach().cpu().double().requires_grad_(True)
    v64 = q64 if ref_dim == 2 else val.detach().cpu().double().requires_grad_(True)
    want2 = reference_module_forward(m, q64, v64, q64, pos, mask, ref_pts, shapes, params=params)
    want2.square().mean().backward()

    def gclose(got, ref, what):
        # float32 GEMMs + float32 kernels against float64 autograd.  The offset branch is looser: a sample that the
        # float32 location arithmetic puts on the other side of a bilinear cell boundary changes ITS grad_loc by O(1)
        got, ref = got.detach().cpu().double(), ref.double()
        err = (got - ref).abs().max().item()
        tol = 2e-3 if "sampling_offsets" in what else 2e-4
        assert err <= tol * ref.abs().max().item() + 1e-9, (what, err, ref.abs().max().item())

    gclose(q.grad, q64.grad, "grad query")
    for k, p in m.named_parameters():
        gclose(p.grad, params[k].grad, "grad " + k)


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


# ------------------------------------------------------------------------------------------------
# fused pre-op chain (softmax + sampling-location arithmetic inside the kernels)
# ------------------------------------------------------------------------------------------------
def _compose_unfused(value, shapes, lsi, offsets, logits, ref, P):
    """py:300-332 in PyTorch, then the unfused op."""
    from ir_ads_b200 import MultiScaleDeformableAttnFunction
    B, Q, H, L, _, _ = offsets.shape
    w = logits.softmax(-1).view(B, Q, H, L, P)
    if ref.shape[-1] == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + offsets / P * ref[:, :, None, :, None, 2:] * 0.5
    return MultiScaleDeformableAttnFunction.apply(value, shapes, lsi, loc.contiguous(), w.contiguous(), 64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ref_dim,D,L,P", [(2, 32, 4, 4), (4, 32, 4, 4), (2, 64, 3, 8), (4, 16, 2, 3), (2, 128, 1, 5)])
def test_fused_function_matches_composition(ref_dim, D, L, P, dtype):
    from ir_ads_b200.functional import MSDeformAttnFusedFunction, fused_supported
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(7)
    levels = [(11, 17), (6, 9), (3, 5), (2, 3)][:L]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q, H = 2, 53, 4
    value = torch.randn(B, S, H, D, device=DEV).to(dtype)
    assert fused_supported(value, L, P)
    offsets = torch.randn(B, Q, H, L, P, 2, device=DEV) * 3.0
    logits = torch.randn(B, Q, H, L * P, device=DEV) * 2.0
    ref = torch.rand(B, Q, L, ref_dim, device=DEV) * 1.2 - 0.1          # some samples fall outside the map
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:].abs() * 0.4 + 0.05
    go = torch.randn(B, Q, H * D, device=DEV).to(dtype)

    leaves_a = [t.clone().requires_grad_(True) for t in (value, offsets, logits, ref)]
    out_a = MSDeformAttnFusedFunction.apply(leaves_a[0], shapes, lsi, leaves_a[1], leaves_a[2], leaves_a[3])
    out_a.backward(go)
    leaves_b = [t.clone().requires_grad_(True) for t in (value, offsets, logits, ref)]
    out_b = _compose_unfused(leaves_b[0], shapes, lsi, leaves_b[1], leaves_b[2], leaves_b[3], P)
    out_b.backward(go)

    tol = 1e-5 if dtype == torch.float32 else 1e-2

    def close(a, b, what, t=tol):
        a, b = a.double(), b.double()
        assert (a - b).abs().max() <= t * b.abs().max() + 1e-6, (what, float((a - b).abs().max()), float(b.abs().max()))

    close(out_a, out_b, "out")
    close(leaves_a[0].grad, leaves_b[0].grad, "grad_value")
    # grad wrt offsets is discontinuous where a sample sits on a pixel boundary: same locations in both
    # paths (bit-identical arithmetic), so no masking is needed here
    close(leaves_a[1].grad, leaves_b[1].grad, "grad_offsets", 1e-4 if dtype == torch.float32 else tol)
    close(leaves_a[2].grad, leaves_b[2].grad, "grad_logits", 1e-4 if dtype == torch.float32 else tol)
    close(leaves_a[3].grad, leaves_b[3].grad, "grad_reference_points", 1e-4 if dtype == torch.float32 else tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ref_dim,D,L,P", [(2, 32, 4, 4), (4, 32, 4, 4), (2, 64, 3, 8), (4, 16, 2, 3)])
def test_fused_function_folds_the_padding_mask(ref_dim, D, L, P, dtype):
    """key_padding_mask inside the kernels == value.masked_fill(mask, 0) in front of the op (py:291-292):
    same output, same gradients for offsets / logits / reference points, zero grad_value under the mask."""
    from ir_ads_b200.functional import MSDeformAttnFusedFunction
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(11)
    levels = [(11, 17), (6, 9), (3, 5), (2, 3)][:L]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q, H = 3, 47, 4
    value = torch.randn(B, S, H, D, device=DEV).to(dtype)
    offsets = torch.randn(B, Q, H, L, P, 2, device=DEV) * 3.0
    logits = torch.randn(B, Q, H, L * P, device=DEV) * 2.0
    ref = torch.rand(B, Q, L, ref_dim, device=DEV) * 1.2 - 0.1
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:].abs() * 0.4 + 0.05
    go = torch.randn(B, Q, H * D, device=DEV).to(dtype)
    # image 0: batch-padding pattern (right and bottom margins of every level); image 1: random 30 %; image 2: none
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    start = 0
    for h, w in levels:
        m = torch.zeros(h, w, dtype=torch.bool, device=DEV)
        m[:, (2 * w) // 3:] = True
        m[(3 * h) // 4:, :] = True
        mask[0, start:start + h * w] = m.reshape(-1)
        start += h * w
    mask[1] = torch.rand(S, device=DEV) < 0.3

    leaves_a = [t.clone().requires_grad_(True) for t in (value, offsets, logits, ref)]
    out_a = MSDeformAttnFusedFunction.apply(leaves_a[0], shapes, lsi, leaves_a[1], leaves_a[2], leaves_a[3], mask)
    out_a.backward(go)
    leaves_b = [t.clone().requires_grad_(True) for t in (value, offsets, logits, ref)]
    out_b = MSDeformAttnFusedFunction.apply(leaves_b[0].masked_fill(mask[..., None, None], 0.0), shapes, lsi,
                                            leaves_b[1], leaves_b[2], leaves_b[3])
    out_b.backward(go)

    assert torch.equal(out_a, out_b)                       # same products in the same order: bit-identical
    assert leaves_a[0].grad[mask].abs().max() == 0
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b, what in zip(leaves_a, leaves_b, ("grad_value", "grad_offsets", "grad_logits", "grad_reference_points")):
        a, b = a.grad.double(), b.grad.double()
        assert (a - b).abs().max() <= tol * b.abs().max() + 1e-6, (what, float((a - b).abs().max()), float(b.abs().max()))
    # uint8 masks are taken as they are
    out_c = MSDeformAttnFusedFunction.apply(value, shapes, lsi, offsets, logits, ref, mask.to(torch.uint8) * 255)
    assert torch.equal(out_c, out_a)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_mask_discards_non_finite_values(dtype):
    """masked_fill(key_padding_mask, 0) DISCARDS whatever value_proj produced under the mask (py:291-292) -- an
    overflowed bf16 activation, say.  The fused kernels never load a masked corner, so NaN / Inf there cannot reach
    the output or any gradient: bit-identical to the same call with finite values under the mask."""
    from ir_ads_b200.functional import MSDeformAttnFusedFunction
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(5)
    levels = [(11, 17), (6, 9), (3, 5)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q, H, D, L, P = 2, 53, 4, 32, 3, 4
    value = torch.randn(B, S, H, D, device=DEV).to(dtype)
    offsets = torch.randn(B, Q, H, L, P, 2, device=DEV) * 3.0
    logits = torch.randn(B, Q, H, L * P, device=DEV)
    ref = torch.rand(B, Q, L, 2, device=DEV) * 1.2 - 0.1
    go = torch.randn(B, Q, H * D, device=DEV).to(dtype)
    mask = torch.rand(B, S, device=DEV) < 0.4
    poisoned = value.clone()
    poisoned[mask] = float("nan")
    poisoned[mask & (torch.rand(B, S, device=DEV) < 0.5)] = float("inf")
    res = []
    for val in (poisoned, value):
        leaves = [t.clone().requires_grad_(True) for t in (val, offsets, logits)]
        out = MSDeformAttnFusedFunction.apply(leaves[0], shapes, lsi, leaves[1], leaves[2], ref, mask)
        out.backward(go)
        res.append([out.detach()] + [t.grad for t in leaves])
    assert all(torch.isfinite(t).all() for t in res[0])
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])
    assert res[0][1][mask].abs().max() == 0
    gv_a, gv_b = res[0][1].float(), res[1][1].float()          # atomics: equal up to summation order
    assert (gv_a - gv_b).abs().max() <= (1e-5 if dtype == torch.float32 else 1e-2) * gv_b.abs().max() + 1e-6


@pytest.mark.parametrize("seed", list(range(8)))
def test_fused_function_random_shapes_against_composition(seed):
    """Seeded fuzz of the fused path (softmax + location affine + optional padding mask inside the kernels) against
    the step-by-step composition around the plain op, over levels / B / Q / H / D / P / ref_dim / dtype."""
    import numpy as np
    from ir_ads_b200.functional import MSDeformAttnFusedFunction, fused_supported
    from ir_ads_b200.workloads import level_tensors

    rng = np.random.default_rng(500 + seed)
    L = int(rng.integers(1, 5))
    levels, h, w = [], int(rng.integers(3, 30)), int(rng.integers(3, 30))
    for _ in range(L):
        levels.append((h, w))
        h, w = max(1, (h + 1) // 2), max(1, (w + 1) // 2)
    B, H = int(rng.integers(1, 4)), int(rng.choice([1, 2, 4, 8]))
    D = int(rng.choice([16, 32, 64, 128]))
    P = int(rng.integers(1, 9))
    Q = int(rng.integers(1, 150))
    ref_dim = int(rng.choice([2, 4]))
    dtype = torch.bfloat16 if rng.integers(0, 3) == 0 else torch.float32
    masked = bool(rng.integers(0, 2))
    torch.manual_seed(seed)
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(a * b for a, b in levels)
    value = torch.randn(B, S, H, D, device=DEV).to(dtype)
    assert fused_supported(value, L, P)
    offsets = torch.randn(B, Q, H, L, P, 2, device=DEV) * 3.0
    logits = torch.randn(B, Q, H, L * P, device=DEV) * 2.0
    ref = torch.rand(B, Q, L, ref_dim, device=DEV) * 1.3 - 0.15
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:].abs() * 0.4 + 0.05
    mask = (torch.rand(B, S, device=DEV) < 0.35) if masked else None
    go = torch.randn(B, Q, H * D, device=DEV).to(dtype)

    la = [t.clone().requires_grad_(True) for t in (value, offsets, logits)]
    MSDeformAttnFusedFunction.apply(la[0], shapes, lsi, la[1], la[2], ref, mask).backward(go)
    out_a = MSDeformAttnFusedFunction.apply(value, shapes, lsi, offsets, logits, ref, mask)
    lb = [t.clone().requires_grad_(True) for t in (value, offsets, logits)]
    vb = lb[0] if mask is None else lb[0].masked_fill(mask[..., None, None], 0.0)
    out_b = _compose_unfused(vb, shapes, lsi, lb[1], lb[2], ref, P)
    out_b.backward(go)
    what = f"levels={levels} B={B} Q={Q} H={H} D={D} P={P} ref_dim={ref_dim} {dtype} masked={masked}"
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    gtol = 1e-4 if dtype == torch.float32 else 1e-2
    for a, b, t, name in ((out_a, out_b, tol, "out"), (la[0].grad, lb[0].grad, tol, "grad_value"),
                          (la[1].grad, lb[1].grad, gtol, "grad_offsets"), (la[2].grad, lb[2].grad, gtol, "grad_logits")):
        a, b = a.double(), b.double()
        assert (a - b).abs().max() <= t * b.abs().max() + 1e-6, (name, what, float((a - b).abs().max()), float(b.abs().max()))


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_function_deterministic_backward(ref_dim, masked):
    """Deterministic mode: the fused Function keeps its fused forward and runs a bit-reproducible backward (the
    unfused deterministic kernels + the chain rule through softmax / affine): identical bits across runs on a
    contended shape, and the same gradients as the regular fused backward up to float summation order."""
    from ir_ads_b200 import MultiScaleDeformableAttention, functional
    from ir_ads_b200.functional import MSDeformAttnFusedFunction
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(5)
    levels = [(6, 9), (3, 5)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q, H, D, L, P = 2, 900, 2, 32, 2, 4
    value = torch.randn(B, S, H, D, device=DEV)
    offsets = torch.randn(B, Q, H, L, P, 2, device=DEV) * 2.0
    logits = torch.randn(B, Q, H, L * P, device=DEV)
    ref = torch.rand(B, Q, L, ref_dim, device=DEV)
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.4 + 0.05
    mask = (torch.rand(B, S, device=DEV) < 0.3) if masked else None
    go = torch.randn(B, Q, H * D, device=DEV)

    def grads():
        leaves = [t.clone().requires_grad_(True) for t in (value, offsets, logits)]
        MSDeformAttnFusedFunction.apply(leaves[0], shapes, lsi, leaves[1], leaves[2], ref, mask).backward(go)
        return [t.grad for t in leaves]

    base = grads()
    functional.set_deterministic(True)
    try:
        runs = [grads() for _ in range(3)]
        for r in runs[1:]:
            assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
        for a, b, name in zip(runs[0], base, ("grad_value", "grad_offsets", "grad_logits")):
            assert (a - b).abs().max() <= 1e-4 * b.abs().max() + 1e-6, name
        if masked:
            assert float(runs[0][0][mask].abs().max()) == 0.0
        m = MultiScaleDeformableAttention(embed_dim=64, num_heads=2, num_levels=2, num_points=4, dropout=0.0,
                                          batch_first=True).to(DEV)
        q = torch.randn(1, S, 64, device=DEV, requires_grad=True)
        runs = []
        for _ in range(2):
            q.grad = None
            m(q, reference_points=torch.rand(1, S, 2, 2, device=DEV, generator=torch.Generator(DEV).manual_seed(1)),
              spatial_shapes=shapes, level_start_index=lsi).square().sum().backward()
            runs.append(q.grad.clone())
        assert torch.equal(runs[0], runs[1])
    finally:
        functional.set_deterministic(False)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_fused_and_unfused_paths_agree(ref_dim):
    from ir_ads_b200 import MultiScaleDeformableAttention
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(3)
    levels = [(12, 17), (6, 9), (3, 5), (2, 3)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    B, Q = 2, (S if ref_dim == 2 else 31)
    m = MultiScaleDeformableAttention(dropout=0.0, batch_first=True).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.attention_weights.weight.normal_(0, 0.2)
    q = torch.randn(B, Q, 256, device=DEV)
    val = None if ref_dim == 2 else torch.randn(B, S, 256, device=DEV)
    ref_pts = torch.rand(B, Q, 4, ref_dim, device=DEV)
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    mask[0, :5] = True
    mask[1] = torch.rand(S, device=DEV) < 0.25            # the fused path folds the mask into the kernels
    res = {}
    for fused in (True, False):
        m.fuse_pre_ops = fused
        m.zero_grad(set_to_none=True)
        qq = q.clone().requires_grad_(True)
        out = m(qq, value=val, key_padding_mask=mask, reference_points=ref_pts, spatial_shapes=shapes,
                level_start_index=lsi)
        out.square().mean().backward()
        res[fused] = [out.detach(), qq.grad] + [p.grad.clone() for p in m.parameters()]
    for a, b in zip(res[True], res[False]):
        assert (a - b).abs().max() <= 2e-5 * b.abs().max() + 1e-7, float((a - b).abs().max() / b.abs().max())


def test_six_layer_stack_is_cuda_graph_capturable():
    """SURVEY 8f-2 (host side): no per-call device->host sync and no library-side allocation, so a
    6-layer encoder-style stack of module forward+backward replays from ONE CUDA graph."""
    from ir_ads_b200 import MultiScaleDeformableAttention
    from ir_ads_b200.workloads import level_tensors

    torch.manual_seed(1)
    levels = [(12, 17), (6, 9), (3, 5), (2, 3)]
    shapes, lsi = level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    layers = torch.nn.ModuleList([MultiScaleDeformableAttention(dropout=0.0, batch_first=True) for _ in range(6)]).to(DEV)
    for m in layers:
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.05)
            m.attention_weights.weight.normal_(0, 0.2)
    x = torch.randn(2, S, 256, device=DEV, requires_grad=True)
    ref_pts = torch.rand(2, S, 4, 2, device=DEV)

    def step():
        h = x
        for m in layers:
            h = m(h, reference_points=ref_pts, spatial_shapes=shapes, level_start_index=lsi)
        h.square().mean().backward()
        return h

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up: shape check cached, allocator primed
        for _ in range(3):
            layers.zero_grad(set_to_none=True); x.grad = None
            step()
    torch.cuda.current_stream().wait_stream(side)
    layers.zero_grad(set_to_none=True); x.grad = None
    eager = step().detach().clone()
    eager_grad = x.grad.detach().clone()
    layers.zero_grad(set_to_none=True); x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = step()
    x.grad.zero_()
    for p in layers.parameters():
        p.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager)
    assert torch.allclose(x.grad, eager_grad, rtol=1e-4, atol=1e-7)     # grad_value atomics reorder
