"""Generate tests/golden/msda_golden.npz from the REFERENCE itself (run in the authoring container).

The reference's ``multi_scale_deformable_attn_pytorch`` is loaded straight from
/root/reference/detrex/layers/multi_scale_deform_attn.py (it needs only torch; ``import detrex``
as a package does not work here -- SURVEY.md section 8c) and evaluated, with autograd for the three
gradients, on seeded inputs.  Inputs are stored as float32 so the fp32 and fp64 tests see
identical bits; golden outputs are stored for the float64 evaluation (the yardstick) and for the
reference's own float32 evaluation (to pin the torch port bit-for-bit).

/root/reference does not exist on the GPU box, so the vectors are committed; re-run this script
only when the case list changes:   python oracle/make_golden.py
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ir_ads_b200.workloads import make_inputs  # noqa: E402

REF_FILE = "/root/reference/detrex/layers/multi_scale_deform_attn.py"

# name -> (levels, B, Q, H, D, P, kind, dist, seed)
CASES = {
    # the reference unit test's own shape (tests/test_ms_deform_attn.py:34-38)
    "ref_test": ([(6, 4), (3, 2)], 1, 2, 2, 2, 2, "decoder", "test", 1),
    # the channel counts its gradcheck walks (tests/test_ms_deform_attn.py:132); 1025 kept tiny
    "d30": ([(6, 4), (3, 2)], 1, 2, 2, 30, 2, "decoder", "test", 2),
    "d32": ([(6, 4), (3, 2)], 1, 2, 2, 32, 2, "decoder", "test", 3),
    "d64": ([(6, 4), (3, 2)], 1, 2, 2, 64, 2, "decoder", "test", 4),
    "d71": ([(6, 4), (3, 2)], 1, 2, 2, 71, 2, "decoder", "test", 5),
    "d1025": ([(3, 2), (2, 2)], 1, 2, 1, 1025, 2, "decoder", "test", 6),
    # out-of-range locations, exact pixel centres, integer coordinates, zero weights
    "edge": ([(6, 4), (3, 2), (2, 3)], 2, 6, 3, 8, 3, "decoder", "edge", 7),
    "edge_d32": ([(9, 7), (5, 4), (3, 2), (2, 1)], 1, 7, 8, 32, 4, "decoder", "edge", 8),
    # DINO-shaped (8 heads x 32 ch, 4 levels x 4 points), encoder (Q = S) and decoder forms
    "enc_mini": ([(6, 10), (3, 5), (2, 3), (1, 2)], 1, 0, 8, 32, 4, "encoder", "model", 9),
    "dec_mini": ([(7, 11), (4, 6), (2, 3), (1, 2)], 2, 12, 8, 32, 4, "decoder", "model", 10),
    # 5 levels x 8 points (cfg 5 form)
    "stress_mini": ([(4, 4), (2, 2), (2, 1), (1, 1), (1, 1)], 1, 0, 4, 32, 8, "encoder", "model", 11),
}


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_msda", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.multi_scale_deformable_attn_pytorch


def run_reference(fn, value, shapes, loc, w, go, dtype):
    v = value.to(dtype).clone().requires_grad_(True)
    lo = loc.to(dtype).clone().requires_grad_(True)
    ww = w.to(dtype).clone().requires_grad_(True)
    out = fn(v, shapes, lo, ww)
    out.backward(go.to(dtype))
    return out.detach(), v.grad, lo.grad, ww.grad


def main():
    fn = load_reference()
    blob = {}
    for name, (levels, B, Q, H, D, P, kind, dist, seed) in CASES.items():
        value, shapes, lsi, loc, w = make_inputs(levels, B, Q, H, D, P, kind, dist, seed)
        g = torch.Generator().manual_seed(1000 + seed)
        go = torch.randn(B, loc.shape[1], H * D, generator=g)
        blob[f"{name}/value"] = value.numpy()
        blob[f"{name}/shapes"] = shapes.numpy()
        blob[f"{name}/lsi"] = lsi.numpy()
        blob[f"{name}/loc"] = loc.numpy()
        blob[f"{name}/w"] = w.numpy()
        blob[f"{name}/grad_out"] = go.numpy()
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            out, gv, gl, gw = run_reference(fn, value, shapes, loc, w, go, dt)
            blob[f"{name}/{tag}/out"] = out.numpy()
            blob[f"{name}/{tag}/grad_value"] = gv.numpy()
            blob[f"{name}/{tag}/grad_loc"] = gl.numpy()
            blob[f"{name}/{tag}/grad_w"] = gw.numpy()
        print(f"{name}: value{tuple(value.shape)} loc{tuple(loc.shape)} out{tuple(out.shape)}")
    dst = os.path.join(ROOT, "tests", "golden", "msda_golden.npz")
    np.savez_compressed(dst, **blob)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB; torch", torch.__version__)


if __name__ == "__main__":
    main()
