"""tests/golden/dcnv3_golden.npz from the REFERENCE's dcnv3_core_pytorch + autograd
(/root/reference/detrex/layers/dcn_v3.py:121-166, loaded standalone: it needs only torch).
Run in the authoring container:  python oracle/make_golden_dcnv3.py"""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILE = "/root/reference/detrex/layers/dcn_v3.py"

# name -> (N, H, W, group, group_channels, kernel, stride, pad, dilation, offset_scale, seed)
CASES = {
    "k3_s1": (2, 7, 9, 2, 16, 3, 1, 1, 1, 1.0, 1),
    "k3_s2": (1, 9, 8, 4, 16, 3, 2, 1, 1, 2.0, 2),
    "k3_d2": (1, 8, 8, 2, 32, 3, 1, 2, 2, 0.5, 3),
    "k5": (1, 9, 9, 2, 16, 5, 1, 2, 1, 1.0, 4),
    "k1": (1, 5, 6, 2, 16, 1, 1, 0, 1, 1.5, 5),
}


def main():
    spec = importlib.util.spec_from_file_location("ref_dcn", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fn = mod.dcnv3_core_pytorch
    blob = {}
    for name, (N, H, W, G, C, k, s, p, d, scale, seed) in CASES.items():
        g = torch.Generator().manual_seed(seed)
        Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
        Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
        K = k * k
        inp = torch.randn(N, H, W, G * C, generator=g)
        off = torch.randn(N, Ho, Wo, G * K * 2, generator=g) * 1.5
        mask = torch.softmax(torch.randn(N, Ho, Wo, G, K, generator=g), -1).reshape(N, Ho, Wo, G * K)
        go = torch.randn(N, Ho, Wo, G * C, generator=g)
        geom = (k, k, s, s, p, p, d, d, G, C, scale)
        blob[f"{name}/geom"] = np.array([N, H, W, G, C, k, s, p, d], dtype=np.int64)
        blob[f"{name}/scale"] = np.array([scale], dtype=np.float64)
        for key, t in (("input", inp), ("offset", off), ("mask", mask), ("grad_out", go)):
            blob[f"{name}/{key}"] = t.numpy()
        i = inp.double().requires_grad_(True)
        o = off.double().requires_grad_(True)
        m = mask.double().requires_grad_(True)
        out = fn(i, o, m, *geom)
        out.backward(go.double())
        blob[f"{name}/out"] = out.detach().numpy()
        blob[f"{name}/grad_input"] = i.grad.numpy()
        blob[f"{name}/grad_offset"] = o.grad.numpy()
        blob[f"{name}/grad_mask"] = m.grad.numpy()
        print(name, tuple(inp.shape), "->", tuple(out.shape))
    dst = os.path.join(ROOT, "tests", "golden", "dcnv3_golden.npz")
    np.savez_compressed(dst, **blob)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
