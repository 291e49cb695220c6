#!/usr/bin/env python
"""Forward / backward results of a kernel-variant library against the product library on the same inputs
(both run in their own process; MSDA_B200_LIB selects the library).
Usage (GPU box): python tools/compare_variant.py build/variants/lib_<name>.so [workloads]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def produce(path, workloads):
    import torch
    import ir_ads_b200
    from ir_ads_b200.workloads import WORKLOADS, make_workload_inputs
    res = {}
    for name in workloads:
        wl = WORKLOADS[name]
        for dist in ("model", "test", "edge"):
            value, shapes, lsi, loc, w = make_workload_inputs(wl, dist, seed=3, device="cuda:0", batch=1)
            go = torch.randn(1, wl.queries, wl.num_heads * wl.head_dim, device="cuda:0",
                             generator=torch.Generator(device="cuda:0").manual_seed(5)).to(value.dtype)
            out = ir_ads_b200.ms_deform_attn_forward(value, shapes, lsi, loc, w, 64)
            gv, gl, gw = ir_ads_b200.ms_deform_attn_backward(value, shapes, lsi, loc, w, go, 64)
            res[f"{name}/{dist}"] = [t.float().cpu() for t in (out, gv, gl, gw)]
    torch.save(res, path)


def main():
    if sys.argv[1] == "--produce":
        produce(sys.argv[2], sys.argv[3].split(","))
        return
    import torch
    variant = sys.argv[1]
    workloads = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    with tempfile.TemporaryDirectory() as tmp:
        paths = {}
        for tag, lib in (("product", ""), ("variant", variant)):
            env = dict(os.environ)
            if lib:
                env["MSDA_B200_LIB"] = lib
            else:
                env.pop("MSDA_B200_LIB", None)
            paths[tag] = os.path.join(tmp, tag + ".pt")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--produce", paths[tag], workloads], env=env, check=True)
        a, b = torch.load(paths["product"]), torch.load(paths["variant"])
    worst = 0.0
    for key in a:
        for nm, x, y in zip(("out", "grad_value", "grad_loc", "grad_w"), a[key], b[key]):
            err = (x - y).abs().max().item()
            rel = err / max(x.abs().max().item(), 1e-30)
            worst = max(worst, rel)
            print(f"{key:14s} {nm:10s} max|diff| {err:.3e}  rel-to-max {rel:.3e}  bit-equal {bool(torch.equal(x, y))}")
    print("worst rel-to-max", worst)
    sys.exit(0 if worst < 2e-6 else 1)


if __name__ == "__main__":
    main()
