#!/usr/bin/env python
"""Step-by-step probe of a libmsda_b200.so on the GPU box (used to localise a host-side hang in round 2).
Usage: python tools/debug/probe_lib.py <lib.so> <step>   steps: book | generic | fast | bwd"""
import ctypes
import sys
import time

import torch

lib = ctypes.CDLL(sys.argv[1])
step = sys.argv[2]
vp, i = ctypes.c_void_p, ctypes.c_int
dev = "cuda:0"
torch.zeros(1, device=dev)
torch.cuda.synchronize()
print("abi", lib.msda_abi_version(), flush=True)
levels = [(8, 10), (4, 5)]
shapes = torch.tensor(levels, dtype=torch.long, device=dev)
lsi = torch.tensor([0, 80], dtype=torch.long, device=dev)
B, S, H, L, Q, P = 1, 100, 2, 2, 7, 4
D = 30 if step == "generic" else 32
value = torch.randn(B, S, H, D, device=dev)
loc = torch.rand(B, Q, H, L, P, 2, device=dev)
w = torch.rand(B, Q, H, L, P, device=dev)
out = torch.empty(B, Q, H * D, device=dev)
st = vp(torch.cuda.current_stream().cuda_stream)
p = lambda t: vp(t.data_ptr())
torch.cuda.synchronize()
t0 = time.time()
print("calling", step, flush=True)
if step == "book":
    offs = torch.empty(B * Q * H * L * P, 4, dtype=torch.long, device=dev)
    frac = torch.empty(B * Q * H * L * P, 2, device=dev)
    lib.msda_debug_bookkeeping.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, vp, vp]
    s = lib.msda_debug_bookkeeping(st, p(loc), p(shapes), p(lsi), B, S, H, D, L, Q, P, p(offs), p(frac))
elif step in ("generic", "fast"):
    lib.msda_forward.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, i, ctypes.c_uint]
    s = lib.msda_forward(st, p(value), p(shapes), p(lsi), p(loc), p(w), B, S, H, D, L, Q, P, p(out), 0, 0)
print("returned", s, round(time.time() - t0, 3), flush=True)
torch.cuda.synchronize()
print("synced", round(time.time() - t0, 3), float(out.abs().sum()) if step != "book" else "", flush=True)
