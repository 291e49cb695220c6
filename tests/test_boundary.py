"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports what
include/msda.h declares, the Python surface has the reference's names and signatures, the
product never touches oracle/, and the workload byte model matches SURVEY.md section 8(d)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import ir_ads_b200
from ir_ads_b200 import _lib, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "msda.h")).read()
    declared = set(re.findall(r"\b(msda_[a-z_]+)\s*\(", header))
    assert {"msda_forward", "msda_backward", "msda_backward_workspace_bytes"} <= declared
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/msda.h but not exported"
    assert _lib.lib().msda_abi_version() == _lib.ABI_VERSION
    assert _lib.lib().msda_status_string(0) == b"MSDA_OK"


def test_dispatch_names():
    name = lambda D, dt, flags=0, bwd=0, L=4, P=4: _lib.lib().msda_dispatch_name(D, L, P, 22223, 8, dt, flags, bwd).decode()
    assert name(32, _lib.MSDA_F32) == "fwd_fast_d32_f32"
    assert name(32, _lib.MSDA_BF16, bwd=1) == "bwd_fast_d32_bf16"
    assert name(30, _lib.MSDA_F32) == "fwd_generic_f32"
    assert name(32, _lib.MSDA_F64) == "fwd_generic_f64"
    assert name(32, _lib.MSDA_F32, _lib.FLAG_FORCE_GENERIC) == "fwd_generic_f32"
    assert name(32, _lib.MSDA_F32, L=17) == "fwd_generic_f32"


def test_argument_errors_are_returned_not_printed():
    h = _lib.lib()
    st = h.msda_forward(None, None, None, None, None, None, -1, 1, 1, 1, 1, 1, 1, None, 0, 0)
    assert st == 1 and b"negative" in h.msda_last_error_message()
    st = h.msda_forward(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, None, 9, 0)
    assert st == 1 and b"dtype" in h.msda_last_error_message()
    with pytest.raises(_lib.MSDAError):
        _lib.check(st, "probe")
    # empty problem: a no-op, not a launch error
    assert h.msda_forward(None, None, None, None, None, None, 0, 5, 8, 32, 4, 7, 4, None, 0, 0) == 0
    assert h.msda_backward_workspace_bytes(2, 10, 8, 32, 4, 3, 4, _lib.MSDA_BF16, 0) == 2 * 10 * 8 * 32 * 4
    assert h.msda_backward_workspace_bytes(2, 10, 8, 32, 4, 3, 4, _lib.MSDA_F32, 0) == 0
    # deterministic: fixed-point accumulators on generic shapes, bins + entries of the sorted path on fast shapes
    assert h.msda_backward_workspace_bytes(2, 10, 8, 30, 4, 3, 4, _lib.MSDA_F32, _lib.FLAG_DETERMINISTIC) == 2 * 10 * 8 * 30 * 8 + 64
    det = _lib.FLAG_DETERMINISTIC
    assert h.msda_backward_workspace_bytes(2, 10, 8, 32, 4, 3, 4, _lib.MSDA_F32, det | _lib.FLAG_DET_ATOMIC) == 2 * 10 * 8 * 32 * 8 + 64
    assert h.msda_backward_workspace_bytes(2, 10, 8, 32, 4, 3, 4, _lib.MSDA_F32, det) >= 2 * 3 * 8 * 4 * 4 * 16


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ir_ads_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "msda_oracle" not in text or f.endswith(".cuh"), f  # .cuh may cite it in comments only
                assert "grid_sample(" not in text, f


def test_public_surface_matches_reference():
    from ir_ads_b200 import MultiScaleDeformableAttention, MultiScaleDeformableAttnFunction

    sig = inspect.signature(MultiScaleDeformableAttention.__init__)
    assert list(sig.parameters)[1:] == ["embed_dim", "num_heads", "num_levels", "num_points", "img2col_step",
                                        "dropout", "batch_first"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [256, 8, 4, 4, 64, 0.1, False]
    fsig = inspect.signature(MultiScaleDeformableAttention.forward)
    assert list(fsig.parameters)[1:] == ["query", "key", "value", "identity", "query_pos", "key_padding_mask",
                                         "reference_points", "spatial_shapes", "level_start_index", "kwargs"]
    asig = inspect.signature(MultiScaleDeformableAttnFunction.forward)
    assert list(asig.parameters) == ["ctx", "value", "value_spatial_shapes", "value_level_start_index",
                                     "sampling_locations", "attention_weights", "im2col_step"]
    assert callable(ir_ads_b200.ms_deform_attn_forward) and callable(ir_ads_b200.ms_deform_attn_backward)
    # the detrex._C replacement exposes exactly the four functions of csrc/vision.cpp:54-59
    from ir_ads_b200 import _C
    assert sorted(_C.__all__) == ["dcnv3_backward", "dcnv3_forward", "ms_deform_attn_backward", "ms_deform_attn_forward"]
    assert list(inspect.signature(_C.ms_deform_attn_backward).parameters) == [
        "value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight", "grad_output", "im2col_step"]
    assert list(inspect.signature(_C.dcnv3_forward).parameters)[:4] == ["input", "offset", "mask", "kernel_h"]


def test_module_parameters_and_init():
    from ir_ads_b200 import MultiScaleDeformableAttention

    torch.manual_seed(0)
    m = MultiScaleDeformableAttention()
    sd = m.state_dict()
    assert list(sd) == ["sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight",
                        "attention_weights.bias", "value_proj.weight", "value_proj.bias", "output_proj.weight",
                        "output_proj.bias"]
    assert sum(p.numel() for p in m.parameters()) == 230272
    assert (m.im2col_step, m.embed_dim, m.num_heads, m.num_levels, m.num_points, m.batch_first) == (64, 256, 8, 4, 4, False)
    assert torch.count_nonzero(sd["sampling_offsets.weight"]) == 0 and torch.count_nonzero(sd["attention_weights.weight"]) == 0
    bias = sd["sampling_offsets.bias"].view(8, 4, 4, 2)
    # head 0 points along +x, head 2 along +y, scaled by (p+1); identical on every level (py:205-217)
    assert torch.allclose(bias[0, :, :, 0], torch.arange(1.0, 5.0).expand(4, 4)) and torch.allclose(bias[0, :, :, 1], torch.zeros(4, 4), atol=1e-6)
    assert torch.allclose(bias[2, :, :, 1], torch.arange(1.0, 5.0).expand(4, 4))
    assert torch.allclose(bias[1, 0, 3], torch.tensor([4.0, 4.0]), atol=1e-5)
    bound = (6.0 / 512) ** 0.5
    assert sd["value_proj.weight"].abs().max() <= bound and sd["value_proj.weight"].std() > 0.5 * bound / 3 ** 0.5
    with pytest.raises(ValueError):
        MultiScaleDeformableAttention(embed_dim=250, num_heads=8)


def test_cpu_tensors_raise():
    from ir_ads_b200 import MultiScaleDeformableAttnFunction

    v, shapes, lsi, loc, w = workloads.make_inputs([(4, 4)], 1, 3, 2, 16, 2, "decoder", "test", 0)
    with pytest.raises(RuntimeError, match="CPU"):
        MultiScaleDeformableAttnFunction.apply(v, shapes, lsi, loc, w, 64)


def test_workload_byte_model_matches_survey():
    w2 = workloads.WORKLOADS["cfg2"]
    assert w2.spatial_size == 22223 and w2.points == 22_756_352
    fwd, bwd = w2.algorithmic_bytes()
    assert round(fwd / 1e6, 1) == 637.2 and round(bwd / 1e6, 1) == 1092.3
    w3 = workloads.WORKLOADS["cfg3"]
    assert w3.points == 2_048_000
    f3, b3 = w3.algorithmic_bytes()
    assert round((f3 + b3) / 1e6, 1) == 363.2
    w5 = workloads.WORKLOADS["cfg5"]
    f5, b5 = w5.algorithmic_bytes()
    assert w5.points == 55_869_440 and round((f5 + b5) / 1e6, 1) == 2905.2


def test_workload_inputs_are_seeded_and_shaped():
    a = workloads.make_inputs([(6, 10), (3, 5)], 2, 0, 4, 16, 4, "encoder", "model", 5)
    b = workloads.make_inputs([(6, 10), (3, 5)], 2, 0, 4, 16, 4, "encoder", "model", 5)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    value, shapes, lsi, loc, w = a
    assert value.shape == (2, 75, 4, 16) and loc.shape == (2, 75, 4, 2, 4, 2) and w.shape == (2, 75, 4, 2, 4)
    assert lsi.tolist() == [0, 60] and torch.allclose(w.sum((-1, -2)), torch.ones(2, 75, 4), atol=1e-5)


def test_misaligned_views_are_copied_before_the_c_abi_sees_them():
    """include/msda.h: every tensor pointer must be 16-byte aligned (128-bit accesses); the wrappers copy a view that is not."""
    from ir_ads_b200.functional import _aligned
    flat = torch.arange(65, dtype=torch.float32)
    view = flat[1:]
    assert view.data_ptr() % 16 != 0
    fixed = _aligned(view)
    assert fixed.data_ptr() % 16 == 0 and torch.equal(fixed, view) and fixed.data_ptr() != view.data_ptr()
    ok = torch.zeros(64)
    assert _aligned(ok) is ok and _aligned(None) is None
    empty = flat[1:1]
    assert _aligned(empty) is empty          # nothing to dereference


def test_valid_corner_fraction_matches_the_oracle_bookkeeping():
    """bench.py scales the on-chip ceilings by the share of corners that lie inside their level (the rows the kernels
    load and scatter); that share is the oracle's count of in-map corners."""
    import numpy as np
    from oracle import msda_c
    levels = [(9, 13), (5, 7), (3, 4)]
    for dist in ("model", "test", "edge"):
        value, shapes, lsi, loc, w = workloads.make_inputs(levels, 2, 50, 4, 32, 4, "decoder", dist, 7)
        B, S, H, D = value.shape
        offs, _ = msda_c.bookkeeping(loc.numpy(), shapes.numpy(), lsi.numpy(), B, S, H, D, False)
        want = float((offs >= 0).sum()) / offs.size
        got = workloads.valid_corner_fraction(loc, levels)
        assert abs(got - want) < 1e-12, (dist, got, want)


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/msda.h must compile as C99 (no C++ in the signatures) and a plain C
    program must link against libmsda_b200.so and get error codes back -- here without a GPU: calls that fail their
    argument checks never reach the device."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "msda.h"
int main(void) {
  if (msda_abi_version() != MSDA_ABI_VERSION) return 1;
  if (strcmp(msda_status_string(MSDA_OK), "MSDA_OK") != 0) return 2;
  /* negative dimension: refused before anything touches the device */
  int s = msda_forward(NULL, NULL, NULL, NULL, NULL, NULL, -1, 4, 2, 32, 1, 1, 1, NULL, MSDA_F32, 0u);
  if (s != MSDA_ERR_INVALID_ARGUMENT) return 3;
  if (strlen(msda_last_error_message()) == 0) return 4;
  /* unknown dtype tag */
  if (msda_forward(NULL, NULL, NULL, NULL, NULL, NULL, 1, 4, 2, 32, 1, 1, 1, NULL, 77, 0u) != MSDA_ERR_INVALID_ARGUMENT) return 5;
  /* an empty problem is a no-op */
  if (msda_forward(NULL, NULL, NULL, NULL, NULL, NULL, 0, 4, 2, 32, 1, 1, 1, NULL, MSDA_F32, 0u) != MSDA_OK) return 6;
  if (msda_backward_workspace_bytes(2, 10, 2, 32, 1, 3, 2, MSDA_BF16, 0u) != (size_t)2 * 10 * 2 * 32 * 4) return 7;
  printf("c abi ok\n");
  return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir,
           "-Wl,-rpath,/usr/local/cuda/lib64", "-Wl,--allow-shlib-undefined"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "c abi ok" in run.stdout, (run.returncode, run.stdout, run.stderr)


def test_count_pass_division_constants_are_exact():
    """The deterministic path's count pass divides point indices by L*P, P, H and Q with multiply-high constants
    (csrc/msda_det.cuh, FastDiv) instead of `/`: the quotient must be exact for every n < 2^31.  Host arithmetic only."""
    import random
    h = _lib.lib()
    rng = random.Random(7)
    top = (1 << 31) - 1
    divisors = [1, 2, 3, 4, 5, 7, 8, 16, 20, 40, 64, 100, 255, 256, 257, 300, 900, 2000, 17821, 21824, 22223, 65535, 65536,
                65537, 1000003, 1 << 30, top] + [rng.randrange(1, 1 << 20) for _ in range(40)]
    for d in divisors:
        edge = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, top - 1, top, (top // d) * d, (top // d) * d - 1]
        for n in edge + [rng.randrange(0, 1 << 31) for _ in range(200)]:
            if 0 <= n <= top:
                assert h.msda_debug_fastdiv(n, d) == n // d, (n, d)


def test_backward_workspace_sizes():
    """Host logic of msda_backward_workspace_bytes (no GPU work): nothing for the float path, float accumulators for
    bf16, and for the deterministic sorted path room for the 16-byte entries, the bin tables, the per-warp maxima of the
    entry-filing backward (two floats per warp) and -- dense problems only -- the 64-bit accumulators."""
    h = _lib.lib()
    B, S, H, D, L, Q, P = 8, 22223, 8, 32, 4, 22223, 4
    n_value, n_pts, rows = B * S * H * D, B * Q * H * L * P, B * Q * H
    assert h.msda_backward_workspace_bytes(B, S, H, D, L, Q, P, _lib.MSDA_F32, 0) == 0
    assert h.msda_backward_workspace_bytes(B, S, H, D, L, Q, P, _lib.MSDA_BF16, 0) == n_value * 4
    det = h.msda_backward_workspace_bytes(B, S, H, D, L, Q, P, _lib.MSDA_F32, _lib.FLAG_DETERMINISTIC)
    warps = rows * D // 128
    assert det >= n_pts * 16 + n_value * 8 + warps * 8 and det % 256 == 0          # dense: encoder self-attention
    sparse = h.msda_backward_workspace_bytes(B, S, H, D, L, 300, P, _lib.MSDA_F32, _lib.FLAG_DETERMINISTIC)
    assert B * 300 * H * L * P * 16 <= sparse < n_value * 8                         # sparse: no accumulators
    atomic = h.msda_backward_workspace_bytes(B, S, H, D, L, Q, P, _lib.MSDA_F32,
                                             _lib.FLAG_DETERMINISTIC | _lib.FLAG_DET_ATOMIC)
    assert atomic == n_value * 8 + 64                                              # fixed-point reds: accumulators only
    generic = h.msda_backward_workspace_bytes(B, S, H, 30, L, Q, P, _lib.MSDA_F32, _lib.FLAG_DETERMINISTIC)
    assert generic == B * S * H * 30 * 8 + 64
