"""Property tests of the CPU oracle (hypothesis): they pin the claims the GPU parity tests build on --
the compensated float coordinate split picks the exact cell, the operator is linear in value and its
backward is the adjoint of its forward."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import msda_c

LEVELS = np.array([[7, 9], [4, 5], [2, 3]], dtype=np.int64)
LSI = np.array([0, 63, 83], dtype=np.int64)
S = 89


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([1, 3, 21, 167, 1333]), st.floats(-0.3, 1.3))
def test_compensated_split_is_the_exact_floor(seed, size, centre):
    """Locations clustered within a few float ulps of cell boundaries -- where a rounded product picks the
    wrong cell -- still get the cell and (to 2^-23) the fraction of the exact rational value."""
    rng = np.random.default_rng(seed)
    k = np.floor(centre * size)
    base = (k + rng.choice([0.0, 0.5, 1.0], size=64)) / size               # boundaries, centres
    loc = (base + rng.integers(-3, 4, size=64) * np.spacing(np.float32(abs(centre) + 1e-3))).astype(np.float32)
    pts = np.zeros((1, 64, 1, 1, 1, 2), dtype=np.float32)
    pts[0, :, 0, 0, 0, 0] = loc
    pts[0, :, 0, 0, 0, 1] = 0.5
    shapes = np.array([[1, size]], dtype=np.int64)
    lsi = np.array([0], dtype=np.int64)
    comp, comp_frac = msda_c.bookkeeping(pts, shapes, lsi, 1, size, 1, 16, True, msda_c.COORD_COMPENSATED)
    coord = loc.astype(np.float64) * size - 0.5                   # exact: float32 inputs, 53-bit arithmetic
    tol = 2.0 ** -23
    for i in range(64):
        inside = -1 < coord[i] < size
        gated = bool((comp[i] == -1).all())
        if gated:
            # gated out: exactly outside, or so close to the gate that the float fraction cannot tell
            assert (not inside) or min(abs(coord[i] + 1), abs(coord[i] - size)) <= tol or comp_frac[i, 0] == 0.0
            if inside and comp_frac[i, 0] == 0.0 and min(abs(coord[i] + 1), abs(coord[i] - size)) > tol:
                # in range but both x corners of the y0 row report invalid cannot happen for H_l = 1
                raise AssertionError(("in-range point gated", float(coord[i])))
            continue
        assert inside or min(abs(coord[i] + 1), abs(coord[i] - size)) <= tol
        x0 = comp[i, 0] // 16 if comp[i, 0] >= 0 else comp[i, 1] // 16 - 1
        got = x0 + float(comp_frac[i, 0])
        # cell + fraction is the exact coordinate to the last float bit of the fraction; when the exact fraction
        # rounds up to 1.0 the pair is (cell, 1.0) == (cell + 1, 0.0): the same sample
        assert abs(got - coord[i]) <= tol, (float(loc[i]), size, got, float(coord[i]))
        assert x0 == np.floor(coord[i]) or abs(coord[i] - round(coord[i])) <= tol


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(1, 4), st.integers(1, 3))
def test_forward_is_linear_and_backward_is_its_adjoint(seed, B, H, P):
    rng = np.random.default_rng(seed)
    D, Q, L = 5, 6, 3
    value = rng.standard_normal((B, S, H, D))
    value2 = rng.standard_normal((B, S, H, D))
    loc = rng.uniform(-0.2, 1.2, (B, Q, H, L, P, 2))
    w = rng.uniform(0, 1, (B, Q, H, L, P))
    go = rng.standard_normal((B, Q, H * D))
    f = lambda v: msda_c.forward(v, LEVELS, LSI, loc, w, np.float64)
    a, b = 1.7, -0.4
    assert np.allclose(f(a * value + b * value2), a * f(value) + b * f(value2), rtol=1e-12, atol=1e-12)
    gv, gl, gw = msda_c.backward(go, value, LEVELS, LSI, loc, w, np.float64)
    # <f(v), go> == <v, f^T(go)>   and   d/dw <f, go> == gw (f is linear in w as well)
    assert np.isclose((f(value) * go).sum(), (value * gv).sum(), rtol=1e-10)
    assert np.isclose((gw * w).sum(), (f(value) * go).sum(), rtol=1e-10)
    # finite difference of one location component (away from cell boundaries the op is smooth)
    i = tuple(rng.integers(0, n) for n in loc.shape)
    eps = 1e-6
    lp, lm = loc.copy(), loc.copy()
    lp[i] += eps
    lm[i] -= eps
    fd = ((msda_c.forward(value, LEVELS, LSI, lp, w, np.float64) - msda_c.forward(value, LEVELS, LSI, lm, w, np.float64)) * go).sum() / (2 * eps)
    size = LEVELS[i[3], 1 - i[5]]
    px = loc[i] * size - 0.5
    if abs(px - round(px)) > 1e-3:
        assert np.isclose(fd, gl[i], rtol=1e-4, atol=1e-6)
