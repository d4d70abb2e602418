"""The error bounds E and F of the pooled kernel's division-free plane pre-filter (traverse_pooled.cuh).

The pre-filter compares  a = fma(n.x,d.x, fma(n.y,d.y, n.z*d.z))  and  b = fma(-n.x,o.x, fma(-n.y,o.y, fma(-n.z,o.z, n.v0)))
with the cell's parameter range; it may only discard a triangle if the reference's own  denom = n.d  and  nom = n.(v0 - o)
(lib/intersection.h:40-49, unfused fp32, left to right) cannot give  lo <= nom/denom <= hi.  That holds if
|a - denom| <= F = 16u |d|_1  and  |b - nom| <= E = 32u (3 S + |o|_1),  u = 2^-24, S = largest |coordinate| of the scene.
Here both pairs are evaluated in numpy float32 (FMA emulated through float64, exact for one fused step up to double
rounding) over adversarial inputs -- far cameras, tiny and huge scenes, grazing rays, heavy cancellation -- and the
observed errors must stay below the bounds with the safety factor the kernel claims (~3).
"""
import numpy as np

U = np.float32(2.0 ** -24)
f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def reference_nom_denom(n, v0, o, d):
    denom = ((n[:, 0] * d[:, 0]) + (n[:, 1] * d[:, 1])) + (n[:, 2] * d[:, 2])
    nom = ((n[:, 0] * (v0[:, 0] - o[:, 0])) + (n[:, 1] * (v0[:, 1] - o[:, 1]))) + (n[:, 2] * (v0[:, 2] - o[:, 2]))
    return nom.astype(np.float32), denom.astype(np.float32)


def prefilter_a_b(n, v0, o, d):
    dp = (n[:, 0].astype(np.float64) * v0[:, 0] + n[:, 1].astype(np.float64) * v0[:, 1] + n[:, 2].astype(np.float64) * v0[:, 2]).astype(np.float32)
    a = fma(n[:, 0], d[:, 0], fma(n[:, 1], d[:, 1], (n[:, 2] * d[:, 2]).astype(np.float32)))
    b = fma(-n[:, 0], o[:, 0], fma(-n[:, 1], o[:, 1], fma(-n[:, 2], o[:, 2], dp)))
    return a, b


def unit(v):
    v = v.astype(np.float64)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def test_error_bounds_hold_with_margin():
    rng = np.random.RandomState(11)
    N = 400000
    worst_e = worst_f = 0.0
    for scale in (1e-3, 1.0, 50.0, 1e4):
        for far in (0.0, 1.0, 30.0, 1000.0):          # camera distance in units of the scene scale
            for dir_len in (1.0, 1e-3, 37.0):         # primaries are not normalised (main.cpp:203-209)
                S = f32(scale)
                v0 = (rng.uniform(-1, 1, (N, 3)) * scale).astype(np.float32)
                S_eff = f32(np.abs(v0).max())
                n = unit(rng.normal(size=(N, 3)))
                o = (rng.uniform(-1, 1, (N, 3)) * scale + unit(rng.normal(size=(N, 3))) * far * scale).astype(np.float32)
                d = unit(rng.normal(size=(N, 3))) * f32(dir_len)
                # a third of the rays graze their plane, a third start (nearly) on it: the cancellation cases
                k = N // 3
                t = np.cross(n[:k].astype(np.float64), rng.normal(size=(k, 3)))
                d[:k] = (unit(t.astype(np.float32)).astype(np.float64) + 1e-6 * rng.normal(size=(k, 3))).astype(np.float32) * f32(dir_len)
                o[k:2 * k] = (v0[k:2 * k].astype(np.float64) + 1e-4 * scale * n[k:2 * k]).astype(np.float32)
                nom, denom = reference_nom_denom(n, v0, o, d)
                a, b = prefilter_a_b(n, v0, o, d)
                E = f32(1.9073486e-6) * (f32(3) * S_eff + np.abs(o).sum(1, dtype=np.float32))
                F = f32(9.5367432e-7) * np.abs(d).sum(1, dtype=np.float32)
                re = float(np.max(np.abs(b.astype(np.float64) - nom.astype(np.float64)) / E.astype(np.float64)))
                rf = float(np.max(np.abs(a.astype(np.float64) - denom.astype(np.float64)) / F.astype(np.float64)))
                worst_e, worst_f = max(worst_e, re), max(worst_f, rf)
                assert re <= 1.0 and rf <= 1.0, (scale, far, dir_len, re, rf)
    # the kernel documents a safety factor of ~3 on both bounds
    assert worst_e <= 0.5 and worst_f <= 0.5, (worst_e, worst_f)


def test_range_implication():
    # lo <= nom/denom <= hi (reference arithmetic, IEEE division)  ==>  A <= F or (B + E >= lo (A - F) and B - E <= hi (A + F))
    rng = np.random.RandomState(5)
    N = 1000000
    v0 = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    n = unit(rng.normal(size=(N, 3)))
    o = rng.uniform(-1.2, 1.2, (N, 3)).astype(np.float32)
    d = unit(rng.normal(size=(N, 3)))
    nom, denom = reference_nom_denom(n, v0, o, d)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (nom / denom).astype(np.float32)
    a, b = prefilter_a_b(n, v0, o, d)
    E = f32(1.9073486e-6) * (f32(3) * f32(np.abs(v0).max()) + np.abs(o).sum(1, dtype=np.float32))
    F = f32(9.5367432e-7) * np.abs(d).sum(1, dtype=np.float32)
    A = np.abs(a)
    B = np.where(np.signbit(a), -b, b).astype(np.float32)
    for width in (1e-6, 1e-3, 0.1):
        # a cell around the exact distance, the exact distance at its lower / upper end or in the middle
        for pos in (0.0, 0.5, 1.0):
            lo = np.maximum(r - f32(width * pos), 0).astype(np.float32)
            hi = (r + f32(width * (1 - pos))).astype(np.float32)
            inside = (denom != 0) & (r >= 0) & (r >= lo) & (r <= hi) & np.isfinite(r)
            c1 = fma(-lo, F, -E)
            c2 = fma(hi, F, E)
            keep = (A <= F) | ((B >= fma(lo, A, c1)) & (B <= fma(hi, A, c2)))
            assert inside.sum() > N // 3
            assert not (inside & ~keep).any(), (width, pos, int((inside & ~keep).sum()))
