"""CPU: the arithmetic of the streamline kernel's far form, restated in NumPy with FP32 rounding after every FMA, against
float64 -- the error bound the per-launch classification (topo.cu: k2x_class, `k2_amax`) relies on.

For a charge (x, q) and a point p, with alpha = sgn(q)/q^2, a = alpha x, b = alpha |x|^2 (formed in double, rounded once):
    t = alpha |p|^2 + b - 2 p.a      (four FP32 FMAs)        = alpha |p - x|^2 up to rounding
    u = |t|^(-3/2),  S += u alpha,  T += u a,  E = p S - T   = q (p - x) / |p - x|^3
The expansion cancels: the rounding of t relative to alpha |p - x|^2 is about 2^-24 (|x| + |p|)^2 / |p - x|^2 per
operation, which is why a charge takes this form only when that amplification is at most `amax` (8) over the whole region
a line can visit (the box inflated by three steps)."""
import numpy as np

F32 = np.float32


def fma32(a, b, c):
    """FP32 fused multiply-add of float32 arrays: the product of two float32 is exact in float64, one rounding at the end."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def classify_far(x, q, box, h, amax=8.0):
    """topo.cu: k2x_box + k2x_class for seeds inside the box (extent = box)."""
    m = F32(3.0) * F32(abs(h)) * F32(1.0001) + F32(1e-6)
    b = box.astype(F32) + m
    pmax = np.sqrt((b * b).sum(dtype=F32), dtype=F32)
    e = np.maximum(np.abs(x) - b, F32(0.0)).astype(F32)
    r2min = (e * e).sum(axis=1, dtype=F32)
    xn = np.sqrt((x * x).sum(axis=1, dtype=F32), dtype=F32) + pmax
    far = (xn * xn <= F32(amax) * r2min) & (xn < F32(1.0e6))
    far &= (q == 0) | (np.abs(q) >= F32(1.0e-12))
    return far


def far_form_field(p, x, q):
    """Sum over charges of the far-form terms at point p (float32 arithmetic as in common.cuh: evalp_far), FP64 totals."""
    al64 = np.sign(q.astype(np.float64)) / q.astype(np.float64) ** 2
    a = (al64[:, None] * x.astype(np.float64)).astype(F32)
    b = (al64 * (x.astype(np.float64) ** 2).sum(axis=1)).astype(F32)
    al = al64.astype(F32)
    c = (F32(-2.0) * p).astype(F32)
    c3 = fma32(p[2:3], p[2:3], fma32(p[1:2], p[1:2], (p[0:1] * p[0:1]).astype(F32)))
    n = len(q)
    t = fma32(np.repeat(c3, n), al, b)
    for k in range(3):
        t = fma32(np.repeat(c[k:k + 1], n), a[:, k], t)
    inv = (1.0 / np.sqrt(np.abs(t).astype(np.float64))).astype(F32)            # MUFU.RSQ, taken as correctly rounded here
    u = ((inv * inv).astype(F32) * inv).astype(F32)
    S = (u.astype(np.float64) * al.astype(np.float64)).sum()
    T = (u.astype(np.float64)[:, None] * a.astype(np.float64)).sum(axis=0)
    return p.astype(np.float64) * S - T, t, al64


def test_far_form_error_is_bounded_by_the_classification():
    rng = np.random.default_rng(12)
    box = np.array([0.5, 0.5, 0.5], F32)
    h = 0.1
    # a protein-like frame: charges from 0.6 to 40 A around the box, including many close to it
    r = np.concatenate([rng.uniform(0.6, 4.0, 4000), rng.uniform(4.0, 40.0, 4000)])
    v = rng.normal(size=(len(r), 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    x = (v * r[:, None]).astype(F32)
    q = rng.uniform(-0.8, 0.8, len(r)).astype(F32)
    q[np.abs(q) < 0.01] = F32(0.3)
    far = classify_far(x, q, box, h)
    assert 0.3 < far.mean() < 0.99                      # both classes are populated
    # every charge beyond a few box sizes is far; everything within one inflated box of the region is near
    assert far[np.linalg.norm(x, axis=1) > 8.0].all() and not far[np.linalg.norm(x, axis=1) < 1.2].any()
    xf, qf = x[far], q[far]
    worst_t, worst_e = 0.0, 0.0
    reach = 0.5 + 3 * h * 1.0001 + 1e-6                 # the region a line can visit
    for _ in range(40):
        p = rng.uniform(-reach, reach, 3).astype(F32)
        e_far, t, al64 = far_form_field(p, xf, qf)
        d = p.astype(np.float64)[None] - xf.astype(np.float64)
        r2 = (d * d).sum(axis=1)
        # (1) the expanded t against alpha |p - x|^2: within amax x a few ulps
        rel_t = np.abs(t.astype(np.float64) - al64 * r2) / np.abs(al64 * r2)
        worst_t = max(worst_t, float(rel_t.max()))
        # (2) the field of the far charges against the direct float64 sum
        e_ref = (qf.astype(np.float64)[:, None] * d / r2[:, None] ** 1.5).sum(axis=0)
        terms = (np.abs(qf.astype(np.float64)) / r2).sum()                     # sum of |terms|
        worst_e = max(worst_e, float(np.abs(e_far - e_ref).max() / terms))
    assert worst_t <= 2 * 8.0 * 2.0 ** -24              # 2 x amax ulps (measured: 6.3 ulps)
    assert worst_e <= 1e-7                              # relative to the sum of |terms| (measured 7e-9): far inside the 1e-5 budget


def test_far_form_would_fail_without_the_classification():
    """The same arithmetic on a charge 0.05 A outside the region loses digits: the reason for the near class."""
    p = np.array([0.79, 0.0, 0.0], F32)
    x = np.array([[0.85, 0.0, 0.0]], F32)
    q = np.array([0.5], F32)
    assert not classify_far(x, q, np.array([0.5, 0.5, 0.5], F32), 0.1)[0]
    _, t, al64 = far_form_field(p, x, q)
    r2 = float(((p.astype(np.float64) - x[0].astype(np.float64)) ** 2).sum())
    rel = abs(float(t[0]) - al64[0] * r2) / abs(al64[0] * r2)
    assert rel > 2 * 8.0 * 2.0 ** -24                   # amplification (|x| + |p|)^2 / r^2 ~ 750
