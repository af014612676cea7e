"""CPU: randomised checks of the float64 oracle -- against the compiled reference (oracle/_ref) on
drawn frames, and through properties the domain offers (superposition in Q, translation and
permutation invariance, the streamline's own bookkeeping, pure-NumPy restatement of the sums)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import f64, ref

K = float(np.float32(14.3996451))        # C:412 `float factor = 14.3996451`


def frame(seed, m, box=1.0, span=20.0):
    """m charges outside the sampling box (like filter_in_box, SC:349-358), float32."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-span, span, (m, 3)).astype(np.float32)
    inside = np.all(np.abs(x) < 1.25 * box, axis=1)
    x[inside, 0] += np.float32(3.0 * box)
    q = rng.uniform(-0.8, 0.8, m).astype(np.float32)
    return x, q


def relmax(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


common = dict(deadline=None, derandomize=True, database=None, max_examples=25, suppress_health_check=[HealthCheck.too_slow])


@settings(**common)
@given(seed=st.integers(0, 10_000), m=st.integers(1, 400), n=st.integers(1, 40))
def test_field_and_esp_equal_a_plain_numpy_sum(seed, m, n):
    x, q = frame(seed, m)
    pts = np.random.default_rng(seed + 1).uniform(-1, 1, (n, 3)).astype(np.float32)
    d = pts[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    r2 = np.sum(d * d, axis=2)
    e_raw = K * np.sum(q[None, :, None].astype(np.float64) * d / r2[:, :, None] ** 1.5, axis=1)
    r2s = np.maximum(r2, float(np.float32(1e-6)))                       # C:433-436
    e_soft = K * np.sum(q[None, :, None].astype(np.float64) * d / r2s[:, :, None] ** 1.5, axis=1)
    phi = K * np.sum(q[None, :].astype(np.float64) / np.sqrt(r2), axis=1)
    # the E-field is a sum of cancelling vectors: scale by the sum of |terms|, not by the result
    mag = K * np.sum(np.abs(q)[None, :] / r2, axis=1).max()
    assert np.max(np.abs(f64.field_grid(pts, x, q, False) - e_raw)) <= 1e-13 * mag
    assert np.max(np.abs(f64.field_grid(pts, x, q, True) - e_soft)) <= 1e-13 * mag
    assert np.max(np.abs(f64.esp_grid(pts, x, q) - phi)) <= 1e-13 * K * np.sum(np.abs(q)[None, :] / np.sqrt(r2), axis=1).max()


@settings(**common)
@given(seed=st.integers(0, 10_000), m=st.integers(2, 300))
def test_superposition_and_permutation(seed, m):
    x, q = frame(seed, m)
    pts = np.random.default_rng(seed + 2).uniform(-1, 1, (7, 3)).astype(np.float32)
    full = f64.field_grid(pts, x, q, False)
    h = m // 2
    parts = f64.field_grid(pts, x[:h], q[:h], False) + f64.field_grid(pts, x[h:], q[h:], False)
    scale = np.max(np.abs(full)) + 1e-300
    assert np.max(np.abs(full - parts)) <= 1e-12 * max(scale, 1.0)
    perm = np.random.default_rng(seed + 3).permutation(m)
    assert np.max(np.abs(f64.esp_grid(pts, x[perm], q[perm]) - f64.esp_grid(pts, x, q))) <= 1e-11 * max(
        1.0, float(np.max(np.abs(f64.esp_grid(pts, x, q)))))
    # doubling every charge doubles field and potential exactly (power of two)
    np.testing.assert_array_equal(f64.field_grid(pts, x, 2 * q, False), 2 * full)


@settings(**common)
@given(seed=st.integers(0, 10_000), m=st.integers(1, 200), h=st.sampled_from([0.1, 0.05, 0.01]),
       n_iter=st.integers(1, 60))
def test_streamline_bookkeeping(seed, m, h, n_iter):
    """dist is the seed-to-end distance of the recorded points, the end point is the first one
    outside the box (kept, C:534-557) or the n_iter-th, every step has length h, and the
    curvature is the mean of the two end curvatures built from the recorded triples (C:565-589)."""
    x, q = frame(seed, m)
    dims = np.array([1.0, 1.0, 1.0], dtype=np.float32)
    s = np.random.default_rng(seed + 4).uniform(-0.9, 0.9, 3).astype(np.float32)
    (dist, curv), k, pts = f64.line(s, n_iter, x, q, h, dims, want_points=True)
    assert 1 <= k <= n_iter
    p0, p0a, p0b, pk, pka, pkb = pts
    np.testing.assert_allclose(p0, s.astype(np.float64), rtol=0, atol=0)
    assert dist == pytest.approx(np.linalg.norm(p0 - pk), rel=1e-12, abs=1e-15)
    assert dist <= k * float(np.float32(h)) * (1 + 1e-9)
    for a, b in ((p0, p0a), (p0a, p0b), (pk, pka), (pka, pkb)):
        assert np.linalg.norm(b - a) == pytest.approx(float(np.float32(h)), rel=1e-6)
    outside = np.any(np.abs(pk) > 1.0)
    assert outside or k == n_iter
    # the first look-ahead point is one step from the seed
    np.testing.assert_allclose(p0a, f64.step(p0.astype(np.float32), h, x, q), rtol=0, atol=1e-6)

    def kappa(a0, a1, a2):
        v1, v2 = a1 - a0, a2 - 2 * a1 + a0
        return np.linalg.norm(np.cross(v1, v2)) / np.linalg.norm(v1) ** 3
    assert curv == pytest.approx(0.5 * (kappa(p0, p0a, p0b) + kappa(pk, pka, pkb)), rel=1e-9, abs=1e-12)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@settings(**dict(common, max_examples=15))
@given(seed=st.integers(0, 10_000), m=st.integers(1, 1500))
def test_oracle_vs_compiled_reference_on_drawn_frames(seed, m):
    x, q = frame(seed, m)
    pts = np.random.default_rng(seed + 5).uniform(-1, 1, (12, 3)).astype(np.float32)
    # FP32 sequential sums of cancelling terms: compare on the scale of the summed magnitudes
    d = pts[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    r2 = np.sum(d * d, axis=2)
    mag_e = K * np.sum(np.abs(q)[None, :] / r2, axis=1).max()
    mag_p = K * np.sum(np.abs(q)[None, :] / np.sqrt(r2), axis=1).max()
    assert np.max(np.abs(ref.compute_looped_field(pts, x, q) - f64.field_grid(pts, x, q, True))) <= 2e-5 * mag_e
    e = np.array([ref.calc_field_base(p, x, q) for p in pts])
    assert np.max(np.abs(e - f64.field_grid(pts, x, q, False))) <= 2e-5 * mag_e
    phi = np.array([ref.calc_esp_base(p, x, q)[0] for p in pts])
    assert np.max(np.abs(phi - f64.esp_grid(pts, x, q))) <= 2e-5 * mag_p
    # streamlines: same number of steps, dist within FP32 noise (curvature: second-difference noise)
    dims = np.array([1.0, 1.0, 1.0], dtype=np.float32)
    seeds = pts[:6] * np.float32(0.9)
    n_it = np.random.default_rng(seed + 6).integers(1, 17, len(seeds))
    a = np.array([ref.thread_operation(s, int(n), x, q, 0.1, dims) for s, n in zip(seeds, n_it)])
    b, steps = f64.topo_batch(seeds, n_it, x, q, 0.1, dims)
    # a line whose end point lies within FP32 noise of a box face may stop one step apart
    close = np.abs(a[:, 0] - b[:, 0]) < 1e-5
    assert close.sum() >= len(seeds) - 1
    assert np.max(np.abs(a[close, 1] - b[close, 1])) < 5e-5 + 2e-7 / 0.01 + 1e-3 * np.max(np.abs(b[close, 1]))
