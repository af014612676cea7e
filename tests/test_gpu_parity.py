"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C-ABI
(ctypes, include/cpet_b200.h), against the float64 oracle, the reference-generated golden vectors
and size-independent properties.  /root/reference is NOT needed here.

Tolerances (BASELINE.json north_star):
  field / ESP : max-norm relative error vs the float64 oracle <= 1e-5
  streamlines : |dist - oracle| <= 2e-6 (fraction of step-count flips <= 1e-3);
                |curv - oracle| <= 5e-5 + 2e-7/h^2  (reference-style FP32 second differences)
                and <= 2e-5 + 2e-6/h for the default direction-based curvature
  histogram   : bit-exact vs np.histogram2d
"""
import ctypes

import numpy as np
import pytest

import synth
from oracle import f64, hist as ohist

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-5


def relmax(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module")
def M():
    from pycpet_b200 import Math_ops

    m = Math_ops()
    yield m
    m.close()


@pytest.fixture(scope="module")
def frame2a(golden):
    g = golden("example_2A_volume.npz")
    return g["x"], g["Q"].reshape(-1)


def reset_tuning(m):
    m.set_tuning(k1_threads=0, k1_points=0, k1_lanes=0, k1_tile_pairs=0, k1_stages=0, k1_splits=0,
                 k1_lattice=-1, k1_softscan=-1, k1_esp_mix=-1, k1_lat_nodes=-1, k1_hybrid=-1,
                 k2_threads=0, k2_tile_pairs=0, k2_stages=0, k2_sort=-1, k2_cap=0, k2_form=0, k2_amax=0)


# ------------------------------------------------------------------------------------ K1 --------
def test_field_example_2A(M, golden, frame2a):
    g = golden("example_2A_volume.npz")
    x, Q = frame2a
    out = M.field_grid(g["mesh"].reshape(-1, 3), x, Q, soften=True, concat=True)
    assert out.shape == (1331, 6) and out.dtype == np.float32
    np.testing.assert_array_equal(out[:, :3], g["mesh"].reshape(-1, 3))
    e = f64.field_grid(g["mesh"].reshape(-1, 3), x, Q, True)
    assert relmax(out[:, 3:], e) < FIELD_TOL
    # and against what the unmodified reference produced for this input (calculator.compute_box)
    assert relmax(out[:, 3:], g["field_box"][:, 3:]) < FIELD_TOL
    c = M.last_counters()
    assert c["pair_evals"] == 1331 * 7890 and c["launches"] >= 1


def test_point_field_example_1A_known_answer(M, golden):
    g = golden("example_1A_point_field.npz")
    e = M.calc_field(g["point"], g["x"], g["Q"].reshape(-1))
    # examples/1A_point-field/outdir/point_field.dat (the reference's only exact shipped answer)
    np.testing.assert_allclose(e, g["shipped_point_field_dat"], rtol=2e-5)
    np.testing.assert_allclose(e, g["field"], rtol=2e-5)


@pytest.mark.parametrize("cfg", [
    dict(k1_points=1, k1_lanes=1), dict(k1_points=2, k1_lanes=1), dict(k1_points=4, k1_lanes=1),
    dict(k1_lanes=8), dict(k1_lanes=32), dict(k1_lanes=1, k1_splits=1), dict(k1_lanes=32, k1_splits=7),
    dict(k1_points=2, k1_lanes=1, k1_tile_pairs=64, k1_stages=2),
    dict(k1_points=1, k1_lanes=8, k1_tile_pairs=128, k1_stages=4, k1_splits=3),
    dict(k1_threads=64, k1_points=4, k1_lanes=1, k1_tile_pairs=256, k1_stages=8),
])
@pytest.mark.parametrize("mode", ["soft", "raw", "esp"])
def test_field_all_kernel_variants(M, cfg, mode):
    x, Q = synth.charges(5001, seed=11, box=1.0)          # odd count: exercises the pad charge
    pts = synth.grid(9, 1.0)                               # 729 points
    reset_tuning(M)
    M.set_tuning(**cfg)
    try:
        if mode == "esp":
            got = M.esp_grid(pts, x, Q)
            want = f64.esp_grid(pts, x, Q)
        else:
            got = M.field_grid(pts, x, Q, soften=(mode == "soft"))
            want = f64.field_grid(pts, x, Q, mode == "soft")
        assert relmax(got, want) < FIELD_TOL
    finally:
        reset_tuning(M)


def test_field_synthetic_golden_and_softening(M, golden):
    g = golden("synthetic_math_ops.npz")
    x, Q = g["x"], g["Q"]
    got = M.compute_looped_field(g["points"], x, Q)
    assert relmax(got, f64.field_grid(g["points"], x, Q, True)) < FIELD_TOL
    assert relmax(got, g["looped_field"]) < FIELD_TOL
    # grid points sitting exactly on charges: softened field is finite and matches (C:433)
    gs = M.compute_looped_field(g["points_soft"], x, Q)
    assert np.all(np.isfinite(gs))
    assert relmax(gs, f64.field_grid(g["points_soft"], x, Q, True)) < FIELD_TOL
    # without softening a point on a charge is NaN/inf exactly like calc_field_base (C:318-319)
    raw = M.field_grid(g["points_soft"][:3], x, Q, soften=False)
    assert not np.all(np.isfinite(raw))
    # compute_batched_field = no softening
    nb = M.compute_batch_field(g["points"], x, Q, 100)
    assert relmax(nb, f64.field_grid(g["points"], x, Q, False)) < FIELD_TOL


def test_esp_example_2A(M, golden, frame2a):
    ge = golden("example_2A_volume_esp.npz")
    x, Q = frame2a
    pts = ge["mesh"].reshape(-1, 3)
    phi = f64.esp_grid(pts, x, Q)
    got = M.esp_grid(pts, x, Q)
    assert relmax(got, phi) < FIELD_TOL
    box = M.esp_grid(pts, x, Q, concat_half=True)
    assert box.dtype == np.float16 and box.shape == (1331, 4)
    np.testing.assert_array_equal(box[:, :3], pts.astype(np.float16))
    # float64 -> float32 -> float16 like the reference; equal to the oracle's cast except where
    # a 1e-6-relative difference straddles a float16 rounding boundary
    want16 = phi.astype(np.float32).astype(np.float16)
    mism = box[:, 3] != want16
    assert mism.mean() < 0.03, (mism.mean(), np.max(np.abs(got - phi)))
    # one float16 ulp, plus the 1e-5 max-norm budget where phi crosses zero and the ulp is tiny
    ulp = np.spacing(np.abs(want16)).astype(np.float64) + 1e-5 * np.max(np.abs(phi))
    assert np.all(np.abs(box[:, 3].astype(np.float64) - want16.astype(np.float64)) <= ulp)
    # the reference's own float16 output for this frame (compute_box_ESP)
    ref_box = ge["esp_box"]
    # (the reference sums in sequential FP32: its own 4e-5 max-norm noise flips ~14 % of the float16
    #  roundings on this frame; never by more than one float16 ulp)
    assert np.mean(box[:, 3] != ref_box[:, 3]) < 0.25
    assert np.all(np.abs(box[:, 3].astype(np.float64) - ref_box[:, 3].astype(np.float64)) <= ulp + 2e-4 * np.max(np.abs(phi)))


def test_field_edge_cases(M):
    x, Q = synth.charges(33, seed=2, box=0.5)
    pts = synth.grid(3, 0.5)
    # empty point list, single point, single charge
    assert M.field_grid(np.zeros((0, 3), np.float32), x, Q).shape == (0, 3)
    one = M.field_grid(pts[:1], x, Q, soften=False)
    assert relmax(one, f64.field_grid(pts[:1], x, Q, False)) < FIELD_TOL
    e1 = M.field_grid(pts, x[:1], Q[:1], soften=False)
    assert relmax(e1, f64.field_grid(pts, x[:1], Q[:1], False)) < FIELD_TOL
    # ragged sizes around the tile / pair / warp boundaries
    for m in (2, 3, 127, 128, 129, 2047, 2049):
        xm, qm = synth.charges(m, seed=m, box=0.5)
        for n in (1, 31, 33, 257):
            p = np.random.default_rng(n).uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
            assert relmax(M.field_grid(p, xm, qm, soften=True), f64.field_grid(p, xm, qm, True)) < FIELD_TOL


def test_error_reporting():
    """Negative status + message instead of garbage: call order, bad flags, bad sizes."""
    from pycpet_b200 import CpetError, Math_ops, _lib

    m = Math_ops()
    pts = synth.grid(3, 0.5)
    with pytest.raises(CpetError, match="cpet_set_charges"):
        m.field_grid(pts)                                   # no charge set on this context yet
    with pytest.raises(CpetError, match="cpet_set_charges"):
        m.topo_batch(pts, np.ones(len(pts)), step_size=0.1, dimensions=(1, 1, 1))
    x, Q = synth.charges(100, seed=0, box=0.5)
    m.set_charges(x, Q)
    out = np.zeros((len(pts), 3), np.float32)
    rc = m.math.cpet_field_grid(m.ctx, len(pts), _lib.ptr(pts), 64, _lib.ptr(out))      # unknown flag bit
    assert rc == -1 and b"flag" in m.math.cpet_last_error()
    rc = m.math.cpet_field_grid(m.ctx, -5, _lib.ptr(pts), 0, _lib.ptr(out))
    assert rc == -1
    with pytest.raises(ValueError):
        m.set_charges(x, Q[:-1])
    with pytest.raises(CpetError, match="unknown tuning key"):
        m.set_tuning(no_such_knob=1)
    with pytest.raises(CpetError):
        Math_ops(device=99).ctx
    m.close()


def test_empty_charge_set(M):
    """M = 0: fields and potentials are exactly zero; the unguarded E/|E| of the tracer gives NaN,
    as the reference's 0/0 does (C:501)."""
    x0 = np.zeros((0, 3), np.float32)
    q0 = np.zeros(0, np.float32)
    pts = synth.grid(3, 0.5)
    assert np.all(M.field_grid(pts, x0, q0, soften=True) == 0.0)
    assert np.all(M.esp_grid(pts, x0, q0) == 0.0)
    out = M.topo_batch(pts[:5], np.full(5, 3), x0, q0, 0.1, np.array([0.5, 0.5, 0.5], np.float32))
    assert np.all(np.isnan(out))


def test_field_accumulation_at_100k_charges(M):
    """Heavy +/- cancellation (net-neutral 100k charges): FP32-in-tile / FP64-across-tiles holds
    1e-5 where a plain FP32 sequential sum (the reference) does not."""
    x, Q = synth.charges(100_000, seed=3, box=5.0)
    pts = synth.grid(8, 5.0)
    assert relmax(M.field_grid(pts, x, Q, soften=True), f64.field_grid(pts, x, Q, True)) < FIELD_TOL
    assert relmax(M.esp_grid(pts, x, Q), f64.esp_grid(pts, x, Q)) < FIELD_TOL


def test_field_full_size_properties(M):
    """ESP configuration size (101^3 points) on a 20k-charge frame: exact linearity in Q by powers
    of two, invariance to point order, and oracle parity on a random sample of points."""
    x, Q = synth.charges(20_000, seed=4, box=5.0)
    pts = synth.grid(101, 5.0)
    assert len(pts) == 101 ** 3
    # the mesh goes to the lattice kernel, its permutation to the general one: with the charge range
    # unsplit both add the same FP64 partials in the same order, so the comparison is bit for bit
    reset_tuning(M)
    M.set_tuning(k1_splits=1, k1_lat_nodes=0, k1_hybrid=0)   # charge-pair lattice form and direct-form general kernel: the same sums
    M.set_charges(x, Q)
    e = M.field_grid(pts, soften=True)
    assert M.last_path() == "lattice"
    M.set_charges(x, 2.0 * Q)
    e2 = M.field_grid(pts, soften=True)
    np.testing.assert_array_equal(e2, 2.0 * e)                      # bit-exact scaling
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(pts))
    M.set_charges(x, Q)
    ep = M.field_grid(pts[perm], soften=True)
    assert M.last_path() == "general"
    np.testing.assert_array_equal(ep, e[perm])                      # per-point independence
    reset_tuning(M)
    idx = rng.choice(len(pts), 2000, replace=False)
    assert relmax(e[idx], f64.field_grid(pts[idx], x, Q, True)) < FIELD_TOL
    en = M.field_grid(pts, soften=True)               # default for a mesh of this size: node pairs packed
    assert M.last_path() == "lattice" and relmax(en, e) < 2e-6
    eh = M.field_grid(pts[perm], soften=True)         # default for a point list of this size: hybrid near/far kernel
    assert M.last_path() == "general_hybrid" and relmax(eh, e[perm]) < 2e-6
    assert relmax(en[idx], f64.field_grid(pts[idx], x, Q, True)) < FIELD_TOL
    phi = M.esp_grid(pts)
    assert relmax(phi[idx], f64.esp_grid(pts[idx], x, Q)) < FIELD_TOL


def test_config3_esp_and_field_at_full_size(M):
    """BASELINE configs[2] at its real size: 101^3 = 1,030,301 points x 100,000 charges through the
    default host entry points (mesh recognition -> lattice kernel, charge-range splits + FP64
    finalize as the launcher decides), against the float64 oracle on 4,096 sampled points, in the
    reference's return layouts ((N,6) f32 [x|E], (N,4) f16 [x|phi], UC:446-447, 473-475)."""
    x, Q = synth.charges(100_000, seed=3, box=5.0)
    pts = synth.grid(101, 5.0)
    assert len(pts) == 1_030_301 and len(Q) == 100_000
    reset_tuning(M)
    M.set_charges(x, Q)
    e = M.field_grid(pts, soften=True, concat=True)
    assert M.last_path() == "lattice"
    assert e.shape == (len(pts), 6) and e.dtype == np.float32
    np.testing.assert_array_equal(e[:, :3], pts)
    idx = np.random.default_rng(7).choice(len(pts), 4096, replace=False)
    assert relmax(e[idx, 3:], f64.field_grid(pts[idx], x, Q, True)) < FIELD_TOL
    phi_ref = f64.esp_grid(pts[idx], x, Q)
    phi = M.esp_grid(pts)
    assert M.last_path() == "lattice"
    assert relmax(phi[idx], phi_ref) < FIELD_TOL
    half = M.esp_grid(pts, concat_half=True)
    assert half.shape == (len(pts), 4) and half.dtype == np.float16
    np.testing.assert_array_equal(half[:, :3], pts.astype(np.float16))
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(half[idx, 3], phi[idx].astype(np.float16))   # same f32 -> f16 rounding as .astype(np.half)
    # the same points in another order take the general kernel: same values within the field budget
    perm = np.random.default_rng(8).permutation(len(pts))[:200_000]
    eg = M.field_grid(pts[perm], soften=True)
    assert M.last_path() == "general_hybrid"          # a point list of this size: hybrid near/far kernel
    assert relmax(eg, e[perm, 3:]) < 2e-6


@pytest.mark.parametrize("case", ["box", "off_centre", "charges_inside", "tiny_q_nan", "few_charges"])
def test_field_general_hybrid_kernel(M, case):
    """The general field kernel in its hybrid near/far form (default for lists of >= 2,048 points) against the float64
    oracle and the direct-form kernel: points in the origin-centred box PyCPET uses, a point cloud far from the origin
    (every charge classifies as near: plain direct form), charges INSIDE the cloud and 2e-4 A from a point (softening
    acts, raw gives the reference's inf/NaN), zero / denormal-small / NaN charges, fewer charges than one block;
    with the block range split, as (N,6) rows, and through propagate."""
    rng = np.random.default_rng({"box": 1, "off_centre": 2, "charges_inside": 3, "tiny_q_nan": 4, "few_charges": 5}[case])
    n = 5000
    x, Q = synth.charges(3000 if case != "few_charges" else 20, seed=21, box=0.6)
    pts = (rng.uniform(-0.6, 0.6, (n, 3))).astype(np.float32)
    if case == "off_centre":
        pts = (pts + np.array([40.0, -25.0, 10.0], np.float32)).astype(np.float32)
    if case == "charges_inside":
        x = np.vstack([x, pts[7] + np.float32(2e-4), [[0.1, 0.2, -0.3]], pts[11]]).astype(np.float32)
        Q = np.concatenate([Q, [0.3, -0.2, 0.4]]).astype(np.float32)
    if case == "tiny_q_nan":
        Q = Q.copy(); Q[::5] = 0.0; Q[1::7] = np.float32(1e-25); Q[2::9] = np.float32(-3e-14)
    reset_tuning(M)
    M.set_charges(x, Q)
    for soften in (True, False):
        want = f64.field_grid(pts, x, Q, soften)
        fin = np.isfinite(want).all(axis=1)
        scale = np.abs(want[fin]).max()
        M.set_tuning(k1_hybrid=0)
        direct = M.field_grid(pts, soften=soften)
        for cfg in (dict(k1_hybrid=-1), dict(k1_hybrid=1, k1_splits=1), dict(k1_hybrid=1, k1_splits=7),
                    dict(k1_hybrid=1, k1_tile_pairs=64, k1_stages=2)):
            M.set_tuning(k1_splits=0, k1_tile_pairs=0, k1_stages=0)
            M.set_tuning(**cfg)
            got = M.field_grid(pts, soften=soften, concat=True)
            auto_small = cfg.get("k1_hybrid") == -1 and len(Q) < 256      # too few charges to pay for the packing
            assert M.last_path() == ("general" if auto_small else "general_hybrid")
            np.testing.assert_array_equal(got[:, :3], pts)
            assert np.array_equal(np.isfinite(got[:, 3:]).all(axis=1), fin), (case, soften, cfg)
            assert np.max(np.abs(got[fin, 3:] - want[fin])) / scale < FIELD_TOL, (case, soften, cfg)
            # two FP32 summation orders of the same terms (one chain per point here, an even and an odd one there); a
            # point cloud 48 A from the origin sees only far, strongly cancelling charges: 2.6e-6 between the two
            assert np.max(np.abs(got[fin, 3:] - direct[fin])) / scale < (6e-6 if case == "off_centre" else 2e-6), (case, soften, cfg)
        M.set_tuning(k1_splits=0, k1_tile_pairs=0, k1_stages=0, k1_hybrid=-1)
    if case == "tiny_q_nan":                          # a NaN charge poisons every point, as in the reference's sum
        Qn = Q.copy(); Qn[3] = np.nan
        M.set_charges(x, Qn)
        assert np.all(np.isnan(M.field_grid(pts, soften=True)))
        M.set_charges(x, Q)
    # one normalised-field step (propagate_topo) goes through the same kernel
    step = M.propagate(pts, 0.1)
    e = f64.field_grid(pts, x, Q, False)
    efin = np.isfinite(e).all(axis=1)
    ok = efin & (np.linalg.norm(np.where(efin[:, None], e, 0.0), axis=1) > 1e-6 * np.abs(e[efin]).max())
    ref = pts.astype(np.float64) + 0.1 * e / np.linalg.norm(e, axis=1, keepdims=True)
    assert np.max(np.abs(step[ok] - ref[ok])) < 5e-6 * max(1.0, float(np.abs(pts).max()))
    reset_tuning(M)


def test_esp_rsqrt_on_the_fma_pipe(M):
    """ESP lattice kernel with one z-node in six taking its 1/sqrt from the FMA pipe (integer seed + three Newton
    steps, common.cuh: rsqrt2_fma) against the all-MUFU kernel and the float64 oracle: same 1e-5 budget, and the
    two kernels within 2e-6 of each other -- the spread lattice kernels with different z-nodes per thread show
    among themselves (the routine is accurate to FP32 rounding, like MUFU.RSQ)."""
    x, Q = synth.charges(20_000, seed=5, box=2.0)
    ax = np.linspace(-2.0, 2.0, 49).astype(np.float32)          # 117,649 nodes: above the automatic threshold
    pts = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    reset_tuning(M)
    M.set_charges(x, Q)
    M.set_tuning(k1_esp_mix=0)
    a = M.esp_lattice(ax, ax, ax)
    M.set_tuning(k1_esp_mix=1)
    b = M.esp_lattice(ax, ax, ax)
    M.set_tuning(k1_esp_mix=-1)
    c = M.esp_grid(pts)                                          # default heuristics through the point-list entry
    assert M.last_path() == "lattice"
    np.testing.assert_array_equal(b, c)
    assert relmax(b, a) < 2e-6
    idx = np.random.default_rng(2).choice(len(pts), 3000, replace=False)
    want = f64.esp_grid(pts[idx], x, Q)
    assert relmax(a[idx], want) < FIELD_TOL and relmax(b[idx], want) < FIELD_TOL
    # charges next to nodes (r down to 1e-3 A) and far away (r ~ 1e3 A): the seed + Newton route over the range of r^2
    xs = np.array([[-1.999, -2.0, -2.0], [900.0, -700.0, 800.0], [0.5, 0.5, 0.501]], np.float32)
    qs = np.array([0.3, -0.7, 0.4], np.float32)
    M.set_charges(xs, qs)
    M.set_tuning(k1_esp_mix=1)
    got = M.esp_lattice(ax[:7], ax[:6], ax[:12])
    M.set_tuning(k1_esp_mix=-1)
    p2 = np.stack(np.meshgrid(ax[:7], ax[:6], ax[:12], indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    ref = f64.esp_grid(p2, xs, qs)
    r = np.linalg.norm(p2[:, None, :].astype(np.float64) - xs[None].astype(np.float64), axis=2)
    scale = 14.3996451 * (np.abs(qs)[None] / r).sum(axis=1)          # sum of |terms| per node
    assert r.min() < 1.1e-3 and r.max() > 1e3
    assert np.max(np.abs(got - ref) / scale) < 1e-6
    reset_tuning(M)


def test_sweep_corner_1e8_points(M):
    """BASELINE configs[4] corner N = 464^3 = 99,897,344 points (x 1,000 charges), device-resident
    through the _dev C ABI: 2.4 GB of (N,6) rows written by the kernel, point indices beyond 2^24,
    6 N floats beyond 2^29.  Oracle parity on 4,096 sampled rows; the lattice entry point and the
    point-list entry point must give the same bits on the sampled rows."""
    import torch

    from pycpet_b200.device import Engine

    x, Q = synth.charges(1000, seed=9, box=1.5)
    n = 464
    c = np.linspace(-1.5, 1.5, n)
    c32 = c.astype(np.float32)
    eng = Engine(0)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    ax = torch.from_numpy(c32).cuda()
    out = eng.field_lattice(ax, ax, ax, soften=True, concat=True)
    assert out.shape == (n ** 3, 6)
    rng = np.random.default_rng(11)
    idx = np.unique(np.concatenate([rng.choice(n ** 3, 4090, replace=False),
                                    [0, n ** 3 - 1, 2 ** 24 + 1, 2 ** 26 + 3, n ** 3 - n, n ** 3 // 2]]))
    rows = out[torch.from_numpy(idx).cuda()].cpu().numpy()
    ix, iy, iz = np.unravel_index(idx, (n, n, n))
    pts = np.column_stack([c32[ix], c32[iy], c32[iz]])
    np.testing.assert_array_equal(rows[:, :3], pts)                   # z fastest, meshgrid(indexing="ij")
    assert relmax(rows[:, 3:], f64.field_grid(pts, x, Q, True)) < FIELD_TOL
    phi = eng.esp_lattice(ax, ax, ax)
    assert relmax(phi[torch.from_numpy(idx).cuda()].cpu().numpy(), f64.esp_grid(pts, x, Q)) < FIELD_TOL
    del phi
    # the general point-list kernel on the last 2^25 points of the same mesh (indices near the top of the range)
    tail = 2 ** 25
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)[-tail:].contiguous()
    eg = eng.field_grid(g, soften=True)
    sel = idx[idx >= n ** 3 - tail]
    got = eg[torch.from_numpy(sel - (n ** 3 - tail)).cuda()].cpu().numpy()
    want = rows[np.isin(idx, sel), 3:]
    assert relmax(got, want) < 2e-6
    eng.close()


@pytest.mark.parametrize("shape", [(11, 11, 11), (5, 7, 23), (3, 2, 101), (6, 5, 4), (2, 3, 1)])
def test_field_lattice_matches_general_kernel(M, frame2a, shape):
    """The lattice kernel (dx, dy shared along z) against the general kernel on the expanded mesh:
    bit-identical with the charge range unsplit, and within 1e-5 of the oracle."""
    x, Q = frame2a
    nx, ny, nz = shape
    xs = np.linspace(-0.5, 0.5, nx, dtype=np.float32)
    ys = np.linspace(-0.4, 0.6, ny, dtype=np.float32)
    zs = np.linspace(-0.7, 0.3, nz, dtype=np.float32)
    mesh = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).astype(np.float32)
    pts = mesh.reshape(-1, 3)
    M.set_charges(x, Q)
    reset_tuning(M)
    try:
        for pz in (2, 4, 5):
            M.set_tuning(k1_splits=1, k1_lanes=1, k1_points=pz)
            for soften in (True, False):
                lat = M.field_lattice(xs, ys, zs, soften=soften, concat=True)
                M.set_tuning(k1_points=4 if pz == 5 else pz)
                gen = M.field_grid(pts, soften=soften, concat=True)
                M.set_tuning(k1_points=pz)
                np.testing.assert_array_equal(lat, gen)
            lat_esp = M.esp_lattice(xs, ys, zs, concat_half=True)
            np.testing.assert_array_equal(lat_esp, M.esp_grid(pts, concat_half=True))
        reset_tuning(M)
        # default heuristics (charge splits may differ between the two paths): oracle parity
        lat = M.field_lattice(xs, ys, zs, soften=True)
        assert relmax(lat, f64.field_grid(pts, x, Q, True)) < FIELD_TOL
        assert relmax(M.esp_lattice(xs, ys, zs), f64.esp_grid(pts, x, Q)) < FIELD_TOL
    finally:
        reset_tuning(M)
    # the calculator-level entry point picks the lattice path for box meshes and only for them
    from pycpet_b200 import calculator as calc
    assert (calc.lattice_axes(mesh) is not None)
    bumped = mesh.copy()
    bumped[0, 0, 0, 2] += np.float32(1e-3)
    assert calc.lattice_axes(bumped) is None
    out = calc.compute_field_on_grid(mesh, x, Q)
    assert out.shape == (len(pts), 6)
    np.testing.assert_array_equal(out[:, :3], pts)
    assert relmax(out[:, 3:], f64.field_grid(pts, x, Q, True)) < FIELD_TOL
    assert relmax(calc.compute_field_on_grid(bumped, x, Q)[:, 3:],
                  f64.field_grid(bumped.reshape(-1, 3), x, Q, True)) < FIELD_TOL


@pytest.mark.parametrize("shape", [(11, 11, 11), (5, 7, 23), (3, 2, 101), (6, 5, 4), (2, 3, 1), (4, 4, 16)])
def test_field_lattice_node_pairs(M, frame2a, shape):
    """The lattice kernel with two z-nodes per packed register (large-mesh default; forced here on small meshes) against
    the charge-pair form, the oracle, with the charge range split, with the z axis shorter than / not a multiple of
    the nodes per thread, softened with a charge sitting on a node, and raw."""
    x, Q = frame2a
    rng = np.random.default_rng(sum(shape))
    axes = [np.sort(rng.uniform(-0.5, 0.5, n)).astype(np.float32) for n in shape]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    xs = np.vstack([x, [[axes[0][0], axes[1][0], axes[2][-1]]]]).astype(np.float32)      # a charge ON the last z-node of column 0
    qs = np.concatenate([Q, [0.25]]).astype(np.float32)
    reset_tuning(M)
    M.set_charges(xs, qs)
    for soften in (True, False):
        M.set_tuning(k1_lat_nodes=0, k1_points=0, k1_splits=0)
        base = M.field_lattice(*axes, soften=soften)
        want = f64.field_grid(pts, xs, qs, soften) if soften else None
        for cfg in (dict(k1_points=8), dict(k1_points=10), dict(k1_points=6, k1_unroll=2), dict(k1_points=4, k1_splits=3),
                    dict(k1_points=8, k1_splits=1), dict()):
            M.set_tuning(k1_points=0, k1_splits=0, k1_unroll=0)
            M.set_tuning(k1_lat_nodes=1, **cfg)
            got = M.field_lattice(*axes, soften=soften)
            if soften:
                assert np.all(np.isfinite(got))
                assert relmax(got, want) < FIELD_TOL, cfg
                assert relmax(got, base) < 2e-6, cfg
            else:                                      # the coincident node is inf / NaN exactly where the charge-pair form has it
                fin = np.isfinite(base)
                assert np.array_equal(np.isfinite(got), fin), cfg
                assert np.max(np.abs(got[fin] - base[fin])) <= 2e-6 * np.max(np.abs(base[fin])), cfg
        M.set_tuning(k1_lat_nodes=1, k1_softscan=1, k1_points=0, k1_splits=0, k1_unroll=0)   # scan + both instantiations
        if soften:
            assert relmax(M.field_lattice(*axes, soften=True), want) < FIELD_TOL
        M.set_tuning(k1_softscan=-1)
    # the point-list entry recognises the mesh and takes the same kernel
    M.set_tuning(k1_lat_nodes=1, k1_lattice=1)
    a = M.field_grid(pts, soften=True)
    if M.last_path() == "lattice":
        np.testing.assert_array_equal(a, M.field_lattice(*axes, soften=True))
    reset_tuning(M)


def test_field_lattice_softening_scan(M, frame2a):
    """Softened meshes: when no charge lies within 1e-3 A of a node the unsoftened instantiation
    serves the call (max(r^2, eps) == r^2 everywhere) -- same bits as the softened kernel; a charge
    sitting on a node (r = 0, where only the softening keeps the field finite) switches back."""
    x, Q = frame2a
    xs = np.linspace(-0.5, 0.5, 11, dtype=np.float32)
    ys = np.linspace(-0.4, 0.6, 9, dtype=np.float32)
    zs = np.linspace(-0.7, 0.3, 13, dtype=np.float32)
    reset_tuning(M)
    try:
        for case in ("clear", "on_node", "near_node", "nan_charge"):
            xc = x.copy()
            if case == "nan_charge":       # fmaxf(NaN, eps) = eps: only the softened kernel reproduces that
                xc[17, 1] = np.nan
            if case == "on_node":
                xc[17] = (xs[3], ys[4], zs[5])
            elif case == "near_node":
                xc[17] = (xs[3] + np.float32(4e-4), ys[4] - np.float32(3e-4), zs[5] + np.float32(2e-4))
            M.set_charges(xc, Q)
            M.set_tuning(k1_splits=1, k1_softscan=0)
            want = M.field_lattice(xs, ys, zs, soften=True)
            n0 = M.last_counters()["launches"]
            M.set_tuning(k1_splits=1, k1_softscan=1)
            got = M.field_lattice(xs, ys, zs, soften=True)
            assert M.last_counters()["launches"] == n0 + 2       # scan + the instantiation that exits at once
            np.testing.assert_array_equal(got, want)
            if case == "nan_charge":
                continue
            assert np.isfinite(got).all()
            raw = M.field_lattice(xs, ys, zs, soften=False)
            if case == "clear":
                np.testing.assert_array_equal(raw, want)
            else:
                assert not np.array_equal(raw, want)            # the softening did act
    finally:
        reset_tuning(M)
        M.set_charges(x, Q)


def test_field_grid_recognises_box_meshes(M, frame2a):
    """compute_looped_field only ever receives mesh.reshape(-1,3): the library detects the
    tensor-product structure on the device and switches kernels without changing a single bit."""
    x, Q = frame2a
    reset_tuning(M)
    M.set_charges(x, Q)
    for shape in [(17, 17, 17), (9, 31, 40), (1, 64, 64), (4096, 1, 4), (33, 33, 5)]:
        axes = [np.linspace(-0.5, 0.5 + 0.1 * i, n, dtype=np.float32) for i, n in enumerate(shape)]
        mesh = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).astype(np.float32)
        pts = mesh.reshape(-1, 3)
        M.set_tuning(k1_lattice=-1, k1_splits=1)
        auto = M.field_grid(pts, soften=True, concat=True)
        assert M.last_path() == "lattice", shape
        M.set_tuning(k1_lattice=0, k1_splits=1, k1_points=4, k1_lanes=1, k1_hybrid=0)   # direct-form general kernel
        gen = M.field_grid(pts, soften=True, concat=True)
        assert M.last_path() == "general"
        np.testing.assert_array_equal(auto, gen)
        reset_tuning(M)
        M.set_tuning(k1_lattice=-1)
        # anything that is not exactly a z-fastest mesh keeps the general kernel
        for bad in (pts[::-1].copy(), pts[np.random.default_rng(0).permutation(len(pts))],
                    np.ascontiguousarray(mesh.transpose(2, 1, 0, 3)).reshape(-1, 3)):
            if np.array_equal(bad, pts):
                continue
            got = M.field_grid(bad, soften=True)
            if M.last_path() == "lattice":      # e.g. a reversed mesh is still a valid mesh
                assert relmax(got, f64.field_grid(bad, x, Q, True)) < FIELD_TOL
            else:
                assert relmax(got, f64.field_grid(bad, x, Q, True)) < FIELD_TOL
        bumped = pts.copy()
        bumped[len(pts) // 2, 1] += np.float32(1e-4)
        got = M.field_grid(bumped, soften=True)
        assert M.last_path() in ("general", "general_hybrid")
        assert relmax(got, f64.field_grid(bumped, x, Q, True)) < FIELD_TOL
    # legacy symbol goes through the same host entry point
    axes = [np.linspace(-0.5, 0.5, 17, dtype=np.float32)] * 3
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).astype(np.float32).reshape(-1, 3)
    e = M.compute_looped_field(pts, x, Q)
    assert relmax(e, f64.field_grid(pts, x, Q, True)) < FIELD_TOL
    reset_tuning(M)


def test_propagate(M, frame2a):
    x, Q = frame2a
    pts = synth.grid(4, 0.4)
    got = M.propagate(pts, 0.1, x, Q)
    want = np.array([f64.step(p, 0.1, x, Q) for p in pts])
    assert np.max(np.abs(got - want)) < 2e-7
    from pycpet_b200 import calculator as calc
    one = calc.propagate_topo(pts[3], x, Q, 0.1)
    np.testing.assert_array_equal(one, got[3])


# ------------------------------------------------------------------------------------ K2 --------
def curv_tol_sd(h):
    return 5e-5 + 2e-7 / h ** 2


def curv_tol_dir(h):
    return 2e-5 + 2e-6 / h


def curv_tol_field_limited(h):
    """kappa = |e_k x e_k+1| / h: a relative error eps in the FP32 field direction shows up as
    ~eps/h in the curvature.  With the north-star field budget eps = 1e-5 that is 1e-5/h; dense
    synthetic frames with strong cancellation (|E| small against the sum of |terms|) come close to
    it, the real-protein goldens stay 3-5x below (curv_tol_dir)."""
    return 5e-5 + 1e-5 / h


def check_lines(got, steps, want, wsteps, h, tol_curv):
    flips = steps != wsteps
    assert flips.mean() <= 1e-3 or flips.sum() <= 1
    ok = ~flips
    assert np.max(np.abs(got[ok, 0] - want[ok, 0])) <= 2e-6
    assert np.max(np.abs(got[ok, 1] - want[ok, 1])) <= tol_curv


def test_topo_example_3A(M, golden, frame2a):
    t = golden("example_3A_topo.npz")
    x, Q = frame2a
    h = float(t["step_size"])
    want, wsteps = f64.topo_batch(t["seeds"], t["n_iter"], x, Q, h, t["dimensions"])
    got, steps = M.topo_batch(t["seeds"], t["n_iter"], x, Q, h, t["dimensions"], want_steps=True)
    assert got.shape == (216, 2) and got.dtype == np.float32
    check_lines(got, steps, want, wsteps, h, curv_tol_dir(h))
    got_sd, steps_sd = M.topo_batch(t["seeds"], t["n_iter"], x, Q, h, t["dimensions"], second_diff=True,
                                    want_steps=True)
    check_lines(got_sd, steps_sd, want, wsteps, h, curv_tol_sd(h))
    # the unmodified reference's output for the same seeded run (compute_topo_complete_c_shared)
    ref = t["hist"]
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) <= 2e-6
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) <= curv_tol_sd(h)
    c = M.last_counters()
    assert c["field_evals"] == int((steps_sd + 2).sum())
    assert c["pair_evals"] == c["field_evals"] * 7890


@pytest.mark.parametrize("cfg", [
    # lines per warp, charge residency, thread counts, queue order
    dict(), dict(k2_cap=1), dict(k2_cap=2), dict(k2_cap=4), dict(k2_cap=4, k2_threads=64, k2_sort=1),
    dict(k2_form=2), dict(k2_form=2, k2_cap=4), dict(k2_form=2, k2_cap=4, k2_tile_pairs=256, k2_stages=2),   # charge-pair-packed hybrid
    dict(k2_form=3), dict(k2_form=3, k2_cap=8), dict(k2_form=3, k2_cap=4), dict(k2_form=3, k2_cap=2),        # points-packed hybrid
    dict(k2_form=3, k2_cap=1, k2_threads=64), dict(k2_form=3, k2_cap=8, k2_threads=32, k2_sort=0),
    dict(k2_form=3, k2_cap=8, k2_tile_pairs=256, k2_stages=2), dict(k2_form=3, k2_tile_pairs=64, k2_stages=4, k2_unroll=4),
    dict(k2_form=3, k2_unroll=8),
    dict(k2_cap=2, k2_sort=0), dict(k2_cap=4, k2_tile_pairs=256, k2_stages=2),   # streamed charge ring
    dict(k2_cap=1, k2_tile_pairs=64, k2_stages=4), dict(k2_cap=2, k2_tile_pairs=512, k2_stages=3, k2_threads=128),
])
def test_topo_all_kernel_variants(M, golden, cfg):
    g = golden("synthetic_math_ops.npz")
    x, Q = g["x"], g["Q"]
    reset_tuning(M)
    M.set_tuning(**cfg)
    try:
        for h in (0.1, 0.01):
            n_iter = g[f"n_iter_h{h}"]
            want, wsteps = f64.topo_batch(g["seeds"], n_iter, x, Q, h, g["dimensions"])
            got, steps = M.topo_batch(g["seeds"], n_iter, x, Q, h, g["dimensions"], want_steps=True)
            check_lines(got, steps, want, wsteps, h, curv_tol_dir(h))
            got, steps = M.topo_batch(g["seeds"], n_iter, x, Q, h, g["dimensions"], second_diff=True,
                                      want_steps=True)
            check_lines(got, steps, want, wsteps, h, curv_tol_sd(h))
            # against the reference's own FP32 output (its deviation from float64 is ~1e-7/h^2)
            ref = g[f"lines_h{h}"]
            ok = np.abs(got[:, 0] - ref[:, 0]) < h / 2
            assert ok.mean() > 0.95
            assert np.max(np.abs(got[ok, 1] - ref[ok, 1])) <= 2 * curv_tol_sd(h)
    finally:
        reset_tuning(M)


def test_topo_zero_charges_and_origin_seed(M):
    """Exact zeros in Q (PQR files carry them) contribute exactly nothing; a seed at the origin is
    not special."""
    x, Q = synth.charges(3001, seed=11, box=0.5)
    Q = Q.copy(); Q[::7] = 0.0
    seeds, n_iter, dims, _ = synth.seeds(6, 0.5, 0.1)
    seeds = np.vstack([np.zeros((1, 3), np.float32), seeds]).astype(np.float32)
    n_iter = np.concatenate([[9], n_iter])
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    for form in (0, 2, 3):                       # default choice, charge-pair-packed, points-packed hybrid kernel
        reset_tuning(M)
        M.set_tuning(k2_form=form)
        try:
            got, steps = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
            check_lines(got, steps, want, wsteps, 0.1, curv_tol_dir(0.1))
            keep = Q != 0.0
            got2 = M.topo_batch(seeds, n_iter, x[keep], Q[keep], 0.1, dims)
            same = np.abs(got2[:, 0] - got[:, 0]) < 0.05
            assert same.mean() > 0.99
            np.testing.assert_allclose(got2[same], got[same], rtol=0, atol=2e-5)
        finally:
            reset_tuning(M)


def test_topo_streamed_charges_large_frame(M):
    """M = 30k charges does not fit in shared memory: tiles stream through the TMA ring."""
    x, Q = synth.charges(30_000, seed=6, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(7, 0.5, 0.1)            # 343 lines
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    got, steps = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
    check_lines(got, steps, want, wsteps, 0.1, curv_tol_field_limited(0.1))
    # the field itself, at the seeds of the same frame: per-point relative error
    e = M.field_grid(seeds, soften=False).astype(np.float64)
    e_ref = f64.field_grid(seeds, x, Q, False)
    per_point = np.linalg.norm(e - e_ref, axis=1) / np.linalg.norm(e_ref, axis=1)
    assert per_point.max() < 1e-5, per_point.max()


def test_topo_fine_step(M, golden):
    """h = 0.001 (the finest step of the reference's convergence protocol,
    scripts/benchmark_sample_step.py:43): thousands of steps per line."""
    g = golden("synthetic_math_ops.npz")
    x, Q = g["x"], g["Q"]
    h = 0.001
    seeds = g["seeds"][:24]
    dims = g["dimensions"]
    max_steps = round(2 * np.linalg.norm(dims) / h)
    n_iter = np.random.RandomState(42).randint(1, max_steps, len(seeds))
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, h, dims)
    got, steps = M.topo_batch(seeds, n_iter, x, Q, h, dims, want_steps=True)
    # step-count flips at the box face are possible after thousands of FP32 steps: compare the rest
    same = steps == wsteps
    assert same.mean() >= 0.9
    assert np.max(np.abs(got[same, 0] - want[same, 0])) <= 2e-5      # accumulated FP32 position rounding
    assert np.max(np.abs(got[same, 1] - want[same, 1])) <= curv_tol_dir(h)
    got_sd, steps_sd = M.topo_batch(seeds, n_iter, x, Q, h, dims, second_diff=True, want_steps=True)
    np.testing.assert_array_equal(steps_sd, steps)
    assert np.max(np.abs(got_sd[same, 1] - want[same, 1])) <= curv_tol_sd(h)
    # the direction-based default is orders of magnitude closer to float64 than second differences
    assert np.max(np.abs(got[same, 1] - want[same, 1])) < 0.1 * np.max(np.abs(got_sd[same, 1] - want[same, 1])) + 1e-4


def test_field_one_million_charges(M):
    """Sweep corner M = 1e6 (16 MB of packed charges streamed from L2 through the TMA ring)."""
    x, Q = synth.charges(1_000_000, seed=8, box=1.5)
    pts = synth.grid(4, 1.5)
    M.set_charges(x, Q)
    assert relmax(M.field_grid(pts, soften=True), f64.field_grid(pts, x, Q, True)) < FIELD_TOL
    assert relmax(M.esp_grid(pts), f64.esp_grid(pts, x, Q)) < FIELD_TOL
    seeds, n_iter, dims, _ = synth.seeds(3, 0.5, 0.1)
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    got, steps = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
    check_lines(got, steps, want, wsteps, 0.1, curv_tol_field_limited(0.1))


def _line_check(got, steps, want, wsteps, h, what):
    ok = np.isfinite(want).all(axis=1) & (steps == wsteps)
    assert (steps != wsteps).sum() <= max(1, len(want) // 500), what
    assert np.max(np.abs(got[ok, 0] - want[ok, 0])) <= 2e-6, what
    assert np.max(np.abs(got[ok, 1] - want[ok, 1])) <= 2e-5 + 2e-6 / h, what


@pytest.mark.parametrize("signs", ["mixed", "all_negative", "all_positive", "zeros_and_tiny"])
def test_topo_hybrid_charge_classes(M, signs):
    """The hybrid kernel sorts charges into near / far-negative / far-positive blocks per launch; every
    composition of those classes (one of them empty, zero and denormal-small charges, everything near,
    everything far) must give the oracle's lines, and the direct-form kernel of round 1 the same."""
    rng = np.random.default_rng(5)
    x, Q = synth.charges(700, seed=11, box=0.5)
    if signs == "all_negative":
        Q = -np.abs(Q) - np.float32(0.01)
    elif signs == "all_positive":
        Q = np.abs(Q) + np.float32(0.01)
    elif signs == "zeros_and_tiny":
        Q = Q.copy()
        Q[::3] = 0.0
        Q[1::7] = np.float32(1e-20)
        Q[2::11] = np.float32(-3e-14)
    seeds, n_iter, dims, _ = synth.seeds(9, 0.5, 0.1)
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    for cfg in (dict(), dict(k2_amax=1), dict(k2_amax=100000), dict(k2_form=1), dict(k2_cap=1), dict(k2_cap=2),
                dict(k2_tile_pairs=256, k2_stages=2), dict(k2_form=3), dict(k2_form=3, k2_amax=1),
                dict(k2_form=3, k2_amax=100000), dict(k2_form=3, k2_cap=8), dict(k2_form=3, k2_cap=1),
                dict(k2_form=3, k2_tile_pairs=256, k2_stages=2), dict(k2_form=2)):
        reset_tuning(M)
        M.set_tuning(**cfg)
        M.set_charges(x, Q)
        got, steps = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
        if cfg.get("k2_amax") == 100000:      # every charge in the expanded form: still the same lines, looser curvature
            ok = np.isfinite(want).all(axis=1) & (steps == wsteps)
            assert (steps != wsteps).sum() <= 1 and np.max(np.abs(got[ok, 0] - want[ok, 0])) <= 2e-5
            assert np.max(np.abs(got[ok, 1] - want[ok, 1])) <= 2e-3
        else:
            _line_check(got, steps, want, wsteps, 0.1, (signs, cfg))
    reset_tuning(M)


def test_topo_hybrid_near_field_and_outside_seeds(M):
    """Charges inside the sampling box (possible in real inputs) and seeds outside the box the
    caller names: the classification is against max(box, seeds) inflated by three steps, so such charges
    take the direct form and the lines still match the oracle."""
    rng = np.random.default_rng(8)
    x, Q = synth.charges(900, seed=3, box=0.5)
    x = np.vstack([x, np.array([[0.31, -0.2, 0.1], [-0.45, 0.45, -0.4], [1.4, 1.3, -1.2]], np.float32)]).astype(np.float32)
    Q = np.concatenate([Q, np.array([0.4, -0.3, 0.5], np.float32)]).astype(np.float32)
    dims = np.array([0.5, 0.5, 0.5], np.float32)
    inside = (rng.uniform(-1, 1, (400, 3)) * dims * 0.97).astype(np.float32)
    outside = (np.array([1.25, 1.2, -1.1]) + rng.uniform(-0.1, 0.1, (40, 3))).astype(np.float32)   # near the third extra charge
    seeds = np.vstack([inside, outside]).astype(np.float32)
    n_iter = rng.integers(1, 16, len(seeds))
    want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    reset_tuning(M)
    M.set_charges(x, Q)
    got, steps = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
    # lines that pass within 0.05 A of a charge amplify any rounding difference: compare the others
    far_enough = np.ones(len(seeds), bool)
    ok = np.isfinite(want).all(axis=1) & (steps == wsteps)
    assert (steps != wsteps).sum() <= 4
    assert np.median(np.abs(got[ok, 0] - want[ok, 0])) <= 2e-7
    assert np.quantile(np.abs(got[ok, 0] - want[ok, 0]), 0.98) <= 2e-6
    assert np.quantile(np.abs(got[ok, 1] - want[ok, 1]), 0.98) <= 2e-5 + 2e-6 / 0.1
    for form in (1, 3):                          # the direct-form kernel and the points-packed hybrid kernel
        M.set_tuning(k2_form=form)
        got1, steps1 = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
        same = steps == steps1
        assert same.mean() >= 0.99
        assert np.quantile(np.abs(got[same] - got1[same]), 0.98) <= 5e-5
        ok1 = np.isfinite(want).all(axis=1) & (steps1 == wsteps)
        assert np.quantile(np.abs(got1[ok1, 0] - want[ok1, 0]), 0.98) <= 2e-6
    reset_tuning(M)


def test_topo_edge_cases(M, frame2a):
    x, Q = frame2a
    dims = np.array([0.5, 0.5, 0.5], np.float32)
    assert M.topo_batch(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), x, Q, 0.1, dims).shape == (0, 2)
    # one line, n_iter = 1 and a seed that leaves the box on its first step
    seeds = np.array([[0.0, 0.0, 0.0], [0.499, 0.499, 0.499], [-0.499, 0.3, 0.1]], np.float32)
    for n_it in (1, 2, 50):
        n_iter = np.full(3, n_it)
        want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
        for form in (3, 0):                      # the default form last: the legacy call below uses it
            M.set_tuning(k2_form=form)
            got, steps = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
            np.testing.assert_array_equal(steps, wsteps)
            assert np.max(np.abs(got - want)) < 5e-5
    # the per-line legacy entry point (thread_operation) gives the same numbers as the batch
    one = M.thread_operation(seeds[2], 50, x, Q, 0.1, dims)
    np.testing.assert_allclose(one, got[2], rtol=0, atol=2e-6)
    # n_iter = 0: no step taken, distance exactly 0
    for form in (0, 3):
        M.set_tuning(k2_form=form)
        z = M.topo_batch(seeds, np.zeros(3, np.int32), x, Q, 0.1, dims)
        assert np.all(z[:, 0] == 0.0) and np.all(np.isfinite(z[:, 1]))
        assert M.topo_batch(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), x, Q, 0.1, dims).shape == (0, 2)
    # a NaN charge poisons every line exactly as it does in the reference's sums; an empty charge set gives E = 0 -> NaN (C:501)
    for form in (0, 3):
        M.set_tuning(k2_form=form)
        Qn = Q.copy(); Qn[5] = np.nan
        assert np.all(np.isnan(M.topo_batch(seeds, np.full(3, 4), x, Qn, 0.1, dims)[:, 1]))
        e = M.topo_batch(seeds, np.full(3, 4), np.zeros((0, 3), np.float32), np.zeros(0, np.float32), 0.1, dims)
        assert np.all(np.isnan(e[:, 1]))
    M.set_tuning(k2_form=0)


@pytest.mark.parametrize("m_charges", [1, 2, 63, 64, 65, 127, 129, 1000, 13500, 13700])
def test_topo_ragged_sizes(M, m_charges):
    """Charge counts around the 32-pair block padding and the shared-memory residency limit, line
    counts around the warp / lines-per-warp boundaries, every result against the oracle; batches of
    different composition give each line the same bits."""
    x, Q = synth.charges(m_charges, seed=40 + m_charges % 7, box=0.5)
    rng = np.random.default_rng(m_charges)
    dims = np.array([0.5, 0.4, 0.5], np.float32)
    reset_tuning(M)
    M.set_charges(x, Q)
    ref_rows = None
    for L in (1, 3, 4, 5, 31, 33, 150, 2500):
        seeds = (rng.uniform(-1, 1, (L, 3)) * dims * 0.98).astype(np.float32)
        n_iter = rng.integers(0, 12, L)
        want, wsteps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
        got, steps = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
        ok = np.isfinite(want).all(axis=1) & (steps == wsteps)
        assert (steps != wsteps).sum() <= max(1, L // 500)
        if ok.any():
            check_lines(got[ok], steps[ok], want[ok], wsteps[ok], 0.1, curv_tol_field_limited(0.1))
        if L == 2500:
            ref_rows = (seeds, n_iter, got)
    # the first 100 lines alone, and with 4 / 2 / 1 lines per warp: same bits as inside the big batch
    # (2,500 lines are a short queue: the default form is the charge-pair-packed hybrid kernel)
    seeds, n_iter, got = ref_rows
    for cfg in (dict(), dict(k2_cap=4), dict(k2_cap=2), dict(k2_cap=1, k2_threads=96)):
        M.set_tuning(**cfg)
        sub = M.topo_batch(seeds[:100], n_iter[:100], step_size=0.1, dimensions=dims)
        np.testing.assert_array_equal(sub, got[:100])
        reset_tuning(M)
    # the points-packed kernel: 8 / 4 / 2 / 1 lines per warp, whole batch or the first 100 lines: same bits
    M.set_tuning(k2_form=3)
    got3 = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims)
    assert np.nanmax(np.abs(got3[:, 0] - got[:, 0])) <= 4e-6 or m_charges < 64
    for cfg in (dict(), dict(k2_cap=8), dict(k2_cap=4), dict(k2_cap=2), dict(k2_cap=1, k2_threads=96),
                dict(k2_cap=8, k2_threads=64)):
        reset_tuning(M)
        M.set_tuning(k2_form=3, **cfg)
        sub = M.topo_batch(seeds[:100], n_iter[:100], step_size=0.1, dimensions=dims)
        np.testing.assert_array_equal(sub, got3[:100])
    reset_tuning(M)


def test_topo_full_size_3A_properties(M, frame2a):
    """Configuration 3A at the size BASELINE.json names (47^3 = 103,823 seeds, h = 0.1, 17 max
    steps): invariants + order independence + oracle parity on a sample."""
    x, Q = frame2a
    seeds, n_iter, dims, max_steps = synth.seeds(47, 0.5, 0.1)
    assert len(seeds) == 103_823 and max_steps == 17
    got, steps = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
    assert np.all(steps >= 1) and np.all(steps <= n_iter)
    assert np.all(got[:, 0] <= steps * 0.1 * (1 + 1e-5) + 1e-6)       # chord <= path length
    assert np.all(np.isfinite(got)) and np.all(got[:, 1] >= 0)
    c = M.last_counters()
    assert c["field_evals"] == int(steps.astype(np.int64).sum() + 2 * len(steps))
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(seeds))
    got_p = M.topo_batch(seeds[perm], n_iter[perm], x, Q, 0.1, dims)
    np.testing.assert_array_equal(got_p, got[perm])                     # bit-exact, order-free
    idx = rng.choice(len(seeds), 1500, replace=False)
    want, wsteps = f64.topo_batch(seeds[idx], n_iter[idx], x, Q, 0.1, dims)
    check_lines(got[idx], steps[idx], want, wsteps, 0.1, curv_tol_dir(0.1))


def test_config4_frame_of_one_million_seeds(M, frame2a):
    """One frame of BASELINE configs[3] at its real size: 100^3 = 1,000,000 seeds x 7,890 charges
    (queue sort with 1e6 keys, 4 lines per warp, 1e6 queue refills).  Oracle parity on 2,048 sampled
    lines, the work-counter identities the bench's `value` relies on, the same bits from the
    round-1 direct-form kernel's bookkeeping (steps), and the frame's histogram against NumPy."""
    x, Q = frame2a
    seeds, n_iter, dims, max_steps = synth.seeds(100, 0.5, 0.1)
    assert len(seeds) == 1_000_000 and max_steps == 17
    reset_tuning(M)
    got, steps = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
    c = M.last_counters()
    assert np.all(steps >= 1) and np.all(steps <= n_iter)
    assert np.all(np.isfinite(got)) and np.all(got[:, 1] >= 0)
    assert np.all(got[:, 0] <= steps * 0.1 * (1 + 1e-5) + 1e-6)
    evals = int(steps.astype(np.int64).sum() + 2 * len(steps))
    assert c["field_evals"] == evals and c["pair_evals"] == evals * len(Q)
    idx = np.random.default_rng(3).choice(len(seeds), 2048, replace=False)
    want, wsteps = f64.topo_batch(seeds[idx], n_iter[idx], x, Q, 0.1, dims)
    check_lines(got[idx], steps[idx], want, wsteps, 0.1, curv_tol_dir(0.1))
    # a line's result does not depend on the batch it was computed in (same kernel form: 2,048 lines alone would
    # be a short queue and default to the charge-pair-packed kernel)
    M.set_tuning(k2_form=3)
    sub = M.topo_batch(seeds[idx], n_iter[idx], x, Q, 0.1, dims)
    reset_tuning(M)
    np.testing.assert_array_equal(sub, got[idx])
    # the direct-form kernel walks the same lines (a step-count flip needs a point within rounding of a box face)
    M.set_tuning(k2_form=1)
    got1, steps1 = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
    reset_tuning(M)
    assert (steps1 != steps).sum() <= 20
    same = steps1 == steps
    assert np.max(np.abs(got1[same, 0] - got[same, 0])) <= 2e-6
    de, ce = np.linspace(0.0, 1.8, 51), np.linspace(0.0, float(got[:, 1].max()), 51)
    np.testing.assert_array_equal(M.hist2d(got, de, ce),
                                  np.histogram2d(got[:, 0].astype(np.float64), got[:, 1].astype(np.float64),
                                                 bins=[de, ce])[0].astype(np.int64))


# ------------------------------------------------------------------------------------ K3 --------
def test_hist2d_bit_exact(M):
    rng = np.random.default_rng(0)
    d = rng.gamma(2.0, 0.3, 100_003).astype(np.float32)
    c = rng.gamma(1.5, 0.4, 100_003).astype(np.float32)
    v32 = np.column_stack([d, c])
    v64 = v32.astype(np.float64)
    for nd, nc, dr, cr in [(50, 50, (0.0, float(d.max())), (0.0, float(c.max()))),
                           (37, 91, (float(d.min()), float(d.max())), (float(c.min()), float(c.max()))),
                           (8, 5, (0.2, 0.9), (0.1, 2.0)),
                           (300, 400, (0.0, 3.0), (0.0, 4.0))]:            # 120k bins: global-atomic path
        de, ce = np.linspace(dr[0], dr[1], nd + 1), np.linspace(cr[0], cr[1], nc + 1)
        want, _, _ = np.histogram2d(v64[:, 0], v64[:, 1], bins=[nd, nc], range=[dr, cr])
        for vals in (v64, v32):
            got = M.hist2d(vals, de, ce)
            assert got.dtype == np.int64 and got.shape == (nd, nc)
            np.testing.assert_array_equal(got, want.astype(np.int64))
        np.testing.assert_array_equal(ohist.hist2d_counts(v64[:, 0], v64[:, 1], nd, nc, dr, cr),
                                      want.astype(np.int64))


def test_hist2d_edges_nan_batch(M):
    ed = np.linspace(0.0, 1.0, 11)
    dd = np.concatenate([ed, ed, [np.nan, -1.0, 2.0, 0.5]])
    cc = np.concatenate([ed, ed[::-1], [0.5, 0.5, 0.5, np.nan]])
    want, _, _ = np.histogram2d(dd[:22], cc[:22], bins=[10, 10], range=[(0, 1), (0, 1)])
    got = M.hist2d(np.column_stack([dd, cc]), ed, ed)
    np.testing.assert_array_equal(got, want.astype(np.int64))          # NaN / outliers dropped
    # batched frames
    rng = np.random.default_rng(3)
    v = rng.random((5, 4001, 2))
    got = M.hist2d(v, np.linspace(0, 1, 21), np.linspace(0, 1, 14))
    for f in range(5):
        w, _, _ = np.histogram2d(v[f, :, 0], v[f, :, 1], bins=[20, 13], range=[(0, 1), (0, 1)])
        np.testing.assert_array_equal(got[f], w.astype(np.int64))
    assert got.sum() == 5 * 4001
    # empty input
    assert M.hist2d(np.zeros((0, 2)), ed, ed).sum() == 0


def test_order_stats_and_device_bin_plan(M):
    """Exact order statistics by device radix select, and the make_histograms bin plan built on
    them (global min/max + scipy.stats.iqr), against NumPy/SciPy on the host."""
    from pycpet_b200 import calculator as calc

    rng = np.random.default_rng(21)
    v = np.concatenate([rng.normal(0, 1, 50_000), rng.gamma(2.0, 0.3, 50_000), [0.0, -0.0, 1e-40, -1e-40],
                        np.full(100, 0.5), [np.inf, -np.inf]]).astype(np.float32)
    rng.shuffle(v)
    s = np.sort(v)
    ranks = np.array([0, 1, 17, len(v) // 4, len(v) // 2, 3 * len(v) // 4, len(v) - 2, len(v) - 1])
    got = M.order_stats(v, ranks)
    np.testing.assert_array_equal(got, s[ranks])
    # NaNs are ordered last, like np.sort
    vn = v.copy()
    vn[::1000] = np.nan
    sn = np.sort(vn)
    got = M.order_stats(vn, [0, len(vn) // 2, len(vn) - 150, len(vn) - 1])
    np.testing.assert_array_equal(got, sn[[0, len(vn) // 2, len(vn) - 150, len(vn) - 1]])
    # strided columns of an (n,2) array
    two = np.column_stack([v, -v]).astype(np.float32)
    np.testing.assert_array_equal(M.order_stats(two, [5, 100], column=1), np.sort(-v)[[5, 100]])
    # single element
    np.testing.assert_array_equal(M.order_stats(np.array([3.5], np.float32), [0]), [3.5])
    # the bin plan of make_histograms (UC:664-685)
    for n_frames, n in [(1, 5832), (4, 5832), (3, 1000)]:
        tops = [np.column_stack([rng.gamma(2.0 + 0.1 * i, 0.3, n), rng.gamma(1.5, 0.4, n)]).astype(np.float32)
                for i in range(n_frames)]
        assert calc.bin_plan_device(tops) == calc.bin_plan(tops)


def test_radix_pass_matches_numpy_twin():
    """Engine.radix_hist (the single select pass a multi-GPU caller all-reduces) against the NumPy
    twin used by the gloo test, and sharding.order_stats_sharded driven by the device pass."""
    import torch

    from pycpet_b200 import sharding
    from pycpet_b200.device import Engine
    from test_sharding_gloo import np_radix_hist

    rng = np.random.default_rng(5)
    v = np.column_stack([rng.gamma(2.0, 0.3, 20_001), rng.normal(0, 2, 20_001)]).astype(np.float32)
    v[7, 1] = np.nan
    eng = Engine(0)
    dv = torch.from_numpy(v).cuda()
    for col in (0, 1):
        h0 = eng.radix_hist(dv, col, [0], 0)
        np.testing.assert_array_equal(h0, np_radix_hist(v[:, col], [0], 0))
        top = int(np.argmax(h0[0]))
        for bits, pre in ((8, [top, (top + 1) % 256]), (16, [top << 8 | 3, top << 8 | 200])):
            np.testing.assert_array_equal(eng.radix_hist(dv, col, pre, bits), np_radix_hist(v[:, col], pre, bits))
        ranks = [0, 5000, 10_000, 20_000]
        got = sharding.order_stats_sharded(lambda pre, bits: eng.radix_hist(dv, col, pre, bits), ranks)
        np.testing.assert_array_equal(got, np.sort(v[:, col])[ranks])
        np.testing.assert_array_equal(eng.order_stats(dv, ranks, column=col), np.sort(v[:, col])[ranks])
    eng.close()


def test_make_histograms_and_chi2(M, tmp_path):
    from pycpet_b200 import calculator as calc

    rng = np.random.default_rng(9)
    tops = [np.column_stack([rng.gamma(2.0 + 0.2 * i, 0.3, 5832), rng.gamma(1.5, 0.4, 5832)]).astype(np.float32)
            for i in range(4)]
    files = []
    for i, t in enumerate(tops):
        p = tmp_path / f"f{i}.top"
        np.savetxt(p, t)                                  # what CPET.run_topo writes (CPET.py:123)
        files.append(str(p))
    H = calc.make_histograms(files)
    # the reference's make_histograms, restated: global range, IQR bin width, np.histogram2d
    allv = np.concatenate(tops).astype(np.float64)
    dr, cr, nd, nc = ohist.bin_plan(allv[:, 0], allv[:, 1], 5832)
    want = np.stack([ohist.normalised_hist(t[:, 0].astype(np.float64), t[:, 1].astype(np.float64), nd, nc, dr, cr)
                     for t in tops])
    np.testing.assert_array_equal(H, want)
    D = calc.construct_distance_matrix(H)
    Dw = ohist.chi2_matrix(want)
    np.testing.assert_allclose(D, Dw, rtol=1e-12, atol=1e-15)
    assert np.all(np.diag(D) == 0) and np.allclose(D, D.T, rtol=0, atol=0)
    assert abs(calc.distance_numpy(H[0], H[1]) - ohist.chi2(want[0], want[1])) < 1e-14


def test_make_histograms_and_chi2_pinned_to_the_reference(M, golden, tmp_path):
    """The drop-in `make_histograms`, `construct_distance_matrix`, `distance_numpy`, the device bin
    plan and the fixed-range histogram against outputs of the UNMODIFIED reference functions
    (tests/golden/make_golden.py --hist): histogram rows bit-exact, chi^2 <= 1e-12."""
    import warnings

    from pycpet_b200 import calculator as calc

    g = golden("histograms_reference.npz")
    for tag, n in (("shipped", 2), ("synth", 3), ("synth_eq", 3)):
        tops = [g[f"{tag}_top{i}"] for i in range(n)]
        files = []
        for i, t in enumerate(tops):
            p = tmp_path / f"{tag}_{i}.top"
            np.savetxt(p, t)                              # CPET.py:123
            files.append(str(p))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")               # the ragged set warns, like the reference (UC:652-657)
            H = calc.make_histograms(files)
            plan_host = calc.bin_plan(tops)
            plan_dev = calc.bin_plan_device(tops)
        want = g[f"{tag}_hist"]
        assert H.shape == want.shape
        np.testing.assert_array_equal(H, want)
        assert plan_dev == plan_host and plan_host[2] * plan_host[3] == want.shape[1]
        np.testing.assert_array_equal(calc.make_histograms_from_arrays(tops, plan=plan_dev), want)
        D = calc.construct_distance_matrix(H)
        np.testing.assert_allclose(D, g[f"{tag}_dist"], rtol=1e-12, atol=1e-15)
        assert abs(calc.distance_numpy(H[0], H[1]) - float(g[f"{tag}_d01"])) < 1e-14
    lo_d, hi_d, nd, lo_c, hi_c, nc = g["grid_fixed_args"]
    t = g["shipped_top0"]
    cnt = calc.histogram2d_counts(t.astype(np.float64), int(nd), int(nc), (lo_d, hi_d), (lo_c, hi_c)).astype(np.float64)
    np.testing.assert_array_equal((cnt / cnt.sum()).flatten(), g["grid_fixed"])


def test_trajectory_pipeline_matches_make_histograms(M, frame2a, tmp_path):
    """pycpet_b200.trajectory (rows stay in HBM: K2 -> radix-select bin plan -> batched K3 -> chi^2 rows)
    against the file-based route the reference takes for the same frames: .top text per frame ->
    calculator.make_histograms (host scipy IQR plan) -> construct_distance_matrix."""
    import torch

    from pycpet_b200 import calculator as calc, trajectory
    from pycpet_b200.device import Engine
    from pycpet_b200.io import save_topology

    x, Q = frame2a
    seeds, n_iter, dims, _ = synth.seeds(16, 0.5, 0.1)

    def frame(f):
        rng = np.random.default_rng(300 + f)
        xf = (x + rng.normal(0, 0.3, x.shape)).astype(np.float32)
        xf[np.all(np.abs(xf) < 0.55, axis=1)] *= 3.0
        return torch.from_numpy(xf).cuda(), torch.from_numpy(Q).cuda()

    eng = Engine(0)
    tr = trajectory.topology_trajectory(eng, 7, frame, seeds, n_iter, 0.1, dims)
    rows = tr["rows"].cpu().numpy()
    assert rows.shape == (7, 4096, 2) and tr["frames"] == list(range(7))
    files = []
    for f in range(7):
        p = str(tmp_path / f"frame{f}.top")
        save_topology(p, rows[f])
        files.append(p)
    want = calc.make_histograms(files)
    assert tr["plan"] == calc.bin_plan([r.astype(np.float64) for r in rows])
    np.testing.assert_array_equal(tr["hists"].cpu().numpy(), want)
    np.testing.assert_allclose(tr["distance"].cpu().numpy(), calc.construct_distance_matrix(want), rtol=1e-13, atol=0)
    np.testing.assert_allclose(tr["distance"].cpu().numpy(), ohist.chi2_matrix(want), rtol=1e-12, atol=0)
    # a row block of the matrix carries the bits of the full matrix, both triangles
    full = eng.chi2_rows(tr["hists"]).cpu().numpy()
    np.testing.assert_array_equal(full, full.T)
    np.testing.assert_array_equal(eng.chi2_rows(tr["hists"], 2, 3).cpu().numpy(), full[2:5])
    np.testing.assert_array_equal(full, M.chi2_matrix(want))
    # work accounting: sum over frames and lines of (K + 2) x M
    total = 0
    for f in range(7):
        xf, Qf = frame(f)
        eng.set_charges(xf, Qf)
        _, st = eng.topo_batch(torch.from_numpy(seeds).cuda(), n_iter, 0.1, dims, want_steps=True)
        total += int(st.sum().item() + 2 * len(seeds)) * len(Q)
    assert tr["pair_evals"] == total
    eng.close()


def test_topo_hist_fused_call(M, frame2a):
    """cpet_topo_hist = cpet_topo_batch + cpet_hist2d with the rows kept on the device."""
    x, Q = frame2a
    seeds, n_iter, dims, _ = synth.seeds(14, 0.5, 0.1)
    de, ce = np.linspace(0, 1.8, 31), np.linspace(0, 3.0, 41)
    M.set_charges(x, Q)
    rows = M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims)
    out = np.zeros((len(seeds), 2), np.float32)
    rows2, counts = M.topo_hist(seeds, n_iter, de, ce, step_size=0.1, dimensions=dims, out=out)
    assert rows2 is out
    np.testing.assert_array_equal(rows2, rows)
    want, _, _ = np.histogram2d(rows[:, 0].astype(np.float64), rows[:, 1].astype(np.float64),
                                bins=[30, 40], range=[(0, 1.8), (0, 3.0)])
    np.testing.assert_array_equal(counts, want.astype(np.int64))
    c = M.last_counters()
    assert c["pair_evals"] > 0 and c["launches"] >= 2
    none_rows, counts2 = M.topo_hist(seeds, n_iter, de, ce, step_size=0.1, dimensions=dims, want_rows=False)
    assert none_rows is None
    np.testing.assert_array_equal(counts2, counts)
    with pytest.raises(ValueError):
        M.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, out=np.zeros((3, 2), np.float32))


# --------------------------------------------------------------------------- legacy symbols -----
def test_topo_hist_frames_matches_single_frame_calls(M):
    """The MD-frame batch call (two internal streams) returns, frame by frame, exactly what
    topo_hist returns for that frame alone: shared and per-frame n_iter, different charge counts per
    frame (one of them streaming its charges), F = 1 and F = 0."""
    seeds, n_iter, dims, max_steps = synth.seeds(9, 0.5, 0.1)
    L = len(seeds)
    de, ce = np.linspace(0, 1.7, 31), np.linspace(0, 4.0, 41)
    rng = np.random.default_rng(5)
    frames = []
    for f, m_f in enumerate([3000, 2999, 30001, 1, 7890]):
        x, Q = synth.charges(m_f, seed=20 + f, box=0.5)
        frames.append((x, Q))
    nit_frames = rng.integers(1, max_steps, size=(len(frames), L)).astype(np.int32)
    reset_tuning(M)
    for nit in (n_iter, nit_frames):
        rows, counts = M.topo_hist_frames(frames, seeds, nit, de, ce, step_size=0.1, dimensions=dims,
                                          want_rows=True)
        c = M.last_counters()
        evals = 0
        for f, (x, Q) in enumerate(frames):
            nf = nit if np.ndim(nit) == 1 else nit[f]
            r1, c1 = M.topo_hist(seeds, nf, de, ce, x=x, Q=Q, step_size=0.1, dimensions=dims)
            evals += M.last_counters()["pair_evals"]
            np.testing.assert_array_equal(rows[f], r1)
            np.testing.assert_array_equal(counts[f], c1)
        assert c["pair_evals"] == evals
        _, counts_only = M.topo_hist_frames(frames, seeds, nit, de, ce, step_size=0.1, dimensions=dims)
        np.testing.assert_array_equal(counts_only, counts)
    rows1, counts1 = M.topo_hist_frames(frames[:1], seeds, n_iter, de, ce, step_size=0.1, dimensions=dims,
                                        want_rows=True)
    np.testing.assert_array_equal(rows1[0], M.topo_hist(seeds, n_iter, de, ce, x=frames[0][0], Q=frames[0][1],
                                                        step_size=0.1, dimensions=dims)[0])
    rows0, counts0 = M.topo_hist_frames([], seeds, n_iter, de, ce, step_size=0.1, dimensions=dims, want_rows=True)
    assert rows0.shape == (0, L, 2) and counts0.shape == (0, 30, 40)
    with pytest.raises(ValueError):
        M.topo_hist_frames(frames, seeds, nit_frames[:2], de, ce, step_size=0.1, dimensions=dims)


def test_legacy_symbols(M, golden):
    g = golden("synthetic_math_ops.npz")
    x, Q, pts = g["x"], g["Q"], g["points"]
    want = f64.field_grid(pts[:4], x, Q, False)
    for i in range(4):
        assert relmax(M.calc_field(pts[i], x, Q), want[i]) < FIELD_TOL
        assert relmax(M.calc_field_base(pts[i], x, Q), want[i]) < FIELD_TOL
        assert abs(M.calc_esp_base(pts[i], x, Q)[0] - f64.esp_grid(pts[i:i + 1], x, Q)[0]) < 1e-5 * 5
    # accumulate-into-output quirk of calc_field_base / calc_esp_base (C:327-332, 482-485)
    acc = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    M.math.calc_field_base(acc, pts[0], len(Q), x, Q)
    np.testing.assert_allclose(acc - np.array([1, 2, 3], np.float32), want[0], rtol=1e-4, atol=1e-5)
    esp = np.array([10.0], dtype=np.float32)
    M.math.calc_esp_base(esp, pts[0], len(Q), x, Q)
    assert abs((esp[0] - 10.0) - f64.esp_grid(pts[:1], x, Q)[0]) < 1e-4
    # helper einsum symbols
    rng = np.random.default_rng(0)
    A = rng.random((50, 3)).astype(np.float32)
    np.testing.assert_allclose(M.einsum_ij_i(A), A.sum(1), rtol=1e-6)
    R = rng.normal(size=(40, 3)).astype(np.float32)
    rm = rng.random(40).astype(np.float32)
    q = rng.normal(size=40).astype(np.float32)
    want_e = 14.3996451 * np.einsum("i,i,ij->j", q.astype(np.float64), rm.astype(np.float64), R.astype(np.float64))
    np.testing.assert_allclose(M.einsum_operation(R, rm, q), want_e, rtol=1e-5)
    Rb = rng.normal(size=(3, 40, 3)).astype(np.float32)
    rmb = rng.random((3, 40)).astype(np.float32)
    want_b = 14.3996451 * np.einsum("i,bi,bij->bj", q.astype(np.float64), rmb.astype(np.float64), Rb.astype(np.float64))
    np.testing.assert_allclose(M.einsum_operation_batch(Rb, rmb, q, 3), want_b, rtol=1e-5)
    a, b = rng.random(100).astype(np.float32), rng.random(100).astype(np.float32)
    np.testing.assert_array_equal(M.vecaddn(a, b), a + b)
    Ad, Bd = rng.random((7, 9)), rng.random(9)
    np.testing.assert_allclose(M.dot(Ad, Bd), Ad @ Bd, rtol=1e-14)
    import scipy.sparse as sp
    S = sp.random(20, 9, density=0.3, format="csr", random_state=1)
    np.testing.assert_allclose(M.sparse_dot(S, Bd), S @ Bd, rtol=1e-14)


def test_dipole_tracer_symbol(M):
    """thread_operation_dipole (development-only in the reference, C:593-660): checked against a
    NumPy transcription of the same expression."""
    rng = np.random.default_rng(4)
    x = rng.uniform(-8, 8, (300, 3)).astype(np.float32)
    x = x[np.linalg.norm(x, axis=1) > 2.0]
    mu = rng.normal(size=(len(x), 3)).astype(np.float32)
    dims = np.array([1, 1, 1], np.float32)
    seed = np.array([0.1, -0.2, 0.3], np.float32)

    def field(p):
        R = p[None, :].astype(np.float64) - x
        rn = np.linalg.norm(R, axis=1)
        proj = 3 * mu[:, 0] * R[:, 0] + mu[:, 1] * R[:, 1] + mu[:, 2] * R[:, 2]
        return (14.3996451 * rn[:, None] ** -5 * (proj[:, None] * R - mu * rn[:, None] ** 2)).sum(0)

    def step(p):
        e = field(p)
        return p + 0.1 * e / np.linalg.norm(e)

    p = seed.astype(np.float64)
    for _ in range(5):
        p = step(p)
        if np.any(np.abs(p) > 1):
            break
    got = M.thread_operation_dipole(seed, 5, x, mu, 0.1, dims)
    assert abs(got[0] - np.linalg.norm(p - seed)) < 1e-5


def test_calculator_level_entry_points(golden, frame2a):
    import types

    import pycpet_b200 as pc

    g = golden("example_2A_volume.npz")
    t = golden("example_3A_topo.npz")
    x, Q = frame2a
    calc = types.SimpleNamespace(x=x, Q=Q.reshape(-1, 1), mesh=g["mesh"], dimensions=t["dimensions"],
                                 step_size=float(t["step_size"]), random_start_points=t["seeds"],
                                 random_max_samples=t["n_iter"], n_samples=216)
    fb, shape = pc.compute_box(calc)
    assert shape == (11, 11, 11, 3) and fb.shape == (1331, 6) and fb.dtype == np.float32
    assert relmax(fb[:, 3:], g["field_box"][:, 3:]) < FIELD_TOL
    eb, shape = pc.compute_box_ESP(calc)
    assert eb.dtype == np.float16 and eb.shape == (1331, 4)
    hist = pc.compute_topo_complete_c_shared(calc)
    assert hist.shape == (216, 2) and calc.hist is hist
    assert np.max(np.abs(hist[:, 0] - t["hist"][:, 0])) <= 2e-6
    hist_gpu = pc.compute_topo_GPU_batch_filter(calc)
    np.testing.assert_array_equal(hist_gpu, hist)
    golden1 = golden("example_1A_point_field.npz")
    c1 = types.SimpleNamespace(x=golden1["x"], Q=golden1["Q"])
    np.testing.assert_allclose(pc.compute_point_field(c1), golden1["field"], rtol=2e-5)


def test_engine_device_pointers_match_host_path(M, frame2a):
    import torch

    from pycpet_b200.device import Engine

    x, Q = frame2a
    eng = Engine(0)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    pts = synth.grid(11, 0.5)
    e_dev = eng.field_grid(torch.from_numpy(pts).cuda(), soften=True, concat=True)
    np.testing.assert_array_equal(e_dev.cpu().numpy(), M.field_grid(pts, x, Q, soften=True, concat=True))
    seeds, n_iter, dims, _ = synth.seeds(12, 0.5, 0.1)
    t_dev, s_dev = eng.topo_batch(torch.from_numpy(seeds).cuda(), n_iter, 0.1, dims, want_steps=True)
    t_host, s_host = M.topo_batch(seeds, n_iter, x, Q, 0.1, dims, want_steps=True)
    np.testing.assert_array_equal(t_dev.cpu().numpy(), t_host)
    np.testing.assert_array_equal(s_dev.cpu().numpy(), s_host)
    de, ce = np.linspace(0, 1.8, 31), np.linspace(0, 3.0, 41)
    h_dev = eng.hist2d(t_dev, de, ce)
    np.testing.assert_array_equal(h_dev.cpu().numpy(), M.hist2d(t_host, de, ce))
    assert eng.fp32_peak_tflops(packed=True, iters=512) > 10.0
    eng.close()
