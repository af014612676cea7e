"""CPU: host-side logic of the mirror (bin planning, .top parsing, partitioners)."""
import numpy as np

from oracle import hist as ohist
from pycpet_b200 import calculator as calc
from pycpet_b200 import sharding


def test_bin_plan_matches_reference_rule(tmp_path):
    rng = np.random.default_rng(5)
    tops = [np.column_stack([rng.gamma(2.0, 0.3, 1000), rng.gamma(1.5, 0.4, 1000)]).astype(np.float32)
            for _ in range(3)]
    dr, cr, nd, nc = calc.bin_plan(tops)
    allv = np.concatenate(tops).astype(np.float64)
    dr2, cr2, nd2, nc2 = ohist.bin_plan(allv[:, 0], allv[:, 1], 1000)
    assert (dr, cr, nd, nc) == (dr2, cr2, nd2, nc2)
    # .top round trip: written like CPET.py:123 (np.savetxt default %.18e)
    p = tmp_path / "a.top"
    np.savetxt(p, tops[0])
    back = calc.read_top_file(str(p))
    np.testing.assert_array_equal(back, tops[0].astype(np.float64))


def test_partitioners():
    for n, size in [(10, 3), (1331, 8), (5, 8), (0, 2)]:
        spans = [sharding.slab(n, r, size) for r in range(size)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(size - 1))
        lens = [b - a for a, b in spans]
        assert max(lens) - min(lens) <= 1
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    n_iter = np.random.RandomState(0).randint(1, 17, 1000)
    parts = [sharding.deal_lines(n_iter, r, 4) for r in range(4)]
    allids = np.sort(np.concatenate(parts))
    np.testing.assert_array_equal(allids, np.arange(1000))
    work = [n_iter[p].sum() for p in parts]
    assert max(work) / min(work) < 1.05


def test_plan_from_order_stats_reproduces_numpy_scipy():
    """The bin plan computed from six order statistics per column (what the device radix select
    returns) equals make_histograms' np.min/np.max/scipy.stats.iqr route bit for bit."""
    rng = np.random.default_rng(11)
    for n_frames, n in [(1, 10), (3, 1000), (2, 5832), (5, 777), (1, 2)]:
        tops = [np.column_stack([rng.gamma(2.0, 0.3, n), rng.gamma(1.5, 0.4, n)]).astype(np.float32)
                for _ in range(n_frames)]
        allv = np.concatenate(tops)
        N = len(allv)
        (p25, n25, g25), (p75, n75, g75) = calc.quartile_ranks(N)
        stats = []
        for col in (0, 1):
            s = np.sort(allv[:, col])
            stats.append(((s[0], s[p25], s[n25], s[p75], s[n75], s[-1]), (g25, g75)))
        got = calc.plan_from_order_stats(stats, n)
        if got[2] > 0 and got[3] > 0:
            assert got == calc.bin_plan(tops)
        # the quartiles themselves against numpy
        for col in (0, 1):
            (lo, a25, b25, a75, b75, hi), _ = stats[col]
            x64 = allv[:, col].astype(np.float64)
            assert calc._lerp(a25, b25, g25) == np.percentile(x64, 25)
            assert calc._lerp(a75, b75, g75) == np.percentile(x64, 75)


def test_compute_curv_and_dist_known_answers():
    """UC:541-562 mirror: curvature of a circle is 1/R, of a straight line 0; a zero first
    difference takes the eps branch (UC:517-518); dist is the end-to-end distance."""
    from pycpet_b200 import calculator as calc

    R, h = 2.0, 0.01
    def on_circle(t0):
        t = t0 + np.arange(3) * (h / R)
        return [np.array([R * np.cos(a), R * np.sin(a), 0.3]) for a in t]
    a0, a1, a2 = on_circle(0.2)
    b0, b1, b2 = on_circle(1.1)
    dist, curv = calc.compute_curv_and_dist(a0, a1, a2, b0, b1, b2)
    assert abs(curv - 1.0 / R) < 1e-4
    assert abs(dist - np.linalg.norm(a0 - b0)) < 1e-15
    p = np.array([0.1, -0.2, 0.3]); d = np.array([0.3, 0.4, 0.5])
    dist, curv = calc.compute_curv_and_dist(p, p + d, p + 2 * d, p + 5 * d, p + 6 * d, p + 7 * d)
    assert curv < 1e-12 and abs(dist - 5 * np.linalg.norm(d)) < 1e-12
    dist, curv = calc.compute_curv_and_dist(p, p, p + d, p, p, p + d)      # v' = 0 -> num / eps = 0 / 1e-5
    assert curv == 0.0 and dist == 0.0


def test_topo_hist_frames_argument_checks_need_no_device():
    """Shape errors of the MD-frame batch call are raised by the Python mirror before anything
    touches the GPU (the reference's ndpointer argtypes reject bad shapes the same way)."""
    import pytest

    from pycpet_b200 import Math_ops

    m = Math_ops()
    seeds = np.zeros((5, 3), np.float32)
    frames = [(np.zeros((4, 3), np.float32), np.zeros(4, np.float32))] * 3
    e = np.linspace(0, 1, 4)
    with pytest.raises(ValueError, match="n_iter"):
        m.topo_hist_frames(frames, seeds, np.ones(4, np.int32), e, e)
    with pytest.raises(ValueError, match="n_iter"):
        m.topo_hist_frames(frames, seeds, np.ones((2, 5), np.int32), e, e)
    with pytest.raises(ValueError, match="rows but Q"):
        m.topo_hist_frames([(np.zeros((4, 3), np.float32), np.zeros(3, np.float32))], seeds, np.ones(5, np.int32), e, e)


def test_gather_helpers_without_a_process_group():
    """world_size 1: the gather helpers copy into the caller's buffer, keep the dtype, and reject a block whose length
    disagrees with the plan; deal_lines_all partitions the lines and deals blocks in serpentine order."""
    import pytest
    import torch

    rows = np.arange(12, dtype=np.int64).reshape(6, 2)
    out = torch.empty((6, 2), dtype=torch.int64)
    got = sharding.all_gather_blocks(rows, [6], out=out)
    assert got is out and got.dtype == torch.int64 and np.array_equal(got.numpy(), rows)
    with pytest.raises(ValueError):
        sharding.all_gather_blocks(rows, [5])
    with pytest.raises(ValueError):
        sharding.all_gather_blocks(rows, [6], out=torch.empty((6, 2), dtype=torch.float32))
    ax = np.linspace(-1, 1, 4)
    full = sharding.lattice_sharded(lambda xs, ys, zs: np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1).reshape(-1, 3), ax, ax, ax)
    assert full.shape == (64, 3) and np.array_equal(full.numpy()[:, 2][:4], ax)          # z fastest
    n_iter = np.random.RandomState(3).randint(1, 17, 1000)
    deal = sharding.deal_lines_all(n_iter, 4)
    assert sorted(np.concatenate(deal).tolist()) == list(range(1000))
    order = np.argsort(-n_iter, kind="stable")
    assert np.array_equal(deal[0][:32], order[:32]) and np.array_equal(deal[3][:32], order[96:128])
    assert np.array_equal(deal[3][32:64], order[128:160])                                # second round runs backwards
    work = [int(n_iter[d].sum()) for d in deal]
    assert max(work) - min(work) <= 0.02 * np.mean(work)
