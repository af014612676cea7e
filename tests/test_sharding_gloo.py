"""CPU, gloo, world_size 2: the N>1 path (partition -> local compute -> gather / all-reduce).
The local compute is the float64 oracle here (no GPU in this tier); on the GPU box the same
functions run with Engine methods over NCCL (tests/test_gpu_parity.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def np_radix_hist(values, prefixes, prefix_bits):
    """NumPy twin of radix_hist_kernel (hist.cu): monotone float32 keys, NaNs last."""
    u = np.ascontiguousarray(values, dtype=np.float32).view(np.uint32)
    key = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    key[np.isnan(values)] = np.uint32(0xFFFFFFFF)
    out = np.zeros((len(prefixes), 256), dtype=np.uint64)
    for t, pre in enumerate(prefixes):
        sel = key if prefix_bits == 0 else key[(key >> np.uint32(32 - prefix_bits)) == np.uint32(pre)]
        bins = (sel >> np.uint32(24 - prefix_bits)) & np.uint32(255)
        out[t] = np.bincount(bins.astype(np.int64), minlength=256)
    return out


class OracleEngine:
    """Host stand-in with the Engine methods pycpet_b200.trajectory uses, computing with the oracle
    (float64 streamlines rounded to float32 rows, NumPy histograms): what is under test is the
    trajectory pipeline's partitioning, its all-reduced radix select and its gathers."""

    def __init__(self):
        from oracle import f64, hist as ohist
        self.f64, self.ohist = f64, ohist
        self.device = torch.device("cpu")
        self.n_charges = 0

    def set_charges(self, x, Q):
        self.x, self.Q = np.asarray(x, np.float32), np.asarray(Q, np.float32)
        self.n_charges = len(self.Q)

    def topo_batch(self, seeds, n_iter, step_size, dimensions, second_diff=False, out=None, steps=None):
        rows, st = self.f64.topo_batch(seeds.numpy(), n_iter.numpy().astype(np.int64), self.x, self.Q, step_size,
                                       dimensions)
        out.copy_(torch.from_numpy(rows.astype(np.float32)))
        if steps is not None:
            steps.copy_(torch.from_numpy(st.astype(np.int32)))
        return out

    def radix_hist(self, values, column, prefixes, prefix_bits):
        return np_radix_hist(values.numpy()[:, column], prefixes, prefix_bits)

    def hist2d(self, values, d_edges, c_edges):
        v = values.numpy().astype(np.float64)
        return torch.from_numpy(np.stack([
            np.histogram2d(f[:, 0], f[:, 1], bins=[d_edges, c_edges])[0].astype(np.int64) for f in v]))

    def chi2_rows(self, hists, row0=0, n_rows=None):
        return torch.from_numpy(self.ohist.chi2_matrix(hists.numpy())[row0:row0 + n_rows].copy())


def _trajectory_inputs():
    x, Q = synth.charges(300, seed=2, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(6, 0.5, 0.1)

    def frame(f):
        rng = np.random.default_rng(50 + f)
        xf = (x + rng.normal(0, 0.3, x.shape)).astype(np.float32)
        xf[np.all(np.abs(xf) < 0.55, axis=1)] *= 3.0
        return xf, Q
    return frame, seeds, n_iter, dims


def _worker(rank, size, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from oracle import f64, hist as ohist
    from pycpet_b200 import sharding

    x, Q = synth.charges(500, seed=1, box=0.5)
    pts = synth.grid(5, 0.5)
    seeds, n_iter, dims, _ = synth.seeds(4, 0.5, 0.1)

    full_field = sharding.grid_sharded(lambda p: f64.field_grid(p, x, Q, True), pts).numpy()
    # the same mesh by slabs of x-planes (5 planes over 2 ranks: ragged 3 + 2), gathered into a caller buffer
    ax = np.linspace(-0.5, 0.5, 5)
    lat_out = torch.empty((125, 3), dtype=torch.float64)
    def lattice_rows(xs, ys, zs):
        g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
        return f64.field_grid(g, x, Q, True)
    got = sharding.lattice_sharded(lattice_rows, ax, ax, ax, out=lat_out)
    assert got is lat_out
    full_topo = sharding.topo_sharded(lambda s, n: f64.topo_batch(s, n, x, Q, 0.1, dims)[0], seeds, n_iter).numpy()
    # histogram of the lines held by this rank, reduced
    ids = sharding.deal_lines(n_iter, rank, size)
    ref_topo = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)[0]
    lo_d, hi_d, lo_c, hi_c = sharding.global_ranges(ref_topo[ids])
    de, ce = ohist.edges(lo_d, hi_d, 7), ohist.edges(lo_c, hi_c, 5)
    counts = sharding.hist_sharded(
        lambda v, a, b: ohist.hist2d_counts(v[:, 0], v[:, 1], 7, 5, (a[0], a[-1]), (b[0], b[-1])),
        ref_topo[ids], de, ce).numpy()
    frames = sharding.frames_sharded(lambda f: np.full((2, 3), float(f)), 5).numpy()
    # one call per rank over its frame ids (the shape of Math_ops.topo_hist_frames): int64 counts
    batch = sharding.frames_batch_sharded(
        lambda ids: np.stack([np.full((3, 2), 10 * f, dtype=np.int64) for f in ids]), 7).numpy()
    # fewer frames than ranks: the idle rank still returns the gathered result with the right dtype
    lone = sharding.frames_batch_sharded(
        lambda ids: np.stack([np.full((2, 2), 7 + f, dtype=np.int64) for f in ids]), 1).numpy()
    # exact global order statistics of values spread over the ranks (radix select + all-reduce)
    mine32 = ref_topo[ids][:, 1].astype(np.float32)
    n_all = len(ref_topo)
    stats = sharding.order_stats_sharded(lambda pre, bits: np_radix_hist(mine32, pre, bits),
                                         [0, n_all // 4, n_all // 2, (3 * n_all) // 4, n_all - 1])
    # the device-resident trajectory pipeline (5 frames over 2 ranks: 3 + 2) with the oracle as the engine
    from pycpet_b200 import trajectory
    frame, tseeds, tn_iter, tdims = _trajectory_inputs()
    tr = trajectory.topology_trajectory(OracleEngine(), 5, frame, tseeds, tn_iter, 0.1, tdims)
    assert tr["frames"] == list(range(rank, 5, size)) and tr["rows"].shape == (len(tr["frames"]), 216, 2)
    np.savez(os.path.join(tmp, f"traj{rank}.npz"), plan=np.array([*tr["plan"][0], *tr["plan"][1], tr["plan"][2], tr["plan"][3]]),
             counts=tr["counts"].numpy(), hists=tr["hists"].numpy(), distance=tr["distance"].numpy(),
             pairs=np.array([tr["pair_evals"]]))
    np.savez(os.path.join(tmp, f"r{rank}.npz"), lattice=lat_out.numpy(), field=full_field, topo=full_topo, counts=counts,
             frames=frames, batch=batch, lone=lone, ranges=np.array([lo_d, hi_d, lo_c, hi_c]), stats=stats)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    from oracle import f64, hist as ohist

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    x, Q = synth.charges(500, seed=1, box=0.5)
    pts = synth.grid(5, 0.5)
    seeds, n_iter, dims, _ = synth.seeds(4, 0.5, 0.1)
    field = f64.field_grid(pts, x, Q, True)
    topo = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)[0]
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["field"], field)
        np.testing.assert_array_equal(z["lattice"], field)
        np.testing.assert_array_equal(z["topo"], topo)
        rg = z["ranges"]
        assert rg[0] == topo[:, 0].min() and rg[1] == topo[:, 0].max()
        expect = ohist.hist2d_counts(topo[:, 0], topo[:, 1], 7, 5, (rg[0], rg[1]), (rg[2], rg[3]))
        np.testing.assert_array_equal(z["counts"], expect)
        assert z["counts"].sum() == len(topo)
        np.testing.assert_array_equal(z["frames"][:, 0, 0], np.arange(5.0))
        assert z["batch"].dtype == np.int64 and z["batch"].shape == (7, 3, 2)
        np.testing.assert_array_equal(z["batch"][:, 2, 1], 10 * np.arange(7))
        assert z["lone"].dtype == np.int64 and z["lone"].shape == (1, 2, 2) and np.all(z["lone"] == 7)
        srt = np.sort(topo[:, 1].astype(np.float32))
        n_all = len(srt)
        np.testing.assert_array_equal(z["stats"], srt[[0, n_all // 4, n_all // 2, (3 * n_all) // 4, n_all - 1]])
    # trajectory pipeline: what calculator.bin_plan / make_histograms / construct_distance_matrix's rule give
    # on the host for the same five frames, on every rank
    from pycpet_b200 import calculator as calc
    frame, tseeds, tn_iter, tdims = _trajectory_inputs()
    rows, pairs = [], 0
    for f in range(5):
        xf, Qf = frame(f)
        r, st = f64.topo_batch(tseeds, tn_iter, xf, Qf, 0.1, tdims)
        rows.append(r.astype(np.float32))
        pairs += int((st.astype(np.int64) + 2).sum()) * len(Qf)
    d_range, c_range, nd, nc = calc.bin_plan(rows)
    want_counts = np.stack([np.histogram2d(r[:, 0].astype(np.float64), r[:, 1].astype(np.float64), bins=[nd, nc],
                                           range=[d_range, c_range])[0].astype(np.int64) for r in rows])
    want_h = want_counts.reshape(5, -1).astype(np.float64)
    want_h = want_h / want_h.sum(axis=1, keepdims=True)
    for r in range(2):
        z = np.load(tmp_path / f"traj{r}.npz")
        np.testing.assert_array_equal(z["plan"], np.array([*d_range, *c_range, nd, nc]))
        np.testing.assert_array_equal(z["counts"], want_counts)
        np.testing.assert_array_equal(z["hists"], want_h)
        np.testing.assert_allclose(z["distance"], ohist.chi2_matrix(want_h), rtol=1e-13, atol=0)
        assert int(z["pairs"][0]) * (1 if r else 1) > 0
    assert int(np.load(tmp_path / "traj0.npz")["pairs"][0]) + int(np.load(tmp_path / "traj1.npz")["pairs"][0]) == pairs


def test_trajectory_rejects_a_degenerate_bin_plan():
    """Lines whose distances are all equal have a zero inter-quartile range: UC:669-685 would ask for an infinite
    number of bins; the pipeline says so instead of launching a histogram with it."""
    from pycpet_b200 import trajectory

    class Flat(OracleEngine):
        def topo_batch(self, seeds, n_iter, step_size, dimensions, second_diff=False, out=None, steps=None):
            out[:, 0] = 0.25
            out[:, 1] = torch.linspace(0.1, 2.0, out.shape[0])
            if steps is not None:
                steps.fill_(1)
            return out

    frame, tseeds, tn_iter, tdims = _trajectory_inputs()
    with pytest.raises((ValueError, ZeroDivisionError, OverflowError)):
        trajectory.topology_trajectory(Flat(), 2, frame, tseeds, tn_iter, 0.1, tdims)
