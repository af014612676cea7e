"""MD-trajectory topology batch, device-resident end to end (BASELINE.json configs[3]; SURVEY.md
section 8 rows a5, a9, a10, f1, f3 chained without leaving the GPUs).

What the reference does per trajectory (CPET/source/CPET.py:96-127 run_topo -> one .top text file per
frame; CPET/source/cluster.py -> make_histograms, CPET/utils/calculator.py:596-718: three parse passes,
global min/max, scipy.stats.iqr bin widths, np.histogram2d per frame; construct_distance_matrix
UC:1003-1015) becomes, with frames dealt round-robin over the ranks (sharding.frames_for_rank):

    K2   per frame: pack charges -> streamlines -> (L,2) [dist|curv] rows kept in HBM  (F/G x L x 8 B per GPU)
    f1   global order statistics (min, max, the four quartile neighbours per column) by radix select over
         the resident rows, the 256-bin histograms all-reduced per pass (sharding.order_stats_sharded)
         -> the bin plan of UC:664-685, bit for bit what calculator.bin_plan computes on the host
    K3   one batched histogram launch over the local frames -> (F/G, nd, nc) int64
         all-gather of the counts by frame -> (F, nd, nc) on every rank; a / a.sum() in float64
    chi2 row blocks of the (F, F) distance matrix split over the ranks, all-gathered

Nothing is written to or parsed from text; `calculator.make_histograms` on the same rows gives the
same histograms (tests/test_gpu_parity.py::test_trajectory_pipeline_matches_make_histograms).
"""
from __future__ import annotations

import time

import numpy as np

from . import sharding
from .calculator import plan_from_order_stats, quartile_ranks


def device_bin_plan(engine, rows, n_ref=None):
    """Bin plan (d_range, c_range, nd, nc) of UC:640-685 from device-resident float32 rows (F_local, L, 2)
    spread over the ranks of the default process group; exact (radix select), no host sort."""
    import torch

    flat = rows.reshape(-1, 2)
    n_local = torch.tensor([flat.shape[0]], dtype=torch.int64)
    n_all = int(sharding.all_reduce_(n_local, "sum")[0])
    if n_ref is None:
        n_ref = rows.shape[1]
    (p25, n25, g25), (p75, n75, g75) = quartile_ranks(n_all)
    ranks = [0, p25, n25, p75, n75, n_all - 1]
    stats = []
    for col in (0, 1):
        vals = sharding.order_stats_sharded(
            lambda pre, bits, col=col: engine.radix_hist(flat, col, pre, bits), ranks)
        stats.append((tuple(vals.tolist()), (g25, g75)))
    return plan_from_order_stats(stats, n_ref)


def topology_trajectory(engine, n_frames, frame_charges, seeds, n_iter, step_size, dimensions,
                        second_diff=False, distance_matrix=True, rows_out=None):
    """Run the trajectory.  `frame_charges(f)` -> (x, Q) of global frame f (CUDA tensors or arrays);
    `n_iter`: (L,) for every frame or a callable f -> (L,).  Returns a dict with
    plan, counts (F, nd, nc) int64, hists (F, nd*nc) float64, distance (F, F) float64 (or None), the
    local rows (F_local, L, 2), the frame ids of this rank, and phase timings in seconds."""
    import torch

    rank, size = sharding.world()
    mine = sharding.frames_for_rank(n_frames, rank, size)
    dev = engine.device
    seeds_d = seeds if torch.is_tensor(seeds) else torch.from_numpy(np.ascontiguousarray(seeds, np.float32)).to(dev)
    L = int(seeds_d.shape[0])
    rows = rows_out if rows_out is not None else torch.empty((len(mine), L, 2), dtype=torch.float32, device=dev)
    fixed_n_iter = None
    if torch.is_tensor(n_iter):
        fixed_n_iter = n_iter.to(device=dev, dtype=torch.int32)
    elif not callable(n_iter):
        fixed_n_iter = torch.as_tensor(np.asarray(n_iter).astype(np.int32)).to(dev)
    t = {}

    def tick():
        torch.cuda.synchronize(dev) if dev.type == "cuda" else None
        return time.perf_counter()

    t0 = tick()
    steps = torch.empty(L, dtype=torch.int32, device=dev)
    field_evals = torch.zeros((), dtype=torch.int64, device=dev)      # sum over lines of K + 2, times M below
    for i, f in enumerate(mine):
        x, Q = frame_charges(f)
        engine.set_charges(x, Q)
        ni = fixed_n_iter if fixed_n_iter is not None else torch.as_tensor(np.asarray(n_iter(f)).astype(np.int32)).to(dev)
        engine.topo_batch(seeds_d, ni, step_size, dimensions, second_diff=second_diff, out=rows[i], steps=steps)
        field_evals += (steps.sum(dtype=torch.int64) + 2 * L) * int(engine.n_charges)
    t1 = tick()
    pairs = int(field_evals.item())
    t["streamlines_s"] = t1 - t0
    plan = device_bin_plan(engine, rows, n_ref=L)
    d_range, c_range, nd, nc = plan
    if nd < 1 or nc < 1 or nd * nc > 50_000_000:
        raise ValueError(f"bin plan of UC:664-685 gives {nd} x {nc} bins (ranges {d_range}, {c_range}): the inter-quartile "
                         "range of a column is degenerate for these lines")
    t2 = tick()
    t["bin_plan_s"] = t2 - t1
    de = np.linspace(d_range[0], d_range[1], nd + 1)
    ce = np.linspace(c_range[0], c_range[1], nc + 1)
    local = engine.hist2d(rows, de, ce) if len(mine) else torch.zeros((0, nd, nc), dtype=torch.int64, device=dev)
    counts_by_rank = [len(sharding.frames_for_rank(n_frames, r, size)) for r in range(size)]
    gathered = sharding.all_gather_blocks(local, counts_by_rank)
    counts = torch.empty_like(gathered)
    lo = 0
    for r, n in enumerate(counts_by_rank):
        counts[r::size] = gathered[lo:lo + n]
        lo += n
    a = counts.reshape(n_frames, -1).to(torch.float64)
    hists = a / a.sum(dim=1, keepdim=True)                       # UC:709-713
    t3 = tick()
    t["histograms_s"] = t3 - t2
    distance = None
    if distance_matrix:
        lo_r, hi_r = sharding.slab(n_frames, rank, size)
        block = engine.chi2_rows(hists, lo_r, hi_r - lo_r)
        distance = sharding.all_gather_blocks(
            block, [sharding.slab(n_frames, r, size)[1] - sharding.slab(n_frames, r, size)[0] for r in range(size)])
    t4 = tick()
    t["distance_s"] = t4 - t3
    t["total_s"] = t4 - t0
    return {"plan": plan, "counts": counts, "hists": hists, "distance": distance, "rows": rows, "frames": mine,
            "timings": t, "pair_evals": pairs}
