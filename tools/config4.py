#!/usr/bin/env python
"""BASELINE.json configs[3] executed for real: an MD-trajectory batch of F frames x L seeds per frame
sharded over the GPUs of one box (frames round-robin), device-resident from the charges of each frame to
the (F, F) chi^2 distance matrix (pycpet_b200.trajectory).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tools/config4.py [--frames 1000] [--axis 100] [--out gpurun_out/config4.json]

Reports wall-clock (barrier to barrier, max over ranks), pair-evals/s, per-rank busy fraction (sum of the
integrator launches' own durations / wall-clock of the streamline phase), the phase split, and checks:
  * 3 sampled frames: 2,048 sampled lines each against the float64 oracle (dist 2e-6, curvature 2e-5 + 2e-6/h)
  * the global order statistics behind the bin plan, verified exactly by counting (#values < s <= rank < #values <= s)
  * the gathered counts of the sampled frames == np.histogram2d of their rows, bit for bit
  * the distance matrix rows of the sampled frames == the float64 NumPy restatement of distance_numpy
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from pycpet_b200 import sharding, trajectory  # noqa: E402
from pycpet_b200.calculator import quartile_ranks  # noqa: E402
from pycpet_b200.device import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--axis", type=int, default=100)
    ap.add_argument("--charges", type=int, default=7890)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/config4.json")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(local)
    eng.set_tuning(timing=1)
    x, Q = synth.charges(args.charges, seed=1, box=0.5)
    seeds, n_iter, dims, max_steps = synth.seeds(args.axis, 0.5, 0.1)
    L = len(seeds)
    dq = torch.from_numpy(Q).to(dev)

    def frame_host(f):
        rng = np.random.default_rng(1000 + f)
        xf = (x + rng.normal(0.0, 0.3, x.shape)).astype(np.float32)
        inside = np.all(np.abs(xf) < 0.5 * 1.05, axis=1)
        xf[inside] *= np.float32(3.0)
        return xf

    # the trajectory is resident before the clock starts ("MD batch = base charge set + per-frame jitter")
    mine = sharding.frames_for_rank(args.frames, rank, world)
    resident = {f: torch.from_numpy(frame_host(f)).to(dev) for f in mine}
    dseeds = torch.from_numpy(seeds).to(dev)
    dnit = torch.from_numpy(n_iter.astype(np.int32)).to(dev)
    # warm-up: one small trajectory (kernels loaded, NCCL rings built) on a strided sample of the seeds -- the first
    # few thousand seeds of the mesh share one x-plane, whose distances have a degenerate inter-quartile range
    stride = max(1, L // 4096)
    trajectory.topology_trajectory(eng, world, lambda f: (resident[mine[0]], dq), dseeds[::stride].contiguous(),
                                   dnit[::stride].contiguous(), 0.1, dims)
    # the trajectory runs twice over the same resident frames: the first pass pays the one-off costs of this process
    # (cudaMalloc of the ~1 GB result buffers, NCCL's first large-message setup), the second is the steady state
    rows_buf = torch.empty((len(mine), L, 2), dtype=torch.float32, device=dev)
    walls, phases = [], []
    for rep in range(args.reps):
        eng.kernel_times()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        tr = trajectory.topology_trajectory(eng, args.frames, lambda f: (resident[f], dq), dseeds, dnit, 0.1, dims,
                                            rows_out=rows_buf)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        walls.append(time.perf_counter() - t0)
        phases.append(dict(tr["timings"]))
        if rep + 1 < args.reps:
            del tr
    wall = walls[-1]
    kt = eng.kernel_times()
    # the library keeps the last 256 timed launches, in launch order: the integrator launches of this rank's frames,
    # then the histogram launch and the chi^2 launch
    k2_ms = kt[:len(mine)] if len(kt) >= len(mine) + 2 else kt[:-2]
    busy = sum(k2_ms) * 1e-3 / max(tr["timings"]["streamlines_s"], 1e-9) * (len(mine) / max(1, len(k2_ms)))

    # ---- checks -------------------------------------------------------------------------------------
    from oracle import f64, hist as ohist
    checks = {}
    rows = tr["rows"]
    sample_frames = [0, args.frames // 2, args.frames - 1]
    d_range, c_range, nd, nc = tr["plan"]
    de, ce = np.linspace(d_range[0], d_range[1], nd + 1), np.linspace(c_range[0], c_range[1], nc + 1)
    worst = {"dist": 0.0, "curv": 0.0, "flips": 0}
    hist_ok, chi_ok, mine_checked = True, True, 0
    hists = tr["hists"].cpu().numpy() if any(f in mine for f in sample_frames) else None
    for f in sample_frames:
        if f not in mine:
            continue
        mine_checked += 1
        r = rows[mine.index(f)].cpu().numpy()
        idx = np.random.default_rng(f).choice(L, 2048, replace=False)
        want, wsteps = f64.topo_batch(seeds[idx], n_iter[idx], frame_host(f), Q, 0.1, dims)
        # a step-count flip shows as a distance off by about one step
        flips = np.abs(r[idx, 0] - want[:, 0]) > 0.05
        ok = ~flips & np.isfinite(want).all(axis=1)
        worst["flips"] += int(flips.sum())
        worst["dist"] = max(worst["dist"], float(np.abs(r[idx, 0] - want[:, 0])[ok].max()))
        worst["curv"] = max(worst["curv"], float(np.abs(r[idx, 1] - want[:, 1])[ok].max()))
        cnt = np.histogram2d(r[:, 0].astype(np.float64), r[:, 1].astype(np.float64), bins=[de, ce])[0].astype(np.int64)
        hist_ok = hist_ok and np.array_equal(cnt, tr["counts"][f].cpu().numpy())
        drow = np.array([ohist.chi2(hists[f], hists[g]) for g in range(0, args.frames, max(1, args.frames // 50))])
        got = tr["distance"][f].cpu().numpy()[::max(1, args.frames // 50)]
        chi_ok = chi_ok and np.allclose(got, drow, rtol=1e-12, atol=0)
    # order statistics, verified by counting over all ranks
    flat = rows.reshape(-1, 2)
    n_all = args.frames * L
    (p25, n25, g25), (p75, n75, g75) = quartile_ranks(n_all)
    ranks = [0, p25, n25, p75, n75, n_all - 1]
    stat_ok = True
    for col in (0, 1):
        vals = sharding.order_stats_sharded(lambda pre, bits, col=col: eng.radix_hist(flat, col, pre, bits), ranks)
        for s, k in zip(vals, ranks):
            below = torch.tensor([int((flat[:, col] < float(s)).sum()), int((flat[:, col] <= float(s)).sum())], dtype=torch.int64)
            below = sharding.all_reduce_(below, "sum")
            stat_ok = stat_ok and (int(below[0]) <= k < int(below[1]))
    stat_ok = stat_ok and d_range == (float(sharding.order_stats_sharded(lambda pre, bits: eng.radix_hist(flat, 0, pre, bits), [0])[0]),
                                      float(sharding.order_stats_sharded(lambda pre, bits: eng.radix_hist(flat, 0, pre, bits), [n_all - 1])[0]))
    flags = torch.tensor([worst["dist"], worst["curv"], float(worst["flips"]), 0.0 if hist_ok else 1.0,
                          0.0 if chi_ok else 1.0, 0.0 if stat_ok else 1.0, float(tr["pair_evals"]), float(mine_checked)],
                         dtype=torch.float64, device=dev)
    per_rank = torch.tensor([[busy, tr["timings"]["streamlines_s"], tr["timings"]["bin_plan_s"], tr["timings"]["histograms_s"],
                              tr["timings"]["distance_s"], float(len(mine))]], dtype=torch.float64, device=dev)
    allr = per_rank
    if world > 1:
        mx = flags.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = flags.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        allr = torch.empty((world, 6), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, per_rank)
    else:
        mx = sm = flags
    if rank == 0:
        pairs = float(sm[6])
        h = 0.1
        rec = {
            "config": f"BASELINE configs[3]: {args.frames} MD frames x {L} seeds ({args.axis}^3), {len(Q)} charges per frame, "
                      f"box 0.5 A, step 0.1 A, frames round-robin over {world} B200",
            "n_gpus": world, "frames": args.frames, "lines_per_frame": L, "charges": int(len(Q)),
            "wall_s": wall, "wall_s_every_pass": walls, "phases_s_rank0_every_pass": phases, "pair_evals": pairs, "pair_evals_per_s": pairs / wall,
            "streamlines_per_s": args.frames * L / wall,
            "fp32_frac_of_nominal_whole_job": pairs / wall * 20 / 1e12 / (world * 148 * 128 * 2 * 1.965e9 / 1e12),
            "plan": {"d_range": d_range, "c_range": c_range, "nd": nd, "nc": nc},
            "phases_s_rank0": tr["timings"],
            "per_rank": {"busy_fraction_streamline_phase": [round(float(v), 4) for v in allr[:, 0]],
                         "streamlines_s": [round(float(v), 4) for v in allr[:, 1]],
                         "bin_plan_s": [round(float(v), 4) for v in allr[:, 2]],
                         "histograms_s": [round(float(v), 4) for v in allr[:, 3]],
                         "distance_s": [round(float(v), 4) for v in allr[:, 4]],
                         "frames": [int(v) for v in allr[:, 5]]},
            "checks": {"sampled_frames": sample_frames, "frames_checked": int(sm[7]),
                       "lines_vs_oracle_max_abs_dist": float(mx[0]), "lines_vs_oracle_max_abs_curv": float(mx[1]),
                       "tolerances": {"dist": 2e-6, "curv": 2e-5 + 2e-6 / h}, "step_count_flips_of_6144": int(sm[2]),
                       "lines_ok": bool(mx[0] <= 2e-6 and mx[1] <= 2e-5 + 2e-6 / h and sm[2] <= 6),
                       "histogram_counts_bit_exact": bool(mx[3] == 0), "chi2_rows_match_float64_numpy": bool(mx[4] == 0),
                       "order_statistics_exact_by_counting": bool(mx[5] == 0)},
        }
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(rec, fh, indent=1)
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
