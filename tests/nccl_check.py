#!/usr/bin/env python
"""Multi-GPU check of pycpet_b200.sharding over NCCL (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/nccl_check.py

Every rank computes its shard with the CUDA kernels (Engine, device pointers), the results are
gathered / reduced with NCCL, and rank 0 compares them bit for bit with a single-GPU run of the same
inputs.  The same sharding functions are exercised on CPU over gloo in tests/test_sharding_gloo.py.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from pycpet_b200 import sharding  # noqa: E402
from pycpet_b200.device import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    # K1's launch heuristics depend on the shard size (lanes per point, charge splits), which changes
    # the order of the FP64 partial sums; pin them so sharded == unsharded bit for bit.  The streamline
    # kernel sums every line in the same order whatever the shard size.
    # Likewise the streamline kernel FORM follows the queue length (points-packed for long queues, charge-pair-
    # packed for short ones); within a form every line is summed in the same order whatever the shard size.
    # The hybrid general field kernel classifies the charges against the bounding box of ITS point list, so a slab and the
    # whole list differ in the last bits: the bit-for-bit check runs on the direct-form kernel.
    eng.set_tuning(k1_lanes=8, k1_splits=1, k2_form=3, k1_hybrid=0)
    x, Q = synth.charges(7890, seed=1, box=0.5)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())

    # grid points by slabs -> all_gather
    pts = synth.grid(41, 0.5)
    field = sharding.grid_sharded(lambda p: eng.field_grid(torch.from_numpy(np.ascontiguousarray(p)).cuda(),
                                                           soften=True), pts)
    # one frame's streamlines by an LPT deal -> all_gather + restore seed order
    seeds, n_iter, dims, _ = synth.seeds(30, 0.5, 0.1)
    topo = sharding.topo_sharded(
        lambda s, n: eng.topo_batch(torch.from_numpy(np.ascontiguousarray(s)).cuda(), n, 0.1, dims), seeds, n_iter)
    # histogram of the lines held by this rank -> all_reduce(sum); global ranges -> all_reduce(min/max)
    ids = sharding.deal_lines(n_iter, rank, world)
    mine = topo[torch.as_tensor(ids, device=topo.device)]
    lo_d, hi_d, lo_c, hi_c = sharding.global_ranges(mine)
    de, ce = np.linspace(lo_d, hi_d, 41), np.linspace(lo_c, hi_c, 31)
    counts = sharding.hist_sharded(lambda v, a, b: eng.hist2d(v, a, b), mine, de, ce)
    # MD frames round-robin -> gather by frame id
    def frame(f):
        rng = np.random.default_rng(100 + f)
        xf = (x + rng.normal(0, 0.3, x.shape)).astype(np.float32)
        xf[np.all(np.abs(xf) < 0.55, axis=1)] *= 3.0
        eng.set_charges(torch.from_numpy(xf).cuda(), torch.from_numpy(Q).cuda())
        rows = eng.topo_batch(torch.from_numpy(seeds[:4096]).cuda(), n_iter[:4096], 0.1, dims)
        return eng.hist2d(rows, np.linspace(0, 1.8, 51), np.linspace(0, 5, 51)).clone()
    frames = sharding.frames_sharded(frame, 6)
    # the same frames through the MD-frame batch call (host C ABI, one call per rank) -> gather
    def frame_charges(f):
        rng = np.random.default_rng(100 + f)
        xf = (x + rng.normal(0, 0.3, x.shape)).astype(np.float32)
        xf[np.all(np.abs(xf) < 0.55, axis=1)] *= 3.0
        return xf, Q
    from pycpet_b200 import Math_ops
    m = Math_ops(device=local)
    batch = sharding.frames_batch_sharded(
        lambda ids: m.topo_hist_frames([frame_charges(f) for f in ids], seeds[:4096], n_iter[:4096],
                                       np.linspace(0, 1.8, 51), np.linspace(0, 5, 51), step_size=0.1,
                                       dimensions=dims)[1], 6)
    # a box mesh by slabs of x-planes (lattice kernel per slab) -> all_gather
    ax = torch.linspace(-0.5, 0.5, 43, device="cuda")
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    esp = sharding.lattice_sharded(lambda xs, ys, zs: eng.esp_lattice(xs, ys, zs, concat_half=True), ax, ax, ax)
    # exact global order statistics of the lines spread over the ranks (radix select, all-reduce per pass)
    n_all = len(seeds)
    want_ranks = [0, n_all // 4, n_all // 2, (3 * n_all) // 4, n_all - 1]
    stats = [sharding.order_stats_sharded(lambda pre, bits, c=c: eng.radix_hist(mine, c, pre, bits), want_ranks)
             for c in (0, 1)]
    # the device-resident trajectory pipeline: 11 frames over the ranks, global bin plan, gathered counts,
    # chi^2 row blocks
    from pycpet_b200 import trajectory
    def traj_frame(f):
        xf, q = frame_charges(f)
        return torch.from_numpy(xf).cuda(), torch.from_numpy(q).cuda()
    tr = trajectory.topology_trajectory(eng, 11, traj_frame, seeds[:4096], n_iter[:4096], 0.1, dims)
    torch.cuda.synchronize()

    ok = True
    if rank == 0:
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        f1 = eng.field_grid(torch.from_numpy(pts).cuda(), soften=True)
        t1 = eng.topo_batch(torch.from_numpy(seeds).cuda(), n_iter, 0.1, dims)
        c1 = eng.hist2d(t1, de, ce)
        fr1 = torch.stack([frame(f) for f in range(6)])
        checks = {
            "field all_gather": torch.equal(field, f1),
            "topo gather + seed order": torch.equal(topo, t1),
            "global ranges": (lo_d, hi_d, lo_c, hi_c) == (float(t1[:, 0].min()), float(t1[:, 0].max()),
                                                         float(t1[:, 1].min()), float(t1[:, 1].max())),
            "hist all_reduce": torch.equal(counts.to(c1.device), c1) and int(counts.sum()) == len(seeds),
            "frames gather": torch.equal(frames.to(fr1.device).to(fr1.dtype), fr1),
            "frames batch call + gather": torch.equal(batch.to(fr1.device).to(fr1.dtype).reshape(fr1.shape), fr1),
        }
        eng.set_tuning(k1_lanes=0, k1_splits=0, k1_hybrid=-1)
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        esp1 = eng.esp_lattice(ax, ax, ax, concat_half=True)
        d = (esp1[:, 3].float() - esp[:, 3].float()).abs().max() / esp1[:, 3].float().abs().max()
        checks["lattice slabs all_gather (coordinates exact, phi within one float16 ulp)"] = (
            torch.equal(esp1[:, :3], esp[:, :3]) and float(d) <= 1.5e-3)
        srt = [torch.sort(t1[:, c]).values.cpu().numpy() for c in (0, 1)]
        checks["order statistics over ranks"] = all(
            np.array_equal(stats[c], srt[c][want_ranks]) for c in (0, 1))
        # the same trajectory on one GPU, through the host-side plan
        from pycpet_b200 import calculator as calc
        rows1 = []
        for f in range(11):
            eng.set_charges(*traj_frame(f))
            rows1.append(eng.topo_batch(torch.from_numpy(seeds[:4096]).cuda(), n_iter[:4096], 0.1, dims).cpu().numpy())
        plan1 = calc.bin_plan([r.astype(np.float64) for r in rows1])
        h1 = calc.make_histograms_from_arrays(rows1, plan=plan1)
        checks["trajectory: bin plan == host plan"] = tr["plan"] == plan1
        checks["trajectory: histograms"] = np.array_equal(tr["hists"].cpu().numpy(), h1)
        checks["trajectory: chi2 matrix"] = np.array_equal(tr["distance"].cpu().numpy(), eng.chi2_rows(torch.from_numpy(h1).cuda()).cpu().numpy())
        for k, v in checks.items():
            print(f"[nccl_check world={world}] {k}: {'OK' if v else 'MISMATCH'}")
            ok = ok and v
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
