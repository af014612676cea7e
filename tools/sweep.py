#!/usr/bin/env python
"""Exploratory kernel sweep on one GPU (not the bench): K1 / K2 throughput over the tuning knobs,
timed with CUDA events inside the library (cpet_last_kernel_ms).  Writes gpurun_out/sweep.json."""
import itertools
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from pycpet_b200.device import Engine  # noqa: E402

PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12


def timed(eng, fn, reps=3):
    best = 1e30
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
        best = min(best, eng.last_kernel_ms())
    return best


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    eng = Engine(0)
    eng.set_tuning(timing=1)
    out = {"gpu": torch.cuda.get_device_name(0)}
    out["fp32_probe_tflops"] = {"ffma": eng.fp32_peak_tflops(False), "ffma2": eng.fp32_peak_tflops(True)}
    print(out, flush=True)
    res = []

    if which in ("all", "k1"):
        for m, n_axis, half in [(7890, 101, 5.0), (100_000, 101, 5.0)]:
            x, Q = synth.charges(m, seed=1, box=half)
            pts = torch.from_numpy(synth.grid(n_axis, half)).cuda()
            eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
            pairs = float(len(pts)) * len(Q)
            for mode in ("soft", "raw", "esp"):
                for P, thr, tile, stages in itertools.product((1, 2, 4), (128, 256), (512, 1024, 2048), (2, 3)):
                    if m < 50_000 and (tile != 1024 or stages != 3 or thr != 256):
                        if not (tile == 1024 and stages == 3):
                            continue
                    eng.set_tuning(k1_points=P, k1_lanes=1, k1_threads=thr, k1_tile_pairs=tile, k1_stages=stages)
                    if mode == "esp":
                        ms = timed(eng, lambda: eng.esp_grid(pts))
                    else:
                        ms = timed(eng, lambda: eng.field_grid(pts, soften=(mode == "soft")))
                    rate = pairs / (ms * 1e-3)
                    rec = dict(kernel="k1", mode=mode, M=len(Q), N=len(pts), P=P, threads=thr, tile=tile,
                               stages=stages, ms=ms, pairs_per_s=rate,
                               frac_nominal_fp32=rate * 20 / 1e12 / PEAK_NOMINAL)
                    res.append(rec)
                    print(json.dumps(rec), flush=True)
        eng.set_tuning(k1_points=0, k1_lanes=0, k1_threads=0, k1_tile_pairs=0, k1_stages=0)

    if which in ("all", "k2"):
        for m, n_axis, h in [(7890, 47, 0.1), (7890, 100, 0.1), (7890, 18, 0.1), (30_000, 47, 0.1), (7890, 47, 0.01)]:
            x, Q = synth.charges(m, seed=1, box=0.5)
            seeds, n_iter, dims, max_steps = synth.seeds(n_axis, 0.5, h)
            sd = torch.from_numpy(seeds).cuda()
            ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
            eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
            for cap, thr, srt in itertools.product((0, 1, 2, 4), (256, 384, 512), (1, 0)):
                if srt == 0 and (cap != 4 or thr != 512):     # unsorted queue: one reference point per case
                    continue
                eng.set_tuning(k2_cap=cap, k2_threads=thr, k2_sort=(-1 if cap == 0 else srt))
                ms = timed(eng, lambda: eng.topo_batch(sd, ni, h, dims), reps=2)
                c = eng.last_counters()
                rate = c["pair_evals"] / (ms * 1e-3)
                rec = dict(kernel="k2", M=len(Q), L=len(seeds), h=h, cap=cap, threads=thr, sort=srt, ms=ms,
                           field_evals=c["field_evals"], pairs_per_s=rate, lines_per_s=len(seeds) / (ms * 1e-3),
                           frac_nominal_fp32=rate * 20 / 1e12 / PEAK_NOMINAL)
                res.append(rec)
                print(json.dumps(rec), flush=True)
        eng.set_tuning(k2_cap=0, k2_threads=0, k2_sort=-1)

    out["results"] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"sweep_{which}.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
