#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200 import Math_ops

m = Math_ops()
for M_ in (301, 20_000):                       # resident and streamed charge staging
    x, Q = synth.charges(M_, seed=1, box=0.5)
    m.set_charges(x, Q)
    pts = synth.grid(17, 0.5)                  # 4913 points: lattice path via auto-detection
    m.field_grid(pts, soften=True, concat=True)
    m.field_grid(pts[::-1][:1000].copy() + np.float32(0.001), soften=False)
    rp = (np.random.default_rng(2).uniform(-0.5, 0.5, (2500, 3))).astype(np.float32)     # >= 2,048 points: hybrid near/far general kernel
    m.field_grid(rp, soften=True, concat=True)
    m.set_tuning(k1_splits=5)
    m.field_grid(rp, soften=False)
    m.set_tuning(k1_splits=0)
    m.esp_grid(pts[:777], concat_half=True)
    m.propagate(pts[:100], 0.1)
    for cfg in (dict(), dict(k1_lanes=8, k1_splits=3), dict(k1_lanes=32), dict(k1_points=2, k1_lanes=1, k1_tile_pairs=64, k1_stages=2)):
        m.set_tuning(k1_points=0, k1_lanes=0, k1_splits=0, k1_tile_pairs=0, k1_stages=0)
        m.set_tuning(**cfg)
        m.field_grid(pts[:333] * np.float32(0.9), soften=True)
    m.set_tuning(k1_points=0, k1_lanes=0, k1_splits=0, k1_tile_pairs=0, k1_stages=0)
    ax = np.linspace(-0.5, 0.5, 9, dtype=np.float32)
    m.set_tuning(k1_softscan=1)                # scan + both lattice instantiations (one exits at once)
    m.field_lattice(ax, ax, ax, soften=True)
    m.set_tuning(k1_lat_nodes=1)               # field lattice kernel with node pairs packed (+ scan, both instantiations)
    m.field_lattice(ax, ax, ax, soften=True)
    m.set_tuning(k1_points=6, k1_splits=2)
    m.field_lattice(ax, ax[:5], ax[:7], soften=False)
    m.set_tuning(k1_points=0, k1_splits=0, k1_lat_nodes=-1)
    m.set_tuning(k1_softscan=-1, k1_esp_mix=1)  # ESP lattice kernel, 6 z-nodes per thread, one rsqrt in six on the FMA pipe
    m.esp_lattice(ax, ax, ax, concat_half=True)
    m.set_tuning(k1_esp_mix=-1)
    seeds, n_iter, dims, _ = synth.seeds(6, 0.5, 0.1)
    for cfg in (dict(), dict(k2_cap=1), dict(k2_cap=2, k2_threads=128), dict(k2_cap=4, k2_tile_pairs=64, k2_stages=2),
                dict(k2_form=1), dict(k2_form=2, k2_cap=4), dict(k2_form=3), dict(k2_form=3, k2_cap=8, k2_threads=64),
                dict(k2_form=3, k2_cap=4, k2_tile_pairs=64, k2_stages=2), dict(k2_form=3, k2_cap=1)):
        m.set_tuning(k2_cap=0, k2_threads=0, k2_tile_pairs=0, k2_stages=0, k2_form=0)
        m.set_tuning(**cfg)
        rows, steps = m.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, want_steps=True)
        m.topo_batch(seeds, n_iter, step_size=0.1, dimensions=dims, second_diff=True)
    m.set_tuning(k2_cap=0, k2_threads=0, k2_tile_pairs=0, k2_stages=0, k2_form=0)
    de, ce = np.linspace(0, 1.8, 21), np.linspace(0, 5, 31)
    m.hist2d(rows, de, ce)
    m.topo_hist(seeds, n_iter, de, ce, step_size=0.1, dimensions=dims)
    m.hist2d(np.random.default_rng(0).random((3, 1000, 2)), np.linspace(0, 1, 301), np.linspace(0, 1, 401))
    m.order_stats(rows, [0, 10, len(rows) - 1], column=1)
frames = [synth.charges(n_f, seed=3 + n_f, box=0.5) for n_f in (500, 20_001, 77)]
m.topo_hist_frames(frames, seeds, n_iter, np.linspace(0, 1.8, 21), np.linspace(0, 5, 31), step_size=0.1,
                   dimensions=dims, want_rows=True)
H = np.random.default_rng(1).random((5, 100)); H /= H.sum(1, keepdims=True)
m.chi2_matrix(H)
m.calc_field(np.zeros(3, np.float32), x, Q); m.calc_esp_base(np.zeros(3, np.float32), x, Q)
m.thread_operation(seeds[0], 5, x, Q, 0.1, dims)
import torch
from pycpet_b200.device import Engine
eng = Engine(0)
eng.chi2_rows(torch.from_numpy(H).cuda(), 1, 3)
eng.radix_hist(torch.from_numpy(rows).cuda(), 1, np.array([0], np.uint32), 0)
torch.cuda.synchronize()
print("sanitize target done")
