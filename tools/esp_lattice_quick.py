#!/usr/bin/env python
"""ESP on a box mesh: general kernel on the flat point list against the lattice kernel (kernel time)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis, half in [(100_000, 101, 5.0), (7890, 101, 0.5), (7890, 41, 0.5)]:
    x, Q = synth.charges(m, seed=1, box=half)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    ax = torch.linspace(-half, half, n_axis, device="cuda")
    pts = torch.from_numpy(synth.grid(n_axis, half)).cuda()
    pairs = float(n_axis) ** 3 * len(Q)
    ref = None
    for name, fn, cfg in [("general", lambda: eng.esp_grid(pts), dict()),
                          ("lattice pz5", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_points=5)),
                          ("lattice pz4", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_points=4)),
                          ("lattice pz2", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_points=2)),
                          ("lattice pz5 u2", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_points=5, k1_unroll=2)),
                          ("lattice pz5 splits1", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_points=5, k1_splits=1)),
                          ("lattice pz6, 1 node in 6 with its rsqrt on the FMA pipe", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_esp_mix=1)),
                          ("same, unroll 2", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_esp_mix=1, k1_unroll=2)),
                          ("default heuristics", lambda: eng.esp_lattice(ax, ax, ax), dict(k1_esp_mix=-1))]:
        eng.set_tuning(k1_points=0, k1_unroll=0, k1_splits=0, k1_esp_mix=0); eng.set_tuning(**cfg)
        best = 1e30
        for _ in range(3):
            out = fn(); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        if ref is None: ref = out.clone()
        err = float((out.double() - ref.double()).abs().max() / ref.double().abs().max())
        if name.startswith("default"):      # against float64 on a sample of nodes
            from oracle import f64
            idx = np.random.default_rng(0).choice(n_axis ** 3, 2048, replace=False)
            p_h = pts[torch.from_numpy(idx).cuda()].cpu().numpy()
            want = f64.esp_grid(p_h, x, Q)
            got = out[torch.from_numpy(idx).cuda()].cpu().numpy()
            print(json.dumps(dict(default_vs_float64_maxrel=float(np.abs(got - want).max() / np.abs(want).max()))), flush=True)
        print(json.dumps(dict(M=len(Q), n=n_axis, kernel=name, ms=round(best, 3), pairs_per_s="%.3e" % (pairs / (best * 1e-3)),
                              maxrel_vs_general=err)), flush=True)
