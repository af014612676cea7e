#!/usr/bin/env python
"""Mid-size box meshes: z-nodes per thread of the lattice kernel against the general kernel (kernel ms)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
x, Q = synth.charges(7890, seed=1, box=0.5)
eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
for n_axis in (17, 21, 31, 41, 51, 61, 81):
    ax = torch.linspace(-0.5, 0.5, n_axis, device="cuda")
    pts = torch.from_numpy(synth.grid(n_axis, 0.5)).cuda()
    for mode in ("field", "esp"):
        row = dict(n=n_axis, mode=mode)
        for name, cfg in [("general", None), ("auto", dict()), ("pz2", dict(k1_points=2)), ("pz4", dict(k1_points=4)), ("pz5", dict(k1_points=5))]:
            eng.set_tuning(k1_points=0, k1_lattice=-1)
            if cfg is None:
                eng.set_tuning(k1_lattice=0)
                fn = (lambda: eng.field_grid(pts, soften=True)) if mode == "field" else (lambda: eng.esp_grid(pts))
            else:
                eng.set_tuning(**cfg)
                fn = (lambda: eng.field_lattice(ax, ax, ax, soften=True)) if mode == "field" else (lambda: eng.esp_lattice(ax, ax, ax))
            best = 1e30
            for _ in range(4):
                fn(); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
            row[name] = round(best * 1e3, 1)
        print(json.dumps(row), flush=True)
