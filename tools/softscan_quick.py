#!/usr/bin/env python
"""Lattice `volume` kernel with and without the softening scan (kernel time, best of 3)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis, half in [(100_000, 100, 1.5), (7890, 101, 0.5), (7890, 41, 0.5)]:
    x, Q = synth.charges(m, seed=1, box=half)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    ax = torch.linspace(-half, half, n_axis, device="cuda")
    pairs = float(n_axis) ** 3 * len(Q)
    ref = None
    for cfg in [dict(k1_softscan=0), dict(k1_softscan=1), dict(k1_softscan=-1), dict(k1_softscan=1, k1_points=4),
                dict(k1_softscan=0, k1_points=4)]:
        eng.set_tuning(k1_points=0, k1_softscan=-1); eng.set_tuning(**cfg)
        best = 1e30; wall = 1e30
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.field_lattice(ax, ax, ax, soften=True); e1.record(); torch.cuda.synchronize()
            best = min(best, eng.last_kernel_ms()); wall = min(wall, e0.elapsed_time(e1))
        if ref is None: ref = out.clone()
        print(json.dumps(dict(M=len(Q), n=n_axis, cfg=cfg, kernel_ms=round(best, 3), call_ms=round(wall, 3),
                              pairs_per_s="%.3e" % (pairs / (best * 1e-3)), pairs_per_s_call="%.3e" % (pairs / (wall * 1e-3)),
                              identical=bool(torch.equal(out, ref)))), flush=True)
