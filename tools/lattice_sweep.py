#!/usr/bin/env python
"""Lattice kernel variants (z-nodes per thread, unroll, tile/stages, splits) on 100^3 / 101^3 meshes."""
import os, sys, json, itertools
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis in [(100_000, 100), (7890, 101)]:
    x, Q = synth.charges(m, seed=1, box=1.5)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    ax = torch.linspace(-1.5, 1.5, n_axis, device="cuda")
    pairs = float(n_axis) ** 3 * len(Q)
    for pz, u, tile, st, sp in itertools.product((2, 4, 5), (1, 2, 4), (512, 1024), (2, 3), (0, 1)):
        if (tile, st) not in ((1024, 3), (512, 2)):
            continue
        eng.set_tuning(k1_points=pz, k1_unroll=u, k1_tile_pairs=tile, k1_stages=st, k1_splits=sp)
        best = 1e30
        for _ in range(3):
            eng.field_lattice(ax, ax, ax, soften=True); torch.cuda.synchronize()
            best = min(best, eng.last_kernel_ms())
        print(json.dumps(dict(M=len(Q), n=n_axis, pz=pz, unroll=u, tile=tile, stages=st, splits=sp, ms=round(best, 3),
                              pairs_per_s="%.3e" % (pairs / (best * 1e-3)))), flush=True)
