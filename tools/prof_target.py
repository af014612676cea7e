#!/usr/bin/env python
"""Small driver for ncu: a few launches of each hot kernel with the default heuristics."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from pycpet_b200.device import Engine  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
x, Q = synth.charges(7890, seed=1, box=0.5)
eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
if which in ("all", "k1"):
    pts = torch.from_numpy(synth.grid(101, 0.5)).cuda()
    ax = torch.linspace(-0.5, 0.5, 101, device="cuda")
    eng.set_tuning(k1_lattice=0)                      # general kernels on the flat point list
    for _ in range(reps):
        eng.field_grid(pts, soften=False)
        eng.field_grid(pts, soften=True)
        eng.esp_grid(pts)
    eng.set_tuning(k1_lattice=-1)
    for _ in range(reps):
        eng.field_lattice(ax, ax, ax, soften=True)    # softening scan -> unsoftened instantiation serves the call
        eng.set_tuning(k1_softscan=0)
        eng.field_lattice(ax, ax, ax, soften=True)    # softened instantiation
        eng.set_tuning(k1_softscan=-1)
if which in ("all", "k2"):
    seeds, n_iter, dims, _ = synth.seeds(47, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda()
    ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    for _ in range(reps):
        out = eng.topo_batch(sd, ni, 0.1, dims)
    de, ce = np.linspace(0, 1.7, 51), np.linspace(0, 5, 51)
    for _ in range(reps):
        eng.hist2d(out, de, ce)
torch.cuda.synchronize()
print("done")
