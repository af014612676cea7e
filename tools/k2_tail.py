#!/usr/bin/env python
"""End-of-queue policy of the points-packed streamline kernel on the 3A frame (58 lines per warp of the chip):
from how many lines per warp left in the queue a warp tops up to 4 (tail4) and to 2 (tail2) lines only."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis in ((7890, 47), (7890, 60), (7890, 38)):
    x, Q = synth.charges(m, seed=1, box=0.5)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    for t4, t2 in ((4, 0), (0, 0), (8, 0), (2, 0), (4, 2), (4, 1), (6, 2), (8, 2), (8, 4), (6, 3), (12, 4)):
        eng.set_tuning(k2_form=3, k2_tail4=t4, k2_tail2=t2)
        best = 1e30
        for _ in range(7):
            eng.topo_batch(sd, ni, 0.1, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        c = eng.last_counters()
        print(json.dumps(dict(L=len(seeds), tail4=t4, tail2=t2, ms=round(best, 4),
                              frac_nominal=round(c["pair_evals"] * 20 / (best * 1e-3) / 74.45e12, 4))), flush=True)
