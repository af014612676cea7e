#!/usr/bin/env python
"""K2: lines per thread (P) x lanes per line (G) at a given threads/CTA (arg 1)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for m, n_axis, h in [(7890, 47, 0.1), (7890, 100, 0.1)]:
    x, Q = synth.charges(m, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, h)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    for P in (1, 2):
        for G in (1, 2, 4):
            eng.set_tuning(k2_points=P, k2_lanes=G, k2_threads=T)
            best = 1e30
            for _ in range(3):
                eng.topo_batch(sd, ni, h, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
            c = eng.last_counters()
            print(json.dumps(dict(L=len(seeds), T=T, P=P, G=G, ms=round(best, 3),
                                  pairs_per_s="%.3e" % (c["pair_evals"] / (best * 1e-3)))), flush=True)
