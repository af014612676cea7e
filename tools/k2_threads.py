#!/usr/bin/env python
"""K2 timing vs threads per CTA (needs a build with -DCPET_K2_MAXT>=the largest value tried)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
threads = [int(v) for v in sys.argv[1:]] or [512]
for m, n_axis, h in [(7890, 47, 0.1), (7890, 100, 0.1), (30_000, 47, 0.1), (7890, 47, 0.01)]:
    x, Q = synth.charges(m, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, h)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    for T in threads:
        for G in (0, 1, 2, 4):
            eng.set_tuning(k2_points=0, k2_lanes=G, k2_threads=T)
            best = 1e30
            try:
                for _ in range(3):
                    eng.topo_batch(sd, ni, h, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
            except Exception as e:
                print("T", T, "G", G, "failed:", str(e)[:80]); continue
            c = eng.last_counters()
            print(json.dumps(dict(M=len(Q), L=len(seeds), h=h, T=T, G=G, ms=round(best, 3),
                                  pairs_per_s="%.3e" % (c["pair_evals"] / (best * 1e-3)))), flush=True)
