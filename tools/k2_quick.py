#!/usr/bin/env python
"""Quick K2 timing of the bench cases with the default heuristics and a few overrides.  Kernel time only (CUDA events around the integrator launch), best of 3."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
CASES = [(7890, 47, 0.1), (7890, 100, 0.1), (7890, 18, 0.1), (30_000, 47, 0.1), (7890, 47, 0.01), (1000, 47, 0.1),
         (100_000, 30, 0.1)]
CFGS = [dict(), dict(k2_cap=4), dict(k2_cap=2), dict(k2_cap=1), dict(k2_sort=0)]
if len(sys.argv) > 1:
    CFGS = [json.loads(a) for a in sys.argv[1:]]
ref = {}
for m, n_axis, h in CASES:
    x, Q = synth.charges(m, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, h)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    base = None
    for cfg in CFGS:
        eng.set_tuning(k2_threads=0, k2_cap=0, k2_sort=-1)
        eng.set_tuning(**cfg)
        best = 1e30
        for _ in range(3):
            out = eng.topo_batch(sd, ni, h, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        c = eng.last_counters()
        out = out[0] if isinstance(out, tuple) else out
        if base is None:
            base = out.clone()
        dmax = float((out - base).abs().nan_to_num().max())
        print(json.dumps(dict(M=len(Q), L=len(seeds), h=h, cfg=cfg, ms=round(best, 3),
                              pairs_per_s="%.3e" % (c["pair_evals"] / (best * 1e-3)), maxdiff_vs_first=dmax)), flush=True)
# small line sets (shipped example 3A: 18^3; a 6^3 probe): one line per warp spread over all SMs
for n_axis in (18, 6, 3):
    x, Q = synth.charges(7890, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    for cfg in [dict(), dict(k2_threads=512)]:
        eng.set_tuning(k2_threads=0, k2_cap=0, k2_sort=-1); eng.set_tuning(**cfg)
        best = 1e30
        for _ in range(3):
            eng.topo_batch(sd, ni, 0.1, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        print(json.dumps(dict(M=len(Q), L=len(seeds), cfg=cfg, ms=round(best, 4))), flush=True)
