import os, sys, json
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis in [(30_000, 47), (100_000, 30), (100_000, 47)]:
    x, Q = synth.charges(m, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    for cfg in [dict(), dict(k2_tile_pairs=1024, k2_stages=6), dict(k2_tile_pairs=2048, k2_stages=3), dict(k2_tile_pairs=3072, k2_stages=2), dict(k2_tile_pairs=1024, k2_stages=3), dict(k2_tile_pairs=512, k2_stages=8)]:
        eng.set_tuning(k2_tile_pairs=0, k2_stages=0); eng.set_tuning(**cfg)
        best = 1e30
        for _ in range(3):
            eng.topo_batch(sd, ni, 0.1, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        c = eng.last_counters()
        print(json.dumps(dict(M=len(Q), L=len(seeds), cfg=cfg, ms=round(best, 3), pairs_per_s="%.3e" % (c["pair_evals"] / (best * 1e-3)))), flush=True)
