#!/usr/bin/env python
"""Points-packed streamline kernel: far-loop unroll (k2_unroll) on the 3A and 1M-line frames, kernel time and error against
the float64 oracle on 512 sampled lines.  Rebuild with CPET_NVCC_EXTRA=-DCPET_K2P_CHUNK=128|256|512 to vary the FP32 chain
length (profiles/round2_k2p_tuning.txt)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
from oracle import f64
eng = Engine(0); eng.set_tuning(timing=1)
for m, n_axis in ((7890, 47), (7890, 100)):
    x, Q = synth.charges(m, seed=1, box=0.5)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    sample = np.random.default_rng(0).choice(len(seeds), 512, replace=False)
    ref, _ = f64.topo_batch(seeds[sample], n_iter[sample], x, Q, 0.1, dims)
    for u in (6, 8, 12):
        eng.set_tuning(k2_form=3, k2_unroll=u)
        best = 1e30
        for _ in range(7):
            out = eng.topo_batch(sd, ni, 0.1, dims); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        c = eng.last_counters()
        o = out.cpu().numpy()[sample]
        print(json.dumps(dict(L=len(seeds), unroll=u, ms=round(best, 4), frac=round(c["pair_evals"] * 20 / (best * 1e-3) / 74.45e12, 4),
                              dist_err=float(np.abs(o[:, 0] - ref[:, 0]).max()), curv_err=float(np.abs(o[:, 1] - ref[:, 1]).max()))), flush=True)
