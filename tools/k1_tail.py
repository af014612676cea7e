#!/usr/bin/env python
"""Wave-quantisation check for K1: throughput vs number of CTAs and charge splits."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine


def main():
    eng = Engine(0); eng.set_tuning(timing=1)
    rng = np.random.default_rng(0)
    for m in (7890, 100_000):
        x, Q = synth.charges(m, seed=1, box=1.5)
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        for n in (296 * 3 * 1024, 1_000_000, 1_030_301, 296 * 4 * 1024):
            pts = torch.from_numpy(rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)).cuda()
            for splits in (1, 2, 4, 8, 16):
                if m < 50_000 and splits > 2:
                    continue
                eng.set_tuning(k1_points=4, k1_lanes=1, k1_splits=splits)
                best = 1e30
                for _ in range(3):
                    eng.field_grid(pts, soften=False); torch.cuda.synchronize()
                    best = min(best, eng.last_kernel_ms())
                print(json.dumps(dict(M=len(Q), N=n, ctas=(n + 1023) // 1024, splits=splits, ms=round(best, 3),
                                      pairs_per_s=float(n) * len(Q) / (best * 1e-3))), flush=True)


if __name__ == "__main__":
    main()
