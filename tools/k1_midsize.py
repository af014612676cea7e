#!/usr/bin/env python
"""General K1 kernel on mid-size point lists: lanes per point (G) and points per thread (P)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine


def main():
    eng = Engine(0); eng.set_tuning(timing=1, k1_lattice=0)
    for m in (7890, 100_000):
        x, Q = synth.charges(m, seed=1, box=0.5)
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        for n_axis in (13, 17, 21, 27, 31, 41, 47, 61):
            pts = torch.from_numpy(synth.grid(n_axis, 0.5)).cuda()
            row = dict(M=m, N=n_axis ** 3)
            for mode in ("raw", "esp"):
                for name, cfg in [("auto", dict()), ("g1p1", dict(k1_lanes=1, k1_points=1)), ("g1p2", dict(k1_lanes=1, k1_points=2)),
                                  ("g1p4", dict(k1_lanes=1, k1_points=4)), ("g8", dict(k1_lanes=8)), ("g32", dict(k1_lanes=32))]:
                    eng.set_tuning(k1_lanes=0, k1_points=0); eng.set_tuning(**cfg)
                    fn = (lambda: eng.field_grid(pts, soften=False)) if mode == "raw" else (lambda: eng.esp_grid(pts))
                    best = 1e30
                    for _ in range(3):
                        fn(); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
                    row[f"{mode}_{name}"] = round(best * 1e3, 1)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
