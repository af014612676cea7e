import os, sys, json
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
eng = Engine(0); eng.set_tuning(timing=1, k1_lattice=0)
x, Q = synth.charges(7890, seed=1, box=0.5)
eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
for n_axis in (5, 9, 11, 12, 13, 21, 41, 61):
    pts = torch.from_numpy(synth.grid(n_axis, 0.5)).cuda()
    row = dict(N=n_axis ** 3)
    for name, cfg in [("auto", dict()), ("g1p1", dict(k1_lanes=1, k1_points=1)), ("g8", dict(k1_lanes=8)), ("g32", dict(k1_lanes=32))]:
        eng.set_tuning(k1_lanes=0, k1_points=0); eng.set_tuning(**cfg)
        for mode in ("soft", "esp"):
            fn = (lambda: eng.field_grid(pts, soften=True)) if mode == "soft" else (lambda: eng.esp_grid(pts))
            best = 1e30
            for _ in range(4):
                fn(); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
            row[f"{mode}_{name}"] = round(best * 1e3, 1)
    print(json.dumps(row), flush=True)
