#!/usr/bin/env python
"""Field lattice kernel: charge pairs packed (shipped) against node pairs packed (k1_lat_nodes=1), kernel time on the
`volume` frame (100^3 nodes x 100,000 charges) and a 464-node z axis, error against the float64 oracle on a sample."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
from oracle import f64
eng = Engine(0); eng.set_tuning(timing=1)
x, Q = synth.charges(100_000, seed=1, box=1.5)
eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
for shape in ((100, 100, 100), (48, 48, 464)):
    axes = [torch.linspace(-1.5, 1.5, n, device="cuda") for n in shape]
    n_pts = shape[0] * shape[1] * shape[2]
    idx = np.random.default_rng(0).choice(n_pts, 1024, replace=False)
    ii = np.unravel_index(idx, shape)
    pts = np.column_stack([axes[k].cpu().numpy()[ii[k]] for k in range(3)]).astype(np.float32)
    want = f64.field_grid(pts, x, Q, True)
    ref = None
    for name, cfg in (("charge pairs, default", dict()), ("nodes pz8 u4", dict(k1_lat_nodes=1, k1_points=8)),
                      ("nodes pz8 u2", dict(k1_lat_nodes=1, k1_points=8, k1_unroll=2)),
                      ("nodes pz6 u4", dict(k1_lat_nodes=1, k1_points=6)), ("nodes pz6 u2", dict(k1_lat_nodes=1, k1_points=6, k1_unroll=2)),
                      ("nodes pz4 u4", dict(k1_lat_nodes=1, k1_points=4)), ("nodes pz10 u4", dict(k1_lat_nodes=1, k1_points=10)),
                      ("nodes auto", dict(k1_lat_nodes=1))):
        eng.set_tuning(k1_lat_nodes=0, k1_points=0, k1_unroll=0)
        eng.set_tuning(**cfg)
        best = 1e30
        for _ in range(3):
            out = eng.field_lattice(*axes, soften=True); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
        got = out[torch.from_numpy(idx).cuda()].cpu().numpy()
        err = float(np.abs(got - want).max() / np.abs(want).max())
        if ref is None:
            ref = out.clone()
        d = float((out.double() - ref.double()).abs().max() / ref.double().abs().max())
        pairs = float(n_pts) * len(Q)
        print(json.dumps(dict(shape=shape, kernel=name, ms=round(best, 3), frac_nominal=round(pairs * 20 / (best * 1e-3) / 74.45e12, 4),
                              maxrel_vs_float64=err, maxrel_vs_shipped=d)), flush=True)
