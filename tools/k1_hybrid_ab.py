#!/usr/bin/env python
"""General field kernel (point lists): the shipped direct form against the experimental hybrid near/far form
(k1_hybrid=1), kernel time and error against the float64 oracle on a sample."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine
from oracle import f64
eng = Engine(0); eng.set_tuning(timing=1)
rng = np.random.default_rng(0)
for m, n, half in ((100_000, 1_000_000, 1.5), (7890, 1_000_000, 0.5), (7890, 100_000, 0.5), (30_000, 5_000, 1.0)):
    x, Q = synth.charges(m, seed=1, box=half)
    pts = (rng.uniform(-1, 1, (n, 3)) * half).astype(np.float32)
    if n == 5_000:                                  # a charge 2e-4 A from a point: the softening acts there
        x = np.vstack([x, pts[17] + np.float32(2e-4)]).astype(np.float32); Q = np.concatenate([Q, [0.3]]).astype(np.float32)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    dp = torch.from_numpy(pts).cuda()
    idx = rng.choice(n, 1024, replace=False)
    if n == 5_000:
        idx[0] = 17
    for soften in (True, False):
        want = f64.field_grid(pts[idx], x, Q, soften)
        ref = None
        for name, cfg in (("direct (shipped)", dict(k1_hybrid=0)), ("hybrid", dict(k1_hybrid=1)), ("hybrid, 3 splits", dict(k1_hybrid=1, k1_splits=3))):
            eng.set_tuning(k1_hybrid=0, k1_splits=0); eng.set_tuning(**cfg)
            best = 1e30
            for _ in range(3):
                out = eng.field_grid(dp, soften=soften, concat=True); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
            got = out[torch.from_numpy(idx).cuda()].cpu().numpy()
            fin = np.isfinite(want).all(axis=1)
            err = float(np.abs(got[fin, 3:] - want[fin]).max() / np.abs(want[fin]).max())
            same_nan = bool(np.array_equal(np.isfinite(got[:, 3:]).all(axis=1), fin))
            if ref is None:
                ref = out.clone()
            both = torch.isfinite(out).all(dim=1) & torch.isfinite(ref).all(dim=1)
            d = float((out[both].double() - ref[both].double()).abs().max() / ref[both].double().abs().max())
            print(json.dumps(dict(M=len(Q), N=n, soften=soften, kernel=name, path=eng.last_path(), ms=round(best, 3),
                                  frac_nominal=round(float(n) * len(Q) * 20 / (best * 1e-3) / 74.45e12, 4),
                                  maxrel_vs_float64=err, nan_pattern_ok=same_nan, maxrel_vs_direct=d,
                                  coords_ok=bool(np.array_equal(got[:, :3], pts[idx])))), flush=True)
eng.set_tuning(k1_hybrid=0, k1_splits=0)
