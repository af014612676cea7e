#!/usr/bin/env python
"""A/B of the streamline kernel forms on the bench cases: k2_form=1 (round-1 direct form), 2 (hybrid near/far,
charge pairs packed) and 3 (hybrid, points packed: the default for long queues), kernel time only (CUDA events around the integrator launch, best of 5),
plus the largest difference of the two outputs and both against the float64 oracle on a sample."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from pycpet_b200.device import Engine  # noqa: E402


def main():
    from oracle import f64
    eng = Engine(0)
    eng.set_tuning(timing=1)
    cases = [(7890, 47, 0.1, 0.5), (7890, 100, 0.1, 0.5), (7890, 47, 0.01, 0.5), (30_000, 47, 0.1, 0.5),
             (100_000, 30, 0.1, 0.5), (7890, 47, 0.1, 1.5), (1000, 47, 0.1, 0.5)]
    extra = [json.loads(a) for a in sys.argv[1:]]
    for m, n_axis, h, box in cases:
        x, Q = synth.charges(m, seed=1, box=box)
        seeds, n_iter, dims, _ = synth.seeds(n_axis, box, h)
        sd = torch.from_numpy(seeds).cuda()
        ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        sample = np.random.default_rng(0).choice(len(seeds), size=min(256, len(seeds)), replace=False)
        ref, _ = f64.topo_batch(seeds[sample], n_iter[sample], x, Q, h, dims)
        outs = {}
        for cfg in [dict(k2_form=1), dict(k2_form=2), dict(k2_form=3)] + extra:
            eng.set_tuning(k2_form=0, k2_threads=0, k2_cap=0, k2_amax=0, k2_unroll=0)
            eng.set_tuning(**cfg)
            best = 1e30
            for _ in range(5):
                out = eng.topo_batch(sd, ni, h, dims)
                torch.cuda.synchronize()
                best = min(best, eng.last_kernel_ms())
            c = eng.last_counters()
            out = (out[0] if isinstance(out, tuple) else out).cpu().numpy()
            outs[json.dumps(cfg)] = out
            d = np.abs(out[sample] - ref)
            print(json.dumps(dict(M=len(Q), L=len(seeds), h=h, box=box, cfg=cfg, ms=round(best, 4),
                                  pairs_per_s="%.4e" % (c["pair_evals"] / (best * 1e-3)),
                                  frac_nominal=round(c["pair_evals"] * 20 / (best * 1e-3) / 74.45e12, 4),
                                  dist_err=float(np.nanmax(d[:, 0])), curv_err=float(np.nanmax(d[:, 1])),
                                  launches=c["launches"])), flush=True)
        ks = list(outs)
        dd = np.abs(outs[ks[0]] - outs[ks[2]])
        print(json.dumps(dict(direct_vs_points_packed_max_dist=float(np.nanmax(dd[:, 0])), max_curv=float(np.nanmax(dd[:, 1])),
                              lines_differing_by_a_step=int((dd[:, 0] > h / 2).sum()))), flush=True)


if __name__ == "__main__":
    main()
