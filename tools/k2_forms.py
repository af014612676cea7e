#!/usr/bin/env python
"""Which streamline kernel form for which queue length?  Kernel time (best of 5) of the charge-pair-packed
hybrid (k2_form=2) and the points-packed hybrid (k2_form=3, lines per warp 8 / 4) over line counts from
8,000 to 1e6, for a resident (7,890 charges) and a streamed (100,000 charges) frame."""
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
    eng = Engine(0)
    eng.set_tuning(timing=1)
    for m in (7890, 100_000):
        x, Q = synth.charges(m, seed=1, box=0.5)
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        for n_axis in (20, 25, 30, 34, 38, 42, 47, 60):
            if m > 10000 and n_axis > 47:
                continue
            seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
            sd = torch.from_numpy(seeds).cuda()
            ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
            row = dict(M=len(Q), L=len(seeds))
            for name, cfg in (("pairs", dict(k2_form=2)), ("points8", dict(k2_form=3, k2_cap=8)),
                              ("points4", dict(k2_form=3, k2_cap=4)), ("auto", dict())):
                eng.set_tuning(k2_form=0, k2_threads=0, k2_cap=0, k2_amax=0, k2_unroll=0)
                eng.set_tuning(**cfg)
                best = 1e30
                for _ in range(5):
                    eng.topo_batch(sd, ni, 0.1, dims)
                    torch.cuda.synchronize()
                    best = min(best, eng.last_kernel_ms())
                c = eng.last_counters()
                row[name] = round(c["pair_evals"] * 20 / (best * 1e-3) / 74.45e12, 4)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
