#!/usr/bin/env python
"""BASELINE config 5 on one GPU: `volume` E-field (softened, box mesh -> lattice kernel) and
`volume_ESP` for M = 1e3..1e6 charges x N = 100^3, 215^3, 464^3 points.  Kernel time (CUDA events
around the dominant kernel), best of 2; cells above 2.5e13 pair-evaluations are skipped."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine


def main():
    eng = Engine(0); eng.set_tuning(timing=1)
    peak = 148 * 128 * 2 * 1.965e9
    for m in (1_000, 10_000, 100_000, 1_000_000):
        x, Q = synth.charges(m, seed=1, box=1.5)
        eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
        for n_axis in (100, 215, 464):
            pairs = float(n_axis) ** 3 * len(Q)
            if pairs > 2.5e13:
                continue
            ax = torch.linspace(-1.5, 1.5, n_axis, device="cuda")
            for mode in ("volume", "volume_ESP"):
                fn = (lambda: eng.field_lattice(ax, ax, ax, soften=True)) if mode == "volume" else (lambda: eng.esp_lattice(ax, ax, ax))
                best = 1e30
                for _ in range(2):
                    out = fn(); torch.cuda.synchronize(); best = min(best, eng.last_kernel_ms())
                del out
                rate = pairs / (best * 1e-3)
                print(json.dumps(dict(mode=mode, M=len(Q), N=n_axis ** 3, ms=round(best, 2), pairs_per_s="%.3e" % rate,
                                      frac_nominal_fp32=round(rate * (20 if mode == "volume" else 11) / peak, 3))), flush=True)


if __name__ == "__main__":
    main()
