#!/usr/bin/env python
"""ncu driver for the streamline kernel: the 3A line set (47^3 lines, 7,890 charges), default
heuristics.  `python tools/prof_k2w.py [reps] [json tuning]`"""
import json, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200.device import Engine


def main():
    eng = Engine(0)
    x, Q = synth.charges(7890, seed=1, box=0.5)
    eng.set_charges(torch.from_numpy(x).cuda(), torch.from_numpy(Q).cuda())
    seeds, n_iter, dims, _ = synth.seeds(47, 0.5, 0.1)
    sd = torch.from_numpy(seeds).cuda(); ni = torch.from_numpy(n_iter.astype(np.int32)).cuda()
    if len(sys.argv) > 2:
        eng.set_tuning(**json.loads(sys.argv[2]))
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
        eng.topo_batch(sd, ni, 0.1, dims)
    torch.cuda.synchronize(); print("done")


if __name__ == "__main__":
    main()
