#!/usr/bin/env python
"""cpet_topo_hist_frames with the reference's calling convention (plain NumPy arrays = pageable memory): wall-clock of
a 20-frame call with the library page-locking the result buffers for the call (frames_pin=1, default) and without
(frames_pin=0: every copy back blocks the enqueueing thread until its frame has finished), against pinned buffers."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from pycpet_b200 import Math_ops
m = Math_ops()
for n_axis in (47, 100):
    x, Q = synth.charges(7890, seed=1, box=0.5)
    seeds, n_iter, dims, _ = synth.seeds(n_axis, 0.5, 0.1)
    K = 20 if n_axis == 47 else 8
    rng = np.random.default_rng(0)
    frames = [((x + rng.normal(0, 0.05, x.shape)).astype(np.float32), Q) for _ in range(K)]
    de, ce = np.linspace(0, 1.8, 51), np.linspace(0, 5, 51)
    nit = np.broadcast_to(n_iter.astype(np.int32), (K, len(seeds))).copy()
    res = {}
    for name, pin, pinned_bufs in (("pageable, frames_pin=0", 0, False), ("pageable, frames_pin=1 (default)", 1, False),
                                   ("caller-pinned buffers", 1, True)):
        m.set_tuning(frames_pin=pin)
        if pinned_bufs:
            rows = torch.empty((K, len(seeds), 2), dtype=torch.float32).pin_memory().numpy()
            counts = torch.empty((K, 50, 50), dtype=torch.int64).pin_memory().numpy()
        else:
            rows = np.zeros((K, len(seeds), 2), np.float32)
            counts = np.zeros((K, 50, 50), np.int64)
        best = 1e30
        for _ in range(4):
            t0 = time.perf_counter()
            m.topo_hist_frames(frames, seeds, nit, de, ce, step_size=0.1, dimensions=dims, want_rows=True,
                               rows_out=rows, counts_out=counts)
            best = min(best, time.perf_counter() - t0)
        res[name] = (round(best * 1e3 / K, 4), rows.copy())
        print(json.dumps(dict(lines=len(seeds), frames=K, buffers=name, ms_per_frame=res[name][0])), flush=True)
    names = list(res)
    assert all(np.array_equal(res[names[0]][1], res[n][1]) for n in names[1:])
m.set_tuning(frames_pin=1)
