#!/usr/bin/env python
"""Time the UNMODIFIED reference (installed once into baseline/_ref with
`pip install --no-deps --target baseline/_ref`) through its own calculator methods on the bench's
synthetic 3A frame: the stock CPU path (multiprocessing.Pool + ctypes, SC:675-712) and the stock
PyTorch `topo_GPU` path (SC:793-978).  Informational only -- bench.py's reference arm drives the
same C code through threads (no per-task pickling), which is several times faster than this.

    python tools/stock_reference.py [n_axis=18] [procs=all] [gpu=0|1]
"""
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(1, ROOT)
os.environ["CPET_BANNER"] = "0"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    # plotting / clustering imports of the reference that this image lacks; untouched by the path
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors = _stub("matplotlib.colors", LinearSegmentedColormap=object, Normalize=object)
    mpl.cm = _stub("matplotlib.cm")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("seaborn")
    _stub("kneed", KneeLocator=object)
    tl = _stub("tensorly")
    tl.decomposition = _stub("tensorly.decomposition", parafac=None, non_negative_parafac=None)
    se = _stub("sklearn_extra")
    se.cluster = _stub("sklearn_extra.cluster", KMedoids=object)


def main():
    n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
    use_gpu = len(sys.argv) > 3 and sys.argv[3] == "1"
    install_stubs()
    import logging

    import synth
    from CPET.source.calculator import calculator   # the reference's class, unmodified

    x, Q = synth.charges(7890, seed=1, box=0.5)
    seeds, n_iter, dims, max_steps = synth.seeds(n_axis, 0.5, 0.1)
    calc = object.__new__(calculator)               # skip __init__ (PDB parsing): set what compute_* reads
    calc.log = logging.LoggerAdapter(logging.getLogger("stock"), {})
    calc.x, calc.Q = x, Q.reshape(-1, 1)
    calc.random_start_points, calc.random_max_samples = seeds, n_iter
    calc.n_samples, calc.step_size, calc.dimensions = len(seeds), 0.1, dims.astype(np.float32)
    calc.concur_slip, calc.GPU_batch_freq, calc.dtype, calc.max_steps = procs, 100, "float32", max_steps
    calc.transformation_matrix, calc.center = np.eye(3), np.zeros(3)
    from oracle import f64
    _, steps = f64.topo_batch(seeds, n_iter, x, Q, 0.1, dims)
    credited = float((steps.astype(np.int64) + 2).sum()) * len(Q)

    t0 = time.perf_counter()
    hist = calc.compute_topo_complete_c_shared()
    t_cpu = time.perf_counter() - t0
    print(f"[stock] CPU Pool({procs}) compute_topo_complete_c_shared: {len(seeds)} lines in {t_cpu:.2f} s "
          f"= {len(seeds) / t_cpu:.0f} lines/s, {credited / t_cpu:.3e} credited pair-evals/s")
    if use_gpu:
        import torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist_gpu = calc.compute_topo_GPU_batch_filter()
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        print(f"[stock] torch topo_GPU compute_topo_GPU_batch_filter: {len(seeds)} lines in {t_gpu:.2f} s "
              f"= {len(seeds) / t_gpu:.0f} lines/s, {credited / t_gpu:.3e} credited pair-evals/s")
        a = np.asarray(hist)[np.lexsort((np.round(hist[:, 1], 3), np.round(hist[:, 0], 3)))]
        b = np.asarray(hist_gpu)[np.lexsort((np.round(hist_gpu[:, 1], 3), np.round(hist_gpu[:, 0], 3)))]
        print(f"[stock] CPU vs torch rows (sorted): max |d dist| {np.max(np.abs(a[:, 0] - b[:, 0])):.2e}")


if __name__ == "__main__":
    main()
