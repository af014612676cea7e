#!/usr/bin/env python
"""MD-batch driver on copies of the shipped 3A structure: frames/s against the number of
preparation workers (the unmodified PyCPET constructor runs in the workers; needs baseline/_ref).

    python tools/md_batch_demo.py [n_frames] [n_samples]
"""
import gzip, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_dropin import _reference_importable, GOLD, REF     # stubs for the absent plotting modules
from pycpet_b200 import md_batch


def main():
    if not os.path.isdir(os.path.join(REF, "CPET")):
        sys.exit("baseline/_ref is not present")
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    opts = json.load(open(os.path.join(GOLD, "dropin_options_topo.json")))
    opts["n_samples"] = n_samples
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "src.pdb")
    with gzip.open(os.path.join(GOLD, "1_alcdehydro_run1.pdb.gz"), "rb") as fi, open(src, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    files = []
    for k in range(n_frames):
        dst = os.path.join(tmp, f"{k}_frame.pdb"); shutil.copyfile(src, dst); files.append(dst)
    _reference_importable()
    t0 = time.perf_counter(); f0 = md_batch.prepare_frame(opts, files[0]); t_prep = time.perf_counter() - t0
    print(json.dumps(dict(what="one constructor call in this process", seconds=round(t_prep, 3), charges=len(f0["Q"]),
                          lines=len(f0["seeds"]))), flush=True)
    de, ce = np.linspace(0, 1.8, 51), np.linspace(0, 5, 51)
    for workers in (0, 4, 8, 15):
        out = os.path.join(tmp, f"out_w{workers}")
        t0 = time.perf_counter()
        res = md_batch.run_topo_frames(opts, files if workers else files[:8], outputpath=out, workers=workers, chunk=8,
                                       initializer=_reference_importable, d_edges=de, c_edges=ce)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(workers=workers, frames=len(res["files"]), seconds=round(dt, 2),
                              frames_per_s=round(len(res["files"]) / dt, 2))), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":        # the workers are spawned: they re-import this file
    main()
