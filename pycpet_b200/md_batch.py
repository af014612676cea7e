"""MD-trajectory batches: host preparation in worker processes, pipelined with the GPU.

The reference walks an input directory one structure at a time (CPET/source/CPET.py:96-127,
`run_topo`): build a `calculator` (PDB/PQR parse -> atom filters -> box-frame transform -> seeds and
`n_iter`, CPET/source/calculator.py:68-415), integrate, `np.savetxt("<protein>.top")`, next file.
On a B200 the integration of a 100k-line frame takes ~2 ms while the reference's own constructor
takes 0.2-0.8 s per file, so the Python preparation is what bounds a trajectory (SURVEY.md 8(f)
row 4).  It stays PyCPET's Python -- this module only runs it in `workers` processes and feeds the
prepared frames to `Math_ops.topo_hist_frames` in chunks, so file k+chunk is being parsed while
chunk k is on the GPU (the ctypes call releases the GIL; the copies of one frame overlap the
kernels of its neighbours inside the call).

`run_topo_frames` keeps the dispatcher's conventions: output name `<protein>.top` with
protein = basename up to the first dot, the per-file skip-if-exists resume rule, rows in seed order,
text bytes identical to `np.savetxt`.
"""
from __future__ import annotations

import functools
import multiprocessing as mp
import os

import numpy as np


def protein_name(path: str) -> str:
    """CPET/source/CPET.py:116: file.split("/")[-1].split(".")[0]."""
    return path.split("/")[-1].split(".")[0]


def prepare_frame(options, path):
    """Worker-side: the reference's own calculator constructor (parse, filter, transform, seeds),
    reduced to the arrays the hot path needs.  Needs an importable PyCPET in the worker."""
    from CPET.source.calculator import calculator

    c = calculator(options, path_to_pdb=path)
    return {
        "path": path,
        "x": np.ascontiguousarray(c.x, dtype=np.float32).reshape(-1, 3),
        "Q": np.ascontiguousarray(c.Q, dtype=np.float32).reshape(-1),
        "seeds": np.ascontiguousarray(c.random_start_points, dtype=np.float32).reshape(-1, 3),
        "n_iter": np.ascontiguousarray(c.random_max_samples, dtype=np.int32).reshape(-1),
        "step_size": float(c.step_size),
        "dimensions": np.ascontiguousarray(c.dimensions, dtype=np.float32).reshape(3),
    }


def _same_geometry(a, b) -> bool:
    return (a["step_size"] == b["step_size"] and np.array_equal(a["dimensions"], b["dimensions"])
            and a["seeds"].shape == b["seeds"].shape and np.array_equal(a["seeds"], b["seeds"]))


def _flush(math, frames, d_edges, c_edges, second_diff):
    """One GPU call per run of frames that share seeds, box and step (normally the whole chunk)."""
    out = []
    i = 0
    while i < len(frames):
        j = i + 1
        while j < len(frames) and _same_geometry(frames[i], frames[j]):
            j += 1
        grp = frames[i:j]
        rows, counts = math.topo_hist_frames(
            [(f["x"], f["Q"]) for f in grp], grp[0]["seeds"], np.stack([f["n_iter"] for f in grp]),
            d_edges, c_edges, step_size=grp[0]["step_size"], dimensions=grp[0]["dimensions"],
            second_diff=second_diff, want_rows=True)
        out += [(f["path"], rows[k], counts[k]) for k, f in enumerate(grp)]
        i = j
    return out


def run_topo_frames(options, files, outputpath=None, d_edges=None, c_edges=None, workers=None,
                    chunk=16, math=None, prepare=prepare_frame, initializer=None, initargs=(),
                    second_diff=None, skip_done=True, keep_rows=False, binary=False):
    """Topology of every structure in `files` (the loop of CPET.run_topo as one pipelined batch).

    options     the reference's options dict, handed unchanged to `prepare(options, path)`
    outputpath  directory for `<protein>.top` files (None: nothing is written)
    binary      also leave `<protein>.top.npy` (float32 rows) next to every text file; make_histograms
                then skips the text parse (same values bit for bit)
    d_edges, c_edges  optional shared bin edges; with them the (nd,nc) int64 counts of every frame
                are returned as well (np.histogram2d binning)
    workers     preparation processes (default: all host cores but one); 0 prepares in-process.
                They are *spawned* (the parent may hold a CUDA context), so a script calling this
                needs the usual `if __name__ == "__main__":` guard
    chunk       frames per GPU call
    second_diff curvature formula: None (default) takes options["curvature"] ("direction", the
                default, or "second_diff" = the reference's literal FP32 second differences)
    initializer/initargs  run once in every worker (e.g. to put PyCPET on sys.path)
    -> {"files": [...done in this call...], "skipped": [...], "counts": (F,nd,nc) or None,
        "rows": [(L,2) float32, ...] if keep_rows}
    """
    from . import io as cio

    if math is None:
        from .calculator import get_math

        math = get_math()
    if second_diff is None:
        mode = str(options.get("curvature", "direction")) if hasattr(options, "get") else "direction"
        if mode not in ("direction", "second_diff"):
            raise ValueError("options['curvature'] must be 'direction' or 'second_diff', got %r" % mode)
        second_diff = mode == "second_diff"
    files = list(files)
    todo, skipped = [], []
    done_names = set()
    if outputpath is not None:
        os.makedirs(outputpath, exist_ok=True)
        if skip_done:                                   # CPET.py:118-119: files ending in "top"
            done_names = {n for n in os.listdir(outputpath) if n.endswith("top")}
    claimed = set()                                     # two inputs with one protein name share one output file
    for f in files:
        name = protein_name(f) + ".top"
        if name in done_names or name in claimed:
            skipped.append(f)
        else:
            claimed.add(name)
            todo.append(f)
    want_counts = d_edges is not None and c_edges is not None
    if not want_counts:                                 # the call needs edges; one catch-all bin
        d_edges = c_edges = np.array([0.0, np.finfo(np.float64).max])

    result = {"files": [], "skipped": skipped, "counts": [] if want_counts else None, "rows": [] if keep_rows else None}

    def emit(batch):
        for path, rows, counts in _flush(math, batch, d_edges, c_edges, second_diff):
            if outputpath is not None:
                dst = os.path.join(outputpath, protein_name(path) + ".top")
                # the reference re-lists the output directory before every file (CPET.py:118-119) so
                # that concurrent runs over one directory skip each other's finished frames
                if skip_done and os.path.exists(dst):
                    skipped.append(path)
                    continue
                cio.save_topology(dst, rows, binary=binary)
            result["files"].append(path)
            if want_counts:
                result["counts"].append(counts)
            if keep_rows:
                result["rows"].append(rows.copy())

    task = functools.partial(prepare, options)
    if workers is None:
        workers = max(1, (os.cpu_count() or 2) - 1)
    if workers == 0 or len(todo) <= 1:
        prepared = map(task, todo)
        pool = None
    else:
        # spawn, not fork: the parent may already hold a CUDA context
        pool = mp.get_context("spawn").Pool(min(workers, len(todo)), initializer=initializer, initargs=initargs)
        prepared = pool.imap(task, todo, chunksize=1)   # ordered; workers run ahead of the consumer
    try:
        batch = []
        for frame in prepared:
            batch.append(frame)
            if len(batch) >= max(1, int(chunk)):
                emit(batch)
                batch = []
        if batch:
            emit(batch)
    except BaseException:
        if pool is not None:
            pool.terminate()        # a failed constructor or GPU call: do not parse the rest
            pool.join()
        raise
    if pool is not None:
        pool.close()
        pool.join()
    if want_counts:
        result["counts"] = (np.stack(result["counts"]) if result["counts"]
                            else np.zeros((0, len(d_edges) - 1, len(c_edges) - 1), np.int64))
    return result
