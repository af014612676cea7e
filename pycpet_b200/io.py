"""Writers and readers for the path's output files, byte-compatible with the reference's np.savetxt
calls and value-identical to its `float(line.split()[k])` parsing loops.

* ``save_topology(path, hist)``         == ``np.savetxt(path, hist)``            (CPET/source/CPET.py:123)
* ``save_numpy_as_dat(meta, volume, name)`` == CPET/utils/io.py:50-109 (same signature, same bytes)
* ``read_rows(path, n_cols)`` / ``read_topology(path)`` == the parse loops of make_histograms
  (CPET/utils/calculator.py:603-633, 690-698): '#' lines skipped, float() of the leading columns

np.savetxt formats row by row in Python; for a 1,000,000-line `.top` that costs seconds, i.e. two
orders of magnitude more than computing the lines on the GPU.  ``cpet_write_rows`` does the same
conversion with snprintf on all host cores.  The 7-line `.dat` header is built here in Python with
the reference's own format strings (it prints NumPy scalars, whose repr is Python's).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from ._lib import check

_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.float16): 2}


def write_rows(path, array, fmt="%.18e", header="", threads=0):
    """Write a 2-D (or 1-D) float array as text, one row per line, space separated."""
    a = np.asarray(array)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("write_rows expects a 1-D or 2-D array")
    if a.dtype not in _DTYPES:
        a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    check(_lib.load().cpet_write_rows(os.fsencode(path), header.encode(), _lib.ptr(a), _DTYPES[a.dtype],
                                      a.shape[0], a.shape[1], fmt.encode(), int(threads)))


def read_rows(path, n_cols=2, threads=0):
    """The data lines of a text file ('#' lines and blank lines skipped) -> (n, n_cols) float64, the
    first n_cols numbers of every line converted exactly as Python's float() converts them."""
    import ctypes

    L = _lib.load()
    n = ctypes.c_int64(0)
    check(L.cpet_count_rows(os.fsencode(path), ctypes.byref(n), int(threads)))
    out = np.empty((n.value, int(n_cols)), dtype=np.float64)
    check(L.cpet_read_rows(os.fsencode(path), int(n_cols), n.value, _lib.ptr(out), int(threads)))
    return out


def _side_channel(path):
    return path + ".npy"


def _fingerprint(path):
    """Cheap identity of a text file: size plus CRC32 of its first and last 64 KiB."""
    import zlib

    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(65536)
        if size > 131072:
            f.seek(size - 65536)
            tail = f.read(65536)
        else:
            tail = f.read()
    return "%d %08x %08x" % (size, zlib.crc32(head), zlib.crc32(tail))


def read_topology(path, use_binary=True):
    """`.top` file -> (n, 2) float64 [dist|curv] (what make_histograms parses, UC:626-633).
    When save_topology(..., binary=True) left a `<name>.top.npy` next to the text, the rows come from
    there -- but only if the fingerprint stored with it (`<name>.top.npy.fp`: size and CRC32 of the
    head and tail of the text) still matches the text file, so a `.top` that was replaced with its
    mtime preserved (cp -p, rsync -t, a restored backup) is parsed again instead of being shadowed by
    stale rows.  '%.18e' prints a float32 exactly, so both routes give the same float64 values bit for
    bit.  A side channel without its text file is never used (use_binary="force" to allow that)."""
    npy = _side_channel(path)
    if use_binary and os.path.exists(npy):
        if os.path.exists(path):
            try:
                with open(npy + ".fp") as f:
                    ok = f.read().strip() == _fingerprint(path)
            except OSError:
                ok = False
            if ok:
                return np.load(npy).astype(np.float64).reshape(-1, 2)
        elif use_binary == "force":
            return np.load(npy).astype(np.float64).reshape(-1, 2)
    return read_rows(path, 2)


def save_topology(path, hist, binary=False):
    """`.top` file: two columns dist, curv in '%.18e' -- the bytes np.savetxt(path, hist) writes.
    binary=True also leaves the float32 rows as `<name>.top.npy` and the text's fingerprint as
    `<name>.top.npy.fp`; neither name ends in "top", so the dispatcher's resume rule
    (CPET/source/CPET.py:118-119) does not see them."""
    write_rows(path, hist, fmt="%.18e")
    if binary:
        a = np.asarray(hist)
        if a.dtype != np.float32:
            raise ValueError("the binary side channel holds float32 rows (what the integrator returns)")
        np.save(_side_channel(path), np.ascontiguousarray(a).reshape(-1, 2))
        with open(_side_channel(path) + ".fp", "w") as f:
            f.write(_fingerprint(path) + "\n")


def dat_header(meta_data):
    """The 7 header lines of `_efield.dat` / `_esp.dat` (CPET/utils/io.py:59-85)."""
    dimensions = meta_data["dimensions"]
    num_steps_list = meta_data["num_steps"]
    trans_mat = meta_data["transformation_matrix"].transpose()
    center = meta_data["center"]
    first_line = "#Sample Density: {} {} {}; Volume: Box: {} {} {}\n".format(
        num_steps_list[0], num_steps_list[1], num_steps_list[2], dimensions[0], dimensions[1], dimensions[2])
    second_line = "#Frame 0\n"
    third_line = "#Center: {} {} {}\n".format(center[0], center[1], center[2])
    basis = "#Basis Matrix:\n# {} {} {}\n# {} {} {}\n# {} {} {}\n".format(
        trans_mat[0][0], trans_mat[0][1], trans_mat[0][2], trans_mat[1][0], trans_mat[1][1], trans_mat[1][2],
        trans_mat[2][0], trans_mat[2][1], trans_mat[2][2])
    return first_line + second_line + third_line + basis


def save_numpy_as_dat(meta_data, volume, name):
    """Same signature and same bytes as the reference's save_numpy_as_dat (CPET/utils/io.py:50-109):
    header + np.savetxt(name, volume, fmt='%.3f')."""
    write_rows(name, volume, fmt="%.3f", header=dat_header(meta_data))
