"""float64 oracle (TEST INFRASTRUCTURE) -- ctypes wrapper around ``cpet_oracle.c``.

Each function names the reference routine it restates (``C`` = CPET/utils/math_module.c).
Parity status: pinned against reference-generated golden vectors (tests/golden/) and
against ``oracle/_ref`` (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcpet_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C")


def build(force: bool = False) -> str:
    """Compile ``cpet_oracle.c`` (gcc, OpenMP) next to its source if needed."""
    src = os.path.join(_HERE, "cpet_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_field_grid.restype = None
        L.orc_field_grid.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p,
                                     ctypes.c_int, _f64p]
        L.orc_esp_grid.restype = None
        L.orc_esp_grid.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f64p]
        L.orc_step.restype = None
        L.orc_step.argtypes = [_f32p, ctypes.c_float, ctypes.c_int, _f32p, _f32p, _f64p]
        L.orc_line.restype = None
        L.orc_line.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32p, _f32p, _f32p,
                               _f32p, _f64p, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
        L.orc_topo_batch.restype = None
        L.orc_topo_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32p, _i64p,
                                     _f32p, _f32p, _f32p, _f64p, ctypes.c_void_p]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the following calls, independent of OMP_NUM_THREADS."""
    lib().orc_set_num_threads(int(n))


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def field_grid(x0, x, Q, soften: bool) -> np.ndarray:
    """E-field at points ``x0`` (N,3).  soften=True restates compute_looped_field (C:405-451,
    r^2 := max(r^2, 1e-6)); soften=False restates calc_field_base / calc_field (C:296-333,
    C:255-293).  Returns (N,3) float64."""
    x0 = _f32(x0, (-1, 3)); x = _f32(x, (-1, 3)); Q = _f32(Q, (-1,))
    out = np.zeros((x0.shape[0], 3), dtype=np.float64)
    lib().orc_field_grid(x0.shape[0], x.shape[0], x0, x, Q, int(bool(soften)), out)
    return out


def esp_grid(x0, x, Q) -> np.ndarray:
    """Potential at points ``x0`` (N,3): restates calc_esp_base (C:453-486) applied per point by
    compute_ESP_on_grid (CPET/utils/calculator.py:450-475).  Returns (N,) float64."""
    x0 = _f32(x0, (-1, 3)); x = _f32(x, (-1, 3)); Q = _f32(Q, (-1,))
    out = np.zeros(x0.shape[0], dtype=np.float64)
    lib().orc_esp_grid(x0.shape[0], x.shape[0], x0, x, Q, out)
    return out


def step(p, h, x, Q) -> np.ndarray:
    """propagate_topo (C:489-503) in float64."""
    out = np.zeros(3, dtype=np.float64)
    x = _f32(x, (-1, 3))
    lib().orc_step(_f32(p, (3,)), float(h), x.shape[0], x, _f32(Q, (-1,)), out)
    return out


def line(seed, n_iter, x, Q, step_size, dimensions, want_points: bool = False):
    """thread_operation (C:523-591) in float64 -> ([dist, curv], K[, six points (6,3)])."""
    x = _f32(x, (-1, 3)); Q = _f32(Q, (-1,))
    ret = np.zeros(2, dtype=np.float64)
    k = ctypes.c_int(0)
    pts = np.zeros((6, 3), dtype=np.float64)
    lib().orc_line(x.shape[0], int(n_iter), float(step_size), _f32(seed, (3,)),
                   _f32(dimensions, (3,)), x, Q, ret, ctypes.byref(k),
                   pts.ctypes.data_as(ctypes.c_void_p))
    return (ret, int(k.value), pts) if want_points else (ret, int(k.value))


def topo_batch(seeds, n_iter, x, Q, step_size, dimensions):
    """All streamlines of a frame in seed order, i.e. what
    calculator.compute_topo_complete_c_shared (CPET/source/calculator.py:675-712) returns.
    -> (out (L,2) float64 [dist|curv], steps (L,) int32)."""
    seeds = _f32(seeds, (-1, 3)); x = _f32(x, (-1, 3)); Q = _f32(Q, (-1,))
    n_iter = np.ascontiguousarray(n_iter, dtype=np.int64).reshape(-1)
    L = seeds.shape[0]
    assert n_iter.shape[0] == L
    out = np.zeros((L, 2), dtype=np.float64)
    steps = np.zeros(L, dtype=np.int32)
    lib().orc_topo_batch(L, x.shape[0], float(step_size), seeds, n_iter,
                         _f32(dimensions, (3,)), x, Q, out,
                         steps.ctypes.data_as(ctypes.c_void_p))
    return out, steps
