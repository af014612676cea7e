"""oracle.ref (TEST INFRASTRUCTURE) -- the reference's own CPU implementation, compiled in place.

``make -C oracle ref`` compiles ``/root/reference/CPET/utils/math_module.c`` (unmodified, from
where it lies) into ``oracle/_ref/math_module_v{3,4}.so``.  This module binds the symbols the
hot path uses with our own ctypes declarations (the reference's ``CPET/utils/c_ops.py`` cannot
travel to the GPU box) following the argument order of c_ops.py:85-159.

Used (a) by tests to validate the float64 restatement and (b) by ``bench.py`` as the
``cpu_baseline`` of kind "reference".  Never imported by ``pycpet_b200``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_lib = None

_f32_1 = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="C")
_f32_2 = np.ctypeslib.ndpointer(dtype=np.float32, ndim=2, flags="C")


def _cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("flags"):
                    fl = set(ln.split(":", 1)[1].split())
                    return {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= fl
    except OSError:
        pass
    return False


def path() -> str | None:
    """Path of the best reference object for this CPU, or None when oracle/_ref was not built."""
    names = ["math_module_v4.so", "math_module_v3.so"] if _cpu_has_avx512() else ["math_module_v3.so"]
    for n in names:
        p = os.path.join(_DIR, n)
        if os.path.exists(p):
            return p
    return None


def available() -> bool:
    return path() is not None


def lib():
    global _lib
    if _lib is None:
        p = path()
        if p is None:
            raise RuntimeError("oracle/_ref not built: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(p)
        # c_ops.py:151-159
        L.compute_looped_field.restype = None
        L.compute_looped_field.argtypes = [ctypes.c_int, ctypes.c_int, _f32_2, _f32_2, _f32_1, _f32_2]
        # c_ops.py:113-138  (E/ESP, x_init, n, x, Q)
        for name in ("calc_field", "calc_field_base", "calc_esp_base"):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [_f32_1, _f32_1, ctypes.c_int, _f32_2, _f32_1]
        # c_ops.py:85-95
        L.thread_operation.restype = None
        L.thread_operation.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32_1, _f32_1,
                                       _f32_2, _f32_1, _f32_1]
        _lib = L
    return _lib


def _prep(x, Q):
    return (np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3),
            np.ascontiguousarray(Q, dtype=np.float32).reshape(-1))


def compute_looped_field(x0, x, Q) -> np.ndarray:
    """Reference `volume` kernel, one single-threaded call (c_ops.py:250-263)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
    x, Q = _prep(x, Q)
    out = np.zeros_like(x0)
    lib().compute_looped_field(x0.shape[0], Q.shape[0], x0, x, Q, out)
    return out


def calc_field(p, x, Q) -> np.ndarray:
    x, Q = _prep(x, Q)
    out = np.zeros(3, dtype=np.float32)
    lib().calc_field(out, np.ascontiguousarray(p, dtype=np.float32).reshape(3), Q.shape[0], x, Q)
    return out


def calc_field_base(p, x, Q) -> np.ndarray:
    x, Q = _prep(x, Q)
    out = np.zeros(3, dtype=np.float32)
    lib().calc_field_base(out, np.ascontiguousarray(p, dtype=np.float32).reshape(3), Q.shape[0], x, Q)
    return out


def calc_esp_base(p, x, Q) -> np.ndarray:
    x, Q = _prep(x, Q)
    out = np.zeros(1, dtype=np.float32)
    lib().calc_esp_base(out, np.ascontiguousarray(p, dtype=np.float32).reshape(3), Q.shape[0], x, Q)
    return out


def thread_operation(seed, n_iter, x, Q, step_size, dimensions) -> np.ndarray:
    x, Q = _prep(x, Q)
    out = np.zeros(2, dtype=np.float32)
    lib().thread_operation(Q.shape[0], int(n_iter), float(step_size),
                           np.ascontiguousarray(seed, dtype=np.float32).reshape(3),
                           np.ascontiguousarray(dimensions, dtype=np.float32).reshape(3), x, Q, out)
    return out


# ---------------------------------------------------------------------------------------------
# Whole-workload drivers (what the reference's calculator-level entry points do).
# Parallelism = forked worker PROCESSES, exactly the reference's model (multiprocessing.Pool,
# CPET/source/calculator.py:690-704) -- Python threads would serialise on the GIL around every
# per-line ctypes call and under-report the reference by ~3x.  Workers inherit the inputs by fork
# and write results straight into an anonymous shared mapping, so nothing is pickled but the spans.
# ---------------------------------------------------------------------------------------------
_WORK = None


def _call_work(span):
    _WORK(*span)
    return span[1] - span[0]


def _shared(shape, dtype):
    """ndarray backed by MAP_SHARED|MAP_ANONYMOUS memory: children forked later write into it."""
    import mmap

    nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    buf = mmap.mmap(-1, nbytes)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _run_chunks(fn, n, threads):
    global _WORK
    threads = max(1, int(threads))
    if threads == 1 or n < 2:
        fn(0, n)
        return
    import multiprocessing as mp
    import resource

    try:    # the C code keeps 16-28 bytes of VLA per charge on the stack (C:259-263, 300-307, 409-410)
        soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
        resource.setrlimit(resource.RLIMIT_STACK, (hard, hard))
    except (ValueError, OSError):
        pass
    step = max(1, -(-n // (threads * 8)))
    spans = [(s, min(n, s + step)) for s in range(0, n, step)]
    _WORK = fn
    import warnings

    try:
        with warnings.catch_warnings():
            # the parent may hold OpenMP / CUDA helper threads; the children only run the reference's
            # single-threaded C and never touch either, which is what the fork() warning is about
            warnings.simplefilter("ignore", DeprecationWarning)
            with mp.get_context("fork").Pool(threads) as pool:
                done = sum(pool.imap_unordered(_call_work, spans))
        assert done == n
    finally:
        _WORK = None


def topo(seeds, n_iter, x, Q, step_size, dimensions, threads: int = 1) -> np.ndarray:
    """compute_topo_complete_c_shared (CPET/source/calculator.py:675-712): one thread_operation
    call per streamline, rows in seed order -> (L,2) float32 [dist|curv]."""
    seeds = np.ascontiguousarray(seeds, dtype=np.float32).reshape(-1, 3)
    n_iter = np.asarray(n_iter).reshape(-1)
    dims = np.ascontiguousarray(dimensions, dtype=np.float32).reshape(3)
    x, Q = _prep(x, Q)
    L = lib()
    m = Q.shape[0]
    out = _shared((seeds.shape[0], 2), np.float32) if threads > 1 else np.zeros((seeds.shape[0], 2), np.float32)
    h = float(step_size)

    def work(a, b):
        for i in range(a, b):
            L.thread_operation(m, int(n_iter[i]), h, seeds[i], dims, x, Q, out[i])

    _run_chunks(work, seeds.shape[0], threads)
    return out


def field_grid(x0, x, Q, threads: int = 1) -> np.ndarray:
    """compute_field_on_grid (CPET/utils/calculator.py:430-447) without the coordinate concat;
    threads>1 splits the points into slabs, one compute_looped_field call per slab."""
    x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
    x, Q = _prep(x, Q)
    L = lib()
    out = _shared(x0.shape, np.float32) if threads > 1 else np.zeros_like(x0)

    def work(a, b):
        L.compute_looped_field(b - a, Q.shape[0], x0[a:b], x, Q, out[a:b])

    _run_chunks(work, x0.shape[0], threads)
    return out


def esp_grid(x0, x, Q, threads: int = 1) -> np.ndarray:
    """compute_ESP_on_grid (CPET/utils/calculator.py:450-475): a Python loop with one
    calc_esp_base call per grid point -> (N,) float64 (the array the reference fills before its
    float16 cast)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
    x, Q = _prep(x, Q)
    L = lib()
    m = Q.shape[0]
    out = _shared((x0.shape[0],), np.float64) if threads > 1 else np.zeros(x0.shape[0], dtype=np.float64)

    def work(a, b):
        for i in range(a, b):
            tmp = np.zeros(1, dtype=np.float32)
            L.calc_esp_base(tmp, x0[i], m, x, Q)
            out[i] = tmp[0]

    _run_chunks(work, x0.shape[0], threads)
    return out
