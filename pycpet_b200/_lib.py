"""ctypes loader for libcpetb200.so (the C-ABI declared in include/cpet_b200.h).

Mirrors how the reference finds and opens its C-shared math module
(CPET/utils/calculator.py:19-29: locate the shared object, ``ctypes.CDLL`` it once, keep a
process-wide handle).  There is no fallback of any kind: if the library has not been built, or it
cannot create a context on an sm_100 device, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcpetb200.so")

CPET_OK = 0
CPET_FIELD_SOFTEN = 1
CPET_OUT_CONCAT = 2
CPET_TOPO_CURV_SECOND_DIFF = 1


class CpetError(RuntimeError):
    """A libcpetb200 entry point returned a negative status."""


_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_uint = ctypes.c_uint
c_float = ctypes.c_float
c_int64 = ctypes.c_int64

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "cpet_abi_version": (c_int, []),
    "cpet_last_error": (ctypes.c_char_p, []),
    "cpet_last_status": (c_int, []),
    "cpet_clear_error": (None, []),
    "cpet_device_count": (c_int, []),
    "cpet_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "cpet_create_on_stream": (c_int, [c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    "cpet_destroy": (c_int, [c_void_p]),
    "cpet_sync": (c_int, [c_void_p]),
    "cpet_device_of": (c_int, [c_void_p]),
    "cpet_last_path": (c_int, [c_void_p]),
    "cpet_set_tuning": (c_int, [c_void_p, ctypes.c_char_p, c_int]),
    "cpet_last_counters": (c_int, [c_void_p, ctypes.POINTER(c_int64)]),
    "cpet_set_charges": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cpet_set_charges_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cpet_field_grid": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_void_p]),
    "cpet_field_grid_dev": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_void_p]),
    "cpet_esp_grid": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_void_p]),
    "cpet_esp_grid_dev": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_void_p]),
    "cpet_field_lattice": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "cpet_field_lattice_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "cpet_esp_lattice": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "cpet_esp_lattice_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "cpet_propagate": (c_int, [c_void_p, c_int, c_void_p, c_float, c_void_p]),
    "cpet_propagate_dev": (c_int, [c_void_p, c_int, c_void_p, c_float, c_void_p]),
    "cpet_topo_batch": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_uint,
                                c_void_p, c_void_p]),
    "cpet_topo_batch_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_uint,
                                    c_void_p, c_void_p]),
    "cpet_hist2d": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p,
                            c_void_p]),
    "cpet_hist2d_f32": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int,
                                c_void_p, c_void_p]),
    "cpet_hist2d_dev": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int,
                                c_void_p, c_void_p]),
    "cpet_topo_hist": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_uint, c_void_p,
                               c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "cpet_topo_hist_frames": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int64, c_float, c_void_p, c_uint, c_void_p, c_int, c_void_p, c_int,
                                      c_void_p, c_void_p]),
    "cpet_order_stats": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpet_order_stats_dev": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpet_radix_hist_dev": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "cpet_chi2_matrix": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "cpet_chi2_rows_dev": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "cpet_write_rows": (c_int, [ctypes.c_char_p, ctypes.c_char_p, c_void_p, c_int, c_int64, c_int,
                                ctypes.c_char_p, c_int]),
    "cpet_count_rows": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_int64), c_int]),
    "cpet_read_rows": (c_int, [ctypes.c_char_p, c_int, c_int64, c_void_p, c_int]),
    "cpet_fp32_peak_probe": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(ctypes.c_double)]),
    "cpet_last_kernel_ms": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "cpet_kernel_times": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double), c_int, ctypes.POINTER(c_int)]),
}

PATH_NAMES = {0: "general", 1: "lattice", 3: "general_hybrid", 11: "k2w", 12: "k2x", 13: "k2p"}     # cpet_last_path codes

# the reference's own symbol names (include/cpet_b200.h section (A)); argtypes are set by Math_ops
LEGACY_SYMBOLS = [
    "compute_looped_field", "compute_batched_field", "calc_field", "calc_field_base",
    "calc_esp_base", "thread_operation", "thread_operation_dipole", "einsum_ij_i",
    "einsum_ij_i_batch", "einsum_operation", "einsum_operation_batch", "vecaddn", "dot",
    "sparse_dot",
]


def lib_path() -> str:
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Open libcpetb200.so (once).  Raises if it was never built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpetError(
                f"{LIB_PATH} is missing: build it with `python -m pycpet_b200.build` "
                "(nvcc, sm_100a).  pycpet_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, lib=None) -> None:
    """Raise CpetError for a negative status; the message comes from the thread-local error state of
    `lib` -- the handle the failing call went through (a Math_ops(shared_loc=...) instance may hold
    another copy of the library than the in-tree one)."""
    if status != CPET_OK:
        L = lib if lib is not None else load()
        msg = L.cpet_last_error()
        text = msg.decode() if msg else "unknown error"
        L.cpet_clear_error()
        raise CpetError(f"libcpetb200 status {status}: {text}")


def check_legacy(lib=None) -> None:
    """The reference's `void` symbols have no error channel; ours record one thread-locally (in the
    library instance the call went through, see check())."""
    L = lib if lib is not None else load()
    st = L.cpet_last_status()
    if st != CPET_OK:
        msg = L.cpet_last_error()
        L.cpet_clear_error()
        raise CpetError(f"libcpetb200 status {st}: {msg.decode() if msg else 'unknown error'}")


def ptr(a: np.ndarray) -> c_void_p:
    return c_void_p(a.ctypes.data)


def f32c(a, shape=None) -> np.ndarray:
    """float32, C-contiguous view/copy -- what the reference's wrappers do with np.array(...,
    dtype='float32') before crossing the boundary (c_ops.py:257-259)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a
