"""``Math_ops`` -- host-side mirror of the reference's ctypes binding class
(CPET/utils/c_ops.py:7-375) on top of libcpetb200.so.

Same constructor, same method names, same argument meaning, same return shapes/dtypes and the
same error behaviour at the boundary (``numpy.ctypeslib.ndpointer`` argtypes reject wrong
dtype / ndim / non-contiguous arrays with ``ctypes.ArgumentError`` before the call, c_ops.py:15-23).
Every method runs CUDA kernels; nothing here computes on the CPU.

On top of the per-point / per-line reference methods it exposes the batched entry points the
calculator-level functions use (``field_grid``, ``esp_grid``, ``topo_batch``, ``hist2d``,
``chi2_matrix``): one call per grid / per frame instead of a Python loop or a process Pool.
"""
from __future__ import annotations

import ctypes

import numpy as np
import numpy.ctypeslib as npct

from . import _lib
from ._lib import CpetError, check, check_legacy, f32c, ptr


class Math_ops:
    def __init__(self, shared_loc=None, device=None):
        # reference: ctypes.CDLL(shared_loc) (c_ops.py:9-12).  A path to libcpetb200.so may be
        # given; by default the in-tree build is used.
        if shared_loc is None:
            self.math = _lib.load()
        else:
            self.math = ctypes.CDLL(shared_loc)
            for name, (res, args) in _lib.SIGNATURES.items():
                fn = getattr(self.math, name)
                fn.restype, fn.argtypes = res, args
        m = self.math
        self.array_1d_int = npct.ndpointer(dtype=np.int32, ndim=1, flags="C")
        self.array_1d_float = npct.ndpointer(dtype=np.float32, ndim=1, flags="C")
        self.array_2d_float = npct.ndpointer(dtype=np.float32, ndim=2, flags="C")
        self.array_3d_float = npct.ndpointer(dtype=np.float32, ndim=3, flags="C")
        self.array_1d_double = npct.ndpointer(dtype=np.double, ndim=1, flags="C")
        self.array_2d_double = npct.ndpointer(dtype=np.double, ndim=2, flags="C")
        f1, f2, f3, ci, cf = (self.array_1d_float, self.array_2d_float, self.array_3d_float,
                              ctypes.c_int, ctypes.c_float)
        legacy = {
            # name: argtypes                                       reference binding
            "einsum_ij_i": [ci, ci, f2, f1],                        # c_ops.py:52-57
            "einsum_ij_i_batch": [ci, ci, ci, f3, f2],              # c_ops.py:58-64
            "einsum_operation_batch": [ci, ci, f2, f1, f3, f2],     # c_ops.py:66-74
            "einsum_operation": [ci, f1, f1, f2, f1],               # c_ops.py:76-83
            "thread_operation": [ci, ci, cf, f1, f1, f2, f1, f1],   # c_ops.py:85-95
            "thread_operation_dipole": [ci, ci, cf, f1, f1, f2, f2, f1],  # c_ops.py:101-111
            "calc_field": [f1, f1, ci, f2, f1],                     # c_ops.py:113-120
            "calc_field_base": [f1, f1, ci, f2, f1],                # c_ops.py:122-129
            "calc_esp_base": [f1, f1, ci, f2, f1],                  # c_ops.py:131-138
            "compute_batched_field": [ci, ci, ci, f2, f2, f1, f2],  # c_ops.py:140-149
            "compute_looped_field": [ci, ci, f2, f2, f1, f2],       # c_ops.py:151-159
            "vecaddn": [f1, f1, f1, ci],                            # c_ops.py:38-43
            # the reference declares float arrays for these two although the C takes double*
            # (c_ops.py:26-50 vs math_module.c:15,49); the C prototypes are authoritative here
            "dot": [self.array_1d_double, self.array_2d_double, self.array_1d_double, ci, ci],
            "sparse_dot": [self.array_1d_double, self.array_1d_int, ci, self.array_1d_int, ci,
                           self.array_1d_double, ci, self.array_1d_double, ci],
        }
        for name, args in legacy.items():
            fn = getattr(m, name)          # AttributeError if a symbol is missing, like the reference
            fn.restype = None
            fn.argtypes = args
        self._device = device
        self._ctx = None
        self._charges_key = None

    # error state lives in the library instance the calls go through (shared_loc may be another copy)
    def _check(self, status):
        check(status, self.math)

    def _check_legacy(self):
        check_legacy(self.math)

    # ------------------------------------------------------------------ context handling ------
    @property
    def ctx(self):
        if self._ctx is None:
            import os
            dev = self._device
            if dev is None:
                dev = int(os.environ.get("CPET_B200_DEVICE", "0"))
            h = ctypes.c_void_p()
            self._check(self.math.cpet_create(int(dev), ctypes.byref(h)))
            self._ctx = h
        return self._ctx

    def close(self):
        if self._ctx is not None:
            self.math.cpet_destroy(self._ctx)
            self._ctx = None
            self._charges_key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tuning(self, **kv):
        for k, v in kv.items():
            self._check(self.math.cpet_set_tuning(self.ctx, k.encode(), int(v)))

    def last_counters(self):
        out = (ctypes.c_int64 * 3)()
        self._check(self.math.cpet_last_counters(self.ctx, out))
        return {"launches": int(out[0]), "pair_evals": int(out[1]), "field_evals": int(out[2])}

    def last_path(self) -> str:
        """Which kernel served the last call: "general" / "lattice" (field, ESP), "k2w" / "k2x" / "k2p" (streamlines)."""
        return _lib.PATH_NAMES.get(int(self.math.cpet_last_path(self.ctx)), "general")

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_double(0.0)
        self._check(self.math.cpet_last_kernel_ms(self.ctx, ctypes.byref(ms)))
        return float(ms.value)

    def fp32_peak_tflops(self, packed=True, iters=4096) -> float:
        t = ctypes.c_double(0.0)
        self._check(self.math.cpet_fp32_peak_probe(self.ctx, int(bool(packed)), int(iters), ctypes.byref(t)))
        return float(t.value)

    # ------------------------------------------------------------------ batched entry points ---
    def set_charges(self, x, Q):
        """Upload one frame's charges (x (M,3), Q (M,) or (M,1)); reused by the calls below."""
        x = f32c(x, (-1, 3))
        Q = f32c(Q, (-1,))
        if x.shape[0] != Q.shape[0]:
            raise ValueError(f"x has {x.shape[0]} rows but Q has {Q.shape[0]} entries")
        self._check(self.math.cpet_set_charges(self.ctx, x.shape[0], ptr(x), ptr(Q)))
        return x.shape[0]

    @staticmethod
    def _out(out, shape, dtype):
        """Caller-supplied output buffer (e.g. pinned host memory) or a fresh array."""
        if out is None:
            return np.zeros(shape, dtype=dtype)
        if out.shape != tuple(shape) or out.dtype != np.dtype(dtype) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a C-contiguous {np.dtype(dtype)} array of shape {tuple(shape)}")
        return out

    def field_grid(self, x_0, x=None, Q=None, soften=True, concat=False, out=None):
        """E at points x_0 (N,3) -> (N,3) float32, or (N,6) [x_0|E] with concat=True."""
        if x is not None:
            self.set_charges(x, Q)
        x_0 = f32c(x_0, (-1, 3))
        n = x_0.shape[0]
        out = self._out(out, (n, 6 if concat else 3), np.float32)
        flags = (_lib.CPET_FIELD_SOFTEN if soften else 0) | (_lib.CPET_OUT_CONCAT if concat else 0)
        self._check(self.math.cpet_field_grid(self.ctx, n, ptr(x_0), flags, ptr(out)))
        return out

    def esp_grid(self, x_0, x=None, Q=None, concat_half=False, out=None):
        """phi at points x_0 (N,3) -> (N,) float32, or (N,4) float16 [x_0|phi] with concat_half."""
        if x is not None:
            self.set_charges(x, Q)
        x_0 = f32c(x_0, (-1, 3))
        n = x_0.shape[0]
        out = self._out(out, (n, 4), np.float16) if concat_half else self._out(out, (n,), np.float32)
        self._check(self.math.cpet_esp_grid(self.ctx, n, ptr(x_0), _lib.CPET_OUT_CONCAT if concat_half else 0,
                                      ptr(out)))
        return out

    def field_lattice(self, xs, ys, zs, x=None, Q=None, soften=True, concat=False, out=None):
        """E on the tensor-product grid xs x ys x zs (z fastest) -> (nx*ny*nz, 3 or 6) float32; ~25 % fewer
        instructions than field_grid on the expanded point list.  Bit-identical to it in the charge-pair form
        (meshes below 1e5 nodes, or k1_lat_nodes=0); the node-pair form large meshes take sums a node's charges in
        index order instead of an even and an odd chain and agrees to FP32 rounding (~1e-6)."""
        if x is not None:
            self.set_charges(x, Q)
        xs, ys, zs = f32c(xs, (-1,)), f32c(ys, (-1,)), f32c(zs, (-1,))
        n = xs.shape[0] * ys.shape[0] * zs.shape[0]
        out = self._out(out, (n, 6 if concat else 3), np.float32)
        flags = (_lib.CPET_FIELD_SOFTEN if soften else 0) | (_lib.CPET_OUT_CONCAT if concat else 0)
        self._check(self.math.cpet_field_lattice(self.ctx, xs.shape[0], ys.shape[0], zs.shape[0], ptr(xs), ptr(ys),
                                           ptr(zs), flags, ptr(out)))
        return out

    def esp_lattice(self, xs, ys, zs, x=None, Q=None, concat_half=False, out=None):
        """phi on the tensor-product grid xs x ys x zs -> (N,) float32 or (N,4) float16."""
        if x is not None:
            self.set_charges(x, Q)
        xs, ys, zs = f32c(xs, (-1,)), f32c(ys, (-1,)), f32c(zs, (-1,))
        n = xs.shape[0] * ys.shape[0] * zs.shape[0]
        out = self._out(out, (n, 4), np.float16) if concat_half else self._out(out, (n,), np.float32)
        self._check(self.math.cpet_esp_lattice(self.ctx, xs.shape[0], ys.shape[0], zs.shape[0], ptr(xs), ptr(ys),
                                         ptr(zs), _lib.CPET_OUT_CONCAT if concat_half else 0, ptr(out)))
        return out

    def propagate(self, x_0, step_size, x=None, Q=None):
        """One normalised-field step for each point: p + h E/|E| -> (N,3) float32."""
        if x is not None:
            self.set_charges(x, Q)
        x_0 = f32c(x_0, (-1, 3))
        out = np.zeros_like(x_0)
        self._check(self.math.cpet_propagate(self.ctx, x_0.shape[0], ptr(x_0), float(step_size), ptr(out)))
        return out

    def topo_batch(self, seeds, n_iter, x=None, Q=None, step_size=0.1, dimensions=(1, 1, 1),
                   second_diff=False, want_steps=False, out=None):
        """All streamlines of a frame -> (L,2) float32 [dist|curv] in seed order."""
        if x is not None:
            self.set_charges(x, Q)
        seeds = f32c(seeds, (-1, 3))
        n = seeds.shape[0]
        n_iter = np.ascontiguousarray(np.asarray(n_iter).reshape(-1), dtype=np.int32)
        if n_iter.shape[0] != n:
            raise ValueError(f"{n} seeds but {n_iter.shape[0]} n_iter entries")
        dims = f32c(dimensions, (3,))
        out = self._out(out, (n, 2), np.float32)
        steps = np.zeros(n, dtype=np.int32) if want_steps else None
        self._check(self.math.cpet_topo_batch(
            self.ctx, n, ptr(seeds), ptr(n_iter), float(step_size), ptr(dims),
            _lib.CPET_TOPO_CURV_SECOND_DIFF if second_diff else 0, ptr(out),
            ptr(steps) if want_steps else None))
        return (out, steps) if want_steps else out

    def topo_hist(self, seeds, n_iter, d_edges, c_edges, x=None, Q=None, step_size=0.1,
                  dimensions=(1, 1, 1), second_diff=False, want_rows=True, out=None, counts_out=None):
        """One frame end to end: streamlines, then their 2-D histogram, with the (L,2) rows staying
        on the device in between.  -> (rows (L,2) float32 or None, counts (nd,nc) int64)."""
        if x is not None:
            self.set_charges(x, Q)
        seeds = f32c(seeds, (-1, 3))
        n = seeds.shape[0]
        n_iter = np.ascontiguousarray(np.asarray(n_iter).reshape(-1), dtype=np.int32)
        if n_iter.shape[0] != n:
            raise ValueError(f"{n} seeds but {n_iter.shape[0]} n_iter entries")
        dims = f32c(dimensions, (3,))
        de = np.ascontiguousarray(d_edges, dtype=np.float64)
        ce = np.ascontiguousarray(c_edges, dtype=np.float64)
        nd, nc = de.shape[0] - 1, ce.shape[0] - 1
        rows = self._out(out, (n, 2), np.float32) if want_rows else None
        counts = self._out(counts_out, (nd, nc), np.int64)
        self._check(self.math.cpet_topo_hist(
            self.ctx, n, ptr(seeds), ptr(n_iter), float(step_size), ptr(dims),
            _lib.CPET_TOPO_CURV_SECOND_DIFF if second_diff else 0,
            ptr(rows) if want_rows else None, None, nd, ptr(de), nc, ptr(ce), ptr(counts)))
        return rows, counts

    def topo_hist_frames(self, frames, seeds, n_iter, d_edges, c_edges, step_size=0.1,
                         dimensions=(1, 1, 1), second_diff=False, want_rows=False, rows_out=None,
                         counts_out=None):
        """A batch of MD frames sharing seeds, box, step and bin edges.  `frames` is a sequence of
        (x (M_f,3), Q (M_f,)) pairs; `n_iter` is (L,) shared by all frames or (F,L) per frame.
        -> (rows (F,L,2) float32 or None, counts (F,nd,nc) int64); copies and kernels of neighbouring
        frames overlap (two internal streams).  Frame by frame identical to topo_hist()."""
        seeds = f32c(seeds, (-1, 3))
        n = seeds.shape[0]
        F = len(frames)
        xs = [f32c(fx, (-1, 3)) for fx, _ in frames]
        qs = [f32c(np.asarray(fq).reshape(-1)) for _, fq in frames]
        for fx, fq in zip(xs, qs):
            if fx.shape[0] != fq.shape[0]:
                raise ValueError(f"x has {fx.shape[0]} rows but Q has {fq.shape[0]} entries")
        n_iter = np.ascontiguousarray(np.asarray(n_iter), dtype=np.int32)
        if n_iter.ndim == 1:
            if n_iter.shape[0] != n:
                raise ValueError(f"{n} seeds but {n_iter.shape[0]} n_iter entries")
            stride = 0
        else:
            if n_iter.shape != (F, n):
                raise ValueError(f"n_iter must be ({n},) or ({F},{n}), got {n_iter.shape}")
            stride = n
        dims = f32c(dimensions, (3,))
        de = np.ascontiguousarray(d_edges, dtype=np.float64)
        ce = np.ascontiguousarray(c_edges, dtype=np.float64)
        nd, nc = de.shape[0] - 1, ce.shape[0] - 1
        rows = self._out(rows_out, (F, n, 2), np.float32) if want_rows else None
        counts = self._out(counts_out, (F, nd, nc), np.int64)
        m_arr = np.array([fx.shape[0] for fx in xs], dtype=np.int32)
        xp = (ctypes.c_void_p * max(F, 1))(*[fx.ctypes.data for fx in xs])
        qp = (ctypes.c_void_p * max(F, 1))(*[fq.ctypes.data for fq in qs])
        self._check(self.math.cpet_topo_hist_frames(
            self.ctx, F, ptr(m_arr), ctypes.cast(xp, ctypes.c_void_p), ctypes.cast(qp, ctypes.c_void_p),
            n, ptr(seeds), ptr(n_iter), stride, float(step_size), ptr(dims),
            _lib.CPET_TOPO_CURV_SECOND_DIFF if second_diff else 0,
            ptr(rows) if want_rows else None, nd, ptr(de), nc, ptr(ce), ptr(counts)))
        return rows, counts

    def hist2d(self, values, d_edges, c_edges):
        """Batched np.histogram2d counts.  values: (F, n, 2) or (n, 2), float64 or float32.
        -> (F, nd, nc) (or (nd, nc)) int64."""
        v = np.asarray(values)
        single = v.ndim == 2
        if single:
            v = v[None]
        if v.ndim != 3 or v.shape[2] != 2:
            raise ValueError("values must be (frames, n, 2) or (n, 2)")
        f64 = v.dtype != np.float32
        v = np.ascontiguousarray(v, dtype=np.float64 if f64 else np.float32)
        de = np.ascontiguousarray(d_edges, dtype=np.float64)
        ce = np.ascontiguousarray(c_edges, dtype=np.float64)
        nd, nc = de.shape[0] - 1, ce.shape[0] - 1
        counts = np.zeros((v.shape[0], nd, nc), dtype=np.int64)
        fn = self.math.cpet_hist2d if f64 else self.math.cpet_hist2d_f32
        self._check(fn(self.ctx, v.shape[0], v.shape[1], ptr(v), nd, ptr(de), nc, ptr(ce), ptr(counts)))
        return counts[0] if single else counts

    def order_stats(self, values, ranks, column=0):
        """Exact ranks-th smallest (0-based) entries of column `column` of a float32 (n, k) array
        (or of a flat array), selected on the device -> float32 array, NaNs ordered last."""
        v = np.ascontiguousarray(values, dtype=np.float32)
        stride = 1 if v.ndim == 1 else int(np.prod(v.shape[1:]))
        n = v.shape[0]
        r = np.ascontiguousarray(ranks, dtype=np.int64).reshape(-1)
        out = np.zeros(r.shape[0], dtype=np.float32)
        self._check(self.math.cpet_order_stats(self.ctx, n, ptr(v), stride, int(column), r.shape[0], ptr(r), ptr(out)))
        return out

    def chi2_matrix(self, H):
        H = np.ascontiguousarray(H, dtype=np.float64)
        if H.ndim != 2:
            raise ValueError("H must be (n_hists, n_bins)")
        out = np.zeros((H.shape[0], H.shape[0]), dtype=np.float64)
        self._check(self.math.cpet_chi2_matrix(self.ctx, H.shape[0], H.shape[1], ptr(H), ptr(out)))
        return out

    # ------------------------------------------------------------------ reference methods ------
    def compute_looped_field(self, x_0, x, Q):                    # c_ops.py:250-263
        res = np.zeros_like(x_0, dtype="float32")
        Q = Q.reshape(-1)
        self.math.compute_looped_field(int(x_0.shape[0]), len(Q), np.array(x_0, dtype="float32"),
                                       np.array(x, dtype="float32"), np.array(Q, dtype="float32"), res)
        self._check_legacy()
        return res

    def compute_batch_field(self, x_0, x, Q, batch_size):          # c_ops.py:265-279
        res = np.zeros_like(x_0, dtype="float32")
        Q = Q.reshape(-1)
        self.math.compute_batched_field(int(x_0.shape[0]), batch_size, len(Q),
                                        np.array(x_0, dtype="float32"), np.array(x, dtype="float32"),
                                        np.array(Q, dtype="float32"), res)
        self._check_legacy()
        return res

    def thread_operation(self, x_0, n_iter, x, Q, step_size, dimensions):   # c_ops.py:281-302
        res = np.zeros(2, dtype="float32")
        n_charges = len(Q)
        Q = Q.reshape(-1)
        self.math.thread_operation(n_charges, n_iter, step_size, x_0, dimensions, x, Q, res)
        self._check_legacy()
        return res

    def thread_operation_dipole(self, x_0, n_iter, x, mu, step_size, dimensions):  # c_ops.py:304-324
        res = np.zeros(2, dtype="float32")
        self.math.thread_operation_dipole(len(mu), n_iter, step_size, x_0, dimensions, x, mu, res)
        self._check_legacy()
        return res

    def calc_esp_base(self, x_0, x, Q):                            # c_ops.py:326-341
        res = np.zeros(1, dtype="float32")
        self.math.calc_esp_base(res, x_0, len(Q), x, Q.reshape(len(Q)))
        self._check_legacy()
        return res

    def calc_field_base(self, x_0, x, Q):                          # c_ops.py:343-358
        res = np.zeros(3, dtype="float32")
        self.math.calc_field_base(res, x_0, len(Q), x, Q.reshape(len(Q)))
        self._check_legacy()
        return res

    def calc_field(self, x_0, x, Q):                               # c_ops.py:360-375
        res = np.zeros(3, dtype="float32")
        self.math.calc_field(res, x_0, len(Q), x, Q.reshape(len(Q)))
        self._check_legacy()
        return res

    def einsum_ij_i(self, A):                                      # c_ops.py:208-212
        res = np.zeros((A.shape[0]), dtype="float32")
        self.math.einsum_ij_i(A.shape[0], A.shape[1], A, res)
        self._check_legacy()
        return res

    def einsum_ij_i_batch(self, A):                                # c_ops.py:214-220
        res = np.zeros((len(A), A[0].shape[0]), dtype="float32")
        self.math.einsum_ij_i_batch(len(A), A[0].shape[0], A[0].shape[1], A, res)
        self._check_legacy()
        return res.reshape(res.shape[1], res.shape[0])

    def einsum_operation(self, R, r_mag, Q):                       # c_ops.py:222-235
        res = np.zeros(3, dtype="float32")
        r_mag = r_mag.reshape(-1)
        R = R.reshape(r_mag.shape[0], 3)
        Q = Q.reshape(-1)
        self.math.einsum_operation(len(Q), np.array(r_mag, dtype="float32"), np.array(Q, dtype="float32"),
                                   np.array(R, dtype="float32"), res)
        self._check_legacy()
        return res

    def einsum_operation_batch(self, R, r_mag, Q, batch_size):     # c_ops.py:237-248
        res = np.zeros((batch_size, 3), dtype="float32")
        Q = Q.reshape(-1)
        self.math.einsum_operation_batch(batch_size, len(Q), np.array(r_mag, dtype="float32"),
                                         np.array(Q, dtype="float32"), np.array(R, dtype="float32"), res)
        self._check_legacy()
        return res

    def vecaddn(self, A, B):                                       # c_ops.py:197-206
        res = np.zeros(len(A), dtype="float32")
        self.math.vecaddn(res, A, B, len(A))
        self._check_legacy()
        return res

    def dot(self, A, B):
        """Dense mat-vec in float64 (math_module.c:49-68)."""
        A = np.ascontiguousarray(A, dtype=np.double)
        B = np.ascontiguousarray(B, dtype=np.double).reshape(-1)
        res = np.zeros(A.shape[0], dtype=np.double)
        self.math.dot(res, A, B, A.shape[0], A.shape[1])
        self._check_legacy()
        return res

    def sparse_dot(self, A, B):
        """CSR mat-vec in float64 (math_module.c:15-46); A is a scipy.sparse.csr_matrix."""
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ind = np.ascontiguousarray(A.indices, dtype=np.int32)
        data = np.ascontiguousarray(A.data, dtype=np.double)
        B = np.ascontiguousarray(B, dtype=np.double).reshape(-1)
        res = np.zeros(len(indptr) - 1, dtype=np.double)
        self.math.sparse_dot(res, indptr, len(indptr), ind, len(ind), data, len(data), B, len(B))
        self._check_legacy()
        return res


__all__ = ["Math_ops", "CpetError"]
