"""Device-pointer API: torch CUDA tensors in, torch CUDA tensors out, no host copies, no syncs.

PyTorch is plumbing only here (device memory, streams, torch.distributed buffers); the math is
the ``*_dev`` entry points of libcpetb200.so, run on torch's current stream.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class Engine:
    """One context bound to (device, torch's current stream on that device)."""

    def __init__(self, device=None, stream=None):
        import torch

        if not torch.cuda.is_available():
            raise _lib.CpetError("Engine needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
            h = ctypes.c_void_p()
            check(self.lib.cpet_create_on_stream(self.device.index, ctypes.c_void_p(self.stream.cuda_stream),
                                                 ctypes.byref(h)))
        self.ctx = h
        self.n_charges = 0

    def close(self):
        if getattr(self, "ctx", None) is not None:
            self.lib.cpet_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _f32(self, t, cols=None):
        torch = self.torch
        t = torch.as_tensor(t, device=self.device) if not torch.is_tensor(t) else t
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if cols is not None:
            t = t.reshape(-1, cols)
        return t

    def set_tuning(self, **kv):
        for k, v in kv.items():
            check(self.lib.cpet_set_tuning(self.ctx, k.encode(), int(v)))

    def last_counters(self):
        out = (ctypes.c_int64 * 3)()
        check(self.lib.cpet_last_counters(self.ctx, out))
        return {"launches": int(out[0]), "pair_evals": int(out[1]), "field_evals": int(out[2])}

    def last_path(self) -> str:
        """Which kernel served the last call: "general" / "lattice" (field, ESP), "k2w" / "k2x" / "k2p" (streamlines)."""
        return _lib.PATH_NAMES.get(int(self.lib.cpet_last_path(self.ctx)), "general")

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_double(0.0)
        check(self.lib.cpet_last_kernel_ms(self.ctx, ctypes.byref(ms)))
        return float(ms.value)

    def kernel_times(self):
        """Durations (ms) of the dominant-kernel launches since the last call (timing=1)."""
        buf = (ctypes.c_double * 256)()
        n = ctypes.c_int(0)
        check(self.lib.cpet_kernel_times(self.ctx, buf, 256, ctypes.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def fp32_peak_tflops(self, packed=True, iters=4096) -> float:
        t = ctypes.c_double(0.0)
        check(self.lib.cpet_fp32_peak_probe(self.ctx, int(bool(packed)), int(iters), ctypes.byref(t)))
        return float(t.value)

    # -- the path --------------------------------------------------------------------------------
    def set_charges(self, x, Q):
        x = self._f32(x, 3)
        Q = self._f32(Q).reshape(-1)
        assert x.shape[0] == Q.shape[0]
        check(self.lib.cpet_set_charges_dev(self.ctx, x.shape[0], _p(x), _p(Q)))
        self.n_charges = x.shape[0]
        self._keep = (x, Q)          # the pack kernel reads them asynchronously

    def field_grid(self, x0, soften=True, concat=False, out=None):
        torch = self.torch
        x0 = self._f32(x0, 3)
        n = x0.shape[0]
        if out is None:
            out = torch.empty((n, 6 if concat else 3), dtype=torch.float32, device=self.device)
        flags = (_lib.CPET_FIELD_SOFTEN if soften else 0) | (_lib.CPET_OUT_CONCAT if concat else 0)
        check(self.lib.cpet_field_grid_dev(self.ctx, n, _p(x0), flags, _p(out)))
        return out

    def esp_grid(self, x0, concat_half=False, out=None):
        torch = self.torch
        x0 = self._f32(x0, 3)
        n = x0.shape[0]
        if out is None:
            out = (torch.empty((n, 4), dtype=torch.float16, device=self.device) if concat_half
                   else torch.empty(n, dtype=torch.float32, device=self.device))
        check(self.lib.cpet_esp_grid_dev(self.ctx, n, _p(x0), _lib.CPET_OUT_CONCAT if concat_half else 0,
                                         _p(out)))
        return out

    def field_lattice(self, xs, ys, zs, soften=True, concat=False, out=None):
        torch = self.torch
        xs, ys, zs = self._f32(xs).reshape(-1), self._f32(ys).reshape(-1), self._f32(zs).reshape(-1)
        n = xs.shape[0] * ys.shape[0] * zs.shape[0]
        if out is None:
            out = torch.empty((n, 6 if concat else 3), dtype=torch.float32, device=self.device)
        flags = (_lib.CPET_FIELD_SOFTEN if soften else 0) | (_lib.CPET_OUT_CONCAT if concat else 0)
        check(self.lib.cpet_field_lattice_dev(self.ctx, xs.shape[0], ys.shape[0], zs.shape[0], _p(xs), _p(ys),
                                              _p(zs), flags, _p(out)))
        self._keep4 = (xs, ys, zs)
        return out

    def esp_lattice(self, xs, ys, zs, concat_half=False, out=None):
        torch = self.torch
        xs, ys, zs = self._f32(xs).reshape(-1), self._f32(ys).reshape(-1), self._f32(zs).reshape(-1)
        n = xs.shape[0] * ys.shape[0] * zs.shape[0]
        if out is None:
            out = (torch.empty((n, 4), dtype=torch.float16, device=self.device) if concat_half
                   else torch.empty(n, dtype=torch.float32, device=self.device))
        check(self.lib.cpet_esp_lattice_dev(self.ctx, xs.shape[0], ys.shape[0], zs.shape[0], _p(xs), _p(ys),
                                            _p(zs), _lib.CPET_OUT_CONCAT if concat_half else 0, _p(out)))
        self._keep4 = (xs, ys, zs)
        return out

    def propagate(self, x0, step_size, out=None):
        torch = self.torch
        x0 = self._f32(x0, 3)
        if out is None:
            out = torch.empty_like(x0)
        check(self.lib.cpet_propagate_dev(self.ctx, x0.shape[0], _p(x0), float(step_size), _p(out)))
        return out

    def topo_batch(self, seeds, n_iter, step_size, dimensions, second_diff=False, out=None,
                   steps=None, want_steps=False):
        torch = self.torch
        seeds = self._f32(seeds, 3)
        n = seeds.shape[0]
        if not torch.is_tensor(n_iter):
            n_iter = torch.as_tensor(np.asarray(n_iter).astype(np.int32), device=self.device)
        if n_iter.dtype != torch.int32 or n_iter.device != self.device:
            n_iter = n_iter.to(device=self.device, dtype=torch.int32)
        n_iter = n_iter.contiguous().reshape(-1)
        assert n_iter.shape[0] == n
        dims = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(dimensions).reshape(3)])
        if out is None:
            out = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        if want_steps and steps is None:
            steps = torch.empty(n, dtype=torch.int32, device=self.device)
        check(self.lib.cpet_topo_batch_dev(
            self.ctx, n, _p(seeds), _p(n_iter), float(step_size), dims,
            _lib.CPET_TOPO_CURV_SECOND_DIFF if second_diff else 0, _p(out),
            _p(steps) if steps is not None else None))
        self._keep2 = (seeds, n_iter)
        return (out, steps) if want_steps else out

    def radix_hist(self, values, column, prefixes, prefix_bits):
        """One radix-select pass over column `column` of a float32 CUDA tensor (n, k):
        -> (len(prefixes), 256) uint64 histogram of the next 8 key bits (host array)."""
        v = self._f32(values)
        stride = 1 if v.dim() == 1 else int(v.shape[1])
        pre = np.ascontiguousarray(prefixes, dtype=np.uint32).reshape(-1)
        hist = np.zeros((pre.shape[0], 256), dtype=np.uint64)
        check(self.lib.cpet_radix_hist_dev(self.ctx, int(v.shape[0]), _p(v), stride, int(column), pre.shape[0],
                                           _lib.ptr(pre), int(prefix_bits), _lib.ptr(hist)))
        return hist

    def order_stats(self, values, ranks, column=0):
        """Exact ranks-th smallest entries (0-based) of a float32 CUDA tensor column -> float32 ndarray."""
        v = self._f32(values)
        stride = 1 if v.dim() == 1 else int(v.shape[1])
        r = np.ascontiguousarray(ranks, dtype=np.int64).reshape(-1)
        out = np.zeros(r.shape[0], dtype=np.float32)
        check(self.lib.cpet_order_stats_dev(self.ctx, int(v.shape[0]), _p(v), stride, int(column), r.shape[0],
                                            _lib.ptr(r), _lib.ptr(out)))
        return out

    def hist2d(self, values, d_edges, c_edges, out=None):
        """values: (F, n, 2) or (n, 2) float32 CUDA tensor -> (F, nd, nc) / (nd, nc) int64."""
        torch = self.torch
        v = self._f32(values)
        single = v.dim() == 2
        if single:
            v = v.unsqueeze(0)
        assert v.dim() == 3 and v.shape[2] == 2
        de = np.ascontiguousarray(d_edges, dtype=np.float64)
        ce = np.ascontiguousarray(c_edges, dtype=np.float64)
        nd, nc = de.shape[0] - 1, ce.shape[0] - 1
        if out is None:
            out = torch.empty((v.shape[0], nd, nc), dtype=torch.int64, device=self.device)
        check(self.lib.cpet_hist2d_dev(self.ctx, v.shape[0], v.shape[1], _p(v), nd, _lib.ptr(de), nc,
                                       _lib.ptr(ce), _p(out)))
        self._keep3 = (v, de, ce)
        return out[0] if single else out

    def chi2_rows(self, hists, row0=0, n_rows=None, out=None):
        """Rows [row0, row0 + n_rows) of the pairwise chi^2 distance matrix (UC:975-978, 1003-1015) of
        device-resident float64 histograms (F, n_bins) -> (n_rows, F) float64 CUDA tensor."""
        torch = self.torch
        h = hists
        if h.dtype != torch.float64 or not h.is_contiguous() or h.device != self.device:
            h = h.to(device=self.device, dtype=torch.float64).contiguous()
        h = h.reshape(h.shape[0], -1)
        f = int(h.shape[0])
        n_rows = f - row0 if n_rows is None else int(n_rows)
        if out is None:
            out = torch.empty((n_rows, f), dtype=torch.float64, device=self.device)
        check(self.lib.cpet_chi2_rows_dev(self.ctx, f, int(h.shape[1]), _p(h), int(row0), n_rows, _p(out)))
        self._keep5 = h
        return out
