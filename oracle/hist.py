"""oracle.hist (TEST INFRASTRUCTURE) -- histogram + chi^2 distance restatements.

The reference delegates the binning arithmetic to a third-party dependency that is not vendored
in /root/reference: **NumPy** ``np.histogram2d`` (called at CPET/utils/calculator.py:702-707;
reference pins ``numpy<2`` in pyproject.toml:19 without a lock file, CPET_ENV.yml lists 1.25.0)
and **SciPy** ``scipy.stats.iqr`` (calculator.py:669-670).  The published algorithm
(numpy/lib/_histograms_impl.py::histogramdd, unchanged between 1.25 and 2.3) is restated here:

  edges_d = linspace(lo_d, hi_d, n_d + 1)                      (float64)
  idx_d   = searchsorted(edges_d, v_d, side="right")           (0 .. n_d+1, 0/n_d+1 = outliers)
  idx_d  -= 1 where v_d == edges_d[-1]                         (right edge is inclusive)
  counts  = bincount(ravel_multi_index(idx, n+2))[1:-1, 1:-1]  (outliers dropped)

No reference test pins histogram VALUES ("parity unpinned by the reference's tests"); the pin is
NumPy itself: tests/test_oracle.py asserts this restatement == np.histogram2d bit-for-bit.
"""
from __future__ import annotations

import numpy as np


def edges(lo: float, hi: float, nbins: int) -> np.ndarray:
    return np.linspace(float(lo), float(hi), int(nbins) + 1)


def hist2d_counts(d, c, nd: int, nc: int, d_range, c_range) -> np.ndarray:
    """Integer counts (nd, nc) exactly as np.histogram2d(d, c, bins=[nd,nc], range=[d_range,c_range])."""
    d = np.asarray(d, dtype=np.float64).reshape(-1)
    c = np.asarray(c, dtype=np.float64).reshape(-1)
    ed = edges(d_range[0], d_range[1], nd)
    ec = edges(c_range[0], c_range[1], nc)
    i = np.searchsorted(ed, d, side="right")
    j = np.searchsorted(ec, c, side="right")
    i[d == ed[-1]] -= 1
    j[c == ec[-1]] -= 1
    flat = np.ravel_multi_index((i, j), (nd + 2, nc + 2))
    full = np.bincount(flat, minlength=(nd + 2) * (nc + 2)).reshape(nd + 2, nc + 2)
    return full[1:-1, 1:-1].astype(np.int64)


def bin_plan(dist_all, curv_all, n_per_frame: float):
    """Global ranges + bin counts the way make_histograms derives them
    (CPET/utils/calculator.py:664-685): Freedman-Diaconis-like width 2*IQR/n^(1/3) with
    n = lines per frame, nbins = int(range / width)."""
    from scipy.stats import iqr

    dist_all = np.asarray(dist_all, dtype=np.float64)
    curv_all = np.asarray(curv_all, dtype=np.float64)
    dmin, dmax = float(np.min(dist_all)), float(np.max(dist_all))
    cmin, cmax = float(np.min(curv_all)), float(np.max(curv_all))
    dres = 2 * iqr(dist_all) / (n_per_frame ** (1 / 3))
    cres = 2 * iqr(curv_all) / (n_per_frame ** (1 / 3))
    nd = int((dmax - dmin) / dres)
    nc = int((cmax - cmin) / cres)
    return (dmin, dmax), (cmin, cmax), nd, nc


def normalised_hist(d, c, nd, nc, d_range, c_range) -> np.ndarray:
    """`a / a.sum()` flattened row-major (calculator.py:709-713)."""
    a = hist2d_counts(d, c, nd, nc, d_range, c_range).astype(np.float64)
    return (a / np.sum(a)).flatten()


def chi2(h1, h2) -> float:
    """distance_numpy (CPET/utils/calculator.py:975-978; duplicated in the reference's
    tests/test_topology.py:46-49): 1/2 * sum_{h1+h2 != 0} (h1-h2)^2 / (h1+h2)."""
    h1 = np.asarray(h1, dtype=np.float64)
    h2 = np.asarray(h2, dtype=np.float64)
    a = (h1 - h2) ** 2
    b = h1 + h2
    return float(np.sum(np.divide(a, b, out=np.zeros_like(a), where=b != 0)) / 2.0)


def chi2_matrix(H) -> np.ndarray:
    """construct_distance_matrix (calculator.py:1003-1015): symmetric, zero diagonal."""
    H = np.asarray(H, dtype=np.float64)
    n = H.shape[0]
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            out[i, j] = out[j, i] = chi2(H[i], H[j])
    return out
