"""Synthetic inputs of SURVEY.md section 8(d): protein-like charge sets, grids, seeds."""
import numpy as np


def charges(m, seed=0, box=0.5, density=0.1):
    """M point charges uniform in a ball at ~0.1 atoms/A^3, q ~ U(-0.8, 0.8) shifted to zero net
    charge; charges inside the sampling box are dropped (mimics filter_in_box,
    CPET/source/calculator.py:349-358)."""
    rng = np.random.default_rng(seed)
    radius = (3.0 * m / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    n = int(m * 1.05) + 64
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = radius * rng.uniform(size=(n, 1)) ** (1.0 / 3.0)
    x = (v * r)
    x = x[~np.all(np.abs(x) < box * 1.05, axis=1)][:m]
    q = rng.uniform(-0.8, 0.8, size=len(x))
    q -= q.mean()
    return np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(q, dtype=np.float32)


def grid(n_per_axis, half):
    """Inclusive box grid, last axis fastest (initialize_box_points_uniform, UC:218-233)."""
    c = np.linspace(-half, half, n_per_axis)
    g = np.stack(np.meshgrid(c, c, c, indexing="ij"), axis=-1)
    return np.ascontiguousarray(g.reshape(-1, 3), dtype=np.float32)


def seeds(n_per_axis, half, step, rng_seed=42):
    """Non-inclusive uniform seeds + n_iter = RandomState(42).randint(1, max_steps, n)
    (UC:222-245 with max_streamline_init == 'fixed_rand'; max_steps per SC:272)."""
    c = np.linspace(-half, half, n_per_axis + 1, endpoint=False)[1:]
    g = np.stack(np.meshgrid(c, c, c, indexing="ij"), axis=-1).reshape(-1, 3)
    dims = np.array([half, half, half], dtype=np.float64)
    max_steps = round(2 * np.linalg.norm(dims) / step)
    n_iter = np.random.RandomState(rng_seed).randint(1, max_steps, len(g))
    return (np.ascontiguousarray(g, dtype=np.float32), n_iter.astype(np.int64),
            dims.astype(np.float32), int(max_steps))
