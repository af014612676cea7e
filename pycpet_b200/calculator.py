"""Calculator-level entry points of the hot path, same names and return layouts as the reference
(CPET/utils/calculator.py free functions, CPET/source/calculator.py ``compute_*`` methods), backed
by libcpetb200.so.  PDB/PQR parsing, atom filtering and the box-frame transform are NOT here: they
stay in PyCPET's own Python (north_star); these functions start from the arrays that
``calculator.__init__`` produces (``x``, ``Q``, ``mesh``, ``random_start_points``,
``random_max_samples``, ``dimensions``, ``step_size``).

Drop-in use with an unmodified PyCPET checkout: ``pycpet_b200.patch_reference()`` (see
INTEGRATION.md) rebinds ``CPET.utils.calculator.Math`` and the four ``calculator.compute_*``
methods to the functions below.
"""
from __future__ import annotations

import sys

import numpy as np

from .c_ops import Math_ops

_MATH = None


def get_math() -> Math_ops:
    """Process-wide ``Math_ops`` handle (the reference's module-level ``Math``,
    CPET/utils/calculator.py:29), created on first use so that forked workers never inherit a
    CUDA context."""
    global _MATH
    if _MATH is None:
        _MATH = Math_ops()
    return _MATH


def __getattr__(name):
    if name == "Math":
        return get_math()
    raise AttributeError(name)


# ------------------------------------------------------------------------------------------------
# free functions (CPET/utils/calculator.py)
# ------------------------------------------------------------------------------------------------
def lattice_axes(grid_coords):
    """If grid_coords is the (nx,ny,nz,3) float32 tensor-product mesh that
    initialize_box_points_uniform builds (UC:218-233: meshgrid(indexing="ij") of three coordinate
    vectors, last axis fastest), return its axis vectors (xs, ys, zs); otherwise None.  The check is
    exact (every node compared), so the lattice kernel sees precisely the points the general kernel
    would."""
    g = np.asarray(grid_coords)
    if g.ndim != 4 or g.shape[3] != 3 or g.dtype != np.float32 or min(g.shape[:3]) < 1:
        return None
    xs, ys, zs = g[:, 0, 0, 0], g[0, :, 0, 1], g[0, 0, :, 2]
    if not (np.array_equal(g[..., 0], np.broadcast_to(xs[:, None, None], g.shape[:3])) and
            np.array_equal(g[..., 1], np.broadcast_to(ys[None, :, None], g.shape[:3])) and
            np.array_equal(g[..., 2], np.broadcast_to(zs[None, None, :], g.shape[:3]))):
        return None
    return np.ascontiguousarray(xs), np.ascontiguousarray(ys), np.ascontiguousarray(zs)


def compute_field_on_grid(grid_coords, x, Q):
    """UC:430-447.  grid_coords (..., 3) -> (N,6) float32 rows [point | E] with the `volume`
    softening max(r^2, 1e-6); one kernel launch instead of one C loop.  The library recognises
    box meshes in the flat point list on the device and gives them the lattice kernel (same
    numbers, fewer instructions); anything else takes the general kernel."""
    x_0 = np.asarray(grid_coords).reshape(-1, 3)
    return get_math().field_grid(x_0, x, Q, soften=True, concat=True)


def compute_ESP_on_grid(grid_coords, x, Q):
    """UC:450-475.  -> (N,4) float16 rows [point | ESP]; replaces the Python loop with one
    ctypes call per grid point."""
    x_0 = np.asarray(grid_coords).reshape(-1, 3)
    return get_math().esp_grid(x_0, x, Q, concat_half=True)


def calculate_electric_field_c_shared_full_alt(x_0, x, Q):
    """UC:296-310 -> (3,) float32 field at one point (no softening)."""
    x_0 = np.asarray(x_0).astype(np.float32)
    x = np.asarray(x).astype(np.float32)
    Q = np.asarray(Q).astype(np.float32)
    return get_math().calc_field(x_0=x_0, x=x, Q=Q)


def calculate_esp_c_shared_full(x_0, x, Q):
    """UC:330-341 -> (1,) float32 potential at one point."""
    return get_math().calc_esp_base(x_0=x_0, x=x, Q=Q)


def calculate_thread_c_shared(x_0, n_iter, x, Q, step_size, dimensions):
    """UC:344-348 -> (2,) float32 [dist, curv] of one streamline."""
    return get_math().thread_operation(x_0=x_0, n_iter=n_iter, x=x, Q=Q, step_size=step_size,
                                       dimensions=dimensions)


def calculate_thread_c_shared_dipole(x_0, n_iter, x, mu, step_size, dimensions):
    """UC:351-358 (marked in-development in the reference; used by CPET/utils/parallel.py:121):
    one streamline in the field of point dipoles mu at x -> (2,) float32 [dist, curv]."""
    return get_math().thread_operation_dipole(x_0=x_0, n_iter=n_iter, x=x, mu=mu, step_size=step_size,
                                              dimensions=dimensions)


def propagate_topo(x_0, x, Q, step_size, debug=False):
    """One normalised-field step p + h E(p)/|E(p)| on the device (the C propagate_topo,
    math_module.c:489-503; UC:35-54 is its Python twin).  x_0 (3,) or (N,3) -> same shape."""
    p = np.asarray(x_0, dtype=np.float32)
    out = get_math().propagate(p.reshape(-1, 3), step_size, x, Q)
    return out.reshape(p.shape)


def compute_topo_batch(seeds, n_iter, x, Q, step_size, dimensions, second_diff=False,
                       want_steps=False):
    """The whole-frame call that replaces both Pool.starmap(task_complete_thread, ...)
    (SC:690-704) and the torch window/filter loop (SC:866-937): (L,2) float32 [dist|curv] in
    seed order."""
    return get_math().topo_batch(seeds, n_iter, x, Q, step_size, dimensions,
                                 second_diff=second_diff, want_steps=want_steps)


def compute_curv_and_dist(x_init, x_init_plus, x_init_plus_plus, x_0, x_0_plus, x_0_plus_plus):
    """UC:541-562 -> [dist, curv_mean] from the first three and the last three points of ONE
    streamline: kappa = |v' x v''| / |v'|^3 with v' = a1 - a0, v'' = a2 - 2 a1 + a0 (UC:500-521,
    eps = 10e-6 replaces a zero denominator), mean of both ends, dist = |x_init - x_0|.
    Six points of host arithmetic in the caller's dtype, exactly as the reference writes it; the
    integrator applies the same two formulas per line on the device (topo_batch, second_diff=True
    for this literal second-difference form)."""
    def curv(v1, v2, eps=10e-6):
        den = np.linalg.norm(v1, axis=0) ** 3
        num = np.linalg.norm(np.cross(v1, v2), axis=0)
        return num / eps if den == 0 else num / den

    curv_init = curv(x_init_plus - x_init, x_init_plus_plus - 2 * x_init_plus + x_init)
    curv_final = curv(x_0_plus - x_0, x_0_plus_plus - 2 * x_0_plus + x_0)
    return [np.linalg.norm(x_init - x_0, axis=-1), (curv_init + curv_final) / 2]


def distance_numpy(hist1, hist2):
    """UC:975-978: chi^2 distance between two flattened normalised histograms (device)."""
    H = np.stack([np.asarray(hist1, dtype=np.float64).ravel(), np.asarray(hist2, dtype=np.float64).ravel()])
    return float(get_math().chi2_matrix(H)[0, 1])


def construct_distance_matrix(histograms):
    """UC:1003-1015: symmetric pairwise chi^2 matrix with zero diagonal."""
    return get_math().chi2_matrix(np.asarray(histograms, dtype=np.float64))


def histogram2d_counts(values, nd, nc, d_range, c_range):
    """np.histogram2d(dist, curv, bins=[nd,nc], range=[d_range,c_range]) counts on the device,
    bit-exact with NumPy.  values: (n,2) or (F,n,2) [dist|curv]."""
    d_edges = np.linspace(float(d_range[0]), float(d_range[1]), int(nd) + 1)
    c_edges = np.linspace(float(c_range[0]), float(c_range[1]), int(nc) + 1)
    return get_math().hist2d(values, d_edges, c_edges)


def read_top_file(path):
    """Parse a .top file (two columns dist, curv; '#' comments) into an (n,2) float64 array: the
    values the reference's `float(line.split()[k])` loops produce (UC:626-633), read by
    cpet_read_rows on all host cores."""
    from .io import read_topology

    return read_topology(path)


def bin_plan(topologies):
    """Global ranges and bin counts as make_histograms derives them (UC:640-685): min/max over all
    frames, bin width 2*IQR/n^(1/3) with n = lines per frame (mean if ragged)."""
    from scipy.stats import iqr
    import warnings

    lens = np.array([len(t) for t in topologies])
    if lens.size and not np.all(lens == lens[0]):
        warnings.warn(f"Topologies provided are of different sizes, using the mean value of "
                      f"{np.mean(lens)} to represent binning for {lens}")
        n_ref = np.mean(lens)
    else:
        n_ref = lens[0]
    dist_all = np.concatenate([np.asarray(t, dtype=np.float64)[:, 0] for t in topologies])
    curv_all = np.concatenate([np.asarray(t, dtype=np.float64)[:, 1] for t in topologies])
    d_range = (float(np.min(dist_all)), float(np.max(dist_all)))
    c_range = (float(np.min(curv_all)), float(np.max(curv_all)))
    d_res = 2 * iqr(dist_all) / (n_ref ** (1 / 3))
    c_res = 2 * iqr(curv_all) / (n_ref ** (1 / 3))
    nd = int((d_range[1] - d_range[0]) / d_res)
    nc = int((c_range[1] - c_range[0]) / c_res)
    return d_range, c_range, nd, nc


def quartile_ranks(n):
    """0-based ranks of the order statistics np.percentile(x, [25, 75]) (default 'linear' method)
    interpolates between, plus the interpolation weights, using NumPy's own virtual-index formula
    n*q + (alpha + q*(1 - alpha - beta)) - 1 with alpha = beta = 1."""
    out = []
    for q in (0.25, 0.75):
        virtual = n * q + (1 + q * (1 - 1 - 1)) - 1
        prev = int(np.floor(virtual))
        out.append((prev, min(prev + 1, n - 1), float(virtual - prev)))
    return out


def _lerp(a, b, t):
    """numpy.lib._function_base_impl._lerp for scalars (float64)."""
    a, b = np.float64(a), np.float64(b)
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def plan_from_order_stats(stats, n_ref):
    """(d_range, c_range, nd, nc) from the six order statistics per column
    [min, q25_lo, q25_hi, q75_lo, q75_hi, max] and the interpolation weights of quartile_ranks."""
    ranges, nbins = [], []
    for col in (0, 1):
        (lo, a25, b25, a75, b75, hi), (g25, g75) = stats[col]
        iqr = _lerp(a75, b75, g75) - _lerp(a25, b25, g25)
        res = 2 * iqr / (n_ref ** (1 / 3))
        lo, hi = float(np.float64(lo)), float(np.float64(hi))
        ranges.append((lo, hi))
        nbins.append(int((hi - lo) / res))
    return ranges[0], ranges[1], nbins[0], nbins[1]


def bin_plan_device(topologies):
    """bin_plan() with the global min / max and the exact inter-quartile range taken from
    device-side radix selection (no host sort): same return value as bin_plan(), bit for bit.
    topologies: (F, n, 2) float32 array or a list of (n_i, 2) float32 arrays."""
    tops = [np.ascontiguousarray(t, dtype=np.float32).reshape(-1, 2) for t in topologies]
    lens = np.array([len(t) for t in tops])
    n_ref = lens[0] if np.all(lens == lens[0]) else np.mean(lens)
    allv = np.concatenate(tops) if len(tops) > 1 else tops[0]
    n = len(allv)
    (p25, n25, g25), (p75, n75, g75) = quartile_ranks(n)
    ranks = [0, p25, n25, p75, n75, n - 1]
    m = get_math()
    stats = [(tuple(m.order_stats(allv, ranks, column=c).tolist()), (g25, g75)) for c in (0, 1)]
    return plan_from_order_stats(stats, n_ref)


def make_histograms_from_arrays(topologies, plan=None):
    """Histograms of a list of (n_i,2) [dist|curv] arrays -> (F, nd*nc) float64, each row
    a/a.sum() flattened row-major (UC:702-713).  Frames of equal length go to the GPU as one
    batched launch; ragged frames one launch each."""
    tops = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 2) for t in topologies]
    if plan is None:
        # .top values are float32 results printed with %.18e (CPET.py:123), so the float64 the reference parses
        # are float32-exact: the global min / max / quartiles then come from the device radix select
        # (bin_plan_device == bin_plan bit for bit); anything else keeps the reference's host rule
        t32 = [t.astype(np.float32) for t in tops]
        exact = all(np.array_equal(a.astype(np.float64), t) for a, t in zip(t32, tops))
        plan = bin_plan_device(t32) if exact and sum(len(t) for t in tops) > 0 else bin_plan(topologies)
    d_range, c_range, nd, nc = plan
    lens = {len(t) for t in tops}
    if len(lens) == 1:
        counts = histogram2d_counts(np.stack(tops), nd, nc, d_range, c_range)
    else:
        counts = np.stack([histogram2d_counts(t, nd, nc, d_range, c_range) for t in tops])
    a = counts.astype(np.float64).reshape(len(tops), -1)
    return a / a.sum(axis=1, keepdims=True)


def make_histograms(topo_files, plot=False):
    """UC:596-718 with the same signature: list of .top paths -> (F, nd*nc) float64."""
    return make_histograms_from_arrays([read_top_file(f) for f in topo_files])


# ------------------------------------------------------------------------------------------------
# calculator-level methods (CPET/source/calculator.py); `calc` is any object carrying the
# attributes the reference's calculator.__init__ sets.
# ------------------------------------------------------------------------------------------------
def compute_point_field(calc):
    """SC:442-463: field at the (already centred) origin -> (3,) float32."""
    return calculate_electric_field_c_shared_full_alt(np.array([0, 0, 0]), calc.x, calc.Q)


def compute_box(calc):
    """SC:489-507 -> ((N,6) float32 [point|E], mesh.shape)."""
    return compute_field_on_grid(calc.mesh, calc.x, calc.Q), calc.mesh.shape


def compute_box_ESP(calc):
    """SC:509-528 -> ((N,4) float16 [point|ESP], mesh.shape)."""
    return compute_ESP_on_grid(calc.mesh, calc.x, calc.Q), calc.mesh.shape


def _second_diff(calc) -> bool:
    """Options key ``"curvature"`` (ours; unknown keys pass the reference's option checks, IO:200-266):
    ``"direction"`` (default) = |e0 x e1| / h from consecutive unit field directions,
    ``"second_diff"`` = the reference's literal FP32 second-difference formula (C:575-580,
    108-121).  patch_reference() records it on the calculator object."""
    mode = str(getattr(calc, "b200_curvature", "direction"))
    if mode not in ("direction", "second_diff"):
        raise ValueError("options['curvature'] must be 'direction' or 'second_diff', got %r" % mode)
    return mode == "second_diff"


def compute_topo_complete_c_shared(calc):
    """SC:675-712 -> (n_samples,2) [dist|curv] in seed order; also stored as calc.hist."""
    hist = compute_topo_batch(calc.random_start_points, calc.random_max_samples, calc.x, calc.Q,
                              calc.step_size, calc.dimensions, second_diff=_second_diff(calc))
    try:
        calc.hist = hist
    except Exception:
        pass
    return hist


def compute_topo_GPU_batch_filter(calc):
    """SC:793-978 -> (n_samples,2) [dist|curv].  The reference returns rows in dump order; this
    returns seed order (a permutation of the same rows; the reference's own tests sort rows before
    comparing, tests/test_topology.py:80-95)."""
    return compute_topo_batch(calc.random_start_points, calc.random_max_samples, calc.x, calc.Q,
                              calc.step_size, calc.dimensions, second_diff=_second_diff(calc))


def patch_reference():
    """Rebind an importable, unmodified PyCPET to this backend: the module-level ``Math`` of
    CPET.utils.calculator plus the hot-path ``compute_*`` methods of the calculator class."""
    import CPET.utils.calculator as UC      # noqa: N811  (raises ImportError if PyCPET is absent)
    import CPET.source.calculator as SC     # noqa: N811

    UC.Math = get_math()
    UC.compute_field_on_grid = compute_field_on_grid
    UC.compute_ESP_on_grid = compute_ESP_on_grid
    UC.distance_numpy = distance_numpy
    UC.construct_distance_matrix = construct_distance_matrix
    UC.make_histograms = make_histograms
    SC.compute_field_on_grid = compute_field_on_grid
    SC.compute_ESP_on_grid = compute_ESP_on_grid
    # CPET/source/cluster.py:14-22 (and the benchmark scripts) import the histogram functions BY NAME,
    # so rebinding the attributes of CPET.utils.calculator alone would leave the dispatcher's
    # `cluster` method on the CPU parse loops and the sklearn pairwise distance
    for modname in ("CPET.source.cluster", "CPET.source.scripts.benchmark_sample_step",
                    "CPET.source.scripts.benchmark_sample_step_dipole"):
        mod = sys.modules.get(modname)
        if mod is None and modname == "CPET.source.cluster":
            try:
                import importlib

                mod = importlib.import_module(modname)
            except ImportError:
                mod = None
        if mod is None:
            continue
        for fname, fn in (("make_histograms", make_histograms),
                          ("construct_distance_matrix", construct_distance_matrix),
                          ("distance_numpy", distance_numpy)):
            if hasattr(mod, fname):
                setattr(mod, fname, fn)
    cls = SC.calculator
    if not getattr(cls.__init__, "_b200_wrapped", False):
        ref_init = cls.__init__

        def init_recording_curvature(self, options, *args, **kwargs):
            ref_init(self, options, *args, **kwargs)
            self.b200_curvature = options.get("curvature", "direction") if hasattr(options, "get") else "direction"

        init_recording_curvature._b200_wrapped = True
        init_recording_curvature.__doc__ = ref_init.__doc__
        cls.__init__ = init_recording_curvature
    cls.compute_box = compute_box
    cls.compute_box_ESP = compute_box_ESP
    cls.compute_topo_complete_c_shared = compute_topo_complete_c_shared
    cls.compute_topo_GPU_batch_filter = compute_topo_GPU_batch_filter
    cls.compute_point_field = compute_point_field
    # the `.dat` writer (CPET/utils/io.py:50-109), also where CPET.py imported it by name
    try:
        import CPET.source.CPET as TOP  # noqa: N811
        import CPET.utils.io as IO      # noqa: N811

        from .io import save_numpy_as_dat

        IO.save_numpy_as_dat = save_numpy_as_dat
        if hasattr(TOP, "save_numpy_as_dat"):
            TOP.save_numpy_as_dat = save_numpy_as_dat
    except ImportError:
        pass
    return cls
