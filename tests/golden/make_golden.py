#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root:   python tests/golden/make_golden.py
Needs /root/reference (read-only).  The GPU box has no reference tree, which is why the
outputs are committed as small fixtures next to this script.

What is driven (all through the reference's own Python, nothing re-implemented here):
  * CPET.source.calculator.calculator.__init__  -- PDB parse, filters, frame transform, mesh /
    seed generation (CPET/source/calculator.py:68-415)
  * calculator.compute_point_field / compute_box / compute_topo_complete_c_shared
    (calculator.py:442-463, 489-507, 675-712)
  * CPET.utils.c_ops.Math_ops on the .so the reference ships for this Python
    (CPET/utils/math_module.cpython-312-x86_64-linux-gnu.so): compute_looped_field,
    calc_field, calc_field_base, calc_esp_base, thread_operation (c_ops.py:250-375)

matplotlib / seaborn / kneed / tensorly / sklearn_extra are imported at module scope by the
reference but are absent from this image; they are replaced by empty stub modules (none of them
is touched by the functions above).
"""
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("CPET_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors = _stub("matplotlib.colors", LinearSegmentedColormap=object, Normalize=object)
    mpl.cm = _stub("matplotlib.cm")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("seaborn")
    _stub("kneed", KneeLocator=object)
    tl = _stub("tensorly")
    tl.decomposition = _stub("tensorly.decomposition", parafac=None, non_negative_parafac=None)
    se = _stub("sklearn_extra")
    se.cluster = _stub("sklearn_extra.cluster", KMedoids=object)


def main():
    os.environ["CPET_BANNER"] = "0"
    install_stubs()
    sys.path.insert(0, REF)
    from CPET.source.calculator import calculator  # noqa: E402
    from CPET.utils.calculator import Math  # the reference's module-level Math_ops singleton
    import CPET.utils.calculator as UC

    ex = os.path.join(REF, "examples")
    pdb = os.path.join(ex, "1A_point-field", "pdb", "1_alcdehydro_run1.pdb")
    meta = {"reference_so": os.path.basename(UC.module_path), "numpy": np.__version__}

    # ---------------- example 1A: point_field (the only exact shipped known-answer) ----------
    opts = json.load(open(os.path.join(ex, "1A_point-field", "options", "options.json")))
    opts["inputpath"] = os.path.dirname(pdb)
    opts["outputpath"] = "/tmp/_golden_out_1A"
    calc = calculator(opts, path_to_pdb=pdb)
    pf = calc.compute_point_field()
    shipped = open(os.path.join(ex, "1A_point-field", "outdir", "point_field.dat")).readline()
    shipped = np.array(shipped.split(":")[1].strip(" []\n").split(), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "example_1A_point_field.npz"),
                        x=np.asarray(calc.x, dtype=np.float32), Q=np.asarray(calc.Q, dtype=np.float32),
                        point=np.zeros(3, dtype=np.float32), field=np.asarray(pf, dtype=np.float32),
                        shipped_point_field_dat=shipped)
    meta["1A"] = {"n_charges": int(len(calc.Q)), "field": [float(v) for v in pf]}

    # ---------------- example 2A: volume on the default 11^3 box grid --------------------------
    opts = json.load(open(os.path.join(ex, "2A_3D-field", "options", "options.json")))
    opts["inputpath"] = os.path.dirname(pdb)
    opts["outputpath"] = "/tmp/_golden_out_2A"
    calc = calculator(opts, path_to_pdb=pdb)
    field_box, mesh_shape = calc.compute_box()
    x2a = np.asarray(calc.x, dtype=np.float32)
    q2a = np.asarray(calc.Q, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "example_2A_volume.npz"),
                        x=x2a, Q=q2a, mesh=np.asarray(calc.mesh), field_box=field_box,
                        mesh_shape=np.array(mesh_shape))
    meta["2A"] = {"n_charges": int(len(q2a)), "mesh_shape": [int(v) for v in mesh_shape],
                  "dtype": str(field_box.dtype)}

    # ESP on the same frame / grid (volume_ESP, compute_box_ESP -> float16 (N,4))
    opts_esp = dict(opts)
    opts_esp["CPET_method"] = "volume_ESP"
    calc = calculator(opts_esp, path_to_pdb=pdb)
    esp_box, _ = calc.compute_box_ESP()
    # also the float32 values before the reference's float16 cast, straight from calc_esp_base
    pts = np.asarray(calc.mesh).reshape(-1, 3)
    esp32 = np.array([Math.calc_esp_base(p, calc.x, calc.Q)[0] for p in pts], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "example_2A_volume_esp.npz"),
                        mesh=np.asarray(calc.mesh), esp_box=esp_box, esp_f32=esp32)

    # ---------------- example 3A: topo, seeded n_iter, 6^3 = 216 lines -------------------------
    opts = json.load(open(os.path.join(ex, "3A_field-topology", "options", "options.json")))
    opts["inputpath"] = os.path.dirname(pdb)
    opts["outputpath"] = "/tmp/_golden_out_3A"
    opts["n_samples"] = 200            # uniform initializer cubes this to 6^3 = 216
    opts["max_streamline_init"] = "fixed_rand"
    opts["concur_slip"] = 4
    calc = calculator(opts, path_to_pdb=pdb)
    hist = calc.compute_topo_complete_c_shared()
    np.savez_compressed(os.path.join(OUT, "example_3A_topo.npz"),
                        seeds=np.asarray(calc.random_start_points, dtype=np.float32),
                        n_iter=np.asarray(calc.random_max_samples, dtype=np.int64),
                        dimensions=np.asarray(calc.dimensions, dtype=np.float32),
                        step_size=np.float32(calc.step_size), max_steps=np.int64(calc.max_steps),
                        hist=np.asarray(hist, dtype=np.float32))
    meta["3A"] = {"n_lines": int(len(hist)), "max_steps": int(calc.max_steps)}
    # 2A and 3A share the same frame and box transform -> same x, Q (checked, stored once in 2A)
    assert np.array_equal(np.asarray(calc.x, dtype=np.float32), x2a)
    assert np.array_equal(np.asarray(calc.Q, dtype=np.float32).ravel(), q2a.ravel())

    # seeds / mesh generator known-answers (initialize_box_points_uniform, UC:171-248)
    dims = np.array([0.5, 0.75, 1.0])
    c0 = np.zeros(3); ax = np.array([1.0, 0, 0]); ay = np.array([0, 1.0, 0])
    mesh_inc, _ = UC.initialize_box_points_uniform(c0, ax, ay, [4, 6, 8], dims, inclusive=True)
    seeds_u, nmax_u, _ = UC.initialize_box_points_uniform(c0, ax, ay, [5, 5, 5], dims, max_steps=27,
                                                          ret_rand_max=True, inclusive=False, seed=42)
    np.savez_compressed(os.path.join(OUT, "seeds_mesh.npz"), dims=dims, mesh_inclusive=mesh_inc,
                        seeds_uniform=seeds_u, n_iter_seed42_max27=nmax_u)

    # ---------------- synthetic, seeded: straight through Math_ops -----------------------------
    rng = np.random.default_rng(7)
    M = 2000
    xs = rng.uniform(-25, 25, size=(M, 3)).astype(np.float32)
    keep = ~np.all(np.abs(xs) < 1.5, axis=1)
    xs = np.ascontiguousarray(xs[keep])
    qs = rng.uniform(-0.8, 0.8, size=len(xs)).astype(np.float32)
    qs -= qs.mean(dtype=np.float64).astype(np.float32)
    pts = rng.uniform(-1.0, 1.0, size=(64, 3)).astype(np.float32)
    looped = Math.compute_looped_field(pts, xs, qs)
    f_alt = np.array([Math.calc_field(p, xs, qs) for p in pts], dtype=np.float32)
    f_base = np.array([Math.calc_field_base(p, xs, qs) for p in pts], dtype=np.float32)
    esp = np.array([Math.calc_esp_base(p, xs, qs)[0] for p in pts], dtype=np.float32)
    # softening known-answer: a grid point that coincides with a charge
    pts_soft = np.vstack([xs[:3], xs[:3] + np.float32(1e-4)]).astype(np.float32)
    looped_soft = Math.compute_looped_field(pts_soft, xs, qs)
    lines = {}
    dims32 = np.array([1.0, 1.0, 1.0], dtype=np.float32)
    seeds = rng.uniform(-0.95, 0.95, size=(48, 3)).astype(np.float32)
    for h in (0.1, 0.01):
        max_steps = round(2 * np.linalg.norm(dims32) / h)
        n_it = np.random.RandomState(42).randint(1, max_steps, len(seeds))
        res = np.array([Math.thread_operation(s, int(n), xs, qs, h, dims32) for s, n in zip(seeds, n_it)],
                       dtype=np.float32)
        lines[f"lines_h{h}"] = res
        lines[f"n_iter_h{h}"] = n_it.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "synthetic_math_ops.npz"),
                        x=xs, Q=qs, points=pts, looped_field=looped, calc_field=f_alt,
                        calc_field_base=f_base, esp=esp, points_soft=pts_soft,
                        looped_field_soft=looped_soft, seeds=seeds, dimensions=dims32, **lines)
    meta["synthetic"] = {"n_charges": int(len(qs))}

    json.dump(meta, open(os.path.join(OUT, "golden_meta.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()


def make_dropin_goldens():
    """Outputs of the UNMODIFIED `CPET(options).run()` on one shipped PDB, for the drop-in test
    (tests/test_gpu_dropin.py): `volume` -> *_efield.dat, `topo` (seeded) -> *.top."""
    import gzip
    import shutil
    import tempfile

    os.environ["CPET_BANNER"] = "0"
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from CPET.source.CPET import CPET

    ex = os.path.join(REF, "examples")
    pdb = os.path.join(ex, "1A_point-field", "pdb", "1_alcdehydro_run1.pdb")
    with open(pdb, "rb") as fi, gzip.open(os.path.join(OUT, "1_alcdehydro_run1.pdb.gz"), "wb", compresslevel=9) as fo:
        shutil.copyfileobj(fi, fo)
    work = tempfile.mkdtemp()
    os.makedirs(os.path.join(work, "pdb"))
    shutil.copy(pdb, os.path.join(work, "pdb"))
    for name, method, extra in [("2A_3D-field", "volume", {}),
                                ("3A_field-topology", "topo", {"n_samples": 200, "max_streamline_init": "fixed_rand",
                                                               "concur_slip": 4})]:
        opts = json.load(open(os.path.join(ex, name, "options", "options.json")))
        opts.update(extra)
        opts["inputpath"] = os.path.join(work, "pdb")
        opts["outputpath"] = os.path.join(work, "out_" + method)
        json.dump({k: v for k, v in opts.items() if k not in ("inputpath", "outputpath")},
                  open(os.path.join(OUT, f"dropin_options_{method}.json"), "w"), indent=1)
        CPET(opts).run()
        produced = sorted(os.listdir(opts["outputpath"]))
        assert len(produced) == 1, produced
        shutil.copy(os.path.join(opts["outputpath"], produced[0]), os.path.join(OUT, "dropin_" + produced[0]))
        print("drop-in golden:", produced[0])


if __name__ == "__main__" and "--dropin" in sys.argv:
    make_dropin_goldens()


def make_hist_goldens():
    """Histogram plan, normalised histograms and chi^2 matrix straight from the UNMODIFIED reference:
    CPET.utils.calculator.make_histograms (UC:596-718), construct_distance_matrix (UC:1003-1015),
    distance_numpy (UC:975-978) and the fixed-range get_hist_grid of
    CPET/source/scripts/residue_breakdown_analysis.py:28-37, run on (a) the two .top files the
    reference ships under examples/3A_field-topology/outdir (stale as tracer outputs, perfectly
    valid as histogram inputs) and (b) a seeded, ragged 3-file synthetic set.  The inputs are stored
    next to the outputs (float32: the reference writes .top from float32 rows, checked here), because
    the GPU box has no reference tree; tests rebuild the text files with np.savetxt('%.18e') as
    CPET.py:123 does."""
    import importlib.util
    import tempfile

    os.environ["CPET_BANNER"] = "0"
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import CPET.utils.calculator as UC

    spec = importlib.util.spec_from_file_location(
        "residue_breakdown_analysis", os.path.join(REF, "CPET", "source", "scripts", "residue_breakdown_analysis.py"))
    rba = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(rba)
        get_hist_grid = rba.get_hist_grid
    except Exception as e:      # its module-scope imports pull optional packages; the function is 8 lines of NumPy
        print("residue_breakdown_analysis not importable here (%s); get_hist_grid golden skipped" % e)
        get_hist_grid = None

    out = {}
    work = tempfile.mkdtemp()

    def run_set(tag, arrays):
        files = []
        for i, a in enumerate(arrays):
            f = os.path.join(work, f"{tag}_{i}.top")
            np.savetxt(f, a)                       # CPET.py:123 (default fmt '%.18e')
            files.append(f)
        H = UC.make_histograms(files)
        D = UC.construct_distance_matrix(H)
        d01 = UC.distance_numpy(H[0], H[1])
        for i, a in enumerate(arrays):
            out[f"{tag}_top{i}"] = a
        out[f"{tag}_hist"] = H
        out[f"{tag}_dist"] = D
        out[f"{tag}_d01"] = np.float64(d01)
        return H

    # (a) shipped example outputs as inputs
    ex = os.path.join(REF, "examples", "3A_field-topology", "outdir")
    shipped = []
    for name in ("1_alcdehydro_run1.top", "2_alcdehydro_run1.top"):
        a64 = np.loadtxt(os.path.join(ex, name))
        a32 = a64.astype(np.float32)
        assert np.array_equal(a32.astype(np.float64), a64), "shipped .top is not float32-exact"
        shipped.append(a32)
    H = run_set("shipped", shipped)
    # the reference's own file -> histogram path on the ORIGINAL shipped text (not our re-written copy)
    H_direct = UC.make_histograms([os.path.join(ex, "1_alcdehydro_run1.top"), os.path.join(ex, "2_alcdehydro_run1.top")])
    assert np.array_equal(H, H_direct), "re-written .top files do not reproduce the shipped ones"

    # (b) seeded synthetic, ragged lengths (the reference then bins with the mean length, UC:650-659)
    rng = np.random.default_rng(11)
    synth = []
    for n, (mu_d, mu_c) in zip((4096, 4096, 3000), ((0.45, 0.6), (0.5, 0.8), (0.4, 0.7))):
        d = np.abs(rng.normal(mu_d, 0.2, n)).astype(np.float32)
        c = rng.gamma(2.0, mu_c / 2.0, n).astype(np.float32)
        synth.append(np.stack([d, c], axis=1))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        run_set("synth", synth)
    # equal-length synthetic set of three (no warning path)
    run_set("synth_eq", [s[:3000] for s in synth])

    if get_hist_grid is not None:
        out["grid_fixed"] = get_hist_grid(shipped[0].astype(np.float64), [0.0, 1.5], 40, [0.0, 3.0], 60)
        out["grid_fixed_args"] = np.array([0.0, 1.5, 40, 0.0, 3.0, 60])

    np.savez_compressed(os.path.join(OUT, "histograms_reference.npz"), **out)
    print("histogram goldens:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__" and "--hist" in sys.argv:
    make_hist_goldens()
