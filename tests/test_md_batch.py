"""CPU: the MD-batch driver's host logic (chunking, grouping by geometry, resume rule, ordered
results from worker processes, byte-identical `.top` text) with a stand-in for the GPU call.
The GPU version with the unmodified PyCPET constructor is tests/test_gpu_dropin.py."""
import os

import numpy as np

from pycpet_b200 import md_batch


def fake_prepare(options, path):
    """Stands in for the reference's calculator constructor: deterministic arrays from the name."""
    k = int(os.path.basename(path).split("_")[0])
    rng = np.random.default_rng(k)
    n_seeds = 5
    seeds = np.arange(3 * n_seeds, dtype=np.float32).reshape(n_seeds, 3) * np.float32(0.01)
    if options.get("odd_geometry") == k:
        seeds = seeds + np.float32(1.0)
    return {"path": path, "x": rng.normal(size=(10 + k, 3)).astype(np.float32),
            "Q": rng.normal(size=10 + k).astype(np.float32), "seeds": seeds,
            "n_iter": rng.integers(1, 9, n_seeds).astype(np.int32), "step_size": 0.1,
            "dimensions": np.array([0.5, 0.5, 0.5], np.float32)}


class FakeMath:
    """topo_hist_frames stand-in: rows depend on the frame's charges and n_iter only."""

    def __init__(self):
        self.calls = []

    def topo_hist_frames(self, frames, seeds, n_iter, d_edges, c_edges, step_size=0.1, dimensions=(1, 1, 1),
                         second_diff=False, want_rows=False, **kw):
        self.calls.append(len(frames))
        n_iter = np.asarray(n_iter)
        assert n_iter.shape == (len(frames), len(seeds))
        rows = np.stack([np.stack([np.full(len(seeds), np.float32(x.sum())) + seeds[:, 0],
                                   n_iter[f].astype(np.float32) * np.float32(q.sum())], axis=1)
                         for f, (x, q) in enumerate(frames)]).astype(np.float32)
        counts = np.zeros((len(frames), len(d_edges) - 1, len(c_edges) - 1), np.int64)
        counts[:, 0, 0] = [len(x) for x, _ in frames]
        return rows, counts


def expected_rows(path, options=None):
    f = fake_prepare(options or {}, path)
    m = FakeMath()
    return m.topo_hist_frames([(f["x"], f["Q"])], f["seeds"], f["n_iter"][None], [0, 1], [0, 1])[0][0]


def test_chunks_order_files_and_resume(tmp_path):
    files = [str(tmp_path / "in" / f"{k}_frame.run1.pdb") for k in range(5)]
    out = tmp_path / "out"
    out.mkdir()
    (out / "1_frame.top").write_text("done earlier\n")
    m = FakeMath()
    res = md_batch.run_topo_frames({}, files, outputpath=str(out), workers=0, chunk=2, math=m,
                                   prepare=fake_prepare, d_edges=np.linspace(0, 1, 4), c_edges=np.linspace(0, 1, 3),
                                   keep_rows=True)
    assert res["skipped"] == [files[1]] and res["files"] == [files[0], files[2], files[3], files[4]]
    assert m.calls == [2, 2]
    assert res["counts"].shape == (4, 3, 2) and list(res["counts"][:, 0, 0]) == [10, 12, 13, 14]
    assert (out / "1_frame.top").read_text() == "done earlier\n"
    for f, rows in zip(res["files"], res["rows"]):
        np.testing.assert_array_equal(rows, expected_rows(f))
        ref = tmp_path / "ref.top"
        np.savetxt(ref, rows)                                       # what CPET.py:123 writes
        assert (out / (md_batch.protein_name(f) + ".top")).read_bytes() == ref.read_bytes()
    # second run: everything is skipped, no GPU call
    m2 = FakeMath()
    res2 = md_batch.run_topo_frames({}, files, outputpath=str(out), workers=0, chunk=2, math=m2, prepare=fake_prepare)
    assert res2["files"] == [] and len(res2["skipped"]) == 5 and m2.calls == [] and res2["counts"] is None


def test_frames_with_another_geometry_get_their_own_call():
    files = [f"/x/{k}_f.pdb" for k in range(6)]
    m = FakeMath()
    res = md_batch.run_topo_frames({"odd_geometry": 2}, files, workers=0, chunk=6, math=m, prepare=fake_prepare,
                                   keep_rows=True)
    assert m.calls == [2, 1, 3] and res["files"] == files
    np.testing.assert_array_equal(res["rows"][2], expected_rows(files[2], {"odd_geometry": 2}))


def test_worker_processes_keep_the_file_order():
    files = [f"/x/{k}_f.pdb" for k in range(7)]
    a = md_batch.run_topo_frames({}, files, workers=0, chunk=3, math=FakeMath(), prepare=fake_prepare, keep_rows=True)
    m = FakeMath()
    b = md_batch.run_topo_frames({}, files, workers=2, chunk=3, math=m, prepare=fake_prepare, keep_rows=True)
    assert b["files"] == files and m.calls == [3, 3, 1]
    for ra, rb in zip(a["rows"], b["rows"]):
        np.testing.assert_array_equal(ra, rb)


def failing_prepare(options, path):
    if "3_" in path:
        raise ValueError(f"cannot parse {path}")
    return fake_prepare(options, path)


def test_a_failing_constructor_surfaces_and_stops_the_batch(tmp_path):
    import pytest

    files = [f"/x/{k}_f.pdb" for k in range(6)]
    for workers in (0, 2):
        m = FakeMath()
        with pytest.raises(ValueError, match="cannot parse"):
            md_batch.run_topo_frames({}, files, outputpath=str(tmp_path / f"o{workers}"), workers=workers, chunk=2,
                                     math=m, prepare=failing_prepare)
        assert m.calls == [2]                       # frames 0,1 were done and written before the failure
        assert sorted(os.listdir(tmp_path / f"o{workers}")) == ["0_f.top", "1_f.top"]
