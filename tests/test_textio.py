"""CPU: the text writers produce the bytes np.savetxt / the reference's save_numpy_as_dat produce."""
import time

import numpy as np

from pycpet_b200 import io as pio


def test_top_file_bytes(tmp_path):
    rng = np.random.default_rng(0)
    hist = np.column_stack([rng.gamma(2.0, 0.3, 20_000), rng.gamma(1.5, 0.4, 20_000)]).astype(np.float32)
    hist[5] = [0.0, -0.0]
    hist[6] = [np.inf, -np.inf]
    hist[7] = [np.nan, 1e-38]
    hist[8] = [3.4e38, 1.4e-45]
    a, b = tmp_path / "a.top", tmp_path / "b.top"
    np.savetxt(a, hist)                       # what CPET.run_topo does (CPET/source/CPET.py:123)
    pio.save_topology(str(b), hist)
    assert a.read_bytes() == b.read_bytes()
    # float64 input too
    np.savetxt(a, hist.astype(np.float64) / 3.0)
    pio.write_rows(str(b), hist.astype(np.float64) / 3.0)
    assert a.read_bytes() == b.read_bytes()


def test_dat_file_bytes(tmp_path):
    rng = np.random.default_rng(1)
    vol = np.column_stack([rng.uniform(-0.5, 0.5, (1331, 3)), rng.normal(0, 3, (1331, 3))]).astype(np.float32)
    vol[3, 3] = -0.0004          # rounds to -0.000
    vol[4, 4] = 0.0005
    meta = {"dimensions": np.array([0.5, 0.5, 0.5]), "num_steps": [11, 11, 11],
            "transformation_matrix": np.array([[0.6, 0.8, 0.0], [-0.8, 0.6, 0.0], [0.0, 0.0, 1.0]]),
            "center": np.array([104.785, 113.388, 117.966])}
    a, b = tmp_path / "a.dat", tmp_path / "b.dat"
    # the reference's writer, restated with its own np.savetxt call (CPET/utils/io.py:98-109)
    np.savetxt(a, vol, fmt="%.3f")
    body = a.read_text()
    a.write_text(pio.dat_header(meta) + body)
    pio.save_numpy_as_dat(meta, vol, str(b))
    assert a.read_bytes() == b.read_bytes()
    assert b.read_text().splitlines()[0] == "#Sample Density: 11 11 11; Volume: Box: 0.5 0.5 0.5"
    # float16 ESP rows (compute_box_ESP returns float16, CPET/utils/calculator.py:473-475)
    esp = np.column_stack([vol[:, :3], rng.normal(0, 1, 1331)]).astype(np.float16)
    np.savetxt(a, esp, fmt="%.3f")
    pio.write_rows(str(b), esp, fmt="%.3f")
    assert a.read_bytes() == b.read_bytes()


def test_writer_is_faster_than_savetxt(tmp_path):
    hist = np.random.default_rng(2).random((200_000, 2)).astype(np.float32)
    t0 = time.perf_counter(); np.savetxt(tmp_path / "a.top", hist); t_np = time.perf_counter() - t0
    t0 = time.perf_counter(); pio.save_topology(str(tmp_path / "b.top"), hist); t_c = time.perf_counter() - t0
    assert (tmp_path / "a.top").read_bytes() == (tmp_path / "b.top").read_bytes()
    assert t_c < t_np
