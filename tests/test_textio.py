"""CPU: the text writers produce the bytes np.savetxt / the reference's save_numpy_as_dat produce."""
import time

import numpy as np
import pytest

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


def _reference_parse(path):
    """The reference's own loop (CPET/utils/calculator.py:626-633), restated; a blank line, on which
    the reference raises IndexError, is skipped here as the reader skips it."""
    d, c = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            line = line.strip().split()
            if not line:
                continue
            d.append(float(line[0]))
            c.append(float(line[1]))
    return np.column_stack([d, c]) if d else np.zeros((0, 2))


def test_top_reader_matches_float_parsing(tmp_path):
    rng = np.random.default_rng(3)
    hist = np.column_stack([rng.gamma(2.0, 0.3, 150_000), rng.gamma(1.5, 0.4, 150_000)]).astype(np.float32)
    hist[5] = [0.0, -0.0]
    hist[6] = [np.inf, -np.inf]
    hist[7] = [np.nan, 1e-38]
    hist[8] = [3.4e38, 1.4e-45]
    p = tmp_path / "a.top"
    np.savetxt(p, hist)                                   # several MB: more than one reader span
    got = pio.read_topology(str(p))
    want = _reference_parse(p)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)) or \
        np.array_equal(got[~np.isnan(want)].view(np.uint64), want[~np.isnan(want)].view(np.uint64))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got.astype(np.float32)[~np.isnan(hist)], hist[~np.isnan(hist)])   # text round trip


def test_top_reader_edge_cases(tmp_path):
    p = tmp_path / "b.top"
    # comments, blank line, CRLF, tabs, '+' sign, third column ignored, doubles that need correct
    # rounding, overflow / underflow, no trailing newline
    p.write_bytes(b"# header\n#another\n\n1.0 2.0\r\n\t+3.5e-1\t  -4.25E+2  99\n"
                  b"0.1 0.30000000000000004\n1e400 -1e400\n1e-400 4.9406564584124654e-324\n"
                  b"2.2250738585072011e-308 1.7976931348623157e308\n7 8")
    got = pio.read_rows(str(p), 2)
    want = _reference_parse(p)
    assert got.shape == (7, 2)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert pio.read_rows(str(p), 1).shape == (7, 1)
    # empty and comment-only files
    e = tmp_path / "e.top"
    e.write_bytes(b"")
    assert pio.read_topology(str(e)).shape == (0, 2)
    e.write_bytes(b"# nothing\n")
    assert pio.read_topology(str(e)).shape == (0, 2)


def test_top_reader_errors(tmp_path):
    import pytest

    from pycpet_b200 import CpetError

    p = tmp_path / "bad.top"
    p.write_bytes(b"1.0 2.0\n3.0\n")
    with pytest.raises(CpetError, match="data line 2"):
        pio.read_topology(str(p))
    p.write_bytes(b"1.0 2.0\n3.0 abc\n")
    with pytest.raises(CpetError):
        pio.read_topology(str(p))
    with pytest.raises(CpetError, match="cannot open"):
        pio.read_topology(str(tmp_path / "missing.top"))


def test_reader_is_faster_than_the_python_loop(tmp_path):
    hist = np.random.default_rng(4).random((200_000, 2)).astype(np.float32)
    p = tmp_path / "c.top"
    pio.save_topology(str(p), hist)
    t0 = time.perf_counter(); want = _reference_parse(p); t_py = time.perf_counter() - t0
    t0 = time.perf_counter(); got = pio.read_topology(str(p)); t_c = time.perf_counter() - t0
    assert np.array_equal(got, want)
    assert t_c < t_py


def test_writer_formats_agree_with_savetxt(tmp_path):
    """'%.Ne' / '%.Nf' take the std::to_chars path, everything else snprintf: same bytes either way."""
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.normal(0, 1, 4000) * 10.0 ** rng.integers(-300, 300, 4000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 7.0, 5e-324, 1.7976931348623157e308, 0.0005, 0.0015,
                         -0.0004, 2.5, 3.5, 1e22, 123456789012345678.0]]).reshape(-1, 2)
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    for fmt in ("%.18e", "%.3f", "%.0f", "%.0e", "%.6e", "%.25e", "%g", "%12.5f", "%+.4e", "%.10g"):
        np.savetxt(a, v, fmt=fmt)
        pio.write_rows(str(b), v, fmt=fmt)
        assert a.read_bytes() == b.read_bytes(), fmt


def test_binary_side_channel_gives_the_same_values(tmp_path):
    import os

    hist = np.random.default_rng(6).gamma(2.0, 0.3, (5000, 2)).astype(np.float32)
    p = str(tmp_path / "f.top")
    pio.save_topology(p, hist, binary=True)
    assert sorted(os.listdir(tmp_path)) == ["f.top", "f.top.npy", "f.top.npy.fp"]
    assert not any(n.endswith("top") for n in ("f.top.npy", "f.top.npy.fp"))   # invisible to the dispatcher's resume rule
    via_npy = pio.read_topology(p)
    via_text = pio.read_topology(p, use_binary=False)
    assert via_npy.dtype == via_text.dtype == np.float64
    assert np.array_equal(via_npy.view(np.uint64), via_text.view(np.uint64))
    # a text file replaced later wins over the stale side channel even when its mtime says otherwise
    # (cp -p, rsync -t, a restored backup): the fingerprint of the text no longer matches
    np.savetxt(p, hist[:10])
    os.utime(p, (1, 1))
    assert pio.read_topology(p).shape == (10, 2)
    # same size, different content
    pio.save_topology(p, hist, binary=True)
    other = hist.copy(); other[::7, 0] += np.float32(0.5)
    pio.write_rows(p, other, fmt="%.18e")
    assert os.path.getsize(p) == len(hist) * 50
    assert np.array_equal(pio.read_topology(p), other.astype(np.float64))
    # a side channel whose text file is gone is not used silently
    pio.save_topology(p, hist, binary=True)
    os.remove(p)
    with pytest.raises(Exception):
        pio.read_topology(p)
    assert pio.read_topology(p, use_binary="force").shape == (5000, 2)
