"""CPU: the C-ABI library loads and exports every symbol include/cpet_b200.h declares; the
reference-shaped binding class constructs on it; compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "cpet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|char)\s*\*?\s*(\w+)\s*\(", txt, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from pycpet_b200 import _lib

    L = _lib.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cpet_b200.h but not exported"
    # every symbol bound by the Python layer is declared in the header
    for n in list(_lib.SIGNATURES) + _lib.LEGACY_SYMBOLS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"
    assert L.cpet_abi_version() == 1


def test_reference_symbols_present():
    """Math_ops.__init__ of the reference sets argtypes on these names (c_ops.py:26-159)."""
    from pycpet_b200 import _lib

    L = _lib.load()
    for n in ["sparse_dot", "vecaddn", "dot", "einsum_ij_i", "einsum_ij_i_batch",
              "einsum_operation_batch", "einsum_operation", "thread_operation",
              "thread_operation_dipole", "calc_field", "calc_field_base", "calc_esp_base",
              "compute_batched_field", "compute_looped_field"]:
        assert hasattr(L, n)


def test_math_ops_constructs_and_rejects_bad_arrays():
    from pycpet_b200 import Math_ops

    m = Math_ops()
    x = np.zeros((4, 3), dtype=np.float64)      # wrong dtype: ndpointer must reject it
    q = np.zeros(4, dtype=np.float32)
    with pytest.raises(ctypes.ArgumentError):
        m.math.calc_field(np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32), 4, x, q)
    with pytest.raises(ctypes.ArgumentError):   # wrong ndim
        m.math.compute_looped_field(1, 4, np.zeros(3, dtype=np.float32), x.astype(np.float32), q,
                                    np.zeros((1, 3), dtype=np.float32))


def test_no_silent_fallback_without_gpu():
    from pycpet_b200 import CpetError, Math_ops, _lib

    if _lib.load().cpet_device_count() > 0:
        pytest.skip("a GPU is present")
    m = Math_ops()
    with pytest.raises(CpetError):
        m.field_grid(np.zeros((2, 3), np.float32), np.ones((4, 3), np.float32), np.ones(4, np.float32))
    with pytest.raises(CpetError):
        m.compute_looped_field(np.zeros((2, 3), np.float32), np.ones((4, 3), np.float32),
                               np.ones(4, np.float32))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pycpet_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pycpet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libcpet_oracle" not in src and "oracle/_ref" not in src, f


C_CLIENT = r"""
/* a plain C99 consumer of include/cpet_b200.h: the boundary holds no C++ or torch types */
#include <stdio.h>
#include <string.h>
#include "cpet_b200.h"

int main(int argc, char **argv) {
    if (argc < 2 || cpet_abi_version() != CPET_ABI_VERSION) return 10;
    /* no device on this box: creation must fail with a status and a message, not crash */
    cpet_ctx *ctx = NULL;
    int rc = cpet_create(0, &ctx);
    if (cpet_device_count() == 0) {
        if (rc != CPET_ERR_NO_DEVICE || ctx != NULL) return 11;
        if (strstr(cpet_last_error(), "no CPU fallback") == NULL) return 12;
        cpet_clear_error();
        if (cpet_last_status() != CPET_OK) return 13;
    } else if (rc == CPET_OK) {
        cpet_destroy(ctx);
    }
    /* host-side text I/O through the C ABI: write two float32 rows, read them back as float64 */
    const float rows[4] = {1.5f, -2.25f, 3.0f, 0.1f};
    if (cpet_write_rows(argv[1], "# header\n", rows, 0, 2, 2, "%.18e", 1) != CPET_OK) return 14;
    int64_t n = -1;
    if (cpet_count_rows(argv[1], &n, 1) != CPET_OK || n != 2) return 15;
    double back[4];
    if (cpet_read_rows(argv[1], 2, n, back, 1) != CPET_OK) return 16;
    for (int i = 0; i < 4; ++i)
        if (back[i] != (double)rows[i]) return 17;
    if (cpet_read_rows(argv[1], 3, n, back, 1) != CPET_ERR_INVALID) return 18;   /* only two columns */
    printf("ok\n");
    return 0;
}
"""


def test_header_is_plain_c_and_links(tmp_path):
    """gcc -std=c99 -pedantic compiles a client against the header and links it to libcpetb200.so."""
    import shutil
    import subprocess

    from pycpet_b200 import lib_path

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.dirname(lib_path())
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L", libdir, "-l:libcpetb200.so", f"-Wl,-rpath,{libdir}"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([str(exe), str(tmp_path / "rows.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)
