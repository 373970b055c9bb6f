"""CPU: the C-ABI library loads, exports every symbol include/dip_b200.h declares with the declared
arity, validates arguments before touching the GPU, and its host-side CSR builder matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import small_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(dip_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_every_declared_symbol_is_exported_with_matching_arity(lib):
    from dip_b200 import _lib
    decl = header_functions()
    assert len(decl) >= 35
    for name, nargs in decl.items():
        assert hasattr(lib, name), name
        assert name in _lib.PROTOTYPES, name
        assert len(_lib.PROTOTYPES[name][1]) == nargs, (name, nargs, len(_lib.PROTOTYPES[name][1]))
    assert set(_lib.PROTOTYPES) == set(decl)
    assert lib.dip_abi_version() == 2


def test_library_holds_sm100a_code_only(lib):
    from dip_b200 import _lib
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_gpu(lib):
    from dip_b200 import _lib
    g = _lib.GemmArgs()
    g.A, g.W, g.C = 256, 256, 256
    g.M, g.N, g.K, g.lda, g.ldc = 4, 100, 128, 128, 100
    assert lib.dip_gemm_nt(C.byref(g), None) != 0
    assert b"multiple of 128" in lib.dip_last_error()
    assert lib.dip_ddim_step(None, None, None, None, 0, None, 0, None, 1, 1, 1, 1, 1, 1, 0, 0, 0, None) != 0
    assert lib.dip_guidance(*([None] * 15), 0.5, 1, 1, 1, 1, 0, None) != 0
    assert lib.dip_engine_workspace_bytes(None, 4) == 0


def test_host_csr_builder_matches_oracle():
    from dip_b200.engine import coo_to_csr_csc
    from oracle import restate
    cs = small_case()
    csr = coo_to_csr_csc(cs.rows, cs.cols, cs.vals, cs.M, cs.n)
    r, c, v = restate.coalesce_coo(cs.rows, cs.cols, cs.vals, cs.M, cs.n)
    assert csr["nnz"] == len(r) < len(cs.rows)                 # the fixture holds duplicates
    assert np.array_equal(csr["csr_col"], c) and np.array_equal(csr["csr_val"], v)
    assert np.array_equal(np.repeat(np.arange(cs.M), np.diff(csr["csr_rowptr"])), r)
    # CSC is the same matrix, rows ascending within each column
    dense = np.zeros((cs.M, cs.n), np.float32)
    dense[r, c] = v
    for j in range(cs.n):
        s, e = csr["csc_colptr"][j], csr["csc_colptr"][j + 1]
        rows_j = csr["csc_row"][s:e]
        assert np.all(np.diff(rows_j) > 0)
        col = np.zeros(cs.M, np.float32); col[rows_j] = csr["csc_val"][s:e]
        assert np.array_equal(col, dense[:, j])
    # empty matrix and out-of-range entries
    z = np.zeros(0, np.int64)
    e = coo_to_csr_csc(z, z, np.zeros(0, np.float32), 3, 4)
    assert e["nnz"] == 0 and not e["csr_rowptr"].any()
    from dip_b200 import _lib
    with pytest.raises(_lib.DipError, match="out of range"):
        coo_to_csr_csc(np.array([5]), np.array([0]), np.array([1.0], np.float32), 3, 4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dip_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DipError, match="no CPU or eager fallback"):
        _lib.load()
