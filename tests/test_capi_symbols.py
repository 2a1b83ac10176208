"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/amrvw_b200.h declares; the host mirror fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "amrvw_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amrvw_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_path():
    syms = header_symbols()
    for need in ("amrvw_create", "amrvw_destroy", "amrvw_roots_rds_f64", "amrvw_roots_css_c64", "amrvw_roots_pencil_f64",
                 "amrvw_roots_rds_f64_dev", "amrvw_last_error"):
        assert need in syms


@pytest.mark.parametrize("strict", [False, True])
def test_library_exports_every_declared_symbol(A, strict):
    lib = ctypes.CDLL(A.lib_path(strict=strict))
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/amrvw_b200.h but not exported"
    assert set(header_symbols()) == set(A.EXPORTS)
    assert lib.amrvw_version() >= 100


def test_library_is_sm100a_and_in_tree(A):
    path = A.lib_path()
    assert os.path.dirname(path) == os.path.join(ROOT, "amrvw.jl_b200")
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_default_options_and_argument_errors_need_no_gpu(A):
    lib = A.load()
    o = A.Options(9, 9, 9, 9, 9)
    lib.amrvw_default_options(ctypes.byref(o))
    assert (o.schedule, o.lanes_per_poly, o.shift_reuse, o.small_window) == (0, 0, 0, 0)
    # the reference's constants (create-bulge.jl:9, AMRVW_algorithm.jl:25,184) are the defaults of the options
    assert (o.it_count, o.it_max_factor, o.stream_ctas_per_sm) == (15, 20, 0) and o.tol == np.finfo(np.float64).eps
    assert ctypes.sizeof(A.Options) == 48
    assert lib.amrvw_set_options(None, ctypes.byref(o)) == -1
    assert lib.amrvw_destroy(None) == 0
    assert lib.amrvw_last_error(None) == b"null context"


def test_no_cpu_fallback(A):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(A.AmrvwError):
        A.Context()
    with pytest.raises(A.AmrvwError):
        A.roots([1.0, 2.0, 3.0, 4.0, 5.0])


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (parity would be void): no import, include,
    dlopen or path reference to it anywhere under amrvw.jl_b200/ (comments may mention it)."""
    pk = os.path.join(ROOT, "amrvw.jl_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[\"<][^\">]*oracle)|(liboracle)|(oracle/_build)|(import_module\([\"']oracle)", re.M)
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), (dirpath, f)
    hdr = open(os.path.join(ROOT, "include", "amrvw_b200.h")).read()
    assert not bad.search(hdr)


def test_host_packing_logic(A):
    from importlib import import_module
    R = import_module("amrvw_jl_b200.roots")
    flat, off = R._pack([[1.0, 2.0], [3.0], [4.0, 5.0, 6.0]])
    assert off.tolist() == [0, 2, 3, 6] and flat.tolist() == [1, 2, 3, 4, 5, 6] and flat.dtype == np.float64
    flat, off = R._pack([[1.0, 2.0], [3.0 + 1j]])
    assert flat.dtype == np.complex128
    assert R._is_batch([[1.0, 2.0], [3.0]]) and not R._is_batch([1.0, 2.0]) and R._is_batch(np.zeros((3, 4)))
    vs, ws = A.basic_pencil([1.0, 2.0, 3.0])
    assert vs.tolist() == [1.0, 2.0] and ws.tolist() == [0.0, 3.0]


def test_partition_batch_is_contiguous_and_balanced_by_squared_length(A):
    """Host logic of the multi-device context (amrvw_create_multi): contiguous cuts, every polynomial on
    exactly one device, work (sum of len^2) balanced; pure host code, runs without a GPU."""
    lib = A.load()
    rng = np.random.default_rng(5)

    def cuts_of(lens, ndev):
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        cuts = np.zeros(ndev + 1, dtype=np.int64)
        assert lib.amrvw_partition_batch(off.ctypes.data, len(lens), ndev, cuts.ctypes.data) == 0
        return cuts

    for B, ndev in ((3000, 8), (512, 8), (7, 8), (1, 2), (0, 4), (11, 3)):
        lens = np.full(B, 3001, dtype=np.int64)
        cuts = cuts_of(lens, ndev)
        assert cuts[0] == 0 and cuts[-1] == B and np.all(np.diff(cuts) >= 0)
        if B >= ndev:
            sizes = np.diff(cuts)
            assert sizes.max() - sizes.min() <= 1, sizes          # equal degrees: an even split (3000 over 8 = 375 each)
    lens = rng.integers(0, 4000, size=2000).astype(np.int64)     # ragged, with empty vectors
    for ndev in (2, 4, 8):
        cuts = cuts_of(lens, ndev)
        assert cuts[0] == 0 and cuts[-1] == len(lens) and np.all(np.diff(cuts) >= 0)
        w = (lens.astype(float) ** 2 + 64.0)
        per = np.array([w[cuts[d]:cuts[d + 1]].sum() for d in range(ndev)])
        assert per.max() <= w.sum() / ndev + w.max() + 1.0, per   # no device exceeds its share by more than one polynomial
    off = np.array([0, 5, 3], dtype=np.int64)
    cuts = np.zeros(3, dtype=np.int64)
    assert lib.amrvw_partition_batch(off.ctypes.data, 2, 2, cuts.ctypes.data) == -1     # decreasing offsets


def test_hessenberg_host_packing_and_argument_errors(A):
    """Host side of the Hessenberg entries (no GPU): column-major packing as Julia stores matrices, dims, and the argument
    checks that need no device."""
    from importlib import import_module
    roots_mod = import_module("amrvw_jl_b200.roots")
    M1 = np.arange(9.0).reshape(3, 3)
    M2 = np.array([[1.0, 2.0], [3.0, 4.0]])
    flat, dims = roots_mod.pack_matrices([M1, np.zeros((0, 0)), M2])
    assert dims.tolist() == [3, 0, 2] and flat.dtype == np.float64
    assert np.array_equal(flat[:9], M1.reshape(-1, order="F")) and np.array_equal(flat[9:], [1.0, 3.0, 2.0, 4.0])
    cflat, cdims = roots_mod.pack_matrices([M2, 1j * M2])
    assert cflat.dtype == np.complex128 and cdims.tolist() == [2, 2]
    with pytest.raises(ValueError):
        roots_mod.pack_matrices([np.zeros((2, 3))])
    with pytest.raises(ValueError):
        A.qr_factorization(np.zeros(4))
    f = A.qr_factorization(M2, unitary=True)
    assert f.unitary and f.H.shape == (2, 2)
    with pytest.raises(TypeError):
        A.eigvals([M2])                 # eigvals takes factorizations, as in the reference


def test_partition_hessenberg_is_contiguous_and_balanced_by_cubed_dimension(A):
    """The split amrvw_eigvals_hessenberg_* apply on a multi-device context (host code): contiguous, every matrix once, shares of
    sum(n^3) within one matrix of equal."""
    lib = A.load()
    rng = np.random.default_rng(8)
    dims = rng.integers(0, 200, 500).astype(np.int32)
    w = dims.astype(np.float64) ** 3
    for ndev in (1, 2, 3, 8):
        cuts = np.zeros(ndev + 1, dtype=np.int64)
        assert lib.amrvw_partition_hessenberg(dims.ctypes.data, len(dims), ndev, cuts.ctypes.data) == 0
        assert cuts[0] == 0 and cuts[-1] == len(dims) and np.all(np.diff(cuts) >= 0)
        shares = np.array([w[cuts[d]:cuts[d + 1]].sum() for d in range(ndev)])
        assert np.all(np.abs(shares - w.sum() / ndev) <= w.max() + 1e-9)
    cuts = np.zeros(3, dtype=np.int64)
    bad = np.array([3, -1], dtype=np.int32)
    assert lib.amrvw_partition_hessenberg(bad.ctypes.data, 2, 2, cuts.ctypes.data) == -1
    assert lib.amrvw_partition_hessenberg(dims.ctypes.data, 0, 2, cuts.ctypes.data) == 0 and cuts.tolist() == [0, 0, 0]
