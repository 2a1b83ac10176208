"""Eigenvalues of upper Hessenberg matrices through qr_factorization (reference src/qrfactorization.jl:63-103,
src/rfactorization.jl:230-357, tests test/test-roots.jl:193-223).

CPU part: the oracle's dense-R and unitary-diagonal twins against dense LAPACK (numpy.linalg.eigvals), which is the
reference's own check (`norm(e1 - e2) <= sqrt(eps())` at n = 15).
GPU part (`-m gpu`): the product through the C ABI (amrvw_eigvals_hessenberg_f64/_c64) against the oracle -- bit for bit with
the library built without FMA contraction, within the reference's tolerance for the default build -- and against LAPACK."""
import numpy as np
import pytest

from conftest import match_roots

SQEPS = np.sqrt(np.finfo(float).eps)


def hess(rng, n, cx, kind="rand"):
    """triu(rand(T, n, n), -1) of test-roots.jl:197, or N(0,1) entries"""
    draw = rng.random if kind == "rand" else rng.standard_normal
    M = draw((n, n)) + (1j * draw((n, n)) if cx else 0.0)
    return np.triu(M, -1)


def rotm(n, i, c, s):
    M = np.eye(n, dtype=complex)
    M[i, i] = c; M[i, i + 1] = s; M[i + 1, i] = -s; M[i + 1, i + 1] = np.conj(c)
    return M


def unitary_hess(rng, n, cx):
    """Qs * D of test-roots.jl:209-221: a descending chain of random rotators times a unitary diagonal"""
    M = np.eye(n, dtype=complex)
    for i in range(n - 1):
        t = rng.uniform(0, 2 * np.pi)
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi)) if cx else 1.0
        M = M @ rotm(n, i, np.cos(t) * ph, np.sin(t))
    D = np.exp(1j * rng.random(n)) if cx else rng.choice([-1.0, 1.0], n)
    M = M @ np.diag(D)
    return M if cx else M.real.copy()


def dist(a, b):
    d, _ = match_roots(np.asarray(a), np.asarray(b))
    return float(np.linalg.norm(d))


# ------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("cx", [False, True])
def test_oracle_qr_factorization_reference_cases(oracle, cx):   # test-roots.jl:195-206
    rng = np.random.default_rng(15 + cx)
    for _ in range(5):
        M = hess(rng, 15, cx)
        e, st = oracle.eigvals_hessenberg(M)
        assert st["unconverged"] == 0
        assert dist(e, np.linalg.eigvals(M)) <= SQEPS


@pytest.mark.parametrize("cx", [False, True])
@pytest.mark.parametrize("unitary", [False, True])
def test_oracle_unitary_hessenberg(oracle, cx, unitary):        # test-roots.jl:208-221 (there with unitary=false only)
    rng = np.random.default_rng(31 + cx)
    for n in (2, 3, 15, 40):
        M = unitary_hess(rng, n, cx)
        e, st = oracle.eigvals_hessenberg(M, unitary=unitary)
        assert st["unconverged"] == 0
        assert dist(e, np.linalg.eigvals(M)) <= SQEPS
        assert np.allclose(np.abs(e), 1.0, atol=1e-12)


@pytest.mark.parametrize("cx", [False, True])
def test_oracle_dense_sizes_and_order(oracle, cx):
    rng = np.random.default_rng(7)
    for n in (1, 2, 3, 4, 7, 33, 64, 120):
        M = hess(rng, n, cx, "randn")
        e, st = oracle.eigvals_hessenberg(M)
        w = np.linalg.eigvals(M)
        assert st["unconverged"] == 0
        assert dist(e, w) <= 1e-11 * max(1.0, np.abs(w).max()) * n
        key = list(zip(e.real, e.imag))
        assert key == sorted(key)                               # sorteig! order (AMRVW_algorithm.jl:134)
        if not cx:
            assert np.sum(e.imag > 0) == np.sum(e.imag < 0)     # conjugate pairs from the 2x2 blocks


def test_oracle_companion_as_dense_matches_roots(oracle):
    """The companion matrix of p6 handed over as a plain Hessenberg matrix has the roots 1..6 (test-roots.jl:113-123 through f3)."""
    p = np.array([720.0, -1764.0, 1624.0, -735.0, 175.0, -21.0, 1.0])
    n = 6
    C = np.zeros((n, n)); C[1:, :-1] = np.eye(n - 1); C[:, -1] = -p[:-1] / p[-1]
    e, _ = oracle.eigvals_hessenberg(C)
    assert np.linalg.norm(np.sort(e.real) - np.arange(1, 7)) <= 1e-8 and np.all(e.imag == 0)


# ------------------------------------------------------------------ product (GPU, through the C ABI)
@pytest.fixture(scope="module")
def ctx(A):
    c = A.Context(seed=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_strict(A):
    c = A.Context(seed=0, strict=True)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cx", [False, True])
def test_gpu_qr_factorization_reference_cases(A, ctx, cx):      # test-roots.jl:195-221, same calls, same tolerance
    rng = np.random.default_rng(15 + cx)
    M = hess(rng, 15, cx)
    e2 = A.eigvals(A.qr_factorization(M, unitary=False), ctx=ctx)
    assert dist(e2, np.linalg.eigvals(M)) <= SQEPS
    U = unitary_hess(rng, 15, cx)
    for unitary in (False, True):
        e2 = A.eigvals(A.qr_factorization(U, unitary=unitary), ctx=ctx)
        assert dist(e2, np.linalg.eigvals(U)) <= SQEPS


@pytest.mark.gpu
@pytest.mark.parametrize("cx", [False, True])
@pytest.mark.parametrize("unitary", [False, True])
def test_gpu_hessenberg_bitwise_vs_oracle(A, ctx_strict, oracle, cx, unitary):
    """Library built without FMA contraction, IEEE turnover: the warp-cooperative kernel (rotations of the active window only,
    split over 32 lanes) gives the oracle's eigenvalues bit for bit for real matrices, ragged batch, sizes 0 and 1 included.
    Complex matrices go through hypot and the complex division, where CUDA's and glibc's last bits differ (as on the companion
    path, where only the real double shift kernel is claimed bit-identical): those are held to 1e-12 n max|lambda|."""
    rng = np.random.default_rng(100 + 2 * cx + unitary)
    # the first batch keeps R in the global workspace (largest matrix 90), the second in shared memory (largest 26)
    for dims in ([15, 1, 2, 0, 3, 40, 33, 7, 64, 5, 90, 2], [15, 1, 2, 0, 3, 26, 7, 5, 22, 2, 19, 11, 4]):
        mats = [(unitary_hess(rng, n, cx) if n >= 2 else hess(rng, n, cx)) if unitary else hess(rng, n, cx, "randn") for n in dims]
        got, nunc = ctx_strict.eigvals_hessenberg(mats, unitary=unitary)
        assert np.all(nunc == 0)
        assert ctx_strict.last_stats["lanes"] == (0 if max(dims) > 39 else 1)      # where R lived (amrvw_stats.lanes of this entry)
        for p, (M, g) in enumerate(zip(mats, got)):
            n = M.shape[0]
            assert len(g) == n
            if n == 0:
                continue
            want, st = oracle.eigvals_hessenberg(M, unitary=unitary, index=p)
            if cx:
                assert dist(g, want) <= 1e-12 * n * max(1.0, np.abs(want).max()), f"n = {n}: |diff| = {dist(g, want):.3g}"
            else:
                assert np.array_equal(g.real, want.real) and np.array_equal(g.imag, want.imag), f"n = {n}: max |diff| = {np.abs(g - want).max():.3g}"
    assert ctx_strict.last_stats["sweeps"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("cx", [False, True])
def test_gpu_hessenberg_default_build_parity(A, ctx, oracle, cx):
    """Default build (FMA, division-free turnovers): eigenvalues within the reference's own tolerance of the oracle's and of
    LAPACK's; R in shared memory (n = 24) and in the global workspace (n = 100, 200; complex: 150)."""
    rng = np.random.default_rng(5 + cx)
    for n, B in ((24, 40), (100, 6), (150 if cx else 200, 3)):
        mats = [hess(rng, n, cx, "randn") for _ in range(B)]
        got, nunc = ctx.eigvals_hessenberg(mats)
        assert np.all(nunc == 0)
        for M, g in zip(mats, got):
            want, _ = oracle.eigvals_hessenberg(M)
            w = np.linalg.eigvals(M)
            scale = max(1.0, np.abs(w).max())
            assert dist(g, want) <= SQEPS * scale and dist(g, w) <= SQEPS * scale     # test-roots.jl:199: norm(e1 - e2) <= sqrt(eps())
            if not cx:
                assert np.sum(g.imag == 0) == np.sum(want.imag == 0)


@pytest.mark.gpu
def test_gpu_hessenberg_large_batch_and_options(A, oracle):
    """2000 small matrices in one call (more than the resident warps: the work counter hands them out), and the reference's
    constants as options reach this kernel too (it_count = 6 still converges)."""
    rng = np.random.default_rng(77)
    mats = [hess(rng, int(n), False) for n in rng.integers(2, 24, 2000)]
    c = A.Context(seed=0, it_count=6)
    try:
        got, nunc = c.eigvals_hessenberg(mats)
    finally:
        c.close()
    assert np.all(nunc == 0)
    for k in range(0, 2000, 97):
        assert dist(got[k], np.linalg.eigvals(mats[k])) <= SQEPS
    with pytest.raises(ValueError):
        A.qr_factorization(np.zeros((3, 4)))


@pytest.mark.gpu
def test_gpu_hessenberg_edge_cases(A, ctx):
    """Empty batch, a matrix beyond the documented size limit (argument error, nothing launched), and non-finite entries (the trip
    cap ends the iteration: slots reported unconverged, no hang)."""
    got, nunc = ctx.eigvals_hessenberg([])
    assert got == [] and len(nunc) == 0
    with pytest.raises(A.AmrvwError):
        ctx.eigvals_hessenberg_packed(np.zeros(8000 * 8000), np.array([8000], dtype=np.int32))
    rng = np.random.default_rng(4)
    G = hess(rng, 6, False, "randn")
    M = G.copy()
    M[2, 3] = np.nan
    got, nunc = ctx.eigvals_hessenberg([M, G])
    assert nunc[0] > 0 and nunc[1] == 0
    assert dist(got[1], np.linalg.eigvals(G)) <= SQEPS
    # the iteration cap as an option: it_max_factor = 1 leaves most of a 40 x 40 matrix unconverged, the default does not
    rng = np.random.default_rng(5)
    H = hess(rng, 40, False, "randn")
    c = A.Context(seed=0, it_max_factor=1)
    try:
        _, few = c.eigvals_hessenberg([H])
    finally:
        c.close()
    _, full = ctx.eigvals_hessenberg([H])
    assert few[0] > 0 and full[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("cx", [False, True])
def test_gpu_hessenberg_thread_per_matrix_kernel(A, ctx_strict, oracle, cx):
    """Batches of at least 64 x #SM small matrices run one THREAD per matrix (hess_thread_kernel, amrvw_stats.lanes == 2): same
    code at cooperation width 1, so the strict build is again bit-identical to the oracle for real matrices; ragged sizes 0..15."""
    rng = np.random.default_rng(900 + cx)
    dims = rng.integers(0, 16, 12000)
    dims[:4] = [15, 15, 2, 1]
    mats = [hess(rng, int(n), cx, "randn") for n in dims]
    got, nunc = ctx_strict.eigvals_hessenberg(mats)
    assert ctx_strict.last_stats["lanes"] == 2 and np.all(nunc == 0)
    for p in list(range(0, 12000, 61)) + [0, 1, 2, 3]:
        n = int(dims[p])
        assert len(got[p]) == n
        if n == 0:
            continue
        want, ost = oracle.eigvals_hessenberg(mats[p], index=p)
        if cx or ost["exceptional"]:      # an exceptional shift goes through sincos, whose last bits differ between CUDA and glibc
            assert dist(got[p], want) <= 1e-12 * n * max(1.0, np.abs(want).max())
        else:
            assert np.array_equal(got[p].real, want.real) and np.array_equal(got[p].imag, want.imag), f"matrix {p}, n = {n}"
    # unitary mode on the same kernel
    um = [unitary_hess(rng, 8 + (p % 8), cx) for p in range(10000)]
    ug, nunc = ctx_strict.eigvals_hessenberg(um, unitary=True)
    assert ctx_strict.last_stats["lanes"] == 2
    # the reference algorithm itself stagnates on a few random orthogonal matrices in 10000 (all eigenvalues on the unit circle; the
    # oracle leaves the same matrices unconverged after 20 n trips, in either R mode): the kernel must agree on which
    bad = np.nonzero(nunc)[0]
    assert len(bad) <= 10
    for p in bad:
        _, ost = oracle.eigvals_hessenberg(um[p], unitary=True, index=int(p))
        assert ost["unconverged"] > 0
    for p in range(0, 10000, 173):
        if nunc[p] == 0:
            assert dist(ug[p], np.linalg.eigvals(um[p])) <= SQEPS


@pytest.mark.gpu
@pytest.mark.parametrize("cx", [False, True])
def test_gpu_hessenberg_device_pointers(A, ctx, cx):
    """amrvw_eigvals_hessenberg_*_dev: matrices, offsets and results stay in HBM (torch owns the buffers), the launch runs on
    torch's current stream; same eigenvalues as the host-buffer call."""
    import torch
    from importlib import import_module
    pack = import_module("amrvw_jl_b200.roots").pack_matrices
    rng = np.random.default_rng(21 + cx)
    mats = [hess(rng, int(n), cx, "randn") for n in rng.integers(1, 50, 64)]
    want, _ = ctx.eigvals_hessenberg(mats)
    flat, dims = pack(mats)
    d64 = dims.astype(np.int64)
    moff = np.concatenate([[0], np.cumsum(d64 * d64)]); eoff = np.concatenate([[0], np.cumsum(d64)])
    host = torch.from_numpy(np.ascontiguousarray(flat).view(np.float64)) if cx else torch.from_numpy(flat)
    d_H = host.cuda(); d_dims = torch.from_numpy(dims).cuda()
    d_moff = torch.from_numpy(moff).cuda(); d_eoff = torch.from_numpy(eoff).cuda()
    d_eig = torch.zeros(2 * int(eoff[-1]) + 2, dtype=torch.float64, device="cuda"); d_nu = torch.ones(len(mats), dtype=torch.int32, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.eigvals_hessenberg_dev(cx, d_H.data_ptr(), d_dims.data_ptr(), d_moff.data_ptr(), d_eoff.data_ptr(), len(mats), int(dims.max()),
                                   d_eig.data_ptr(), d_nu.data_ptr(), want_stats=True)
        assert ctx.last_stats["launches"] == 1 and ctx.last_stats["sweeps"] > 0
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    got = d_eig.cpu().numpy()[: 2 * int(eoff[-1])].view(np.complex128)
    assert int(d_nu.sum().item()) == 0
    for p, w in enumerate(want):
        assert np.array_equal(got[eoff[p]:eoff[p + 1]], w)
