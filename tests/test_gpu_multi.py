"""Multi-GPU inside the library (amrvw_create_multi): ONE host-buffer call roots the whole batch
on several GPUs -- contiguous split by polynomial, one stream per device, results copied straight into the
caller's buffers.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _rows(oracle, seed, B, n, kind=0):
    return [oracle.fill_coeffs(seed, p, kind, n + 1) for p in range(B)]


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_multi_device_context_is_bit_identical_to_one_device(A, oracle, ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    rng = np.random.default_rng(3)
    # ragged real batch with empty vectors and closed-form degrees in it
    rows = []
    for p in range(61):
        n = int(rng.integers(3, 400))
        rows.append(oracle.fill_coeffs(41, p, 0, n + 1))
    rows[5] = np.zeros(0); rows[17] = np.array([1.0, 2.0, 1.0]); rows[40] = np.concatenate([[0.0, 0.0], rows[40]])
    one = A.Context(device=0, seed=7)
    many = A.Context(devices=list(range(ndev)), seed=7)
    try:
        assert many.device_count == ndev and one.device_count == 1
        a = A.roots(rows, ctx=one)
        b = A.roots(rows, ctx=many)
        st = dict(many.last_stats)
        assert len(a) == len(b)
        # same kernels, same per-polynomial seeds; the lane count is chosen per device from its share of the
        # batch, so compare through the parity rule when the shares differ in schedule, bit for bit otherwise
        for x, y in zip(a, b):
            assert len(x) == len(y)
            assert np.array_equal(np.sort_complex(x), np.sort_complex(y)) or np.max(np.abs(np.sort_complex(x) - np.sort_complex(y))) < 1e-9
        assert st["unconverged"] == 0 and st["launches"] >= 2 * 5
        # complex and pencil entry points
        crow = [oracle.fill_coeffs(43, p, 1, 2 * 81) for p in range(24)]
        crow = [r[0::2] + 1j * r[1::2] for r in crow]
        ca = A.roots(crow, ctx=one); cb = A.roots(crow, ctx=many)
        for x, y in zip(ca, cb):
            assert np.max(np.abs(x - y)) < 1e-10
        pairs = [A.basic_pencil(r) for r in _rows(oracle, 44, 24, 90)]
        pa = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=one)
        pb = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=many)
        for x, y in zip(pa, pb):
            assert np.max(np.abs(x - y)) < 1e-9
        # counts entry point
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        c1, _ = one.count_real_roots_csr(flat, off)
        c2, _ = many.count_real_roots_csr(flat, off)
        assert np.array_equal(c1, c2)
        with pytest.raises(A.AmrvwError):
            many.set_stream(0)
    finally:
        one.close(); many.close()


def test_multi_device_same_schedule_is_bitwise(A, oracle):
    """With the lane count pinned (so every device runs the schedule the single device runs) the multi-device
    roots are bit-identical to the single-device call."""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    rows = _rows(oracle, 45, 96, 300)
    one = A.Context(device=0, seed=1, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=16)
    many = A.Context(devices=[0, 1], seed=1, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=16)
    try:
        a = A.roots(rows, ctx=one)
        b = A.roots(rows, ctx=many)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert many.last_stats["turnovers"] == one.last_stats["turnovers"]
    finally:
        one.close(); many.close()


def test_multi_device_c2_share_e2e(A, oracle):
    """One call, 2 GPUs, degree-3000 polynomials: the per-device share goes through whatever schedule its size
    selects; nothing unconverged, exact conjugate pairs."""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    rows = _rows(oracle, 2, 64, 3000)
    many = A.Context(devices="all", seed=2)
    try:
        got = A.roots(rows, ctx=many)
        assert many.last_stats["unconverged"] == 0
        for g in got[::8]:
            pos = np.sort_complex(g[g.imag > 0]); neg = np.sort_complex(np.conj(g[g.imag < 0]))
            assert len(g) == 3000 and np.array_equal(pos, neg)
    finally:
        many.close()


@pytest.mark.parametrize("ndev", [2, 4])
def test_multi_device_hessenberg_split(A, ndev):
    """amrvw_eigvals_hessenberg_* on a multi-device context: contiguous split by matrix balanced by n^3, eigenvalues land in the
    caller's buffer at the single-device slots; identical values (no exceptional shift is drawn on these matrices)."""
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    rng = np.random.default_rng(11)
    mats = [np.triu(rng.standard_normal((int(n), int(n))), -1) for n in rng.integers(0, 70, 90)]
    one = A.Context(device=0, seed=7)
    many = A.Context(devices=list(range(ndev)), seed=7)
    try:
        a, ua = one.eigvals_hessenberg(mats)
        b, ub = many.eigvals_hessenberg(mats)
        assert many.last_stats["launches"] == ndev and np.all(ua == 0) and np.all(ub == 0)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        cm = [m + 1j * np.triu(rng.standard_normal(m.shape), -1) for m in mats[:30]]
        ca, _ = one.eigvals_hessenberg(cm, unitary=False)
        cb, _ = many.eigvals_hessenberg(cm, unitary=False)
        for x, y in zip(ca, cb):
            assert np.array_equal(x, y)
    finally:
        one.close(); many.close()
