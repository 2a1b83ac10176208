"""CPU tests: pin the oracle on the reference's known answers.

Every check here restates a test of /root/reference/test or compares against an output the
reference itself printed (tests/golden/reference_notebook.json)."""
import numpy as np
import pytest

from conftest import pencil_split, wilkinson_coeffs, match_roots

p4 = np.array([24.0, -50.0, 35.0, -10.0, 1.0])
p5 = np.array([-120.0, 274.0, -225.0, 85.0, -15.0, 1.0])
p6 = np.array([720.0, -1764.0, 1624.0, -735.0, 175.0, -21.0, 1.0])
pc = np.array([40.0 + 10.0j, -76.0 + 60.0j, 10.0 - 95.0j, 25.0 + 40.0j, -10.0 - 5.0j, 1.0 + 0.0j])   # test-roots.jl:13
pc_rts = np.array([1j, 1 + 1j, 2 + 1j, 3 + 1j, 4 + 1j])
pc2 = np.array([1j, -1, 0, 0, -1j, 1])                                                                   # test-amrvw-basics.jl:11
SQEPS = np.sqrt(np.finfo(float).eps)


def _g(rows):
    return np.array([complex(float(a), float(b)) for a, b in rows])


def basic_pencil(ps):
    ps = np.asarray(ps); qs = np.zeros(len(ps) - 1, dtype=ps.dtype); qs[-1] = ps[-1]
    return ps[:-1].copy(), qs


# ------------------------------------------------------------------ outputs the reference printed (bit for bit)
def test_notebook_p4_bitwise(oracle, golden):
    r, _ = oracle.roots(p4)
    want = _g(golden["p4"]["roots"])
    assert np.array_equal(r.real, want.real) and np.all(r.imag == 0)


def test_notebook_p4_pencil_bitwise(oracle, golden):
    r, _ = oracle.roots(*basic_pencil(p4)[:1], ws=basic_pencil(p4)[1])
    want = _g(golden["p4_pencil"]["roots"])
    assert np.array_equal(r.real, want.real) and np.all(r.imag == 0)


def test_notebook_pc_bitwise(oracle, golden):
    """Complex single shift: every digit the reference printed, compared as a multiset (the sign of a
    zero real part, which decides the tie order -0.0 < 0.0 in sorteig!, comes from Julia's complex
    division and is outside /root/reference)."""
    r, _ = oracle.roots(pc2)
    want = _g(golden["pc"]["roots"])
    key = lambda z: (abs(z.real) if z.real == 0 else z.real, z.imag)
    assert sorted(map(key, r)) == sorted(map(key, want))


def test_notebook_B_chain_and_rho(oracle, golden):
    st = oracle.FlatState(ps=p4)
    c, s = st._chains()["B"]
    want = golden["p4_B_chain"]["rotators"]
    assert [(float(a), float(b)) for a, b, _ in want] == list(zip(c.tolist(), s.tolist()))
    Ct = st.dense_chain("Ct", 5)
    assert Ct[-1, 0].real == pytest.approx(float(golden["p4_rho"]["value"]), rel=1e-14)


def test_notebook_wilkinson(oracle, golden):
    """Takes one exceptional shift (Base.rand in the reference) -> tolerance only; the agreement to
    1e-9 also pins the coefficient construction and the ill-conditioned spurious complex pairs."""
    ps = wilkinson_coeffs(20)
    r, st = oracle.roots(ps)
    want = _g(golden["wilkinson20"]["roots"])
    d, _ = match_roots(r, want)
    assert st["exceptional"] >= 1
    assert d.max() < 1e-8
    assert np.sum(r.imag == 0) == np.sum(want.imag == 0) == 12
    r2, _ = oracle.roots(*pencil_split(ps, 13)[:1], ws=pencil_split(ps, 13)[1])
    want2 = _g(golden["wilkinson20_pencil13"]["roots"])
    d2, _ = match_roots(r2, want2)
    assert d2.max() < 1e-6 and np.all(r2.imag == 0)


# ------------------------------------------------------------------ test/test-roots.jl
@pytest.mark.parametrize("p", [p4, p5, p6])
def test_rds_known_answers(oracle, p):   # test-roots.jl:113-123
    r, _ = oracle.roots(p)
    d = len(p) - 1
    assert np.allclose(r.real, np.arange(1, d + 1), rtol=SQEPS) and np.all(r.imag == 0)


def test_css_known_answer(oracle):       # test-roots.jl:135
    r, _ = oracle.roots(pc)
    assert np.all(np.abs(r - pc_rts) <= SQEPS)


def test_second_pc(oracle):              # test-amrvw-basics.jl:199-202
    r, _ = oracle.roots(pc2)
    V = np.array([-1, -1j, 1j, 1j, 1])
    assert all(np.min(np.abs(np.round(z, 2) - V)) == 0 for z in r)


@pytest.mark.parametrize("p", [p4, p5, p6])
def test_rds_qz(oracle, p):              # test-roots.jl:146-153
    vs, ws = basic_pencil(p)
    r, _ = oracle.roots(vs, ws=ws)
    assert np.allclose(r.real, np.arange(1, len(p)), rtol=SQEPS) and np.all(r.imag == 0)


def test_css_qz(oracle):                 # test-roots.jl:166-167
    vs, ws = basic_pencil(pc)
    r, _ = oracle.roots(vs, ws=ws)
    assert np.allclose(np.round(r, 12), pc_rts)


@pytest.mark.parametrize("i", [1, 2, 3, 4])
def test_pencil_split(oracle, i):        # test-roots.jl:184-188
    vs, ws = pencil_split(p4, i)
    r, _ = oracle.roots(vs, ws=ws)
    assert np.linalg.norm(r - np.array([1, 2, 3, 4])) <= SQEPS


def test_smoke_random_500(oracle):       # test-roots.jl:125-128,137-139,155-158
    rng = np.random.default_rng(7)
    rs = rng.random(500)
    r, st = oracle.roots(rs)
    assert len(r) == 499 and np.all(r != 0) and st["unconverged"] == 0
    rc = rng.random(500) + 1j * rng.random(500)
    r, st = oracle.roots(rc)
    assert len(r) == 499 and np.all(r != 0) and st["unconverged"] == 0
    vs, ws = basic_pencil(rs)
    r, st = oracle.roots(vs, ws=ws)
    assert len(r) == 499 and np.all(r != 0)


SIMPLE = [(np.array([]), np.array([])), (np.array([1.0]), np.array([])), (np.array([1.0, 1.0]), np.array([-1.0])),
          (np.array([1.0, 2.0, 1.0]), np.array([-1.0, -1.0]))]


@pytest.mark.parametrize("cx", [False, True])
@pytest.mark.parametrize("ps,rts", SIMPLE)
def test_simple_cases(oracle, ps, rts, cx):   # test-roots.jl:18-80
    if cx:
        ps = ps.astype(complex)
    z2 = np.zeros(2, dtype=ps.dtype)

    def check(inp, want):
        r, _ = oracle.roots(inp)
        assert len(r) == len(want) and np.all(r == want)
    check(ps, rts)
    check(np.concatenate([z2, ps]), np.concatenate([rts, [0, 0]]) if len(ps) else rts)
    check(np.concatenate([ps, z2]), rts)
    check(np.concatenate([z2, ps, z2]), np.concatenate([rts, [0, 0]]) if len(ps) else rts)


@pytest.mark.parametrize("ps,rts", SIMPLE[2:])
def test_simple_cases_pencil(oracle, ps, rts):   # test-roots.jl:83-107
    ps = ps.astype(complex)
    z2 = np.zeros(2, dtype=complex)
    for inp, want in [(ps, rts), (np.concatenate([z2, ps]), np.concatenate([rts, [0, 0]])), (np.concatenate([ps, z2]), rts),
                      (np.concatenate([z2, ps, z2]), np.concatenate([rts, [0, 0]]))]:
        vs, ws = basic_pencil(inp)
        r, _ = oracle.roots(vs, ws=ws)
        assert len(r) == len(want) and np.all(r == want)


# ------------------------------------------------------------------ test/test-amrvw-basics.jl "transformations"
def rotm(c, s, i, n):
    M = np.eye(n, dtype=complex)
    M[i - 1:i + 1, i - 1:i + 1] = [[c, s], [-s, np.conj(c)]]
    return M


def random_rotator(rng, cx):   # diagnostics.jl:12-25
    if cx:
        a, b, c = rng.standard_normal(3)
        n = np.sqrt(a * a + b * b + c * c)
        return complex(a / n, b / n), c / n
    t = rng.random() * 2 * np.pi
    return np.cos(t), np.sin(t)


def test_givensrot(oracle):    # test-amrvw-basics.jl:26-32, 56-62
    rng = np.random.default_rng(3)
    for _ in range(20):
        a, b = rng.random(2)
        c, s, r = oracle.givensrot(a, b)
        v = rotm(c, s, 1, 2) @ np.array([a, b])
        assert abs(v[1]) <= SQEPS and abs(v[0] - r) <= 1e-15 and r >= 0
        a, b = rng.random(2) + 1j * rng.random(2)
        c, s, r = oracle.givensrot(a, b)
        v = rotm(c, s, 1, 2) @ np.array([a, b])
        assert abs(v[1]) <= SQEPS and abs(v[0] - r) <= 4e-16 * abs(r)


@pytest.mark.parametrize("cx", [False, True])
def test_turnover_preserves_product(oracle, cx):   # test-amrvw-basics.jl:35-42, 79-86
    rng = np.random.default_rng(11)
    eps = np.finfo(float).eps
    for _ in range(50):
        U, V, W = (random_rotator(rng, cx) for _ in range(3))
        U1, V1, W1 = oracle.turnover(U, V, W)
        lhs = rotm(*U, 2, 4) @ rotm(*V, 3, 4) @ rotm(*W, 2, 4)
        rhs = rotm(*U1, 3, 4) @ rotm(*V1, 2, 4) @ rotm(*W1, 3, 4)
        assert np.max(np.abs(lhs - rhs)) <= 8 * eps
        # the other orientation (i, i-1, i) uses the same formulas
        lhs = rotm(*U, 3, 4) @ rotm(*V, 2, 4) @ rotm(*W, 3, 4)
        rhs = rotm(*U1, 2, 4) @ rotm(*V1, 3, 4) @ rotm(*W1, 2, 4)
        assert np.max(np.abs(lhs - rhs)) <= 8 * eps


def test_turnover_regression_triple(oracle):       # test-amrvw-basics.jl:68-77
    U = (-0.9990813749617048 + 0.042846941864777804j, -0.0007387675316362891)
    V = (0.20398474677465908 - 0.08295034712193844j, -0.9754534653153004)
    W = (0.03541646337113308 + 0.28843087042807297j, -0.9568454980331911)
    U1, V1, W1 = oracle.turnover(U, V, W)
    lhs = rotm(*U, 2, 4) @ rotm(*V, 1, 4) @ rotm(*W, 2, 4)
    rhs = rotm(*U1, 1, 4) @ rotm(*V1, 2, 4) @ rotm(*W1, 1, 4)
    assert np.allclose(lhs, rhs, rtol=SQEPS, atol=1e-15)


def test_turnover_zero_sines(oracle):              # test-amrvw-basics.jl:89-117
    rng = np.random.default_rng(5)
    eps = np.finfo(float).eps
    diag = (complex(np.sin(np.pi / 4), np.cos(np.pi / 4)), 0.0)
    for which in range(3):
        rots = [random_rotator(rng, True) for _ in range(3)]
        rots[which] = diag
        U, V, W = rots
        U1, V1, W1 = oracle.turnover(U, V, W)
        lhs = rotm(*U, 1, 3) @ rotm(*V, 2, 3) @ rotm(*W, 1, 3)
        rhs = rotm(*U1, 2, 3) @ rotm(*V1, 1, 3) @ rotm(*W1, 2, 3)
        assert np.max(np.abs(lhs - rhs)) <= 8 * eps


def test_fuse(oracle):                             # test-amrvw-basics.jl:46-51, 120-128
    rng = np.random.default_rng(9)
    eps = np.finfo(float).eps
    for _ in range(20):
        U, V = random_rotator(rng, False), random_rotator(rng, False)
        UV, _ = oracle.fuse(U, V)
        assert np.max(np.abs(rotm(*UV, 2, 4) - rotm(*U, 2, 4) @ rotm(*V, 2, 4))) <= 4 * eps
        U, V = random_rotator(rng, True), random_rotator(rng, True)
        UV, alpha = oracle.fuse(U, V)
        D = np.diag([1, alpha, np.conj(alpha), 1])
        assert np.linalg.norm(rotm(*UV, 2, 4) @ D - rotm(*U, 2, 4) @ rotm(*V, 2, 4)) <= SQEPS


# ------------------------------------------------------------------ "factorization", "bulge step", "diagonal block"
def basic_decompose(p):   # companion.jl:6-24
    p = np.asarray(p); n = len(p) - 1
    par = 1.0 if n % 2 == 0 else -1.0
    qs = np.zeros(n + 1, dtype=p.dtype)
    for i in range(1, n):
        qs[i - 1] = par * (-1) ** (i + 1) * p[i] / p[-1]
    qs[n - 1] = par * p[0] / p[-1]
    qs[n] = -1
    return qs


@pytest.mark.parametrize("p", [p4, p5, p6])
def test_factorization_real(oracle, p):   # test-amrvw-basics.jl:135-176
    d = len(p) - 1
    st = oracle.FlatState(ps=p)
    n = d + 1
    xs = basic_decompose(p)
    C = st.dense_chain("Ct", n).conj().T
    y = C @ xs
    assert abs(y[0] - np.linalg.norm(xs)) <= SQEPS and np.linalg.norm(y[1:]) <= SQEPS
    assert np.linalg.norm(np.sort(np.linalg.eigvals(st.dense()).real) - np.arange(1, d + 1)) <= 1e-6
    ch = st._chains()
    ctc, cts = ch["Ct"]; bc, bs = ch["B"]
    M = np.eye(n, dtype=complex)
    for i in range(d - 1, 0, -1):   # Ct.x[2:end] = indices N-1..1 (stored descending)
        M = M @ rotm(ctc[i - 1], cts[i - 1], i, n)
    for i in range(1, d):           # B.x[1:end-1]
        M = M @ rotm(bc[i - 1], bs[i - 1], i, n)
    assert np.linalg.norm(M - np.eye(n)) <= SQEPS
    T = rotm(ctc[d - 1], cts[d - 1], d, n) @ rotm(bc[d - 1], bs[d - 1], d, n)
    assert np.linalg.norm(T[-2:, -2:] - np.array([[0, -1], [1, 0]])) <= SQEPS
    Z = st.dense_chain("Ct", n) @ st.dense_chain("B", n)
    en = np.zeros(n); en[-2] = 1
    X = st.dense_chain("Q", n) @ (Z + np.outer(xs, en))
    assert np.linalg.norm(np.sort(np.linalg.eigvals(X).real) - np.arange(0, d + 1)) <= 1e-6


def test_factorization_complex(oracle):   # test-amrvw-basics.jl:181-206
    st = oracle.FlatState(ps=pc2)
    n = len(pc2)
    xs = basic_decompose(pc2.astype(complex))
    C = st.dense_chain("Ct", n).conj().T
    y = C @ xs
    assert abs(abs(y[0]) - np.linalg.norm(xs)) <= 1e-12 and np.linalg.norm(y[1:]) <= SQEPS
    U = np.round(np.linalg.eigvals(st.dense()), 2)
    V = np.array([-1, -1j, 1j, 1j, 1])
    assert all(np.min(np.abs(u - V)) < 1e-9 for u in U)


@pytest.mark.parametrize("p", [p4, p5, p6])
def test_bulge_step_preserves_eigenvalues(oracle, p):   # test-amrvw-basics.jl:213-235
    d = len(p) - 1
    st = oracle.FlatState(ps=p)
    ctr = [0, 1, d - 1, 0, d - 2]     # make_counter: it_count = 0 -> the exceptional (random) shift branch
    ctr = st.bulge_step(ctr, seed=3)
    assert np.linalg.norm(np.sort(np.linalg.eigvals(st.dense()).real) - np.arange(1, d + 1)) <= 100 * SQEPS
    ctr[3] = 14
    st.bulge_step(ctr)
    assert np.linalg.norm(np.sort(np.linalg.eigvals(st.dense()).real) - np.arange(1, d + 1)) <= 100 * SQEPS


def test_bulge_step_complex_and_pencil(oracle):
    st = oracle.FlatState(ps=pc)
    ctr = st.bulge_step([0, 1, 4, 0, 3], seed=1)
    ctr[3] = 14
    st.bulge_step(ctr)
    ev = np.linalg.eigvals(st.dense())
    d, _ = match_roots(ev, pc_rts)
    assert d.max() <= 1e-6
    vs, ws = pencil_split(p6, 3)
    st = oracle.FlatState(vs=vs, ws=ws)
    assert np.linalg.norm(np.sort(np.linalg.eigvals(st.dense()).real) - np.arange(1, 7)) <= 1e-6
    ctr = st.bulge_step([0, 1, 5, 15, 4])
    assert np.linalg.norm(np.sort(np.linalg.eigvals(st.dense()).real) - np.arange(1, 7)) <= 1e-5


@pytest.mark.parametrize("p", [p4, p5, p6, pc])
def test_diagonal_block_matches_dense(oracle, p):   # test-amrvw-basics.jl:259-289
    d = len(p) - 1
    st = oracle.FlatState(ps=p)
    ctr = st.bulge_step([0, 1, d - 1, 0, d - 2], seed=2)
    ctr[3] += 1
    st.bulge_step(ctr)
    M = st.dense()
    R = st.dense_R("V")
    for j in range(1, d):
        A = st.diagonal_block(j)
        assert np.allclose(M[j - 1:j + 1, j - 1:j + 1], A, rtol=1e-8, atol=1e-10)
    assert np.isfinite(R).all()


def test_eigvals_state(oracle):   # AMRVW_algorithm.jl:207-213 on a factorization
    st = oracle.FlatState(ps=p5)
    ev, unconv = st.eigvals()
    assert unconv == 0 and np.allclose(ev.real, np.arange(1, 6)) and np.all(ev.imag == 0)


# ------------------------------------------------------------------ independent cross-checks + statistics
@pytest.mark.parametrize("n", [10, 50, 200])
@pytest.mark.parametrize("kind", ["rds", "css", "pencil"])
def test_against_numpy_roots(oracle, n, kind):
    rng = np.random.default_rng(100 + n)
    if kind == "css":
        p = rng.standard_normal(n + 1) + 1j * rng.standard_normal(n + 1)
        r, st = oracle.roots(p)
    elif kind == "pencil":
        p = rng.standard_normal(n + 1)
        vs, ws = pencil_split(p, max(1, n // 2))
        r, st = oracle.roots(vs, ws=ws)
    else:
        p = rng.standard_normal(n + 1)
        r, st = oracle.roots(p)
    ref = np.roots(p[::-1])
    d, _ = match_roots(r, ref)
    assert st["unconverged"] == 0 and d.max() < 1e-9


def test_exceptional_shift_inputs(oracle):   # x^n - 1 stagnates for 15 sweeps
    p = np.zeros(21); p[0] = -1; p[-1] = 1
    r, st = oracle.roots(p)
    assert st["exceptional"] >= 1
    d, _ = match_roots(r, np.exp(2j * np.pi * np.arange(20) / 20))
    assert d.max() < 1e-12


def test_multishift_schedule_same_roots(oracle):
    rng = np.random.default_rng(42)
    for n in (60, 257):
        p = rng.standard_normal(n + 1)
        r0, s0 = oracle.roots(p)
        for L, reuse in ((4, 1), (8, 2), (16, 2)):
            r1, s1 = oracle.roots(p, sched=1, L=L, small=max(16, 2 * L // reuse + 2), reuse=reuse)
            d, _ = match_roots(r1, r0)
            assert s1["unconverged"] == 0 and d.max() < 1e-10
            assert np.sum(r1.imag == 0) == np.sum(r0.imag == 0)


def test_dense_trailing_block_and_hessenberg_qr(oracle):
    """The shifts of a fat step (oracle ms_shifts_dense = GPU shifts.cuh): the trailing block written out
    entry by entry must equal the dense matrix of the factorization, and the Hessenberg QR must find
    its eigenvalues (numpy) -- which are also what the core-chasing sub-solve returns."""
    rng = np.random.default_rng(7)
    for n, m in ((40, 7), (64, 15), (90, 31)):
        p = rng.standard_normal(n + 1)
        st = oracle.FlatState(ps=p)
        ctr = [0, 1, n - 1, 15, -1]
        for _ in range(6):                    # a few ordinary sweeps so that the block is not the initial companion
            ctr = st.bulge_step(ctr)
            ctr[3] = 15
        H, dense, chase, rc = st.ms_shift_check(ctr, m)
        assert rc == 0
        stop, k0 = ctr[2], ctr[2] - m
        A = st.dense()                          # slot k <-> row k-1
        blk = A[k0:stop + 1, k0:stop + 1].copy()
        # Q[k0] is made diagonal for the shifts: that only rescales the first row by 1/|c_k0| and drops what
        # couples the block to the rows above it, so compare everything except the first row
        assert np.allclose(H[1:, :], blk[1:, :], rtol=1e-9, atol=1e-11)
        assert np.allclose(np.tril(H, -2), 0.0)
        ev = np.linalg.eigvals(H)
        d, _ = match_roots(dense, ev)
        assert d.max() < 1e-9 * max(1.0, np.abs(ev).max())
        d, _ = match_roots(dense, chase)
        assert d.max() < 1e-8 * max(1.0, np.abs(ev).max())
        # conjugate pairs come out adjacent and exactly conjugate (the pairing of the real double shift relies on it)
        k = 0
        while k < len(dense):
            if dense[k].imag != 0:
                assert dense[k + 1] == np.conj(dense[k]); k += 2
            else:
                k += 1


def test_multishift_dense_shifts_same_roots(oracle):
    rng = np.random.default_rng(43)
    for n in (120, 400):
        p = rng.standard_normal(n + 1)
        r0, s0 = oracle.roots(p)
        for L, reuse in ((8, 2), (16, 2), (16, 1), (32, 4)):
            small = max(24, 2 * L // reuse + 8)
            r1, s1 = oracle.roots(p, sched=1, L=L, small=small, reuse=reuse, dense=1)
            r2, s2 = oracle.roots(p, sched=1, L=L, small=small, reuse=reuse, dense=0)
            d, _ = match_roots(r1, r0)
            assert s1["unconverged"] == 0 and d.max() < 1e-10
            assert np.sum(r1.imag == 0) == np.sum(r0.imag == 0)
            assert abs(s1["fat_steps"] - s2["fat_steps"]) <= max(3, s2["fat_steps"] // 10)     # same convergence as the sub-solve's shifts
            assert s1["serial_turnovers"] < s2["serial_turnovers"]


def test_readme_statistic_miniature(oracle):
    """README.md:121-127 in miniature: mean number of exactly-real roots of randn polynomials
    against the Kac-type formula (count(imag == 0) semantics of real_polynomial_roots)."""
    n, B = 200, 300
    coeffs = np.concatenate([oracle.fill_coeffs(2, p, 0, n + 1) for p in range(B)])
    off = np.arange(B + 1, dtype=np.int64) * (n + 1)
    out, roff, nr, st = oracle.roots_csr(coeffs, off)
    counts = np.array([np.sum(out[roff[p]:roff[p] + nr[p]].imag == 0) for p in range(B)])
    expect = 2 / np.pi * np.log(n) + 0.625738072 + 2 / (np.pi * n)
    assert abs(counts.mean() - expect) < 4 * counts.std() / np.sqrt(B)
    assert st["unconverged"] == 0


def test_oracle_pencil_multishift_schedule_matches_reference_schedule(oracle):
    """The sequential twin of the GPU pencil unit schedule (amrvw_algorithm_ms on the pencil state, shifts from
    dense_trailing_block_pencil or the core-chasing sub-solve) finds the roots of the reference's one-bulge
    schedule: minimum-distance matching within the conditioning-scaled tolerance, equal real-root counts."""
    from conftest import match_roots, pencil_split, root_condition_tolerance
    rows = [oracle.fill_coeffs(44, i, 0, d + 1) for i, d in enumerate([120, 97, 64, 33])]
    for split in ("basic", "half"):
        pairs = []
        for r in rows:
            if split == "basic":
                vs_ = r[:-1].copy(); ws_ = np.zeros(len(r) - 1); ws_[-1] = r[-1]      # basic_pencil (companion.jl:44-54)
                pairs.append((vs_, ws_))
            else:
                pairs.append(pencil_split(r, (len(r) - 1) // 2))
        vs = np.concatenate([v for v, _ in pairs]); ws = np.concatenate([w for _, w in pairs])
        off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(v) for v, _ in pairs], out=off[1:])
        ref = oracle.roots_csr(vs, off, ws=ws)
        for L, reuse, small, dense in ((8, 2, 16, 1), (16, 2, 24, 0), (4, 1, 16, 1)):
            out, roff, nr, tot = oracle.roots_csr(vs, off, ws=ws, sched=1, L=L, small=small, reuse=reuse, dense=dense)
            assert tot["fat_steps"] > 0 and tot["unconverged"] == 0
            for p, r in enumerate(rows):
                g = out[roff[p]:roff[p] + nr[p]]; w = ref[0][ref[1][p]:ref[1][p] + ref[2][p]]
                d, idx = match_roots(g, w)
                assert np.max(d / root_condition_tolerance(r, w[idx], C=256.0)) <= 1.0
                assert np.sum(g.imag == 0) == np.sum(w.imag == 0)


def test_oracle_complex_multishift_schedule_matches_reference_schedule(oracle):
    """The sequential twin of the GPU complex unit schedule (amrvw_algorithm_ms_c: dense trailing block + complex
    Hessenberg QR or the core-chasing sub-solve for the shifts) finds the roots of the reference's schedule."""
    from conftest import match_roots, root_condition_tolerance
    rows = []
    for i, d in enumerate([120, 97, 64, 33, 200]):
        r = oracle.fill_coeffs(77, i, 0, 2 * (d + 1))
        rows.append(r[0::2] + 1j * r[1::2])
    flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
    ref = oracle.roots_csr(flat, off)
    for L, reuse, small, dense in ((8, 2, 24, 1), (16, 1, 24, 1), (16, 2, 24, 0), (4, 1, 16, 1)):
        out, roff, nr, tot = oracle.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=dense)
        assert tot["fat_steps"] > 0 and tot["unconverged"] == 0
        for p, r in enumerate(rows):
            g = out[roff[p]:roff[p] + nr[p]]; w = ref[0][ref[1][p]:ref[1][p] + ref[2][p]]
            d, idx = match_roots(g, w)
            assert np.max(d / root_condition_tolerance(r, w[idx])) <= 1.0
