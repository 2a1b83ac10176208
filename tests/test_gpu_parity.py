"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(ctypes over amrvw.jl_b200/libamrvw_b200.so); the oracle is only the checker.

Parity rule (north star / SURVEY.md 8c): roots matched by minimum-distance assignment,
|dz| <= tau * kappa(z) * max(1,|z|) with tau = C n u ||a||/|a_n|, and the number of exactly-real roots
(imag == 0) must agree exactly.  C is held to what is observed (profiles/r02_parity_margins.txt: the worst
needed constant over every parity test of a GPU run): C_RDS = 6 (observed 2.2), C_CSS = 24 (4.1), C_PENCIL =
16 (2.8) -- no more than 8x the observation, so the constants are measurements, not guesses.  For the one-bulge (reference-schedule) kernels
built without FMA contraction the real double shift path is additionally bit-identical to the
oracle, which itself reproduces the reference's printed digits."""
import json
import os

import numpy as np
import pytest

from conftest import match_roots, pencil_split, root_condition_tolerance, wilkinson_coeffs

pytestmark = pytest.mark.gpu

p4 = np.array([24.0, -50.0, 35.0, -10.0, 1.0])
p5 = np.array([-120.0, 274.0, -225.0, 85.0, -15.0, 1.0])
p6 = np.array([720.0, -1764.0, 1624.0, -735.0, 175.0, -21.0, 1.0])
pc = np.array([40.0 + 10.0j, -76.0 + 60.0j, 10.0 - 95.0j, 25.0 + 40.0j, -10.0 - 5.0j, 1.0 + 0.0j])
pc_rts = np.array([1j, 1 + 1j, 2 + 1j, 3 + 1j, 4 + 1j])
pc2 = np.array([1j, -1, 0, 0, -1j, 1])
SQEPS = np.sqrt(np.finfo(float).eps)
# tolerance constants of the parity rule; see check_parity and profiles/r02_parity_margins.txt for what is observed
C_RDS = 6.0          # observed <= 2.2 over 30 checks (0.67 at config 2's full size, 0.61 on the streamed 375-share, 2.2 on the degree classes)
C_CSS = 24.0         # observed <= 4.1
C_PENCIL = 16.0      # observed <= 2.8 (real and complex pencils)


def _g(rows):
    return np.array([complex(float(a), float(b)) for a, b in rows])


@pytest.fixture(scope="module")
def ctx(A):
    c = A.Context(seed=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_ref(A):
    c = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_strict(A):
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_REFERENCE)
    yield c
    c.close()


def batch(oracle, seed, B, n, kind=0, cx=False):
    w = 2 if cx else 1
    rows = [oracle.fill_coeffs(seed, p, kind, w * (n + 1)) for p in range(B)]
    if cx:
        rows = [r[0::2] + 1j * r[1::2] for r in rows]
    return rows


MARGIN_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_margins.jsonl")


def check_parity(coeff_rows, got, want, C=None, real_counts=True):
    """Every root within C n u ||a||/|a_n| kappa(z) max(1,|z|) of its minimum-distance partner, equal counts of
    exactly-real roots.  The worst observed |dz|/tol (in units of the tolerance with C = 1, so the number is
    the constant the data needs) is printed and appended to gpurun_out/parity_margins.jsonl: the constants
    passed by the callers are held to at most 8x what profiles/r02_parity_margins.txt records."""
    worst = 0.0
    C = C_RDS if C is None else C
    for a, g, w in zip(coeff_rows, got, want):
        assert len(g) == len(w)
        d, idx = match_roots(g, w)
        tol1 = root_condition_tolerance(a, w[idx], C=1.0)
        need = np.max(d / np.maximum(tol1, 1e-300))       # the C this polynomial needs
        worst = max(worst, need)
        assert need <= C, f"root parity violated: max |dz|/tol = {need / C:.3g} (needs C = {need:.3g}, allowed {C})"
        if real_counts:
            assert np.sum(g.imag == 0) == np.sum(w.imag == 0)
    name = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    line = {"test": name, "polynomials": len(coeff_rows), "degree": int(max(len(a) for a in coeff_rows)) - 1, "needed_C": worst, "allowed_C": C}
    print("parity margin:", json.dumps(line))
    try:
        os.makedirs(os.path.dirname(MARGIN_LOG), exist_ok=True)
        with open(MARGIN_LOG, "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass
    return worst


# ------------------------------------------------------------------ golden vectors through the C ABI
def test_gpu_golden_p4(A, ctx_strict, golden):
    r = A.roots(p4, ctx=ctx_strict)
    want = _g(golden["p4"]["roots"])
    assert np.array_equal(r.real, want.real) and np.all(r.imag == 0)     # the reference's printed digits, bit for bit


def test_gpu_golden_pencil_p4(A, ctx_strict, golden):
    vs, ws = A.basic_pencil(p4)
    r = A.roots(vs, ws, ctx=ctx_strict)
    want = _g(golden["p4_pencil"]["roots"])
    assert np.array_equal(r.real, want.real) and np.all(r.imag == 0)


def test_gpu_golden_pc(A, ctx_strict, golden):
    r = A.roots(pc2, ctx=ctx_strict)
    want = _g(golden["pc"]["roots"])
    d, _ = match_roots(r, want)
    assert d.max() <= 4e-16


@pytest.mark.parametrize("which", ["default", "reference"])
def test_gpu_known_answers(A, ctx, ctx_ref, which):
    c = ctx if which == "default" else ctx_ref
    for p in (p4, p5, p6):
        r = A.roots(p, ctx=c)
        assert np.allclose(r.real, np.arange(1, len(p)), rtol=SQEPS) and np.all(r.imag == 0)
        rr = A.real_polynomial_roots(p, ctx=c)
        assert np.allclose(rr.real_roots, np.arange(1, len(p))) and len(rr.complex_roots) == 0
        vs, ws = A.basic_pencil(p)
        r = A.roots(vs, ws, ctx=c)
        assert np.allclose(r.real, np.arange(1, len(p)), rtol=SQEPS)
    r = A.roots(pc, ctx=c)
    assert np.all(np.abs(r - pc_rts) <= SQEPS)
    vs, ws = A.basic_pencil(pc)
    assert np.allclose(np.round(A.roots(vs, ws, ctx=c), 12), pc_rts)
    for i in range(1, 5):
        vs, ws = pencil_split(p4, i)
        assert np.linalg.norm(A.roots(vs, ws, ctx=c) - np.array([1, 2, 3, 4])) <= SQEPS


def test_gpu_simple_cases_and_trimming(A, ctx):    # test/test-roots.jl:18-108
    cases = [(np.array([]), np.array([])), (np.array([1.0]), np.array([])), (np.array([1.0, 1.0]), np.array([-1.0])),
             (np.array([1.0, 2.0, 1.0]), np.array([-1.0, -1.0]))]
    for cx in (False, True):
        inputs, wants = [], []
        for ps, rts in cases:
            ps = ps.astype(complex) if cx else ps
            z2 = np.zeros(2, dtype=ps.dtype)
            pad = np.concatenate([rts, [0, 0]]) if len(ps) else rts
            inputs += [ps, np.concatenate([z2, ps]), np.concatenate([ps, z2]), np.concatenate([z2, ps, z2])]
            wants += [rts, pad, rts, pad]
        got = A.roots(inputs, ctx=ctx)             # ragged batch, including empty and all-zero inputs
        for g, w in zip(got, wants):
            assert len(g) == len(w) and np.all(g == w)
        for ps, rts in cases[2:]:
            ps = ps.astype(complex) if cx else ps
            z2 = np.zeros(2, dtype=ps.dtype)
            for inp, want in [(ps, rts), (np.concatenate([z2, ps]), np.concatenate([rts, [0, 0]])), (np.concatenate([ps, z2]), rts),
                              (np.concatenate([z2, ps, z2]), np.concatenate([rts, [0, 0]]))]:
                vs, ws = A.basic_pencil(inp)
                g = A.roots(vs, ws, ctx=ctx)
                assert len(g) == len(want) and np.all(g == want)
    # zero roots are appended after the sorted non-zero ones (companion.jl:266-268)
    r = A.roots(np.concatenate([[0.0, 0.0], p4]), ctx=ctx)
    assert np.allclose(r[:4].real, [1, 2, 3, 4]) and np.all(r[4:] == 0)


# ------------------------------------------------------------------ bit parity of the reference-schedule kernels
def test_gpu_rds_bitwise_vs_oracle(A, ctx_strict, oracle):
    """One bulge at a time, no FMA contraction: the GPU's real double shift path performs the same
    IEEE operations in the same order as the oracle -> identical bits (inputs that take an
    exceptional shift are excluded: sincos differs between libm and CUDA in the last bit)."""
    rows = batch(oracle, 11, 48, 60) + batch(oracle, 12, 8, 301) + [p4, p5, p6]
    flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
    out, roff, nr, _, pp = oracle.roots_csr(flat, off, per_poly=True)
    got = A.roots(rows, ctx=ctx_strict)
    n_checked = 0
    for p, g in enumerate(got):
        if pp[p]["exceptional"]:
            continue
        w = out[roff[p]:roff[p] + nr[p]]
        assert np.array_equal(g.view(np.float64), w.view(np.float64)), f"polynomial {p} differs"
        n_checked += 1
    assert n_checked >= 50
    assert ctx_strict.last_stats["turnovers"] > 0


def test_gpu_pencil_bitwise_vs_oracle(A, ctx_strict, oracle):
    rows = batch(oracle, 13, 24, 40)
    vs, ws = zip(*[pencil_split(r, len(r) // 2) for r in rows])
    got = A.roots(list(vs), list(ws), ctx=ctx_strict)
    for v, w, g in zip(vs, ws, got):
        want, st = oracle.roots(v, ws=w)
        if st["exceptional"]:
            continue
        assert np.array_equal(g.view(np.float64), want.view(np.float64))


@pytest.mark.parametrize("L,reuse,small,dense", [(4, 1, 16, 1), (8, 2, 16, 1), (16, 2, 24, 1), (16, 2, 24, 0), (16, 1, 40, 1), (32, 2, 40, 1), (32, 4, 24, 0), (32, 4, 24, 1)])
def test_gpu_pipelined_bitwise_vs_sequential_multishift(A, oracle, L, reuse, small, dense):
    """The pipelined kernel (L bulges in flight, three indices apart, data handed lane to lane by
    shuffles) against the same multishift schedule executed one bulge after the other on the CPU
    (oracle amrvw_algorithm_ms): bulges three apart commute exactly, so without FMA contraction the
    two must agree bit for bit -- this pins the systolic data flow, the shift computation (dense
    trailing block + Hessenberg QR, or the core-chasing sub-solve) and the small-window path."""
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_reuse=reuse, small_window=small,
                  shift_solver=A.SHIFTS_DENSE if dense else A.SHIFTS_CHASE)
    try:
        rows = batch(oracle, 31, 10, 150) + batch(oracle, 32, 5, 421) + batch(oracle, 33, 6, 64) + batch(oracle, 34, 3, 30) + [p4, p5, p6]
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        out, roff, nr, tot, pp = oracle.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=dense, per_poly=True)
        got = A.roots(rows, ctx=c)
        n_checked = 0
        for p, g in enumerate(got):
            if pp[p]["exceptional"]:
                continue
            w = out[roff[p]:roff[p] + nr[p]]
            assert np.array_equal(g.view(np.float64), w.view(np.float64)), f"polynomial {p} (degree {len(rows[p]) - 1}) differs"
            n_checked += 1
        assert n_checked >= 20
        assert c.last_stats["fat_steps"] == tot["fat_steps"] and c.last_stats["turnovers"] == tot["turnovers"]
    finally:
        c.close()


# ------------------------------------------------------------------ tolerance parity, all variants, both schedules
@pytest.mark.parametrize("n,B", [(100, 64), (333, 16), (1000, 4)])
@pytest.mark.parametrize("which", ["default", "reference"])
def test_gpu_rds_parity(A, ctx, ctx_ref, oracle, n, B, which):
    c = ctx if which == "default" else ctx_ref
    rows = batch(oracle, 1, B, n)                       # C1-style: N(0,1) coefficients
    want = [oracle.roots(r)[0] for r in rows]
    got = A.roots(rows, ctx=c)
    check_parity(rows, got, want, C=C_RDS)
    assert c.last_stats["unconverged"] == 0


@pytest.mark.parametrize("n,B,kind", [(100, 32, 1), (500, 8, 1), (200, 8, 0)])
def test_gpu_css_parity(A, ctx, oracle, n, B, kind):   # C3-style: U[0,1) + i U[0,1), and complex normal
    rows = batch(oracle, 3, B, n, kind=kind, cx=True)
    want = [oracle.roots(r)[0] for r in rows]
    got = A.roots(rows, ctx=ctx)
    check_parity(rows, got, want, C=C_CSS, real_counts=False)


@pytest.mark.parametrize("n,B", [(100, 32), (500, 8)])
def test_gpu_pencil_parity(A, ctx, oracle, n, B):      # C4-style: basic_pencil and a mid split
    rows = batch(oracle, 4, B, n)
    for split in ("basic", "half"):
        pairs = [A.basic_pencil(r) if split == "basic" else pencil_split(r, n // 2) for r in rows]
        vs, ws = zip(*pairs)
        want = [oracle.roots(v, ws=w)[0] for v, w in pairs]
        got = A.roots(list(vs), list(ws), ctx=ctx)
        check_parity(rows, got, want, C=C_PENCIL)


def test_gpu_c5_style_uniform_shifted(A, ctx, oracle):  # C5 input law U[0,1) - 1/2 at reduced degree
    rows = batch(oracle, 5, 8, 600, kind=2)
    want = [oracle.roots(r)[0] for r in rows]
    got = A.roots(rows, ctx=ctx)
    check_parity(rows, got, want, C=C_RDS)


def test_gpu_ragged_batch(A, ctx, oracle):
    degs = [3, 4, 5, 7, 16, 33, 64, 100, 129, 257, 2, 1, 0]
    rows = [oracle.fill_coeffs(21, i, 0, d + 1) for i, d in enumerate(degs)]
    got = A.roots(rows, ctx=ctx)
    for r, g in zip(rows, got):
        w, _ = oracle.roots(r)
        assert len(g) == len(w)
        if len(r) > 3:
            check_parity([r], [g], [w], C=C_RDS)
        else:
            assert np.array_equal(g, w)


def test_gpu_exceptional_shift_inputs(A, ctx, ctx_ref, oracle):   # x^n - 1, x^64 + 1, Wilkinson-20, (x-1)^8
    p1 = np.zeros(21); p1[0] = -1; p1[-1] = 1
    p2 = np.zeros(65); p2[0] = 1; p2[-1] = 1
    for c in (ctx, ctx_ref):
        r = A.roots(p1, ctx=c)
        d, _ = match_roots(r, np.exp(2j * np.pi * np.arange(20) / 20))
        assert d.max() < 1e-12
        r = A.roots(p2, ctx=c)
        d, _ = match_roots(r, np.exp(1j * np.pi * (2 * np.arange(64) + 1) / 64))
        assert d.max() < 1e-12
    r = A.roots(wilkinson_coeffs(20), ctx=ctx_ref)
    assert ctx_ref.last_stats["exceptional"] >= 1
    assert len(r) == 20 and np.all(np.isfinite(r.view(np.float64)))
    r8 = A.roots(np.poly(np.ones(8))[::-1].copy(), ctx=ctx)
    assert np.max(np.abs(r8 - 1)) < 0.1


def test_gpu_wilkinson_golden(A, ctx_strict, ctx_ref, golden):
    """Wilkinson-20 (examples/amrvw.ipynb cell 73): condition numbers ~1e13, so the printed digits
    are only reproducible with identical rounding: the no-FMA build follows the oracle up to the
    exceptional shift (sincos differs in the last bit) and lands within 1e-5 of the reference's
    output, with the same 12 real + 4 complex pairs; the FMA build is checked by backward error."""
    a = wilkinson_coeffs(20)
    r = A.roots(a, ctx=ctx_strict)
    want = _g(golden["wilkinson20"]["roots"])
    d, _ = match_roots(r, want)
    assert d.max() < 1e-5 and np.sum(r.imag == 0) == 12
    r2 = A.roots(a, ctx=ctx_ref)
    assert horner_backward_error(a, r2).max() < 1e-13


# ------------------------------------------------------------------ full-size, size-independent properties
def horner_backward_error(a, z):
    """max over roots of |p(z)| / sum |a_k||z|^k, evaluated stably for |z| > 1 via the reversed polynomial."""
    a = np.asarray(a, dtype=np.complex128)
    z = np.asarray(z, dtype=np.complex128)
    out = np.empty(len(z))
    inside = np.abs(z) <= 1
    for mask, coef, pts in ((inside, a, z[inside]), (~inside, a[::-1], 1.0 / z[~inside])):
        if pts.size == 0:
            continue
        num = np.polyval(coef[::-1], pts)
        den = np.polyval(np.abs(coef[::-1]), np.abs(pts))
        out[mask] = np.abs(num) / den
    return out


def test_gpu_full_size_c2_properties(A, ctx, oracle):
    """BASELINE config 2 at full degree (n = 3000), a slice of the batch: every polynomial converges,
    residuals are at backward-error level, conjugate pairing is exact, real-root statistics follow
    README.md:121-127; and 64 polynomials are checked root by root against the oracle (minimum-distance matching,
    condition-scaled tolerance, exact counts of real roots)."""
    n, B = 3000, 96
    rows = batch(oracle, 2, B, n)
    got = A.roots(rows, ctx=ctx)
    assert ctx.last_stats["unconverged"] == 0
    counts = []
    for r, g in zip(rows, got):
        assert len(g) == n
        be = horner_backward_error(r, g)
        assert be.max() < 1e-10
        counts.append(np.sum(g.imag == 0))
        pos = np.sort_complex(g[g.imag > 0]); neg = np.sort_complex(np.conj(g[g.imag < 0]))
        assert np.array_equal(pos, neg)            # exact conjugate pairs (qdrtc, utils.jl:130-140)
        keys = g.real + 0  # sorted by (re, im)
        assert np.all(np.diff(keys) >= 0)
    expect = 2 / np.pi * np.log(n) + 0.625738072 + 2 / (np.pi * n)
    counts = np.array(counts)
    assert abs(counts.mean() - expect) < 4 * counts.std() / np.sqrt(B) + 0.2
    K = 64
    flat = np.concatenate(rows[:K]); off = np.arange(K + 1, dtype=np.int64) * (n + 1)
    out, roff, nr, _ = oracle.roots_csr(flat, off)       # the reference's schedule, all host cores
    want = [out[roff[p]:roff[p] + nr[p]] for p in range(K)]
    check_parity(rows[:K], got[:K], want, C=C_RDS)


def test_gpu_count_real_roots_dev(A, ctx, oracle):
    import torch
    rows = batch(oracle, 8, 40, 120)
    flat = np.concatenate(rows); off = np.arange(41, dtype=np.int64) * 121
    dev = torch.device("cuda:0")
    d_c = torch.from_numpy(flat).to(dev); d_o = torch.from_numpy(off).to(dev)
    d_r = torch.zeros(2 * int(off[-1]), dtype=torch.float64, device=dev)
    d_n = torch.zeros(40, dtype=torch.int32, device=dev); d_u = torch.zeros(40, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(120, dtype=torch.int32, device=dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.roots_dev("rds_f64", d_c.data_ptr(), d_o.data_ptr(), 40, int(off[-1]), 121, d_r.data_ptr(), d_n.data_ptr(), d_u.data_ptr())
    ctx.count_real_roots_dev(d_r.data_ptr(), d_o.data_ptr(), 40, 0, d_n.data_ptr(), d_cnt.data_ptr())
    torch.cuda.synchronize()
    ctx.set_stream(None)
    cnt = d_cnt.cpu().numpy().reshape(40, 3)
    for p, r in enumerate(rows):
        w, _ = oracle.roots(r)
        assert cnt[p, 0] == np.sum(w.imag == 0) and cnt[p, 1] == cnt[p, 2] and cnt[p].sum() == 120


def test_gpu_fp64_peak_probe(ctx):
    tf = ctx.fp64_peak_tflops()
    assert 5.0 < tf < 80.0


def test_gpu_fast_primitives_selftest(ctx):
    """Division-free turnover and Newton rsqrt on device (test/test-amrvw-basics.jl:35-42: the turnover
    preserves the product; here to 8 eps like the oracle's version of the test)."""
    r = ctx.selftest(400000)
    eps = np.finfo(float).eps
    assert r["rsqrt_rel_err"] <= 2 * eps, r
    assert r["turnover_product_err"] <= 8 * eps, r
    assert r["unit_norm_err"] <= 8 * eps, r


def test_gpu_fast_complex_turnover_selftest(ctx):
    """The division-free complex turnover against the IEEE one (the reference's formulas,
    transformations.jl:158-229) on random triples, every 7th with a sine down to 1e-18: same rotators to
    a few ulp, product preserved (test/test-amrvw-basics.jl:79-86), unit norm."""
    r = ctx.selftest_complex(400000)
    eps = np.finfo(float).eps
    assert r["fast_vs_ieee"] <= 16 * eps, r
    assert r["turnover_product_err"] <= 12 * eps, r
    assert r["unit_norm_err"] <= 8 * eps, r
    assert r["general_path"] <= 4, r


# ------------------------------------------------------------------ complex single shift on the unit queue
@pytest.mark.parametrize("L,reuse,solver", [(4, 0, 0), (8, 0, 0), (16, 0, 0), (32, 0, 0), (16, 1, 0), (32, 1, 0), (8, 4, 0), (16, 2, 1), (8, 1, 1)])
def test_gpu_css_pipelined_lanes(A, oracle, L, reuse, solver):
    """The pipelined complex path for every lane-group width: bulges two indices apart, the creation
    phase of each bulge carried by its lane (diagonal.jl:199-225 without the pass over Q), deferred
    small windows.  Ragged batch so that groups of one warp see different windows; both schedules of
    the library must agree with the oracle's reference schedule root by root."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_reuse=reuse, shift_solver=solver)   # 0 dense, 1 chase
    try:
        degs = [40, 41, 64, 97, 130, 200, 257, 33, 5, 3, 2, 1, 350]
        rows = []
        for i, d in enumerate(degs):
            r = oracle.fill_coeffs(77, i, 0, 2 * (d + 1))
            rows.append(r[0::2] + 1j * r[1::2])
        want = [oracle.roots(r)[0] for r in rows]
        got = A.roots(rows, ctx=c)
        big = [i for i, d in enumerate(degs) if d > 2]
        check_parity([rows[i] for i in big], [got[i] for i in big], [want[i] for i in big], C=C_CSS, real_counts=False)
        for i, d in enumerate(degs):
            if d <= 2:
                assert np.allclose(got[i], want[i], rtol=1e-14, atol=0)
        assert c.last_stats["unconverged"] == 0 and c.last_stats["fat_steps"] > 0
    finally:
        c.close()


def test_gpu_css_structured_inputs(A, ctx, ctx_ref):
    """Complex inputs that stagnate or deflate exactly: x^n - i (exceptional shifts, create-bulge.jl:85-93),
    a polynomial with trimmed zero ends, real coefficients passed as complex (conjugate pairs must still
    be found, without the real path's exact pairing)."""
    n = 48
    p1 = np.zeros(n + 1, dtype=complex); p1[0] = -1j; p1[-1] = 1
    truth = np.exp(1j * (np.pi / 2 + 2 * np.pi * np.arange(n)) / n)
    for c in (ctx, ctx_ref):
        r = A.roots(p1, ctx=c)
        d, _ = match_roots(r, truth)
        assert d.max() < 1e-12
    rng = np.random.default_rng(5)
    core = rng.standard_normal(61) + 1j * rng.standard_normal(61)
    p2 = np.concatenate([np.zeros(3), core, np.zeros(2)])
    r = A.roots(p2, ctx=ctx)
    assert len(r) == 63 and np.all(r[-3:] == 0)
    d, _ = match_roots(r[:-3], np.roots(core[::-1]))
    assert d.max() < 1e-10
    p3 = rng.standard_normal(201).astype(complex)
    r = A.roots(p3, ctx=ctx)
    d, _ = match_roots(r, np.roots(p3[::-1]))
    assert d.max() < 1e-9


def test_gpu_full_size_c3_properties(A, ctx, oracle):
    """BASELINE config 3 at full size (2000 complex polynomials of degree 1000) through the C ABI:
    size-independent properties -- every root slot filled, nothing unconverged, Horner backward error at
    rounding level on every polynomial, and oracle parity on a sample of the batch."""
    B, n = 2000, 1000
    rows = [None] * B
    for p in range(B):
        r = oracle.fill_coeffs(3, p, 1, 2 * (n + 1))
        rows[p] = r[0::2] + 1j * r[1::2]
    got = A.roots(rows, ctx=ctx)
    assert ctx.last_stats["unconverged"] == 0
    worst = 0.0
    for a, g in zip(rows[::8], got[::8]):        # np.polyval is O(n^2) per polynomial: every 8th one
        assert len(g) == n and np.all(np.isfinite(g.view(np.float64)))
        worst = max(worst, horner_backward_error(a, g).max())
    for g in got:
        assert len(g) == n and np.all(np.isfinite(g.view(np.float64)))
    assert worst < 1e-10, worst
    sample = list(range(0, B, 250))
    want = [oracle.roots(rows[p])[0] for p in sample]
    check_parity([rows[p] for p in sample], [got[p] for p in sample], want, C=C_CSS, real_counts=False)


# ------------------------------------------------------------------ companion pencil on the unit queue
@pytest.mark.parametrize("L,solver", [(4, 0), (8, 0), (16, 0), (32, 0), (8, 2), (16, 2)])
def test_gpu_pencil_pipelined_lanes(A, oracle, L, solver):
    """roots(vs, ws) on the pipelined path for every lane-group width and both shift solvers (dense
    V W^{-1} block + Hessenberg QR, and the core-chasing sub-solve): five chains, 11 turnovers per chase
    step (pencil.jl:63-71).  Ragged batch with trimmed zeros (utils.jl:29-58) and closed-form degrees;
    roots must agree with the oracle's one-bulge schedule root by root, with equal real-root counts."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_solver=solver)
    try:
        degs = [40, 41, 64, 97, 130, 200, 257, 33, 5, 3, 2, 1, 350]
        rows = [oracle.fill_coeffs(44, i, 0, d + 1) for i, d in enumerate(degs)]
        for split in ("basic", "half"):
            pairs = [A.basic_pencil(r) if split == "basic" or len(r) <= 3 else pencil_split(r, (len(r) - 1) // 2) for r in rows]
            vs, ws = zip(*pairs)
            want = [oracle.roots(v, ws=w)[0] for v, w in pairs]
            got = A.roots(list(vs), list(ws), ctx=c)
            big = [i for i, d in enumerate(degs) if d > 2]
            check_parity([rows[i] for i in big], [got[i] for i in big], [want[i] for i in big], C=C_PENCIL)
            for i, d in enumerate(degs):
                if d <= 2:
                    assert np.allclose(got[i], want[i], rtol=1e-14, atol=0)
            assert c.last_stats["unconverged"] == 0 and c.last_stats["fat_steps"] > 0
    finally:
        c.close()


def test_gpu_pencil_known_answers_pipelined(A, ctx):
    """test/test-roots.jl:146-188 through the default (pipelined) schedule: basic_pencil and every
    pencil_split of p4, p5, p6; a pencil with zero leading/trailing entries."""
    for p, want in ((p4, [1, 2, 3, 4]), (p5, [1, 2, 3, 4, 5]), (p6, [1, 2, 3, 4, 5, 6])):
        for i in range(0, len(p) - 1):
            vs, ws = A.basic_pencil(p) if i == 0 else pencil_split(p, i)
            r = A.roots(vs, ws, ctx=ctx)
            assert np.all(r.imag == 0) and np.allclose(np.sort(r.real), want, rtol=1e-9, atol=0)
    rng = np.random.default_rng(9)
    core = rng.standard_normal(81)
    vs, ws = A.basic_pencil(np.concatenate([np.zeros(2), core]))
    r = A.roots(vs, ws, ctx=ctx)
    assert len(r) == 82 and np.all(r[-2:] == 0)
    d, _ = match_roots(r[:-2], np.roots(core[::-1]))
    assert d.max() < 1e-9


def test_gpu_full_size_c4_properties(A, ctx, oracle):
    """BASELINE config 4 at full size (1000 pencils of degree 500, basic split and mid split) through the
    C ABI: nothing unconverged, Horner backward error at rounding level, exact conjugate pairing, and
    oracle parity on a sample of the batch."""
    B, n = 1000, 500
    rows = batch(oracle, 4, B, n)
    for split in ("basic", "half"):
        pairs = [A.basic_pencil(r) if split == "basic" else pencil_split(r, n // 2) for r in rows]
        vs, ws = zip(*pairs)
        got = A.roots(list(vs), list(ws), ctx=ctx)
        assert ctx.last_stats["unconverged"] == 0 and ctx.last_stats["fat_steps"] > 0
        for a, g in zip(rows[::4], got[::4]):
            assert len(g) == n
            assert horner_backward_error(a, g).max() < 1e-10
            pos = np.sort_complex(g[g.imag > 0]); neg = np.sort_complex(np.conj(g[g.imag < 0]))
            assert np.array_equal(pos, neg)
        sample = list(range(0, B, 125))
        want = [oracle.roots(pairs[p][0], ws=pairs[p][1])[0] for p in sample]
        check_parity([rows[p] for p in sample], [got[p] for p in sample], want, C=C_PENCIL)


# ------------------------------------------------------------------ chained warps (small batches of large polynomials)
@pytest.mark.parametrize("L,reuse,small,dense", [(64, 2, 72, 1), (64, 4, 40, 1), (128, 4, 72, 1), (128, 4, 72, 0), (256, 8, 72, 1)])
def test_gpu_chain_bitwise_vs_sequential_multishift(A, oracle, L, reuse, small, dense):
    """chain.cuh: one CTA per polynomial, 2 or 4 warps chained into one systolic array of 64 / 128 bulges
    (hand-off through shared memory, one bar.sync per lock-step).  Without FMA contraction it must agree
    bit for bit with the oracle's sequential multishift schedule at the same L -- windows both shorter
    (fill/drain loop only, fewer bulges than lanes) and longer (lean loop) than 3 L."""
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_reuse=reuse, small_window=small,
                  shift_solver=A.SHIFTS_DENSE if dense else A.SHIFTS_CHASE)
    try:
        rows = batch(oracle, 31, 4, 150) + batch(oracle, 32, 3, 700) + batch(oracle, 35, 2, 1300) + [p4, p6]
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        out, roff, nr, tot, pp = oracle.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=dense, per_poly=True)
        got = A.roots(rows, ctx=c)
        n_checked = 0
        for p, g in enumerate(got):
            if pp[p]["exceptional"]:
                continue
            w = out[roff[p]:roff[p] + nr[p]]
            assert np.array_equal(g.view(np.float64), w.view(np.float64)), f"polynomial {p} (degree {len(rows[p]) - 1}) differs"
            n_checked += 1
        assert n_checked >= 8
        if not any(x["exceptional"] for x in pp):    # an exceptional shift draws sincos(rand): last-bit differences change the path
            assert c.last_stats["fat_steps"] == tot["fat_steps"] and c.last_stats["turnovers"] == tot["turnovers"]
    finally:
        c.close()


@pytest.mark.parametrize("L", [64, 128, 256])
def test_gpu_chain_parity(A, oracle, L):
    """Product build of the chained kernel against the oracle's reference (one-bulge) schedule: roots by
    minimum-distance matching within the conditioning-scaled tolerance, equal real-root counts."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L)
    try:
        rows = batch(oracle, 5, 3, 1500, kind=2) + batch(oracle, 2, 2, 900)
        want = [oracle.roots(r)[0] for r in rows]
        got = A.roots(rows, ctx=c)
        check_parity(rows, got, want, C=C_RDS)
        assert c.last_stats["unconverged"] == 0 and c.last_stats["fat_steps"] > 0
    finally:
        c.close()


def test_gpu_robust_turnover_selftest(ctx):
    """The fall-back turnover of the multishift schedules on triples whose sines reach the denormal range
    (a fat step passes bulges through rotators that converged under earlier bulges of the same step; the
    reference deflates first, AMRVW_algorithm.jl:181-202): product preserved to rounding, unit norm, while
    the reference's quotient b = s1 s2 / s5 (transformations.jl:188-191) loses the product on such inputs."""
    r = ctx.selftest_robust(400000)
    assert r["triples"] == 400000
    assert r["robust_product_err"] < 8e-16 and r["unit_norm_err"] < 8e-16
    assert r["reference_formula_product_err"] > 1e-6      # the failure this path exists for


@pytest.mark.parametrize("L,B,n,kind", [(128, 24, 4000, 0), (128, 16, 4000, 2), (32, 16, 2000, 2)])
def test_gpu_many_bulges_backward_error(A, oracle, L, B, n, kind):
    """Up to 128 bulges per fat step pass through rotators that have already converged: Horner backward
    error of every polynomial stays at rounding level (with the reference's turnover formulas as the
    fall-back 8 of 96 such polynomials came back with backward error 1e-2)."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L)
    try:
        rows = batch(oracle, 12 if kind == 0 else 11, B, n, kind=kind)
        got = A.roots(rows, ctx=c)
        assert c.last_stats["unconverged"] == 0
        worst = max(horner_backward_error(r, g).max() for r, g in zip(rows, got))
        assert worst < 1e-10, worst
    finally:
        c.close()


def test_gpu_css_many_bulges_backward_error(A, oracle):
    """Complex single shift with 32 bulges per fat step on long windows: backward error at rounding level on
    every polynomial (complex twin of the fall-back turnover for vanished sines)."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=32)
    try:
        rows = batch(oracle, 13, 12, 2000, kind=1, cx=True)
        got = A.roots(rows, ctx=c)
        assert c.last_stats["unconverged"] == 0
        worst = max(horner_backward_error(r, g).max() for r, g in zip(rows, got))
        assert worst < 1e-10, worst
    finally:
        c.close()


def test_gpu_full_size_c5_properties(A, ctx, oracle):
    """BASELINE config 5, one GPU's share at full size (64 polynomials of degree 10 000, U[0,1) - 1/2): the
    automatic choice is four chained warps per polynomial (128 bulges per fat step).  Nothing unconverged,
    every root finite, conjugate pairing exact, sorted by real part, and Horner backward error at rounding
    level on a sample (np.polyval is O(n^2) per polynomial)."""
    B, n = 64, 10000
    rows = batch(oracle, 5, B, n, kind=2)
    got = A.roots(rows, ctx=ctx)
    assert ctx.last_stats["unconverged"] == 0 and ctx.last_stats["lanes"] == 128
    for g in got:
        assert len(g) == n and np.all(np.isfinite(g.view(np.float64)))
        pos = np.sort_complex(g[g.imag > 0]); neg = np.sort_complex(np.conj(g[g.imag < 0]))
        assert np.array_equal(pos, neg)
        assert np.all(np.diff(g.real) >= 0)
    worst = max(horner_backward_error(rows[p], got[p]).max() for p in range(0, B, 8))
    assert worst < 1e-9, worst
    # two of them root by root against the oracle, at the only size where 128 bulges per fat step run on degree 10 000
    K = 2
    flat = np.concatenate(rows[:K]); off = np.arange(K + 1, dtype=np.int64) * (n + 1)
    out, roff, nr, _ = oracle.roots_csr(flat, off)
    want = [out[roff[p]:roff[p] + nr[p]] for p in range(K)]
    check_parity(rows[:K], got[:K], want, C=C_RDS)


def test_gpu_chain_more_polynomials_than_resident_ctas(A, ctx, oracle):
    """chain.cuh takes polynomials blockIdx.x, blockIdx.x + gridDim.x, ...: a batch larger than the number of
    resident CTAs (148 SMs x <= 3) must give the same roots as the unit kernel, polynomial by polynomial."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=64)
    try:
        rows = batch(oracle, 17, 700, 120)
        got = A.roots(rows, ctx=c)
        ref = A.roots(rows, ctx=ctx)
        assert c.last_stats["unconverged"] == 0 and c.last_stats["lanes"] == 64
        for p in range(0, 700, 7):
            d, _ = match_roots(got[p], ref[p])
            assert d.max() < 1e-9 and np.sum(got[p].imag == 0) == np.sum(ref[p].imag == 0)
        want = [oracle.roots(rows[p])[0] for p in (0, 333, 699)]
        check_parity([rows[p] for p in (0, 333, 699)], [got[p] for p in (0, 333, 699)], want, C=C_RDS)
    finally:
        c.close()


@pytest.mark.parametrize("L,reuse,small,dense", [(4, 1, 16, 1), (8, 2, 16, 1), (16, 2, 24, 1), (16, 2, 24, 0), (32, 2, 40, 1)])
def test_gpu_pencil_pipelined_bitwise_vs_sequential_multishift(A, oracle, L, reuse, small, dense):
    """The pencil unit kernel (five chains, 11 turnovers per step, dense V W^{-1} trailing block) against the
    same multishift schedule executed one bulge after the other on the CPU (oracle amrvw_algorithm_ms on the
    pencil state, dense_trailing_block_pencil): without FMA contraction the two must agree bit for bit."""
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_reuse=reuse, small_window=small,
                  shift_solver=A.SHIFTS_DENSE if dense else A.SHIFTS_CHASE)
    try:
        rows = batch(oracle, 41, 8, 150) + batch(oracle, 42, 4, 321) + batch(oracle, 43, 5, 64) + batch(oracle, 44, 3, 30) + [p4, p5, p6]
        pairs = [A.basic_pencil(r) if i % 2 == 0 or len(r) <= 3 else pencil_split(r, (len(r) - 1) // 2) for i, r in enumerate(rows)]
        vs = np.concatenate([v for v, _ in pairs]); ws = np.concatenate([w for _, w in pairs])
        off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(v) for v, _ in pairs], out=off[1:])
        out, roff, nr, tot, pp = oracle.roots_csr(vs, off, ws=ws, sched=1, L=L, small=small, reuse=reuse, dense=dense, per_poly=True)
        got = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=c)
        n_checked = 0
        for p, g in enumerate(got):
            if pp[p]["exceptional"]:
                continue
            w = out[roff[p]:roff[p] + nr[p]]
            assert np.array_equal(g.view(np.float64), w.view(np.float64)), f"pencil {p} (degree {len(rows[p]) - 1}) differs"
            n_checked += 1
        assert n_checked >= 18
        if not any(x["exceptional"] for x in pp):
            assert c.last_stats["fat_steps"] == tot["fat_steps"] and c.last_stats["turnovers"] == tot["turnovers"]
    finally:
        c.close()


@pytest.mark.parametrize("L,reuse,small", [(8, 2, 24), (16, 2, 24), (16, 1, 24), (32, 2, 40)])
def test_gpu_css_pipelined_vs_sequential_multishift(A, oracle, L, reuse, small):
    """The complex unit kernel (no-FMA build) against its sequential twin in the oracle (amrvw_algorithm_ms_c:
    same fat steps, shifts from the same dense block + complex Hessenberg QR, bulges one after the other with
    the reference's phase push-down).  hypot / complex division / complex sqrt come from different math
    libraries on the two sides, so this is not a bit comparison: same number of fat steps within 2 %, roots
    within the conditioning-scaled tolerance at a constant 16x tighter than against the reference schedule."""
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_reuse=reuse, small_window=small,
                  shift_solver=A.SHIFTS_DENSE)
    try:
        rows = []
        for i, d in enumerate([150, 200, 97, 64, 33, 257, 5, 321]):
            r = oracle.fill_coeffs(77, i, 0, 2 * (d + 1))
            rows.append(r[0::2] + 1j * r[1::2])
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        out, roff, nr, tot, pp = oracle.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=1, per_poly=True)
        got = A.roots(rows, ctx=c)
        want = [out[roff[p]:roff[p] + nr[p]] for p in range(len(rows))]
        check_parity(rows, got, want, C=4.0, real_counts=False)
        assert c.last_stats["unconverged"] == 0
        assert abs(c.last_stats["fat_steps"] - tot["fat_steps"]) <= 0.02 * tot["fat_steps"] + 1
        assert abs(c.last_stats["turnovers"] - tot["turnovers"]) <= 0.02 * tot["turnovers"]
    finally:
        c.close()


# ------------------------------------------------------------------ round 2: the gaps VERDICT r01 named
def test_gpu_empty_vector_in_the_middle_of_a_batch(A, ctx, ctx_ref, oracle):
    """roots(S[]) = S[] (test/test-roots.jl:18-27) inside a batch: an empty coefficient vector owns no root
    slot, so it must not shift its neighbours' results (slot offsets[p] - p overlapped)."""
    for c in (ctx, ctx_ref):
        for cx in (False, True):
            q5 = (p5 + 0j) if cx else p5
            q4 = (p4 + 0j) if cx else p4
            empty = np.zeros(0, dtype=q5.dtype)
            for seq in ([q5, empty, q5], [empty, q5], [q5, empty], [empty, empty, q4, empty, q5, empty]):
                got = A.roots(seq, ctx=c)
                for a, g in zip(seq, got):
                    if len(a) == 0:
                        assert len(g) == 0
                    else:
                        assert np.allclose(g.real, np.arange(1, len(a)), rtol=1e-9, atol=0) and np.all(np.abs(g.imag) <= 1e-9), (seq, g)
    rows = batch(oracle, 21, 6, 60)
    seq = [rows[0], np.zeros(0), rows[1], np.zeros(0), np.zeros(0), rows[2], np.zeros(1), rows[3], rows[4], np.zeros(0), rows[5]]
    got = A.roots(seq, ctx=ctx)
    alone = A.roots(rows, ctx=ctx)
    k = 0
    for a, g in zip(seq, got):
        if len(a) > 1:
            assert np.array_equal(g, alone[k])       # same kernels, same per-polynomial work: the neighbours change nothing
            k += 1
        else:
            assert len(g) == 0
    vs, ws = A.basic_pencil(p5)
    gp = A.roots([vs, np.zeros(0), vs], [ws, np.zeros(0), ws], ctx=ctx)
    assert len(gp[1]) == 0 and np.allclose(gp[0].real, [1, 2, 3, 4, 5]) and np.allclose(gp[2].real, [1, 2, 3, 4, 5])


def test_gpu_chain_regime_c2_split_over_8_gpus(A, ctx, oracle):
    """The north-star split: 3000 polynomials of degree 3000 over 8 GPUs = 375 per GPU, the regime where the
    automatic choice is chained warps (128 bulges per fat step).  One GPU's share through the default path:
    nothing unconverged, exact conjugate pairs, and 16 polynomials root by root against the oracle."""
    n, B = 3000, 375
    rows = batch(oracle, 2, B, n)
    got = A.roots(rows, ctx=ctx)
    assert ctx.last_stats["unconverged"] == 0 and ctx.last_stats["lanes"] >= 64
    for g in got[::25]:
        pos = np.sort_complex(g[g.imag > 0]); neg = np.sort_complex(np.conj(g[g.imag < 0]))
        assert len(g) == n and np.array_equal(pos, neg)
    sel = list(range(0, B, 24))[:16]
    flat = np.concatenate([rows[p] for p in sel]); off = np.arange(len(sel) + 1, dtype=np.int64) * (n + 1)
    out, roff, nr, _ = oracle.roots_csr(flat, off)
    want = [out[roff[i]:roff[i] + nr[i]] for i in range(len(sel))]
    check_parity([rows[p] for p in sel], [got[p] for p in sel], want, C=C_RDS)


@pytest.mark.parametrize("n,B", [(60, 24), (300, 6)])
def test_gpu_complex_pencil_random_parity(A, ctx, oracle, n, B):
    """roots(vs, ws) for ComplexF64 (pencil.jl:63-80 with the complex single shift) on random inputs against the
    oracle: basic_pencil and a mid split of random complex polynomials (the reference tests only known answers)."""
    rows = batch(oracle, 23, B, n, kind=1, cx=True)
    for split in ("basic", "mid"):
        pairs = [A.basic_pencil(r) if split == "basic" else pencil_split(r, n // 2) for r in rows]
        got = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=ctx)
        want = [oracle.roots(v, ws=w)[0] for v, w in pairs]
        check_parity(rows, got, want, C=C_PENCIL, real_counts=False)


def test_gpu_reference_constants_as_options(A, oracle):
    """IT_COUNT, it_max = 20 n and tol = eps are options of the context (create-bulge.jl:9, AMRVW_algorithm.jl:25,184)
    with the reference's values as defaults; non-convergence is not an error: missing roots stay 0 and are counted."""
    rows = batch(oracle, 31, 8, 80)
    base = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE)
    try:
        ref = A.roots(rows, ctx=base)
        explicit = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE, it_count=15, it_max_factor=20, tol=np.finfo(float).eps)
        same = A.roots(rows, ctx=explicit)
        assert all(np.array_equal(a, b) for a, b in zip(ref, same))      # the defaults ARE the reference's constants
        explicit.close()
        capped = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE, it_max_factor=1)
        out, roff, nr, nunc = capped.roots_csr(np.concatenate(rows), np.arange(9, dtype=np.int64) * 81)
        assert np.all(nunc > 0) and capped.last_stats["unconverged"] == int(nunc.sum())       # AMRVW_algorithm.jl:23-30
        for p in range(8):
            g = out[roff[p]:roff[p] + nr[p]]
            assert np.sum(g == 0) >= nunc[p]
        capped.close()
        loose = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE, tol=1e-9)
        got = A.roots(rows, ctx=loose)
        for g, w in zip(got, ref):
            d, _ = match_roots(g, w)
            assert d.max() < 1e-5
        st_loose = dict(loose.last_stats)
        loose.close()
        A.roots(rows, ctx=base)
        assert st_loose["turnovers"] <= base.last_stats["turnovers"]       # earlier deflation: no more work than with eps
        eager = A.Context(seed=0, schedule=A.SCHEDULE_REFERENCE, it_count=6)
        got = A.roots(rows, ctx=eager)
        assert eager.last_stats["exceptional"] > 0 and eager.last_stats["unconverged"] == 0, eager.last_stats      # create-bulge.jl:25-37 taken often
        for r, g in zip(rows, got):
            assert horner_backward_error(r, g).max() < 1e-11
        eager.close()
        piped = A.Context(seed=0, it_max_factor=1)                      # the unit schedule honours the cap too
        out, roff, nr, nunc = piped.roots_csr(np.concatenate(rows), np.arange(9, dtype=np.int64) * 81)
        assert piped.last_stats["unconverged"] == int(nunc.sum())
        piped.close()
        with pytest.raises(A.AmrvwError):
            A.Context(seed=0, tol=0.5)
    finally:
        base.close()


def test_gpu_count_real_roots_host_entry(A, ctx, oracle):
    """real_polynomial_roots for a batch with host buffers (complex_conjugate_root_theorem.jl:32-54, README.md:121-127):
    only 12 bytes per polynomial come back; the counts equal those of the roots call and of the oracle."""
    rows = batch(oracle, 8, 60, 150)
    flat = np.concatenate(rows); off = np.arange(61, dtype=np.int64) * 151
    counts, nunc = ctx.count_real_roots_csr(flat, off)
    got = A.roots(rows, ctx=ctx)
    assert np.all(nunc == 0)
    for p, r in enumerate(rows):
        w, _ = oracle.roots(r)
        assert counts[p, 0] == np.sum(got[p].imag == 0) == np.sum(w.imag == 0)
        assert counts[p, 1] == counts[p, 2] == np.sum(w.imag > 0) and counts[p].sum() == 150


@pytest.mark.parametrize("L,reuse,small,dense", [(128, 8, 40, 1), (128, 4, 72, 1), (64, 4, 40, 1), (16, 2, 24, 1), (16, 1, 40, 1), (32, 2, 40, 1)])
def test_gpu_stream_bitwise_vs_sequential_multishift(A, oracle, L, reuse, small, dense):
    """stream.cuh: chained warps fed by a stream of fat steps of different polynomials -- bulges created one lock-step
    early, the last chase step split over two lock-steps, jobs of several polynomials in the array at once, polynomials
    migrating between CTAs, phase A on service warps.  Per polynomial the operations and their order are unchanged, so
    without FMA contraction the roots must agree bit for bit with the oracle's sequential multishift schedule at the
    same L, whatever the interleaving (more polynomials than CTAs, windows shorter and longer than 3 L)."""
    c = A.Context(seed=0, strict=True, schedule=A.SCHEDULE_STREAMED, lanes_per_poly=L, shift_reuse=reuse, small_window=small,
                  shift_solver=A.SHIFTS_DENSE if dense else A.SHIFTS_CHASE)
    try:
        rows = batch(oracle, 31, 40, 150) + batch(oracle, 32, 30, 700) + batch(oracle, 35, 6, 1300) + [p4, p6, np.zeros(0), np.array([1.0, 2.0, 1.0])]
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        out, roff, nr, tot, pp = oracle.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=dense, per_poly=True)
        for rep in range(2):       # twice: the interleaving differs from run to run, the roots must not
            got = A.roots(rows, ctx=c)
            assert c.last_stats["lanes"] == L
            n_checked = 0
            for p, g in enumerate(got):
                if pp[p]["exceptional"]:
                    continue
                w = out[roff[p]:roff[p] + nr[p]]
                assert np.array_equal(g.view(np.float64), w.view(np.float64)), f"polynomial {p} (degree {len(rows[p]) - 1}) differs"
                n_checked += 1
            assert n_checked >= 70
            if not any(x["exceptional"] for x in pp):
                assert c.last_stats["fat_steps"] == tot["fat_steps"] and c.last_stats["turnovers"] == tot["turnovers"]
    finally:
        c.close()


def test_gpu_stream_parity_and_auto_choice(A, ctx, oracle):
    """Product build of the streamed kernel (the automatic choice for a few hundred large polynomials, e.g. one GPU's
    share of config 2 split over 8) against the oracle's reference schedule on a mixed batch."""
    c = A.Context(seed=0, schedule=A.SCHEDULE_STREAMED)
    try:
        rows = batch(oracle, 5, 6, 1500, kind=2) + batch(oracle, 2, 10, 900) + batch(oracle, 7, 30, 200)
        flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
        out, roff, nr, _ = oracle.roots_csr(flat, off)
        want = [out[roff[p]:roff[p] + nr[p]] for p in range(len(rows))]
        got = A.roots(rows, ctx=c)
        check_parity(rows, got, want, C=C_RDS)
        assert c.last_stats["unconverged"] == 0 and c.last_stats["fat_steps"] > 0 and c.last_stats["lanes"] == 128
    finally:
        c.close()


def test_gpu_large_ragged_batch_multi_cta_ordering(A, ctx, oracle):
    """More than 8192 polynomials: the ordering by degree runs as one launch per comparator stage over the whole grid
    instead of one CTA.  A polynomial's roots do not depend on its neighbours in the batch, so the big call must
    reproduce, bit for bit, the same polynomials rooted in chunks; a sample is checked against the oracle."""
    rng = np.random.default_rng(11)
    B = 20000
    degs = rng.integers(3, 41, size=B)
    rows = [oracle.fill_coeffs(51, p, 0, int(d) + 1) for p, d in enumerate(degs)]
    got = A.roots(rows, ctx=ctx)
    assert ctx.last_stats["unconverged"] == 0
    for lo in (0, 5000, 15000):
        part = A.roots(rows[lo:lo + 5000], ctx=ctx)
        odd = [i for i, (a, b) in enumerate(zip(got[lo:lo + 5000], part)) if not np.array_equal(a, b)]
        assert len(odd) <= 50                       # only an exceptional (random) shift depends on the position in the batch (seen: 0.2 % of small polynomials)
        for i in odd:
            d, _ = match_roots(got[lo + i], part[i])
            assert d.max() < 1e-8
    sel = list(range(0, B, 997))
    want = [oracle.roots(rows[p])[0] for p in sel]
    check_parity([rows[p] for p in sel], [got[p] for p in sel], want, C=C_RDS)


def test_gpu_ragged_degree_classes(A, ctx, oracle):
    """a ragged host batch is split into degree classes, each rooted by its own sub-context with the lane count
    its degree and size call for, side by side -- here hundreds of small polynomials (unit queue, few lanes) next to a few
    large ones (chained warps).  Every polynomial against the oracle; pinning the lane count switches the split off and
    must give the same roots up to the parity tolerance."""
    rng = np.random.default_rng(5)
    degs = np.concatenate([rng.integers(30, 61, size=300), rng.integers(300, 501, size=40), rng.integers(2600, 3001, size=6), [0, 1, 2]])
    rng.shuffle(degs)
    rows = [oracle.fill_coeffs(61, p, 0, int(d) + 1) if d >= 1 else np.zeros(int(d)) for p, d in enumerate(degs)]
    got = A.roots(rows, ctx=ctx)
    st = dict(ctx.last_stats)
    assert st["unconverged"] == 0 and st["lanes"] >= 64          # the large class runs on chained warps ...
    pinned = A.Context(seed=0, schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=16)
    try:
        ref = A.roots(rows, ctx=pinned)
        assert pinned.last_stats["lanes"] == 16                    # ... which one lane count for the whole batch cannot give
    finally:
        pinned.close()
    flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
    out, roff, nr, _ = oracle.roots_csr(flat, off)
    big = [p for p, r in enumerate(rows) if len(r) > 3]
    want = [out[roff[p]:roff[p] + nr[p]] for p in big]
    check_parity([rows[p] for p in big], [got[p] for p in big], want, C=C_RDS)
    check_parity([rows[p] for p in big], [ref[p] for p in big], want, C=C_RDS)
    for p, r in enumerate(rows):
        assert len(got[p]) == max(len(r) - 1, 0)
    # complex and pencil batches go through the same split
    crow = [oracle.fill_coeffs(62, p, 1, 2 * (int(d) + 1)) for p, d in enumerate([40] * 30 + [300] * 6 + [1400] * 2)]
    crow = [r[0::2] + 1j * r[1::2] for r in crow]
    cg = A.roots(crow, ctx=ctx)
    cw = [oracle.roots(r)[0] for r in crow]
    check_parity(crow, cg, cw, C=C_CSS, real_counts=False)
    prow = [oracle.fill_coeffs(63, p, 0, int(d) + 1) for p, d in enumerate([40] * 30 + [300] * 6 + [1400] * 2)]
    pairs = [A.basic_pencil(r) for r in prow]
    pg = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=ctx)
    pw = [oracle.roots(v, ws=w)[0] for v, w in pairs]
    check_parity(prow, pg, pw, C=C_PENCIL)


def test_gpu_unit_compaction_moves_polynomials_and_keeps_roots(A, oracle):
    """North star: "lanes are compacted as polynomials converge".  With two polynomials per warp and more units than resident
    warps, a unit whose first polynomial has finished takes over another half-empty unit's polynomial (rounds.cuh: phase A).  A
    polynomial's own sequence of fat steps does not depend on the unit that hosts it: the roots are bit-identical with the
    compaction forced off, and amrvw_debug_counters[47] counts the polynomials that moved."""
    rows = batch(oracle, 77, 4000, 96)
    flat = np.concatenate(rows); off = np.arange(len(rows) + 1, dtype=np.int64) * 97
    res = {}
    for flag in ("1", "0"):
        os.environ["AMRVW_COMPACT"] = flag
        try:
            c = A.Context(seed=0, lanes_per_poly=16, schedule=A.SCHEDULE_PIPELINED)
            try:
                out, roff, nroots, nunc = c.roots_csr(flat, off, want_stats=True)
                moved = c.debug_counters()[47]
                res[flag] = (out.copy(), nroots.copy(), int(nunc.sum()), int(moved), dict(c.last_stats))
            finally:
                c.close()
        finally:
            del os.environ["AMRVW_COMPACT"]
    on, offr = res["1"], res["0"]
    assert on[2] == 0 and offr[2] == 0
    assert on[3] > 100 and offr[3] == 0, (on[3], offr[3])
    assert np.array_equal(on[1], offr[1]) and np.array_equal(on[0].view(np.float64), offr[0].view(np.float64))
    assert on[4]["turnovers"] == offr[4]["turnovers"] and on[4]["fat_steps"] == offr[4]["fat_steps"]
    got = [on[0][roff[p]: roff[p] + on[1][p]] for p in range(0, 4000, 400)]
    want = [oracle.roots(rows[p])[0] for p in range(0, 4000, 400)]
    check_parity([rows[p] for p in range(0, 4000, 400)], got, want)
