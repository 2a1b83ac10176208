"""Host-side mirror of the reference's user API for the batched-roots path.

Same names and argument meaning as the reference (all under /root/reference/src):
    roots(ps)                      companion.jl:252-272   Float64 -> RDS, ComplexF64 -> CSS
    roots(vs, ws)                  companion.jl:282-303   companion pencil
    basic_pencil(ps)               companion.jl:44-54
    real_polynomial_roots(ps)      complex_conjugate_root_theorem.jl:32-54
plus the one thing the reference lacks: `roots` of a *sequence of coefficient vectors* roots them
all in one call on the GPU (the reference's batch is the comprehension of README.md:121).

Everything numeric happens in the CUDA library behind include/amrvw_b200.h; this file only packs
ragged inputs into CSR buffers and unpacks the result.  No CPU fallback.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import AmrvwError, Options, Stats


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data if a is not None else 0)


class Context:
    """One per host thread (amrvw_ctx).  `device`: one GPU (amrvw_create; -1 = current);
    `devices`: a list of GPUs, or "all" -- one context that splits every host-buffer batch over
    them (amrvw_create_multi)."""

    def __init__(self, device=-1, seed=0, strict=False, devices=None, lib_path=None, **options):
        self._lib = _lib.load(strict=strict, path=lib_path)
        self._h = ctypes.c_void_p()
        if devices is not None:
            devs = [] if devices == "all" else [int(d) for d in devices]
            arr = (ctypes.c_int * max(len(devs), 1))(*devs)
            rc = self._lib.amrvw_create_multi(ctypes.byref(self._h), arr, len(devs), ctypes.c_uint64(seed))
        else:
            rc = self._lib.amrvw_create(ctypes.byref(self._h), int(device), ctypes.c_uint64(seed))
        if rc != 0:
            self._h = ctypes.c_void_p()
            raise AmrvwError(f"amrvw_create failed ({rc}): no usable CUDA device; this package has no CPU path")
        self.last_stats = None
        self._seed = seed
        if options:
            self.set_options(seed=seed, **options)

    @property
    def device_count(self):
        return self._lib.amrvw_device_count(self._h)

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.amrvw_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise AmrvwError(f"amrvw error {rc}: {self._lib.amrvw_last_error(self._h).decode()}")

    def set_options(self, schedule=0, lanes_per_poly=0, shift_reuse=0, small_window=0, shift_solver=0, seed=0, it_count=0, it_max_factor=0, tol=0.0, stream_ctas_per_sm=0):
        """it_count / it_max_factor / tol: the reference's IT_COUNT = 15 (create-bulge.jl:9), it_max = 20 n
        (AMRVW_algorithm.jl:25) and deflation tolerance eps (AMRVW_algorithm.jl:184); 0 = those values."""
        o = Options(int(schedule), int(lanes_per_poly), int(shift_reuse), int(small_window), int(shift_solver), int(it_count), int(seed),
                    int(it_max_factor), int(stream_ctas_per_sm), float(tol))
        self._check(self._lib.amrvw_set_options(self._h, ctypes.byref(o)))

    def set_stream(self, cuda_stream_handle):
        """Run on the caller's stream; 0 is the legacy default stream.  None: back to the private stream."""
        if cuda_stream_handle is None:
            self._check(self._lib.amrvw_use_own_stream(self._h))
        else:
            self._check(self._lib.amrvw_set_stream(self._h, ctypes.c_void_p(int(cuda_stream_handle))))

    def fp64_peak_tflops(self):
        v = ctypes.c_double()
        self._check(self._lib.amrvw_measure_fp64_peak(self._h, ctypes.byref(v)))
        return v.value

    def selftest(self, n=200000):
        out = (ctypes.c_double * 4)()
        self._check(self._lib.amrvw_selftest(self._h, int(n), out))
        return {"rsqrt_rel_err": out[0], "turnover_product_err": out[1], "unit_norm_err": out[2], "general_path": out[3]}

    def selftest_complex(self, n=200000):
        out = (ctypes.c_double * 4)()
        self._check(self._lib.amrvw_selftest_complex(self._h, int(n), out))
        return {"fast_vs_ieee": out[0], "turnover_product_err": out[1], "unit_norm_err": out[2], "general_path": out[3]}

    def selftest_robust(self, n=200000):
        out = (ctypes.c_double * 4)()
        self._check(self._lib.amrvw_selftest_robust(self._h, int(n), out))
        return {"robust_product_err": out[0], "unit_norm_err": out[1], "reference_formula_product_err": out[2], "triples": out[3]}

    def last_profile(self):
        """Warp-cycle accounting of the last pipelined RDS call made with stats (amrvw_last_profile)."""
        out = (ctypes.c_int64 * 16)()
        self._check(self._lib.amrvw_last_profile(self._h, out))
        keys = ("cycles_units", "cycles_phase_a", "cycles_lean", "cycles_general", "steps_lean", "steps_general", "lean_calls", "unit_visits",
                "us_setup", "us_main", "cycles_driver", "us_small", "us_finish", "cycles_pop", "cycles_loop", "lane_cycles_special")
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def debug_counters(self):
        out = (ctypes.c_int64 * 48)()
        self._check(self._lib.amrvw_debug_counters(self._h, out))
        return [int(v) for v in out]

    # -- host buffers (CSR) in, host buffers out
    def roots_csr(self, coeffs, offsets, ws=None, want_stats=True):
        """coeffs: flat float64 or complex128 array; offsets: int64[B+1]; ws: pencil second vector.
        Returns (roots complex128 flat, root_offsets int64[B+1], nroots int32[B], n_unconverged int32[B])."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = len(offsets) - 1
        cx = np.iscomplexobj(coeffs) or (ws is not None and np.iscomplexobj(ws))
        dt = np.complex128 if cx else np.float64
        coeffs = np.ascontiguousarray(coeffs, dtype=dt)
        pencil = ws is not None
        if pencil:
            ws = np.ascontiguousarray(ws, dtype=dt)
            if ws.shape != coeffs.shape:
                raise ValueError("vs and ws must have the same layout")
        total = int(offsets[-1]) if B >= 0 and len(offsets) else 0
        root_off = offsets.copy()              # polynomial p's roots start at complex slot offsets[p]
        out = np.zeros(total + 1, dtype=np.complex128)
        nroots = np.zeros(max(B, 1), dtype=np.int32)
        nunc = np.zeros(max(B, 1), dtype=np.int32)
        st = Stats()
        stp = ctypes.byref(st) if want_stats else None
        L = self._lib
        if pencil:
            fn = L.amrvw_roots_pencil_c64 if cx else L.amrvw_roots_pencil_f64
            rc = fn(self._h, _ptr(coeffs), _ptr(ws), _ptr(offsets), B, _ptr(out), _ptr(nroots), _ptr(nunc), stp)
        else:
            fn = L.amrvw_roots_css_c64 if cx else L.amrvw_roots_rds_f64
            rc = fn(self._h, _ptr(coeffs), _ptr(offsets), B, _ptr(out), _ptr(nroots), _ptr(nunc), stp)
        self._check(rc)
        self.last_stats = st.as_dict() if want_stats else None
        return out, root_off, nroots[:B], nunc[:B]

    def count_real_roots_csr(self, coeffs, offsets, want_stats=False):
        """real_polynomial_roots counts for a Float64 batch, host buffers: int32[B, 3] =
        (imag == 0, imag > 0, imag < 0) per polynomial, and n_unconverged int32[B]."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = len(offsets) - 1
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        counts = np.zeros((max(B, 1), 3), dtype=np.int32)
        nunc = np.zeros(max(B, 1), dtype=np.int32)
        st = Stats()
        self._check(self._lib.amrvw_count_real_roots_rds_f64(self._h, _ptr(coeffs), _ptr(offsets), B, _ptr(counts), _ptr(nunc), ctypes.byref(st) if want_stats else None))
        self.last_stats = st.as_dict() if want_stats else None
        return counts[:B], nunc[:B]

    def eigvals_hessenberg(self, mats, unitary=False, want_stats=True):
        """Eigenvalues of a sequence of upper Hessenberg matrices (qr_factorization + eigvals, qrfactorization.jl:63-103) in one
        GPU call.  Returns (list of complex128 arrays sorted by (real, imag), n_unconverged int32[B])."""
        flat, dims = pack_matrices(mats)
        out, eoff, nunc = self.eigvals_hessenberg_packed(flat, dims, unitary=unitary, want_stats=want_stats)
        return [out[eoff[p]:eoff[p + 1]].copy() for p in range(len(dims))], nunc

    def eigvals_hessenberg_packed(self, flat, dims, unitary=False, want_stats=True):
        """The C call itself: `flat` holds the matrices back to back in column major order (float64 or complex128), matrix p is
        dims[p] x dims[p].  Returns (eigenvalues flat complex128, slot offsets int64[B+1], n_unconverged int32[B])."""
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        B = len(dims)
        cx = np.iscomplexobj(flat)
        flat = np.ascontiguousarray(flat, dtype=np.complex128 if cx else np.float64)
        if flat.size < int((dims.astype(np.int64) ** 2).sum()):
            raise ValueError("matrix buffer shorter than sum(dims^2)")
        if flat.size == 0:
            flat = np.zeros(1, dtype=flat.dtype)
        eoff = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(dims, out=eoff[1:])
        out = np.zeros(int(eoff[-1]) + 1, dtype=np.complex128)
        nunc = np.zeros(max(B, 1), dtype=np.int32)
        st = Stats()
        fn = self._lib.amrvw_eigvals_hessenberg_c64 if cx else self._lib.amrvw_eigvals_hessenberg_f64
        self._check(fn(self._h, _ptr(flat), _ptr(dims if B else np.zeros(1, dtype=np.int32)), B, int(bool(unitary)), _ptr(out), _ptr(nunc),
                       ctypes.byref(st) if want_stats else None))
        self.last_stats = st.as_dict() if want_stats else None
        return out, eoff, nunc[:B]

    def eigvals_hessenberg_dev(self, complex_, d_H, d_dims, d_moff, d_eoff, B, max_dim, d_eig, d_nunconv, unitary=False, want_stats=False):
        """Device pointers (ints) in and out, asynchronous on the context's stream unless want_stats."""
        st = Stats()
        fn = self._lib.amrvw_eigvals_hessenberg_c64_dev if complex_ else self._lib.amrvw_eigvals_hessenberg_f64_dev
        self._check(fn(self._h, d_H, d_dims, d_moff, d_eoff, B, int(max_dim), int(bool(unitary)), d_eig, d_nunconv, ctypes.byref(st) if want_stats else None))
        self.last_stats = st.as_dict() if want_stats else None

    # -- device pointers (ints) in/out: used by bench.py with torch-owned HBM buffers
    def roots_dev(self, kind, d_coeffs, d_offsets, B, total, max_len, d_roots, d_nroots, d_nunconv, d_ws=None, want_stats=False):
        st = Stats()
        stp = ctypes.byref(st) if want_stats else None
        L = self._lib
        P = ctypes.c_void_p
        if kind in ("pencil_f64", "pencil_c64"):
            fn = L.amrvw_roots_pencil_f64_dev if kind == "pencil_f64" else L.amrvw_roots_pencil_c64_dev
            rc = fn(self._h, P(d_coeffs), P(d_ws), P(d_offsets), B, total, max_len, P(d_roots), P(d_nroots), P(d_nunconv), stp)
        else:
            fn = {"rds_f64": L.amrvw_roots_rds_f64_dev, "css_c64": L.amrvw_roots_css_c64_dev}[kind]
            rc = fn(self._h, P(d_coeffs), P(d_offsets), B, total, max_len, P(d_roots), P(d_nroots), P(d_nunconv), stp)
        self._check(rc)
        self.last_stats = st.as_dict() if want_stats else None
        return self.last_stats

    def count_real_roots_dev(self, d_roots, d_offsets, B, shift, d_nroots, d_counts3):
        P = ctypes.c_void_p
        self._check(self._lib.amrvw_count_real_roots_dev(self._h, P(d_roots), P(d_offsets), B, int(shift), P(d_nroots), P(d_counts3)))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _is_batch(ps):
    if isinstance(ps, np.ndarray):
        return ps.ndim == 2 or ps.dtype == object
    if isinstance(ps, (list, tuple)):
        return len(ps) > 0 and isinstance(ps[0], (list, tuple, np.ndarray))
    return False


def _pack(seq):
    arrs = [np.asarray(a) for a in seq]
    cx = any(np.iscomplexobj(a) for a in arrs)
    dt = np.complex128 if cx else np.float64
    lens = np.array([a.shape[0] if a.ndim else 0 for a in arrs], dtype=np.int64)
    offsets = np.zeros(len(arrs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = np.concatenate([a.astype(dt, copy=False).reshape(-1) for a in arrs]) if len(arrs) and offsets[-1] > 0 else np.zeros(0, dtype=dt)
    return flat, offsets


def roots(ps, ws=None, ctx=None):
    """AMRVW.roots: `roots(ps)` or `roots(vs, ws)`; either a single coefficient vector
    (a0..an, lowest degree first) or a sequence of them (rooted in one GPU call).

    Float64 coefficients use the real double shift algorithm, ComplexF64 the complex single shift
    one (companion.jl:252-272); two arguments select the companion pencil (companion.jl:282-303).
    Returns complex128 roots sorted by (real, imag) with the zero roots of trimmed low-order zeros
    appended, exactly like the reference; a batch returns a list of such arrays.
    """
    ctx = ctx or default_context()
    batch = _is_batch(ps)
    seq = list(ps) if batch else [ps]
    flat, offsets = _pack(seq)
    wflat = None
    if ws is not None:
        wseq = list(ws) if batch else [ws]
        wflat, woff = _pack(wseq)
        if not np.array_equal(woff, offsets):
            raise ValueError("roots(vs, ws): vs and ws must have equal lengths")
        if np.iscomplexobj(wflat) != np.iscomplexobj(flat):
            flat = flat.astype(np.complex128)
            wflat = wflat.astype(np.complex128)
    out, roff, nroots, _ = ctx.roots_csr(flat, offsets, wflat)
    res = [out[roff[p]: roff[p] + nroots[p]].copy() for p in range(len(seq))]
    return res if batch else res[0]


def pack_matrices(mats):
    """Square matrices -> (flat column-major buffer, dims int32[B]) as amrvw_eigvals_hessenberg_* take them."""
    mats = [np.asarray(m) for m in mats]
    for m in mats:
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("qr_factorization needs square matrices")
    cx = any(np.iscomplexobj(m) for m in mats)
    dt = np.complex128 if cx else np.float64
    dims = np.array([m.shape[0] for m in mats], dtype=np.int32)
    if len(mats) and int((dims.astype(np.int64) ** 2).sum()):
        flat = np.concatenate([np.asfortranarray(m, dtype=dt).reshape(-1, order="F") for m in mats])
    else:
        flat = np.zeros(0, dtype=dt)
    return flat, dims


class QRFactorization:
    """What `AMRVW.qr_factorization(H; unitary)` returns (qrfactorization.jl:63-103), as far as this package needs it: the
    Hessenberg matrix and the kind of R.  The factorization itself is formed on the GPU inside `eigvals`."""

    def __init__(self, H, unitary=False):
        H = np.asarray(H)
        if H.ndim != 2 or H.shape[0] != H.shape[1]:
            raise ValueError("qr_factorization needs a square matrix")
        self.H = H
        self.unitary = bool(unitary)


def qr_factorization(H, unitary=False):
    """AMRVW.qr_factorization (qrfactorization.jl:63-66): H upper Hessenberg; `unitary=True` keeps R as a unitary diagonal."""
    return QRFactorization(H, unitary)


def eigvals(state, ctx=None):
    """eigvals(state) (AMRVW_algorithm.jl:207-216) of one QRFactorization or of a sequence of them (one GPU call; the
    sequence must share `unitary`).  Complex128 eigenvalues sorted by (real, imag)."""
    batch = isinstance(state, (list, tuple))
    seq = list(state) if batch else [state]
    if not all(isinstance(s, QRFactorization) for s in seq):
        raise TypeError("eigvals needs the result of qr_factorization")
    if len({s.unitary for s in seq}) > 1:
        raise ValueError("one call roots factorizations of one kind (unitary or not)")
    ctx = ctx or default_context()
    res, _ = ctx.eigvals_hessenberg([s.H for s in seq], unitary=seq[0].unitary if seq else False)
    return res if batch else res[0]


def basic_pencil(ps):
    """companion.jl:44-54: vs = [a0..a_{n-1}], ws = [0,..,0,a_n]."""
    ps = np.asarray(ps)
    qs = np.zeros(len(ps) - 1, dtype=ps.dtype)
    qs[-1] = ps[-1]
    return ps[:-1].copy(), qs


@dataclass
class ComplexConjugateRootTheoremRoots:
    """complex_conjugate_root_theorem.jl:7-10: real roots, and the upper-half-plane member of each
    conjugate pair."""
    real_roots: np.ndarray
    complex_roots: np.ndarray

    def __eq__(self, other):
        return np.array_equal(self.real_roots, other.real_roots) and np.array_equal(self.complex_roots, other.complex_roots)


class ComplexConjugateException(Exception):
    """complex_conjugate_root_theorem.jl:22-26."""

    def __init__(self, count_real, count_cc_negimag, count_cc_posimag):
        super().__init__(count_real, count_cc_negimag, count_cc_posimag)
        self.count_real, self.count_cc_negimag, self.count_cc_posimag = count_real, count_cc_negimag, count_cc_posimag


def real_polynomial_roots(coefs, ctx=None):
    """complex_conjugate_root_theorem.jl:32-54: classify by exact imag == 0 / imag > 0."""
    coefs = np.asarray(coefs, dtype=np.float64)
    rts = roots(coefs, ctx=ctx)
    real_roots = rts.real[rts.imag == 0]
    complex_roots = rts[rts.imag > 0]
    cn = int(np.sum(rts.imag < 0))
    if cn != len(complex_roots):
        raise ComplexConjugateException(len(real_roots), cn, len(complex_roots))
    return ComplexConjugateRootTheoremRoots(real_roots, complex_roots)
