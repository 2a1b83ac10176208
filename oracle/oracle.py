"""ctypes binding of the CPU oracle (oracle/amrvw_oracle.cpp).  TEST INFRASTRUCTURE: imported only
by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs -- never by the product."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class OStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in
                "turnovers sweeps exceptional trips unconverged fat_steps serial_turnovers pipe_steps".split()]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    src = os.path.join(HERE, "amrvw_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_state_size.restype = ctypes.c_int64
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def num_threads():
    return lib().orc_num_threads()


def roots_csr(coeffs, offsets, ws=None, seed=0, threads=0, sched=0, L=8, small=24, reuse=1, dense=0, per_poly=False):
    """Batched oracle roots.  Returns (flat complex128 roots, root offsets, nroots, total stats[, per-poly stats])."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    cx = np.iscomplexobj(coeffs) or (ws is not None and np.iscomplexobj(ws))
    dt = np.complex128 if cx else np.float64
    coeffs = np.ascontiguousarray(coeffs, dtype=dt)
    nroots = np.zeros(max(B, 1), dtype=np.int32)
    st = OStats()
    pp = (OStats * max(B, 1))() if per_poly else None
    if ws is None:
        roff = offsets.copy()          # polynomial p's roots start at complex slot offsets[p]
        out = np.zeros(int(offsets[-1]) + 1, dtype=np.complex128)
        lib().orc_roots_batch(_p(coeffs), _p(offsets), ctypes.c_int64(B), int(cx), _p(out), _p(nroots), ctypes.byref(st),
                              pp, ctypes.c_uint64(seed), int(threads), int(sched), int(L), int(small + 1000 * reuse + 100000 * dense))
    else:
        ws = np.ascontiguousarray(ws, dtype=dt)
        roff = offsets.copy()
        out = np.zeros(int(offsets[-1]) + 1, dtype=np.complex128)
        if sched and not cx:      # the multishift schedule of the GPU pencil unit kernel, executed sequentially (real only)
            lib().orc_roots_pencil_batch_ms(_p(coeffs), _p(ws), _p(offsets), ctypes.c_int64(B), _p(out), _p(nroots), ctypes.byref(st),
                                            pp, ctypes.c_uint64(seed), int(threads), int(sched), int(L), int(small + 1000 * reuse + 100000 * dense))
        else:
            lib().orc_roots_pencil_batch(_p(coeffs), _p(ws), _p(offsets), ctypes.c_int64(B), int(cx), _p(out), _p(nroots),
                                         ctypes.byref(st), ctypes.c_uint64(seed), int(threads))
    res = (out, roff, nroots[:B], st.as_dict())
    if per_poly:
        res = res + ([pp[i].as_dict() for i in range(B)],)
    return res


def roots(ps, ws=None, **kw):
    """Oracle roots(ps) / roots(vs, ws) of one polynomial -> (roots, stats)."""
    ps = np.asarray(ps)
    off = np.array([0, len(ps)], dtype=np.int64)
    out, roff, nr, st = roots_csr(ps, off, ws=None if ws is None else np.asarray(ws), **kw)
    return out[:nr[0]].copy(), st


def eigvals_hessenberg(H, unitary=False, seed=0, index=0):
    """Oracle eigvals(qr_factorization(H)) (qrfactorization.jl:63-103): H upper Hessenberg, real or complex -> (eigenvalues, stats)."""
    H = np.asarray(H)
    n = H.shape[0]
    cx = np.iscomplexobj(H)
    Hf = np.asfortranarray(H, dtype=np.complex128 if cx else np.float64)
    out = np.zeros(max(n, 1), dtype=np.complex128)
    st = OStats()
    lib().orc_eigvals_hessenberg(_p(Hf), int(n), int(cx), int(bool(unitary)), _p(out), ctypes.c_uint64(seed), ctypes.c_uint64(index), ctypes.byref(st))
    return out[:n].copy(), st.as_dict()


def fill_coeffs(seed, poly, kind, count):
    """Synthetic inputs shared by every leg: kind 0 = N(0,1), 1 = U[0,1), 2 = U[0,1) - 1/2."""
    out = np.zeros(count, dtype=np.float64)
    lib().orc_fill_coeffs(ctypes.c_uint64(seed), ctypes.c_uint64(poly), int(kind), ctypes.c_int64(count), _p(out))
    return out


# ---- primitives and factorization pieces (mirror test/test-amrvw-basics.jl)
def givensrot(a, b):
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        aa = np.array([complex(a).real, complex(a).imag]); bb = np.array([complex(b).real, complex(b).imag]); o = np.zeros(5)
        lib().orc_givensrot_cplx(_p(aa), _p(bb), _p(o))
        return complex(o[0], o[1]), o[2], complex(o[3], o[4])
    o = np.zeros(3)
    lib().orc_givensrot_real(ctypes.c_double(a), ctypes.c_double(b), _p(o))
    return o[0], o[1], o[2]


def turnover(q1, q2, q3):
    """q = (c, s); returns three (c, s)."""
    if any(isinstance(q[0], complex) for q in (q1, q2, q3)):
        i = np.array([v for q in (q1, q2, q3) for v in (complex(q[0]).real, complex(q[0]).imag, q[1])], dtype=np.float64); o = np.zeros(9)
        lib().orc_turnover_cplx(_p(i), _p(o))
        return [(complex(o[3 * k], o[3 * k + 1]), o[3 * k + 2]) for k in range(3)]
    i = np.array([v for q in (q1, q2, q3) for v in q], dtype=np.float64); o = np.zeros(6)
    lib().orc_turnover_real(_p(i), _p(o))
    return [(o[2 * k], o[2 * k + 1]) for k in range(3)]


def fuse(a, b):
    if isinstance(a[0], complex) or isinstance(b[0], complex):
        i = np.array([complex(a[0]).real, complex(a[0]).imag, a[1], complex(b[0]).real, complex(b[0]).imag, b[1]]); o = np.zeros(5)
        lib().orc_fuse_cplx(_p(i), _p(o))
        return (complex(o[0], o[1]), o[2]), complex(o[3], o[4])
    i = np.array([a[0], a[1], b[0], b[1]], dtype=np.float64); o = np.zeros(2)
    lib().orc_fuse_real(_p(i), _p(o))
    return (o[0], o[1]), 1.0


class FlatState:
    """amrvw(ps) / amrvw(vs, ws) as flat arrays + helpers to densify (tests only)."""

    def __init__(self, ps=None, vs=None, ws=None):
        self.pencil = vs is not None
        src = np.asarray(vs if self.pencil else ps)
        self.cx = np.iscomplexobj(src) or (ws is not None and np.iscomplexobj(ws))
        dt = np.complex128 if self.cx else np.float64
        if self.pencil:
            v = np.ascontiguousarray(vs, dtype=dt); w = np.ascontiguousarray(ws, dtype=dt)
            self.N = len(v)
            self.buf = np.zeros(lib().orc_state_size(self.N, int(self.cx), 1))
            lib().orc_factor_pencil(_p(v), _p(w), self.N, int(self.cx), _p(self.buf))
        else:
            a = np.ascontiguousarray(ps, dtype=dt)
            self.N = len(a) - 1
            self.buf = np.zeros(lib().orc_state_size(self.N, int(self.cx), 0))
            lib().orc_factor(_p(a), len(a), int(self.cx), _p(self.buf))

    def _chains(self):
        N, w = self.N, (3 if self.cx else 2)
        pos = 0
        out = {}

        def take(n):
            nonlocal pos
            blk = self.buf[pos:pos + n * w].reshape(n, w); pos += n * w
            c = blk[:, 0] + 1j * blk[:, 1] if self.cx else blk[:, 0]
            return c, blk[:, -1]
        out["Q"] = take(N - 1); out["B"] = take(N); out["Ct"] = take(N)
        if self.pencil:
            out["WB"] = take(N); out["WCt"] = take(N)
        if self.cx:
            def taked(n):
                nonlocal pos
                blk = self.buf[pos:pos + 2 * n].reshape(n, 2); pos += 2 * n
                return blk[:, 0] + 1j * blk[:, 1]
            out["DQ"] = taked(N); out["DR"] = taked(N + 1)
            if self.pencil:
                out["WD"] = taked(N + 1)
        return out

    @staticmethod
    def _rotm(c, s, i, n):
        M = np.eye(n, dtype=np.complex128)
        M[i - 1:i + 1, i - 1:i + 1] = [[c, s], [-s, np.conj(c)]]
        return M

    def dense_chain(self, name, n):
        """Product of the chain in its natural order: descending = X1 X2 ... , ascending (Ct) = XN ... X1."""
        c, s = self._chains()[name]
        M = np.eye(n, dtype=np.complex128)
        idx = range(1, len(c) + 1)
        order = reversed(idx) if name in ("Ct", "WCt") else idx
        for i in order:
            M = M @ self._rotm(c[i - 1], s[i - 1], i, n)
        return M

    def dense_R(self, which="V"):
        """Matrix(RF) of rfactorization.jl:172-192: Ct (B D) + Ct e1 y^T."""
        ch = self._chains(); n = self.N + 1
        Ct = self.dense_chain("Ct" if which == "V" else "WCt", n); Bm = self.dense_chain("B" if which == "V" else "WB", n)
        D = np.diag(ch["DR" if which == "V" else "WD"]) if self.cx else np.eye(n)
        e1 = np.zeros(n); e1[0] = 1; en1 = np.zeros(n); en1[-1] = 1
        rho = en1 @ Ct @ e1
        yt = -(1 / rho) * (en1 @ (Ct @ (Bm @ D)))
        return Ct @ (Bm @ D) + Ct @ np.outer(e1, yt)

    def dense(self):
        """Matrix(state) (qrfactorization.jl:39-53): (Q D_Q) R, top-left N x N block."""
        ch = self._chains(); n = self.N + 1
        Q = self.dense_chain("Q", n)
        DQ = np.diag(np.concatenate([ch["DQ"], [1.0]])) if self.cx else np.eye(n)
        if self.pencil:
            V = self.dense_R("V"); W = self.dense_R("W")
            R = np.eye(n, dtype=np.complex128)
            R[:-1, :-1] = V[:-1, :-1] @ np.linalg.inv(W[:-1, :-1])
        else:
            R = self.dense_R("V")
        return (Q @ DQ @ R)[:-1, :-1]

    def diagonal_block(self, j):
        A = np.zeros(8)
        lib().orc_diagonal_block(_p(self.buf), self.N, int(self.cx), int(self.pencil), int(j), _p(A))
        if self.cx:
            return (A[0::2] + 1j * A[1::2]).reshape(2, 2)
        return A[:4].reshape(2, 2)

    def bulge_step(self, ctr, seed=0):
        c = np.array(ctr, dtype=np.int32)
        lib().orc_bulge_step(_p(self.buf), self.N, int(self.cx), int(self.pencil), _p(c), ctypes.c_uint64(seed))
        return [int(x) for x in c]

    def ms_shift_check(self, ctr, m):
        """Dense trailing block of the window ctr (order m+1), its eigenvalues by the Hessenberg QR and by the core-chasing sub-solve."""
        c = np.array(ctr, dtype=np.int32)
        M = m + 1
        H = np.zeros(M * M); d = np.zeros(2 * M); e = np.zeros(2 * M)
        rc = lib().orc_ms_shift_check(_p(self.buf), self.N, _p(c), int(m), _p(H), _p(d), _p(e))
        return H.reshape(M, M), d[0::2] + 1j * d[1::2], e[0::2] + 1j * e[1::2], rc

    def eigvals(self, seed=0):
        out = np.zeros(2 * self.N)
        unconv = lib().orc_eigvals(_p(self.buf), self.N, int(self.cx), int(self.pencil), _p(out), ctypes.c_uint64(seed))
        return out[0::2] + 1j * out[1::2], unconv
