import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_notebook.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def A():
    """The product package (ctypes over the CUDA library).  Building is cheap when up to date."""
    import amrvw_jl_b200 as pkg
    pkg.build()
    return pkg


def wilkinson_coeffs(n=20):
    """Polynomials.poly(1.0:n) as Polynomials.jl computed it in Float64 (sequential products)."""
    c = np.zeros(n + 1)
    c[0] = 1.0
    for j in range(n):
        for i in range(j, -1, -1):
            c[i + 1] = c[i + 1] - (j + 1.0) * c[i]
    return c[::-1].copy()


def pencil_split(ps, n):
    """test/test-roots.jl:174-182"""
    ps = np.asarray(ps)
    vs = np.zeros(len(ps) - 1, dtype=ps.dtype)
    ws = np.zeros(len(ps) - 1, dtype=ps.dtype)
    vs[:n] = ps[:n]
    ws[n - 1:] = ps[n:]
    return vs, ws


def match_roots(a, b):
    """Minimum-distance assignment of roots a -> b (Hungarian); returns |a_i - b_pi(i)|, pi."""
    from scipy.optimize import linear_sum_assignment
    a = np.asarray(a); b = np.asarray(b)
    if len(a) <= 600:
        cost = np.abs(a[:, None] - b[None, :])
        r, c = linear_sum_assignment(cost)
        return cost[r, c], c
    # large: both are sorted by (re, im) up to perturbation; nearest-neighbour in a KD-tree, then
    # Hungarian on the (few) contested points
    from scipy.spatial import cKDTree
    tree = cKDTree(np.c_[b.real, b.imag])
    d, idx = tree.query(np.c_[a.real, a.imag])
    uniq, counts = np.unique(idx, return_counts=True)
    if len(uniq) == len(b):
        return d, idx
    contested = np.isin(idx, uniq[counts > 1])
    free_b = np.setdiff1d(np.arange(len(b)), idx[~contested])
    ia = np.where(contested)[0]
    cost = np.abs(a[ia][:, None] - b[free_b][None, :])
    r, c = linear_sum_assignment(cost)
    d = d.copy(); idx = idx.copy()
    d[ia[r]] = cost[r, c]; idx[ia[r]] = free_b[c]
    return d, idx


def root_condition_tolerance(coeffs, roots_, C=64.0):
    """|dz| <= tau * kappa(z) * max(1,|z|), kappa = sum|a_k||z|^k / (|z||p'(z)|),
    tau = C * n * u * ||a||_2 / |a_n|  (backward error grows linearly in ||a||, README.md:104-105).
    Evaluated in a scaled way (|z| > 1 uses the reversed polynomial) so z^n does not overflow."""
    a = np.asarray(coeffs, dtype=np.complex128)
    n = len(a) - 1
    u = 2.0 ** -53
    tau = C * n * u * np.linalg.norm(a) / abs(a[-1])
    z = np.asarray(roots_, dtype=np.complex128)
    tol = np.empty(len(z))
    absa = np.abs(a)
    k = np.arange(n + 1)
    for i, zi in enumerate(z):
        r = abs(zi)
        if r <= 1.0:
            pw = r ** k
            s = np.sum(absa * pw)
            dp = abs(np.polyval((a[1:] * k[1:])[::-1], zi))
            kap = s / max(r * dp, 1e-300)
        else:  # p(z) = z^n q(1/z): kappa is invariant under z -> 1/z with reversed coefficients
            w = 1.0 / zi
            rw = abs(w)
            b = a[::-1]
            pw = rw ** k
            s = np.sum(np.abs(b) * pw)
            dq = abs(np.polyval((b[1:] * k[1:])[::-1], w))
            kap = s / max(rw * dq, 1e-300)
        tol[i] = tau * kap * max(1.0, r)
    return tol
