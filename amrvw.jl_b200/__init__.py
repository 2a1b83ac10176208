"""amrvw.jl_b200 -- B200-native batched core-chasing polynomial root finder.

The directory name carries a dot (it is the reference's name, AMRVW.jl, plus the target), so it is
imported through the shim `amrvw_jl_b200.py` at the repository root:

    import amrvw_jl_b200 as A
    A.roots([24.0, -50.0, 35.0, -10.0, 1.0])          # -> 1, 2, 3, 4
    A.roots([randn(3001) for _ in range(3000)])        # one GPU call for the whole batch
"""
from ._lib import (AmrvwError, Options, Stats, EXPORTS, SCHEDULE_AUTO, SCHEDULE_PIPELINED, SCHEDULE_REFERENCE, SCHEDULE_STREAMED,
                   SHIFTS_AUTO, SHIFTS_CHASE, SHIFTS_DENSE, lib_path, load)
from .roots import (ComplexConjugateException, ComplexConjugateRootTheoremRoots, Context, basic_pencil, default_context,
                    real_polynomial_roots, roots, QRFactorization, qr_factorization, eigvals)
from . import build as _build


def build(**kw):
    return _build.build(**kw)


__all__ = ["roots", "qr_factorization", "eigvals", "QRFactorization", "basic_pencil", "real_polynomial_roots", "ComplexConjugateRootTheoremRoots", "ComplexConjugateException",
           "Context", "default_context", "AmrvwError", "Options", "Stats", "EXPORTS", "load", "lib_path", "build",
           "SCHEDULE_AUTO", "SCHEDULE_REFERENCE", "SCHEDULE_PIPELINED", "SCHEDULE_STREAMED", "SHIFTS_AUTO", "SHIFTS_CHASE", "SHIFTS_DENSE"]
