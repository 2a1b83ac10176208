"""Build the sm_100a shared library in-tree (amrvw.jl_b200/libamrvw_b200.so).

nvcc cross-compiles without a GPU.  `strict=True` additionally builds libamrvw_b200_strict.so with
FMA contraction off (-fmad=false): same sources, used only by the bit-parity tests against the
oracle (which is built with -ffp-contract=off).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "capi.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc"))) if f.endswith((".cu", ".cuh"))] + [
    os.path.join(HERE, "..", "include", "amrvw_b200.h")
]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _stale(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(strict=True, force=False, verbose=False):
    outs = [(os.path.join(HERE, "libamrvw_b200.so"), [])]
    if strict:
        outs.append((os.path.join(HERE, "libamrvw_b200_strict.so"), ["-fmad=false", "-DAMRVW_STRICT"]))
    procs = []
    for out, extra in outs:
        if not force and not _stale(out):
            continue
        cmd = [_nvcc(), *ARCH, "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++", "-shared", "-Xcompiler", "-fPIC",
               *extra, "-o", out, SRC]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))     # the two builds are independent: side by side
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    return [o for o, _ in outs]


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
