"""CPU tests of the plumbing bench.py uses at N > 1: `synth.shard_range` splits one batch
contiguously by polynomial and a world_size-2 gloo group gathers per-rank results on rank 0 -- no data-path
collective.  No product code runs here (there is no GPU): the per-rank "compute" is the oracle, so these
tests can only catch a broken partition or gather.  The product's own multi-GPU path (amrvw_create_multi)
is covered by tests/test_capi_symbols.py (the split, on CPU) and tests/test_gpu_multi.py (two GPUs)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import amrvw_jl_b200  # noqa: F401  (shim)
    from importlib import import_module
    synth = import_module("amrvw_jl_b200.synth")
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, n = 11, 24
    lo, hi = synth.shard_range(B, rank, world)
    rows = synth.fill_coeffs(9, np.arange(lo, hi), 0, n + 1)
    mine = [(lo + i, orc.roots(r)[0]) for i, r in enumerate(rows)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # host gather; never on the timed path
    dist.barrier()
    if rank == 0:
        q.put(sorted((i, r.tolist()) for part in gathered for i, r in part))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from importlib import import_module
    import amrvw_jl_b200  # noqa: F401
    synth = import_module("amrvw_jl_b200.synth")
    for B in (0, 1, 7, 512, 3000):
        for w in (1, 2, 3, 8):
            parts = [synth.shard_range(B, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_world2_gloo_shard_range_and_gather_plumbing(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from importlib import import_module
    synth = import_module("amrvw_jl_b200.synth")
    assert [i for i, _ in res] == list(range(11))
    for i, r in res:
        want, _ = oracle.roots(synth.fill_coeffs(9, [i], 0, 25)[0])
        assert np.array_equal(np.array(r), want)


def test_bench_shards_one_batch_over_the_ranks():
    """bench.py --gpus N roots ONE config-2 batch split by polynomial (strong scaling): the ranks' inputs, concatenated
    in rank order, are the single-GPU batch, whatever N."""
    import importlib, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    synth = importlib.import_module("amrvw_jl_b200.synth")
    kind, batch, degree, law, seed = "rds_f64", 37, 12, 0, 2
    whole, _, off, ncoef = bench.make_inputs(kind, batch, degree, law, seed, 0)
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = synth.shard_range(batch, r, world)
            flat, _, o, _ = bench.make_inputs(kind, hi - lo, degree, law, seed, lo)
            assert o[-1] == (hi - lo) * ncoef
            parts.append(flat)
        assert np.array_equal(np.concatenate(parts), whole)
