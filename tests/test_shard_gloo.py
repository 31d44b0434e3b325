"""World-size-2 (gloo, CPU) tests of the multi-GPU host logic.

The product's compute is CUDA-only; what runs here is the partition / assembly logic that the sharded paths use on the
host (`shard_range`, `shard_batch`, `assemble_slices` in tensortrainnumerics.jl_b200/api.py), driven by the NumPy oracle
as the per-rank compute so that the N>1 data flow is checked end to end on CPU:
  * cfg4 matvec: rank p contracts with the slice H[:, c_p, :] only; the all-gather of the slices must equal the
    unsharded K_matfree (src/solvers/dmrg.jl:239-244);
  * cfg5 batch: every vector is owned by exactly one rank, no collective on the data path.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ttn_b200 as t
        import ttn_oracle as o
        rng = np.random.default_rng(11)          # same inputs on every rank (replicated operands)
        w, chi_l, chi_r, nn = 3, 6, 7, 4         # chi_r = 7 is not divisible by 2: ragged slices
        G = rng.standard_normal((w, chi_l, chi_l)); H = rng.standard_normal((w, chi_r, chi_r))
        Am = rng.standard_normal((w, nn, nn, w)); V = rng.standard_normal((chi_l, nn, chi_r))
        c0, cp = t.shard_range(chi_r, rank, world)
        Yloc = o.dmrg_matvec2(G, Am, V, H[:, c0:c0 + cp, :], symmetrize=False)        # (chi_l, nn, cp)
        # all-gather of ragged slices: pad to the largest slice, gather, trim
        cpmax = max(t.shard_range(chi_r, r, world)[1] for r in range(world))
        pad = np.zeros((chi_l, nn, cpmax)); pad[:, :, :cp] = Yloc
        bufs = [torch.zeros(pad.shape, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(bufs, torch.from_numpy(pad))
        Y = t.assemble_slices([b.numpy()[:, :, :t.shard_range(chi_r, r, world)[1]] for r, b in enumerate(bufs)], axis=2)
        Yref = o.dmrg_matvec2(G, Am, V, H, symmetrize=False)
        err_mv = float(np.linalg.norm(Y - Yref) / np.linalg.norm(Yref))
        # cfg5: batch ownership
        total = 13
        f, n = t.shard_batch(total, rank, world)
        owned = torch.zeros(total, dtype=torch.int64)
        owned[f:f + n] += 1
        dist.all_reduce(owned)
        q.put((rank, err_mv, bool((owned == 1).all()), (c0, cp)))
    finally:
        dist.destroy_process_group()


def test_world2_partition_and_allgather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[3] for r in res] == [(0, 4), (4, 3)]
    for _, err, ok, _ in res:
        assert err < 1e-14
        assert ok


def test_shard_range_covers_index_exactly_once():
    import ttn_b200 as t
    for chi in (1, 5, 64, 1000, 1024):
        for n in (1, 2, 3, 4, 8):
            if n > chi:
                continue
            seen = np.zeros(chi, dtype=int)
            sizes = []
            for r in range(n):
                c0, cp = t.shard_range(chi, r, n)
                seen[c0:c0 + cp] += 1
                sizes.append(cp)
            assert (seen == 1).all() and max(sizes) - min(sizes) <= 1
    with pytest.raises(AssertionError):
        t.shard_range(8, 2, 2)
