"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun.  Every rank builds the sharded two-site operator
(csrc/shard.cu), runs the fused matvec + all-gather exchange over NVLink peer memory, and compares the COMPLETE gathered vector
with the oracle's K_matfree (`o.dmrg_matvec2`, src/solvers/dmrg.jl:239-244) — twice (both epoch parities) and for a ragged
bond (chi not divisible by the rank count); then the replicated Lanczos on the sharded operator against a dense eigensolve."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ttn_b200 as t
    import ttn_oracle as o

    def exchange(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    worst = 0.0
    for cs, w, cplx in ((40, 3, False), (37, 2, False), (24, 3, True)):
        rng = np.random.default_rng(4 + cs)
        nn = 4

        def rnd(*shape):
            a = rng.standard_normal(shape)
            return a + 1j * rng.standard_normal(shape) if cplx else a
        G, H, Am, V = rnd(w, cs, cs), rnd(w, cs, cs), rnd(w, nn, nn, w), rnd(cs, nn, cs)
        op = t.ShardedMatvec(G, Am, H, rank, world, exchange)
        ref = o.dmrg_matvec2(G, Am, V, H, symmetrize=False)
        for rep in range(3):                     # epochs 1, 2, 3: both result buffers, flag reuse
            Y = op.apply(V)
            err = float(np.linalg.norm(Y - ref) / np.linalg.norm(ref))
            worst = max(worst, err)
            assert err < 1e-12, (rank, cs, rep, err)
        assert op.error() == 0
        t.synchronize(); dist.barrier()
        op.free()
    # replicated Lanczos on the sharded operator: symmetric problem, lowest eigenvalue against the dense matrix
    rng = np.random.default_rng(11)
    cs, w, nn = 12, 2, 4
    G = rng.standard_normal((w, cs, cs)); G = G + np.transpose(G, (0, 2, 1))
    H = rng.standard_normal((w, cs, cs)); H = H + np.transpose(H, (0, 2, 1))
    Am = rng.standard_normal((w, nn, nn, w)); Am = Am + np.transpose(Am, (0, 2, 1, 3))
    K = np.einsum("yad,ybez,zcf->abcdef", G, Am, H, optimize=True).reshape(cs * nn * cs, cs * nn * cs, order="F")
    lam = np.linalg.eigvalsh(0.5 * (K + K.T))[0]
    op = t.ShardedMatvec(G, Am, H, rank, world, exchange)
    th, x, _ = op.eigsolve(rng.standard_normal((cs, nn, cs)), krylovdim=40, maxiter=60, tol=1e-12)
    assert abs(th - lam) < 1e-9 * max(1.0, abs(lam)), (th, lam)
    thetas = [None] * world
    dist.all_gather_object(thetas, float(th))
    assert all(v == thetas[0] for v in thetas), thetas       # bit-identical on every rank (no scalar all-reduce needed)
    t.synchronize(); dist.barrier()
    op.free()
    # the sharded matvec inside a DMRG sweep: Heisenberg XYZ d = 12, bond caps 8 / 32 / 64 (64 = exact at the centre), against the
    # dense ground-state energy (examples/heisenberg_xyz_dmrg.jl:9-19); energies identical on every rank
    d = 12
    Hh = o.heisenberg_xyz_tto(d, jx=1.1, jy=0.8, jz=1.2, lam=0.0)
    e0 = np.linalg.eigvalsh(np.real(o.tto_to_matrix(Hh)))[0]
    x0 = o.rand_tt((2,) * d, 8, rng=np.random.default_rng(3), normalise=True)
    sc = t.ShardContext(np.float64, 64 * 64 * 4, rank, world, exchange)
    for sym in (True, False):
        E, x, rh = t.dmrg_eigsolve(Hh, x0, N=2, tol=1e-12, sweep_schedule=[2, 4, 6], rmax_schedule=[8, 32, 64], linsolv_tol=1e-12,
                                   linsolv_maxiter=300, krylovdim=24, symmetrize=sym, shard=sc)
        assert abs(E[-1] - e0) < 1e-10 * abs(e0), (E[-1], e0)
        es = [None] * world
        dist.all_gather_object(es, [float(v) for v in E])
        assert all(v == es[0] for v in es), "energies differ between the ranks"
        E1, _, _ = t.dmrg_eigsolve(Hh, x0, N=2, tol=1e-12, sweep_schedule=[2, 4, 6], rmax_schedule=[8, 32, 64], linsolv_tol=1e-12,
                                   linsolv_maxiter=300, krylovdim=24, symmetrize=sym)
        assert abs(E1[-1] - E[-1]) < 1e-10 * abs(e0)
    t.synchronize(); dist.barrier()
    sc.free()
    if rank == 0:
        print(f"PARITY OK world={world} worst_rel_err={worst:.2e} theta={th:.12f} dmrg_sharded_E={E[-1]:.12f} dense={e0:.12f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
