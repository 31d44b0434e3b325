"""cfg4 two-site matvec sharded over the GPUs of one node: one process per GPU under torchrun.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_bench.py [--chi 1024]

Times (CUDA events on the library stream, max over ranks) the fused matvec + all-gather (tile stores to peer memory in
the GEMM epilogue, epoch-flag handshake) and, for comparison, the same local slice followed by an NCCL all-gather
(torch.distributed.all_gather_into_tensor).  Checks the gathered vector against a dense NumPy contraction at a small size first.
Prints one JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(torch, dist, rank, world, chi=1024, w=5, reps=20, krylovdim=8, with_nccl=True):
    """Sharded cfg4 matvec on an initialised process group (dist may be None for world == 1); returns the result dict
    (identical on every rank)."""
    class A:  # argument holder so the body below reads like the script
        pass
    args = A()
    args.chi, args.w, args.reps, args.krylovdim = chi, w, reps, krylovdim
    import ttn_b200 as t
    from ttn_b200 import _lib
    lib = _lib.lib()
    stream = torch.cuda.ExternalStream(t.stream_handle())

    def exchange(b):
        if world == 1:
            return [b]
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def barrier():
        t.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- parity at a small size against a dense NumPy contraction (every rank checks its complete gathered vector) ------------
    rng = np.random.default_rng(4)
    w, nn, cs = 3, 4, 40
    G = rng.standard_normal((w, cs, cs)); H = rng.standard_normal((w, cs, cs)); Am = rng.standard_normal((w, nn, nn, w))
    V = rng.standard_normal((cs, nn, cs))
    op = t.ShardedMatvec(G, Am, H, rank, world, exchange)
    Y = op.apply(V)
    barrier()
    Yref = np.einsum("yad,ybez,def,zcf->abc", G, Am, V, H, optimize=True)   # K_matfree, dmrg.jl:239-244 (one application)
    err = float(np.linalg.norm(Y - Yref) / np.linalg.norm(Yref))
    werr = op.error()
    barrier()
    op.free()
    errs = [None] * world
    if world > 1:
        dist.all_gather_object(errs, (err, werr))
    else:
        errs = [(err, werr)]

    # ---- cfg4 shapes ----------------------------------------------------------------------------------------------
    chi, w = args.chi, args.w
    rng = np.random.default_rng(4)
    G = np.asfortranarray(rng.standard_normal((w, chi, chi))); H = np.asfortranarray(rng.standard_normal((w, chi, chi)))
    Am = np.asfortranarray(rng.standard_normal((w, nn, nn, w))); V = np.asfortranarray(rng.standard_normal((chi, nn, chi)))
    flops = 4.0 * w * nn * chi ** 3 + 2.0 * w * w * nn * nn * chi ** 2
    dV = C.c_void_p()
    _lib.check(lib.ttn_dev_alloc(V.nbytes, C.byref(dV)))
    _lib.check(lib.ttn_h2d(dV, V.ctypes.data, V.nbytes))

    def time_loop(fn, reps):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            tm = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        return ms

    fused = t.ShardedMatvec(G, Am, H, rank, world, exchange)
    ms_fused = time_loop(lambda: fused.apply_dev(dV), args.reps)
    werr2 = fused.error()

    # NCCL variant: unbound operator (local slice only) + all_gather_into_tensor on torch's stream
    plain = t.ShardedMatvec(G, Am, H, rank, world, None)
    ms_nccl = None
    if with_nccl and world > 1 and chi % world == 0:
        n_el = chi * nn * chi
        full = torch.empty(n_el, dtype=torch.float64, device="cuda")
        lib_stream = stream

        def nccl_step():
            dY = plain.apply_dev(dV)
            ptr = dY.value + 8 * chi * nn * plain.c0     # the local slice inside the library-owned buffer
            with torch.cuda.stream(lib_stream):
                t_slice = _as_tensor(torch, ptr, chi * nn * plain.cp)
                dist.all_gather_into_tensor(full, t_slice)
        ms_nccl = time_loop(nccl_step, args.reps)
    ms_local = time_loop(lambda: plain.apply_dev(dV), args.reps)

    # ---- sharded Lanczos (one bond solve of the DMRG sweep: krylovdim matvecs + replicated vector algebra) -------
    dx = C.c_void_p()
    _lib.check(lib.ttn_dev_alloc(V.nbytes, C.byref(dx)))

    def eig_step():
        th, mv = C.c_double(), C.c_int()
        _lib.check(lib.ttn_shard_eigsolve(fused.h, dx, args.krylovdim, 1, 1e-10, C.byref(th), C.byref(mv)))
        return th.value
    _lib.check(lib.ttn_h2d(dx, V.ctypes.data, V.nbytes))
    ms_eig = time_loop(eig_step, 3)
    theta = eig_step()
    thetas = [None] * world
    if world > 1:
        dist.all_gather_object(thetas, theta)
    else:
        thetas = [theta]
    res = {
        "metric": "sharded local-matvec FP64 TFLOP/s", "n_gpus": world, "chi": chi, "w": w, "gflop": flops / 1e9,
        "fused_ms": ms_fused, "fused_tflops": flops / ms_fused / 1e9,
        "nccl_allgather_ms": ms_nccl, "nccl_tflops": (flops / ms_nccl / 1e9) if ms_nccl else None,
        "local_slice_only_ms": ms_local,
        "lanczos_bond_solve_ms": ms_eig, "krylovdim": args.krylovdim,
        "theta_identical_on_all_ranks": bool(all(x == thetas[0] for x in thetas)),
        "parity_small_rel_err_per_rank": [e[0] for e in errs], "wait_timeouts": [e[1] for e in errs] + [werr2],
        "exchange": "GEMM epilogue stores to peer buffers over NVLink (CUDA IPC), epoch flags; %d bytes per rank per matvec"
                    % (V.nbytes // world * (world - 1))}
    barrier()
    fused.free(); plain.free()
    _lib.check(lib.ttn_dev_free(dV)); _lib.check(lib.ttn_dev_free(dx))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--w", type=int, default=5)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--krylovdim", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    res = run(torch, dist if world > 1 else None, rank, world, args.chi, args.w, args.reps, args.krylovdim)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _as_tensor(torch, ptr, n):
    """zero-copy torch view of n doubles at device address ptr (library-owned memory)"""
    class _W:
        pass
    w = _W()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(w, device="cuda")


if __name__ == "__main__":
    main()
